// Stand-in for dune-common's visibility.hh (absent from this image): the reference headers only need the macro.
#ifndef DUNE_EXPORT
#define DUNE_EXPORT __attribute__((visibility("default")))
#endif
