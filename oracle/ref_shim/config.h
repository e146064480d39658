/* oracle/ref_shim/config.h -- empty stand-in for DUNE's generated config.h (build glue of oracle/_ref, test infrastructure). */
