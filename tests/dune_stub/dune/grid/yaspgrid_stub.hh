// tests/dune_stub: the slice of YaspGrid< dim, EquidistantOffsetCoordinates >, GridPart, DiscreteFunctionSpace and
// AdaptiveDiscreteFunction that the binding touches (one rank).  Test infrastructure only.
#ifndef B200FEM_DUNE_STUB_YASPGRID_HH
#define B200FEM_DUNE_STUB_YASPGRID_HH
#include <array>
#include <cstddef>
#include <utility>
#include <vector>
namespace Dune {
template <int dim>
struct StubYaspGrid {
  static constexpr int dimension = dim;
  struct Torus { std::array<int, dim> d; int dims(int i) const { return d[i]; } };
  struct Coords { std::array<double, dim> o; double origin(int i) const { return o[i]; } };
  struct Level { Coords coords; };
  StubYaspGrid(std::array<double, dim> lo, std::array<double, dim> hi, std::array<int, dim> n) : n_(n) { for (int i = 0; i < dim; ++i) { size_[i] = hi[i] - lo[i]; level_.coords.o[i] = lo[i]; torus_.d[i] = 1; } }
  int maxLevel() const { return 0; }
  int levelSize(int, int i) const { return n_[i]; }
  const std::array<double, dim>& domainSize() const { return size_; }
  const Torus& torus() const { return torus_; }
  const Level* begin() const { return &level_; }
  std::array<int, dim> n_; std::array<double, dim> size_; Torus torus_; Level level_;
};
// the slice of an unstructured cube grid (ALUGrid< 2, 2, cube, conforming >) behind an adaptive leaf grid part that the binding touches:
// element iteration, geometry().corner( i ), indexSet().index / subIndex / size
template <int dim>
struct StubCubeGrid {
  static constexpr int dimension = dim;
  static constexpr bool b200Unstructured = true;       // the opt-in read by Dune::Fem::B200IsCartesian
  StubCubeGrid(std::vector<std::array<double, dim>> v, std::vector<std::array<int, 1 << dim>> c) : vertices(std::move(v)), cubes(std::move(c)) {}
  std::vector<std::array<double, dim>> vertices; std::vector<std::array<int, 1 << dim>> cubes;
};
namespace Fem {
template <class Grid>
struct StubLeafGridPart {
  typedef Grid GridType; static constexpr int dim = Grid::dimension;
  struct Geometry { const Grid* g; int e; std::array<double, dim> corner(int i) const { return g->vertices[(std::size_t)g->cubes[(std::size_t)e][(std::size_t)i]]; } };
  struct Entity { const Grid* g; int e; Geometry geometry() const { return Geometry{g, e}; } };
  struct Iterator { const Grid* g; int e; Entity operator*() const { return Entity{g, e}; } Iterator& operator++() { ++e; return *this; } bool operator!=(const Iterator& o) const { return e != o.e; } };
  struct IndexSet {
    const Grid* g;
    std::size_t size(int codim) const { return codim == 0 ? g->cubes.size() : g->vertices.size(); }
    std::size_t index(const Entity& en) const { return (std::size_t)en.e; }
    std::size_t subIndex(const Entity& en, int i, int) const { return (std::size_t)g->cubes[(std::size_t)en.e][(std::size_t)i]; }
  };
  struct Comm { int rank() const { return 0; } int size() const { return 1; } template <class T> void broadcast(T*, int, int) const {} };
  explicit StubLeafGridPart(const Grid& g) : g_(g), is_{&g} {}
  const Grid& grid() const { return g_; } const Comm& comm() const { return c_; } const IndexSet& indexSet() const { return is_; }
  template <int cd> Iterator begin() const { return Iterator{&g_, 0}; }
  template <int cd> Iterator end() const { return Iterator{&g_, (int)g_.cubes.size()}; }
  const Grid& g_; IndexSet is_; Comm c_;
};
// a Lagrange space over it: only the dof count is the space's own business here
template <class GridPart>
struct StubLagrangeSpace {
  typedef GridPart GridPartType;
  static constexpr int b200SpaceKind = 0;              // B200FEM_LAGRANGE
  static constexpr int localBlockSize = 1;
  StubLagrangeSpace(const GridPart& gp, int order, std::size_t size) : gp_(gp), order_(order), size_(size) {}
  const GridPart& gridPart() const { return gp_; }
  int order() const { return order_; }
  std::size_t size() const { return size_; }
  const GridPart& gp_; int order_; std::size_t size_;
};
struct StubComm { int rank() const { return 0; } int size() const { return 1; } template <class T> void broadcast(T*, int, int) const {} };
template <class Grid>
struct StubGridPart { typedef Grid GridType; explicit StubGridPart(const Grid& g) : g_(g) {} const Grid& grid() const { return g_; } const StubComm& comm() const { return c_; } const Grid& g_; StubComm c_; };
template <class GridPart, int kind>
struct StubDGSpace {
  typedef GridPart GridPartType;
  static constexpr int b200SpaceKind = kind;       // the opt-in read by Dune::Fem::B200SpaceKind
  static constexpr int localBlockSize = 1;
  StubDGSpace(const GridPart& gp, int order, std::size_t nb) : gp_(gp), order_(order) { elements_ = 1; for (int i = 0; i < GridPart::GridType::dimension; ++i) elements_ *= gp.grid().levelSize(0, i); size_ = elements_ * nb; }
  const GridPart& gridPart() const { return gp_; }
  int order() const { return order_; }
  std::size_t size() const { return size_; }
  const GridPart& gp_; int order_; std::size_t elements_, size_;
};
template <class Space>
struct StubDiscreteFunction {
  typedef Space DiscreteFunctionSpaceType;
  typedef typename Space::GridPartType GridPartType;
  typedef double RangeFieldType;
  explicit StubDiscreteFunction(const Space& s) : s_(s), dofs_(s.size(), 0.0) {}
  const Space& space() const { return s_; }
  std::vector<double>& dofVector() { return dofs_; }
  const std::vector<double>& dofVector() const { return dofs_; }
  const Space& s_; std::vector<double> dofs_;
};
}}
#endif
