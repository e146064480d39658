// dune/fem/schemes/b200galerkin.hh -- the reference-side binding of the B200 operator.
//
// A DUNE-FEM maintainer drops this header next to dune/fem/schemes/molgalerkin.hh.  It follows the existing `MOL` precedent of
// an operator-name prefix: python/dune/fem/operator/__init__.py:148-150,169,178 builds the C++ type name
// 'Dune::Fem::' + operatorPrefix + 'DifferentiableGalerkinOperator< Integrands, LinearOperator >' and includes one header per
// prefix (:203-207), so galerkin(form, space, operatorPrefix='B200') selects the classes below and nothing else changes for
// users.  Everything the classes do goes through the C ABI of include/b200fem.h (libb200fem.so).
//
//   B200GalerkinOperator< Integrands, DomainFunction, RangeFunction >   : Dune::Fem::Operator   (operator/common/operator.hh:32-65)
//       same constructors, setCommunicate / setQuadratureOrders / nonlinear / gridSizeInterior as
//       Dune::Fem::GalerkinOperator (schemes/galerkin.hh:1383-1504)
//   B200MOLGalerkinOperator                                              (schemes/molgalerkin.hh:100-197)
//   B200KrylovInverseOperator< DiscreteFunction >                        (solver/krylovinverseoperators.hh:46-281): the whole Krylov
//       loop on the device; bind(op) / operator()(rhs, x) / iterations() with the reference's signed iteration count
//
// Requirements on the template arguments (all part of the reference's own interfaces unless marked NEW):
//   DiscreteFunction: DiscreteFunctionSpaceType, RangeFieldType = double, space(), dofVector().data()  -- a contiguous block
//       vector in the reference layout (function/blockvectors/defaultblockvectors.hh:284-294, 345-346), i.e.
//       AdaptiveDiscreteFunction or a numpy-backed function (function/adaptivefunction/adaptivefunction.hh:81-87)
//   Space: gridPart(), order(); the space kind is taken from B200SpaceKind< Space > (specialise for other spaces)
//   GridPart: grid(), comm(); Grid = Dune::YaspGrid< dim, EquidistantOffsetCoordinates< double, dim > > or
//       EquidistantCoordinates: levelSize(l, i), maxLevel(), domainSize(), torus().dims(i), and -- for the offset variant -- the
//       lower-left corner through B200GridTraits< Grid >::lowerLeft (specialise it for other Cartesian grids)
//   Integrands (NEW, emitted by the UFL code generator for the 'B200' prefix next to the C++ bodies it already writes,
//       python/dune/models/integrands/model.py:72-106): static const char *b200Source() -- the same interior / skeleton /
//       boundary bodies as CUDA C++ functions (interface: include/b200fem.h, b200fem_operator_create_jit);
//       void b200Constants( std::vector< double > & ) const -- the values of the dune.ufl.Constant coefficients in the order
//       the source indexes them; IntegrandsTraits::{skeleton, boundary} as in schemes/integrands.hh:90-110.
#ifndef DUNE_FEM_SCHEMES_B200GALERKIN_HH
#define DUNE_FEM_SCHEMES_B200GALERKIN_HH

#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <dune/common/exceptions.hh>
#include <dune/fem/operator/common/operator.hh>

#include <b200fem.h>

namespace Dune
{
  namespace Fem
  {

    // maps a discrete function space type onto b200fem_space_kind; the primary template reads a static member so that the
    // spaces' own headers can opt in without touching this file
    template< class Space, class = void >
    struct B200SpaceKind;
    template< class Space >
    struct B200SpaceKind< Space, std::void_t< decltype( Space::b200SpaceKind ) > >
    {
      static constexpr int value = Space::b200SpaceKind;
    };

    // Cartesian description of the grid behind a grid part
    template< class Grid >
    struct B200GridTraits
    {
      static constexpr int dimension = Grid::dimension;
      // YaspGrid: cells per direction on the leaf level, domain extent, process grid of the torus
      static void describe ( const Grid &grid, std::int32_t (&cells)[ 3 ], double (&lower)[ 3 ], double (&upper)[ 3 ], std::int32_t (&proc)[ 3 ] )
      {
        for( int i = 0; i < 3; ++i ) { cells[ i ] = 1; lower[ i ] = 0.0; upper[ i ] = 1.0; proc[ i ] = 1; }
        const auto size = grid.domainSize();
        for( int i = 0; i < dimension; ++i )
        {
          cells[ i ] = grid.levelSize( grid.maxLevel(), i );
          lower[ i ] = lowerLeft( grid, i );
          upper[ i ] = lower[ i ] + size[ i ];
          proc[ i ] = grid.torus().dims( i );
        }
      }
      // EquidistantOffsetCoordinates keep the origin in the level's coordinate container; EquidistantCoordinates start at 0
      template< class G >
      static auto lowerLeftImpl ( const G &grid, int i, int ) -> decltype( grid.begin()->coords.origin( i ) ) { return grid.begin()->coords.origin( i ); }
      template< class G >
      static double lowerLeftImpl ( const G &, int, long ) { return 0.0; }
      static double lowerLeft ( const Grid &grid, int i ) { return lowerLeftImpl( grid, i, 0 ); }
    };

    // Unstructured cube grids (ALUGrid< dim, dim, cube, conforming >): their traits specialisation sets `cartesian = false`; the
    // mesh is then handed over as arrays, collected in ONE walk over the grid part -- corner( i ) of every element and
    // indexSet.subIndex( element, i, dim ) of its vertices (the cube reference element's vertex order is the library's), elements in
    // index-set order, so that the first-touch dof numbering the library applies is the AdaptiveLeafIndexSet's
    // (gridpart/adaptiveleafindexset.hh:884-906).
    template< class Grid, class = void >
    struct B200IsCartesian : std::true_type {};
    template< class Grid >
    struct B200IsCartesian< Grid, std::void_t< decltype( Grid::b200Unstructured ) > > : std::integral_constant< bool, !Grid::b200Unstructured > {};

    template< class GridPart >
    inline void b200DescribeUnstructured ( const GridPart &gridPart, std::vector< double > &coords, std::vector< std::int64_t > &cubes )
    {
      constexpr int dim = GridPart::GridType::dimension, nv = 1 << dim;
      const auto &indexSet = gridPart.indexSet();
      coords.assign( std::size_t( indexSet.size( dim ) ) * dim, 0.0 );
      cubes.assign( std::size_t( indexSet.size( 0 ) ) * nv, 0 );
      for( auto it = gridPart.template begin< 0 >(); it != gridPart.template end< 0 >(); ++it )
      {
        const auto &entity = *it;
        const auto geometry = entity.geometry();
        const std::size_t e = indexSet.index( entity );
        for( int i = 0; i < nv; ++i )
        {
          const std::size_t v = indexSet.subIndex( entity, i, dim );
          cubes[ e*nv + i ] = std::int64_t( v );
          const auto x = geometry.corner( i );
          for( int d = 0; d < dim; ++d )
            coords[ v*dim + d ] = x[ d ];
        }
      }
    }

    namespace B200Impl
    {
      inline void check ( int rc )
      {
        if( rc == B200FEM_OK )
          return;
        if( rc == B200FEM_ERR_NOT_IMPLEMENTED )
          DUNE_THROW( NotImplemented, b200fem_last_error() );
        DUNE_THROW( InvalidStateException, b200fem_last_error() );
      }

      // device context + mesh + space handles shared by operator and solver; one per (space, device)
      template< class Space >
      struct Handles
      {
        explicit Handles ( const Space &space, int device = -1 )
        {
          typedef typename Space::GridPartType::GridType GridType;
          const auto &gridPart = space.gridPart();
          const int rank = gridPart.comm().rank();
          std::int32_t cells[ 3 ] = { 1, 1, 1 }, proc[ 3 ] = { 1, 1, 1 };
          double lower[ 3 ] = { 0, 0, 0 }, upper[ 3 ] = { 1, 1, 1 };
          if constexpr ( B200IsCartesian< GridType >::value )
          {
            B200GridTraits< GridType >::describe( gridPart.grid(), cells, lower, upper, proc );
            // the library deals cells in blocks with the first (n % p) ranks one cell larger; YaspGrid's load balancer is not
            // restated, so the two partitions are only known to agree when every direction divides evenly
            for( int i = 0; i < GridType::dimension; ++i )
              if( cells[ i ] % proc[ i ] != 0 )
                DUNE_THROW( NotImplemented, "B200GalerkinOperator: cells per direction must be a multiple of the process grid" );
          }
          else if( gridPart.comm().size() > 1 )
            DUNE_THROW( NotImplemented, "B200GalerkinOperator: unstructured grids on one rank" );
          // one process per GPU (misc/mpimanager.hh:352-461): rank modulo the visible devices unless a device is named
          int devices = 0;
          check( b200fem_device_count( &devices ) );
          check( b200fem_ctx_create( device >= 0 ? device : rank % devices, nullptr, &ctx ) );
          const int world = gridPart.comm().size();
          if( world > 1 )
          {
            // the NCCL communicator behind halo exchange and scalar products: id from rank 0 through the grid's own communication
            char id[ 128 ] = {};
            if( rank == 0 )
              check( b200fem_nccl_unique_id( id ) );
            gridPart.comm().broadcast( id, 128, 0 );
            check( b200fem_nccl_init( ctx, id, rank, world ) );
          }
          if constexpr ( B200IsCartesian< GridType >::value )
            check( b200fem_mesh_cartesian_distributed( ctx, GridType::dimension, cells, lower, upper, proc, rank, &mesh ) );
          else
          {
            std::vector< double > coords; std::vector< std::int64_t > cubes;
            b200DescribeUnstructured( gridPart, coords, cubes );
            check( b200fem_mesh_unstructured( ctx, GridType::dimension, std::int64_t( coords.size() / GridType::dimension ), coords.data(),
                                              std::int64_t( cubes.size() >> GridType::dimension ), cubes.data(), &mesh ) );
          }
          // localBlockSize = dimRange of the space (space/common/discretefunctionspace.hh): vector-valued spaces keep the scalar
          // block mapper, dof (block, c) = block * dimRange + c (function/blockvectors/defaultblockvectors.hh:284-294)
          check( b200fem_space_create_vector( mesh, B200SpaceKind< Space >::value, space.order(), B200FEM_NUMBERING_YASP, int( Space::localBlockSize ), &this->space ) );
          std::int64_t size = 0;
          check( b200fem_space_size( this->space, &size ) );
          if( std::size_t( size ) != std::size_t( space.size() ) * Space::localBlockSize )
            DUNE_THROW( InvalidStateException, "B200GalerkinOperator: dof count differs from the space's" );
        }
        Handles ( const Handles & ) = delete;
        ~Handles ()
        {
          b200fem_space_destroy( space );
          b200fem_mesh_destroy( mesh );
          b200fem_ctx_destroy( ctx );
        }
        b200fem_ctx *ctx = nullptr;
        b200fem_mesh *mesh = nullptr;
        b200fem_space *space = nullptr;
      };
    } // namespace B200Impl


    // B200GalerkinOperator
    // --------------------

    template< class Integrands, class DomainFunction, class RangeFunction = DomainFunction >
    struct B200GalerkinOperator
      : public virtual Operator< DomainFunction, RangeFunction >
    {
      typedef DomainFunction DomainFunctionType;
      typedef RangeFunction RangeFunctionType;
      typedef typename RangeFunctionType::DiscreteFunctionSpaceType RangeDiscreteFunctionSpaceType;
      typedef typename DomainFunctionType::DiscreteFunctionSpaceType DomainDiscreteFunctionSpaceType;
      typedef typename RangeDiscreteFunctionSpaceType::GridPartType GridPartType;
      typedef Integrands ModelType;

      static_assert( std::is_same< typename DomainFunctionType::RangeFieldType, double >::value, "the device path computes in double" );

      // DifferentiableGalerkinOperator( dSpace, rSpace, integrands ) -- the constructor the Python hook calls
      // (python/dune/fem/operator/__init__.py:179-181)
      template< class... Args >
      B200GalerkinOperator ( const DomainDiscreteFunctionSpaceType &dSpace, const RangeDiscreteFunctionSpaceType &rSpace, Args &&... args )
        : rSpace_( rSpace ), integrands_( std::forward< Args >( args )... ), handles_( rSpace )
      {
        if( static_cast< const void * >( &dSpace ) != static_cast< const void * >( &rSpace ) && dSpace.size() != rSpace.size() )
          DUNE_THROW( NotImplemented, "B200GalerkinOperator: domain and range space must coincide" );
        std::vector< double > constants;
        integrands_.b200Constants( constants );
        B200Impl::check( b200fem_operator_create_jit( handles_.space, Integrands::b200Source(), constants.data(), int( constants.size() ),
                                                      Integrands::hasSkeleton ? 1 : 0, Integrands::hasBoundary ? 1 : 0, &op_ ) );
      }
      B200GalerkinOperator ( const B200GalerkinOperator & ) = delete;
      ~B200GalerkinOperator () { b200fem_operator_destroy( op_ ); }

      // GalerkinOperator::setCommunicate / setQuadratureOrders (schemes/galerkin.hh:1409-1423)
      void setCommunicate ( const bool communicate ) { B200Impl::check( b200fem_operator_set_communicate( op_, communicate ? 1 : 0 ) ); }
      void setQuadratureOrders ( unsigned int interior, unsigned int surface ) { B200Impl::check( b200fem_operator_set_quadrature_orders( op_, interior, surface ) ); }

      virtual bool nonlinear () const final override { return integrands_.nonlinear(); }

      // w = L[u] (schemes/galerkin.hh:1430-1433): host dof vectors, copied to the device, evaluated, copied back
      virtual void operator() ( const DomainFunctionType &u, RangeFunctionType &w ) const final override
      {
        B200Impl::check( b200fem_operator_apply( op_, u.dofVector().data(), w.dofVector().data() ) );
        ++applies_;
      }

      // refreshes the values of the form's Constants on the device (they may have been changed through model())
      void updateConstants () const
      {
        std::vector< double > constants;
        integrands_.b200Constants( constants );
        B200Impl::check( b200fem_operator_set_constants( op_, constants.data(), int( constants.size() ) ) );
      }

      const GridPartType &gridPart () const { return rSpace_.gridPart(); }
      ModelType &model () const { return integrands_; }
      std::size_t gridSizeInterior () const { std::int64_t n = 0; b200fem_space_elements( handles_.space, &n ); return std::size_t( n ); }

      // the C handle, for the device-resident solvers below
      b200fem_operator *handle () const { return op_; }

    protected:
      const RangeDiscreteFunctionSpaceType &rSpace_;
      mutable Integrands integrands_;
      B200Impl::Handles< RangeDiscreteFunctionSpaceType > handles_;
      b200fem_operator *op_ = nullptr;
      mutable std::size_t applies_ = 0;
    };


    // B200MOLGalerkinOperator -- w = M^-1 L[u] (schemes/molgalerkin.hh:100-197), the inverse mass fused into the kernel's store
    template< class Integrands, class DomainFunction, class RangeFunction = DomainFunction >
    struct B200MOLGalerkinOperator
      : public B200GalerkinOperator< Integrands, DomainFunction, RangeFunction >
    {
      typedef B200GalerkinOperator< Integrands, DomainFunction, RangeFunction > BaseType;
      template< class... Args >
      explicit B200MOLGalerkinOperator ( Args &&... args )
        : BaseType( std::forward< Args >( args )... )
      {
        B200Impl::check( b200fem_operator_set_inverse_mass( this->op_, 1 ) );
      }
    };


    // B200KrylovInverseOperator -- KrylovInverseOperator< DF, method > with the whole loop on the device
    // (solver/krylovinverseoperators.hh:46-281; loops: solver/linear/{cg,bicgstab,gmres}.hh)
    template< class DiscreteFunction >
    class B200KrylovInverseOperator
      : public Operator< DiscreteFunction, DiscreteFunction >
    {
    public:
      enum Method { cg = 0, bicgstab = 1, gmres = 2 };          // SolverParameter::{cg,bicgstab,gmres} (solver/parameter.hh)

      explicit B200KrylovInverseOperator ( Method method = cg, double tolerance = 1e-8, int maxIterations = 1000,
                                           int errorMeasure = B200FEM_TOL_ABSOLUTE, int gmresRestart = 20 )
        : method_( method ), tolerance_( tolerance ), maxIterations_( maxIterations ), errorMeasure_( errorMeasure ), restart_( gmresRestart )
      {}

      // bind( op ) (solver/inverseoperatorinterface.hh:109-113); the Krylov loop acts on the homogeneous linear part A = L - L[0],
      // or on the difference-quotient Jacobian after b200fem_operator_linearize
      template< class Integrands, class RF >
      void bind ( const B200GalerkinOperator< Integrands, DiscreteFunction, RF > &op ) { op_ = op.handle(); }
      void unbind () { op_ = nullptr; }

      virtual void operator() ( const DiscreteFunction &rhs, DiscreteFunction &x ) const override
      {
        if( !op_ )
          DUNE_THROW( InvalidStateException, "B200KrylovInverseOperator: no operator bound" );
        int rc = B200FEM_OK;
        if( method_ == cg )
          rc = b200fem_cg_solve( op_, rhs.dofVector().data(), x.dofVector().data(), tolerance_, maxIterations_, errorMeasure_, &iterations_, nullptr );
        else if( method_ == bicgstab )
          rc = b200fem_bicgstab_solve( op_, rhs.dofVector().data(), x.dofVector().data(), tolerance_, maxIterations_, errorMeasure_, &iterations_, nullptr );
        else
          rc = b200fem_gmres_solve( op_, rhs.dofVector().data(), x.dofVector().data(), restart_, tolerance_, maxIterations_, errorMeasure_, &iterations_, nullptr );
        B200Impl::check( rc );
      }

      // signed as in the reference: negative when the tolerance was not reached (solver/linear/cg.hh:116)
      int iterations () const { return iterations_; }
      virtual bool nonlinear () const override { return false; }

    private:
      Method method_;
      double tolerance_;
      int maxIterations_, errorMeasure_, restart_;
      b200fem_operator *op_ = nullptr;
      mutable int iterations_ = 0;
    };

  } // namespace Fem

} // namespace Dune

#endif // #ifndef DUNE_FEM_SCHEMES_B200GALERKIN_HH
