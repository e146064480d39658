// tests/dune_stub: the abstract operator interface of the reference, restated from dune/fem/operator/common/operator.hh:32-65
// (typedefs, pure virtual operator(), finalize(), nonlinear()).  Test infrastructure only.
#ifndef B200FEM_DUNE_STUB_OPERATOR_HH
#define B200FEM_DUNE_STUB_OPERATOR_HH
namespace Dune { namespace Fem {
template <class DomainFunction, class RangeFunction = DomainFunction>
struct Operator {
  typedef DomainFunction DomainFunctionType;
  typedef RangeFunction RangeFunctionType;
  typedef typename DomainFunction::RangeFieldType DomainFieldType;
  typedef typename RangeFunction::RangeFieldType RangeFieldType;
  virtual ~Operator() {}
  virtual void operator()(const DomainFunctionType& u, RangeFunctionType& w) const = 0;
  virtual void finalize() {}
  virtual bool nonlinear() const { return true; }
};
}}
#endif
