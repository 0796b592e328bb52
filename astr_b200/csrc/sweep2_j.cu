// astr_b200/csrc/sweep2_j.cu -- the j-direction instantiations of the register line-solve engine (sweep2_impl.cuh)
#include "sweep2_impl.cuh"

int astr_sweep2_set_plan_j(int optype, const LinePlan& plan, const FilterCoef& fc) { return set_plan_local(optype, plan, fc); }
int astr_sweep2_launch_j(int optype, Sweep2Args& a, const CUtensorMap& tm, const LinePlan& plan, cudaStream_t st) {
  switch (optype) {
    case OP_DERIV: return launch2<1, OP_DERIV>(a, tm, plan, st);
    case OP_FILTER: return launch2<1, OP_FILTER>(a, tm, plan, st);
    case OP_FLUXP: return launch2<1, OP_FLUXP>(a, tm, plan, st);
    default: return launch2<1, OP_FLUXM>(a, tm, plan, st);
  }
}
