// astr_b200/csrc/sweep2_i.cu -- the i-direction instantiations of the register line-solve engine (sweep2_impl.cuh)
#include "sweep2_impl.cuh"

int astr_sweep2_set_plan_i(int optype, const LinePlan& plan, const FilterCoef& fc) { return set_plan_local(optype, plan, fc); }
int astr_sweep2_launch_i(int optype, Sweep2Args& a, const LinePlan& plan, cudaStream_t st) {
  switch (optype) {
    case OP_DERIV: return launch2i<OP_DERIV>(a, plan, st);
    case OP_FILTER: return launch2i<OP_FILTER>(a, plan, st);
    case OP_FLUXP: return launch2i<OP_FLUXP>(a, plan, st);
    default: return launch2i<OP_FLUXM>(a, plan, st);
  }
}
