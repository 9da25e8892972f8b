// mmz_hinst.cu - hybrid kernel instances: -DMMZ_NVP=<registers per Hessian row> -DMMZ_BOX=<box geoms compiled in>
// (14, 0: the Ant family; 16, 1: Ant with a movable block; 4, 1: the Point and its arrow box), one kernel per mode.
#include "mmz_hstep.cuh"

#define MMZ_HCAT_(a, b) a##b
#define MMZ_HCAT(a, b) MMZ_HCAT_(a, b)

namespace mmz {

hkernel_fn MMZ_HCAT(get_hkernel_, MMZ_NVP)(int mode) {
  switch (mode) {
    case TMODE_STEP: return maze_hkernel<MMZ_NVP, MMZ_BOX, TMODE_STEP>;
    case TMODE_FORWARD: return maze_hkernel<MMZ_NVP, MMZ_BOX, TMODE_FORWARD>;
    case TMODE_OBSERVE: return maze_hkernel<MMZ_NVP, MMZ_BOX, TMODE_OBSERVE>;
    case TMODE_RESET: return maze_hkernel<MMZ_NVP, MMZ_BOX, TMODE_RESET>;
    default: return maze_hkernel<MMZ_NVP, MMZ_BOX, TMODE_REFRESH>;
  }
}

}  // namespace mmz
