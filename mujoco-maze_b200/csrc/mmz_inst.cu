// mmz_inst.cu - one (lanes-per-env, padded-nv, features) instance of the maze kernels.
// Compiled once per instance with -DMMZ_G=.. -DMMZ_NVP=.. -DMMZ_FEAT=.. so the instances build in
// parallel; each defines mmz_get_kernel_<G>_<NVP>_<FEAT>(mode).
#include "mmz_kernels.cuh"

#define MMZ_CAT_(a, b, c, d) a##b##_##c##_##d
#define MMZ_CAT(a, b, c, d) MMZ_CAT_(a, b, c, d)

namespace mmz {

kernel_fn MMZ_CAT(get_kernel_, MMZ_G, MMZ_NVP, MMZ_FEAT)(int mode) {
  switch (mode) {
    case MODE_STEP: return maze_kernel<MMZ_G, MMZ_NVP, MMZ_FEAT, MODE_STEP>;
    case MODE_FORWARD: return maze_kernel<MMZ_G, MMZ_NVP, MMZ_FEAT, MODE_FORWARD>;
    case MODE_OBSERVE: return maze_kernel<MMZ_G, MMZ_NVP, MMZ_FEAT, MODE_OBSERVE>;
    case MODE_RESET: return maze_kernel<MMZ_G, MMZ_NVP, MMZ_FEAT, MODE_RESET>;
    default: return maze_kernel<MMZ_G, MMZ_NVP, MMZ_FEAT, MODE_REFRESH>;
  }
}

}  // namespace mmz
