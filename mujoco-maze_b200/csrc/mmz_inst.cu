// mmz_inst.cu - one (lanes-per-env, padded-nv) instance of the maze kernels.
// Compiled once per instance with -DMMZ_G=.. -DMMZ_NVP=.. so the instances build in parallel.
#include "mmz_kernels.cuh"

namespace mmz {

template <>
kernel_fn get_kernel<MMZ_G, MMZ_NVP>(int mode) {
  switch (mode) {
    case MODE_STEP: return maze_kernel<MMZ_G, MMZ_NVP, MODE_STEP>;
    case MODE_FORWARD: return maze_kernel<MMZ_G, MMZ_NVP, MODE_FORWARD>;
    case MODE_OBSERVE: return maze_kernel<MMZ_G, MMZ_NVP, MODE_OBSERVE>;
    case MODE_RESET: return maze_kernel<MMZ_G, MMZ_NVP, MODE_RESET>;
    default: return maze_kernel<MMZ_G, MMZ_NVP, MODE_REFRESH>;
  }
}

}  // namespace mmz
