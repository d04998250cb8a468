// Tensor-core (tcgen05 / TMEM) cell kernel -- placeholder translation unit until the tcgen05
// path lands; reports "unsupported" so AUTO falls through to the SIMT cell kernel.
#include "naf_common.cuh"

namespace naf {

bool xattn_cell_tc_supported(const naf_xattn_params&, const char** why) {
  *why = "library built without the tcgen05 path";
  return false;
}

int launch_xattn_cell_tc(const naf_xattn_params&, cudaStream_t) {
  return fail(NAF_ERR_UNSUPPORTED, "xattn(cell-tc): library built without the tcgen05 path");
}

}  // namespace naf
