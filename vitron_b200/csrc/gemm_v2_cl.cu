// vitron_b200 — gemm_v2 "cluster pair" instantiations (two CTAs share the A tile through TMA multicast); see gemm_v2.cuh.
#include "gemm_v2.cuh"

namespace vb {
int launch_gemm_v2_cl(int bn, int need, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 256: return dispatch_v2_cl<256, 4>(need, ta, tb, p, stream);
    case 160: return dispatch_v2_cl<160, 6>(need, ta, tb, p, stream);
    case 128: return dispatch_v2_cl<128, 6>(need, ta, tb, p, stream);
    default: return VB_ERR_UNSUPPORTED;
  }
}
}  // namespace vb
