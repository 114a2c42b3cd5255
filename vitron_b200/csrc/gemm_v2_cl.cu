// vitron_b200 — gemm_v2 CTA-pair instantiations (tcgen05 cta_group::2: 256 x BN tile over the two SMs of a TPC); see gemm_v2.cuh.
#include "gemm_v2.cuh"

namespace vb {
int launch_gemm_v2_cl(int bn, int need, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 256: return dispatch_v2_cl<256, 6>(need, ta, tb, p, stream);   // 6 x (16 + 16) KB ring
    case 160: return dispatch_v2_cl<160, 8>(need, ta, tb, p, stream);   // 8 x (16 + 10) KB
    case 128: return dispatch_v2_cl<128, 8>(need, ta, tb, p, stream);   // 8 x (16 + 8) KB
    default: return VB_ERR_UNSUPPORTED;
  }
}
}  // namespace vb
