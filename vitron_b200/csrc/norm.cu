// vitron_b200 — RMSNorm / LayerNorm / GroupNorm(NHWC), HBM-bound, 16-byte vectorised, fp32 stats.
//   rmsnorm   : HF LlamaRMSNorm (transformers 4.31 modeling_llama.py, called per decoder layer)
//   layernorm : nn.LayerNorm of CLIPEncoderLayer (languagebind/image/modeling_image.py:136-151),
//               BasicTransformerBlock (i2vgen util.py:510-540), SEEM / GLIGEN blocks
//   groupnorm : nn.GroupNorm(32, C) (+SiLU) of ResBlock / TemporalConvBlock_v2 / transformers
//               (i2vgen util.py:640-655, 1358-1375, 1014) on NHWC [n, spatial, c] tensors
#include "common.cuh"
#include "vitron_b200.h"

namespace vb {

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < THREADS / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// one CTA per row; the row lives in registers between the statistics and the apply pass
template <int THREADS, int VPT, bool LAYER>
__global__ void __launch_bounds__(THREADS)
rownorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w,
               const bf16* __restrict__ b, bf16* __restrict__ out, long long ldo, int d, float eps) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const bf16* xr = x + row * ldx;
  bf16* orow = out + row * ldo;
  float v[VPT * 8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < d) {
      uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
      v[i * 8 + 0] = f0.x; v[i * 8 + 1] = f0.y; v[i * 8 + 2] = f1.x; v[i * 8 + 3] = f1.y;
      v[i * 8 + 4] = f2.x; v[i * 8 + 5] = f2.y; v[i * 8 + 6] = f3.x; v[i * 8 + 7] = f3.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += LAYER ? v[i * 8 + j] : v[i * 8 + j] * v[i * 8 + j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i * 8 + j] = 0.f;
    }
  }
  float mean = 0.f, rstd;
  if (LAYER) {
    mean = block_sum<THREADS>(s, red) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * THREADS + threadIdx.x) * 8;
      if (c < d) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float t = v[i * 8 + j] - mean; q += t * t; }
      }
    }
    rstd = rsqrtf(block_sum<THREADS>(q, red) / d + eps);
  } else {
    rstd = rsqrtf(block_sum<THREADS>(s, red) / d + eps);
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * THREADS + threadIdx.x) * 8;
    if (c < d) {
      uint4 uw = *reinterpret_cast<const uint4*>(w + c);
      float wv[8];
      float2 t;
      t = unpack_bf16(uw.x); wv[0] = t.x; wv[1] = t.y;
      t = unpack_bf16(uw.y); wv[2] = t.x; wv[3] = t.y;
      t = unpack_bf16(uw.z); wv[4] = t.x; wv[5] = t.y;
      t = unpack_bf16(uw.w); wv[6] = t.x; wv[7] = t.y;
      float bv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (LAYER && b != nullptr) {
        uint4 ub = *reinterpret_cast<const uint4*>(b + c);
        t = unpack_bf16(ub.x); bv[0] = t.x; bv[1] = t.y;
        t = unpack_bf16(ub.y); bv[2] = t.x; bv[3] = t.y;
        t = unpack_bf16(ub.z); bv[4] = t.x; bv[5] = t.y;
        t = unpack_bf16(ub.w); bv[6] = t.x; bv[7] = t.y;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i * 8 + j] - mean) * rstd * wv[j] + bv[j];
      *reinterpret_cast<uint4*>(orow + c) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                       pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

// out[row] = rsqrt(mean(x^2) + eps): one warp per row
__global__ void row_rstd_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ out, long long rows,
                                int d, float eps) {
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane * 8; c < d; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
    float2 f;
    f = unpack_bf16(u.x); s += f.x * f.x + f.y * f.y;
    f = unpack_bf16(u.y); s += f.x * f.x + f.y * f.y;
    f = unpack_bf16(u.z); s += f.x * f.x + f.y * f.y;
    f = unpack_bf16(u.w); s += f.x * f.x + f.y * f.y;
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = rsqrtf(s / d + eps);
}

// d <= 1024: one WARP per row (8 rows per CTA); each lane keeps up to 4 x 8 elements in registers.
template <bool LAYER>
__global__ void __launch_bounds__(256)
rownorm_warp_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, const bf16* __restrict__ b,
                    bf16* __restrict__ out, long long ldo, long long rows, int d, float eps) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* xr = x + row * ldx;
  float v[4][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 f = unpack_bf16(uu[j]); v[i][2 * j] = f.x; v[i][2 * j + 1] = f.y; }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += LAYER ? v[i][j] : v[i][j] * v[i][j];
    }
  }
  s = warp_sum(s);
  float mean = 0.f, rstd;
  if (LAYER) {
    mean = s / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < d) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; q += t * t; }
      }
    }
    rstd = rsqrtf(warp_sum(q) / d + eps);
  } else {
    rstd = rsqrtf(s / d + eps);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < d) {
      const uint4 uw = *reinterpret_cast<const uint4*>(w + c);
      const uint32_t ww[4] = {uw.x, uw.y, uw.z, uw.w};
      uint32_t bb[4] = {0, 0, 0, 0};
      if (LAYER && b != nullptr) {
        const uint4 ub = *reinterpret_cast<const uint4*>(b + c);
        bb[0] = ub.x; bb[1] = ub.y; bb[2] = ub.z; bb[3] = ub.w;
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 wf = unpack_bf16(ww[j]), bf = unpack_bf16(bb[j]);
        o[j] = pack_bf16((v[i][2 * j] - mean) * rstd * wf.x + bf.x, (v[i][2 * j + 1] - mean) * rstd * wf.y + bf.y);
      }
      *reinterpret_cast<uint4*>(out + row * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <bool LAYER>
static int launch_rownorm(const void* x, long long ldx, const void* w, const void* b, void* out,
                          long long ldo, long long rows, long long d, float eps, cudaStream_t st) {
  if (rows <= 0) return VB_OK;
  if (d % 8 != 0 || ldx % 8 != 0 || ldo % 8 != 0) return VB_ERR_ARG;
  const bf16* xp = reinterpret_cast<const bf16*>(x);
  const bf16* wp = reinterpret_cast<const bf16*>(w);
  const bf16* bp = reinterpret_cast<const bf16*>(b);
  bf16* op = reinterpret_cast<bf16*>(out);
  unsigned grid = static_cast<unsigned>(rows);
  if (d <= 1024 && rows >= 64)
    rownorm_warp_kernel<LAYER><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(xp, ldx, wp, bp, op, ldo, rows, (int)d, eps);
  else if (d <= 128 * 8) rownorm_kernel<128, 1, LAYER><<<grid, 128, 0, st>>>(xp, ldx, wp, bp, op, ldo, (int)d, eps);
  else if (d <= 256 * 8 * 1) rownorm_kernel<256, 1, LAYER><<<grid, 256, 0, st>>>(xp, ldx, wp, bp, op, ldo, (int)d, eps);
  else if (d <= 256 * 8 * 2) rownorm_kernel<256, 2, LAYER><<<grid, 256, 0, st>>>(xp, ldx, wp, bp, op, ldo, (int)d, eps);
  else if (d <= 256 * 8 * 4) rownorm_kernel<256, 4, LAYER><<<grid, 256, 0, st>>>(xp, ldx, wp, bp, op, ldo, (int)d, eps);
  else if (d <= 512 * 8 * 4) rownorm_kernel<512, 4, LAYER><<<grid, 512, 0, st>>>(xp, ldx, wp, bp, op, ldo, (int)d, eps);
  else return VB_ERR_UNSUPPORTED;
  VB_LAUNCH_CHECK();
  return VB_OK;
}

// ------------------------------------------------------------------ GroupNorm NHWC
// pass 1: per (image, spatial slab) partial sums per group -> atomics into [n, groups, 2] fp32
// pass 2: normalise + affine (+ activation).  x: [n, spatial, c], group = channel / (c / groups).
__global__ void gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats, long long spatial,
                                int c, int groups, int rows_per_cta) {
  // blockDim.x = vec_per_row * rows_per_pass: every thread owns one fixed 8-channel vector, keeps
  // per-channel sums in registers, and folds them into per-group shared accumulators at the end.
  extern __shared__ float sacc[];  // [groups * 2]
  const int n = blockIdx.y;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cpg = c / groups;
  const int vec_per_row = c / 8;
  const int rows_per_pass = blockDim.x / vec_per_row;
  const int my_vec = threadIdx.x % vec_per_row;
  const int my_row = threadIdx.x / vec_per_row;
  const long long rend = min(r0 + rows_per_cta, spatial);
  const bf16* base = x + static_cast<long long>(n) * spatial * c + my_vec * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (long long r = r0 + my_row; r < rend; r += rows_per_pass) {
    uint4 u = *reinterpret_cast<const uint4*>(base + r * c);
    float2 f;
    f = unpack_bf16(u.x); s[0] += f.x; q[0] += f.x * f.x; s[1] += f.y; q[1] += f.y * f.y;
    f = unpack_bf16(u.y); s[2] += f.x; q[2] += f.x * f.x; s[3] += f.y; q[3] += f.y * f.y;
    f = unpack_bf16(u.z); s[4] += f.x; q[4] += f.x * f.x; s[5] += f.y; q[5] += f.y * f.y;
    f = unpack_bf16(u.w); s[6] += f.x; q[6] += f.x * f.x; s[7] += f.y; q[7] += f.y * f.y;
  }
  // merge channels of the same group before touching shared memory
  int g_prev = (my_vec * 8) / cpg;
  float as = 0.f, aq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (my_vec * 8 + j) / cpg;
    if (g != g_prev) {
      atomicAdd(&sacc[g_prev * 2], as);
      atomicAdd(&sacc[g_prev * 2 + 1], aq);
      as = aq = 0.f;
      g_prev = g;
    }
    as += s[j];
    aq += q[j];
  }
  atomicAdd(&sacc[g_prev * 2], as);
  atomicAdd(&sacc[g_prev * 2 + 1], aq);
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sacc[i]);
}

// y = x * scale[n][c] + shift[n][c] (+ activation): the per-channel affine of this image is computed once
// per CTA into smem, then a slab of rows is streamed with 16-byte loads / stores.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, const bf16* __restrict__ w,
                const bf16* __restrict__ b, bf16* __restrict__ out, long long spatial, int c, int groups, float eps,
                int act, int rows_per_cta) {
  extern __shared__ float ss[];  // [c] scale, [c] shift
  const int n = blockIdx.y;
  const int cpg = c / groups;
  const float cnt = static_cast<float>(cpg) * static_cast<float>(spatial);
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int g = ch / cpg;
    const float sum = stats[(static_cast<long long>(n) * groups + g) * 2], sq = stats[(static_cast<long long>(n) * groups + g) * 2 + 1];
    const float mean = sum / cnt;
    const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + eps);
    const float sc = rstd * __bfloat162float(w[ch]);
    ss[ch] = sc;
    ss[c + ch] = __bfloat162float(b[ch]) - mean * sc;
  }
  __syncthreads();
  const int vec_per_row = c / 8;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long nvec = min(static_cast<long long>(rows_per_cta), spatial - r0) * vec_per_row;
  const long long base = (static_cast<long long>(n) * spatial + r0) * c;
  for (long long i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int cv = static_cast<int>(i % vec_per_row) * 8;
    const uint4 u = *reinterpret_cast<const uint4*>(x + base + i * 8);
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
    uint32_t oo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16(uu[j]);
      float y0 = f.x * ss[cv + 2 * j] + ss[c + cv + 2 * j];
      float y1 = f.y * ss[cv + 2 * j + 1] + ss[c + cv + 2 * j + 1];
      if (act == VB_ACT_SILU) { y0 = silu(y0); y1 = silu(y1); }
      else if (act == VB_ACT_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
      oo[j] = pack_bf16(y0, y1);
    }
    *reinterpret_cast<uint4*>(out + base + i * 8) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_rmsnorm(const void* x, int64_t ldx, const void* weight, void* out, int64_t ldo,
                             int64_t rows, int64_t d, float eps, cudaStream_t stream) {
  VB_CHECK_ARG(x && weight && out && d > 0);
  return launch_rownorm<false>(x, ldx, weight, nullptr, out, ldo, rows, d, eps, stream);
}

extern "C" int vb200_row_rstd(const void* x, int64_t ldx, float* out, int64_t rows, int64_t d, float eps,
                              cudaStream_t stream) {
  VB_CHECK_ARG(x && out && rows >= 0 && d > 0 && d % 8 == 0 && ldx % 8 == 0);
  if (rows == 0) return VB_OK;
  row_rstd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x), ldx, out,
                                                                             rows, static_cast<int>(d), eps);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

extern "C" int vb200_layernorm(const void* x, int64_t ldx, const void* weight, const void* bias,
                               void* out, int64_t ldo, int64_t rows, int64_t d, float eps,
                               cudaStream_t stream) {
  VB_CHECK_ARG(x && weight && out && d > 0);
  return launch_rownorm<true>(x, ldx, weight, bias, out, ldo, rows, d, eps, stream);
}

extern "C" size_t vb200_groupnorm_workspace_size(int64_t n, int64_t groups) {
  return static_cast<size_t>(n) * groups * 2 * sizeof(float);
}

extern "C" int vb200_groupnorm_nhwc(const void* x, const void* weight, const void* bias, void* out,
                                    int64_t n, int64_t spatial, int64_t c, int64_t groups, float eps,
                                    int act, void* workspace, size_t workspace_bytes,
                                    cudaStream_t stream) {
  VB_CHECK_ARG(x && weight && bias && out && n > 0 && spatial > 0 && c > 0 && groups > 0);
  VB_CHECK_ARG(c % 8 == 0 && c % groups == 0);
  VB_CHECK_ARG(c / 8 <= 1024);
  size_t need = vb200_groupnorm_workspace_size(n, groups);
  if (!workspace || workspace_bytes < need) return VB_ERR_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream);
  if (e != cudaSuccess) { vb_set_last_error(e); return VB_ERR_CUDA; }
  // slabs sized so that n * slabs ~ 4 CTAs per SM
  long long want = (4LL * vb_num_sms() + n - 1) / n;
  long long rows_per_cta = (spatial + want - 1) / want;
  if (rows_per_cta < 8) rows_per_cta = 8;
  long long slabs = (spatial + rows_per_cta - 1) / rows_per_cta;
  dim3 grid(static_cast<unsigned>(slabs), static_cast<unsigned>(n));
  const int vec_per_row = static_cast<int>(c / 8);
  int rows_per_pass = 256 / vec_per_row;
  if (rows_per_pass < 1) rows_per_pass = 1;
  int threads = vec_per_row * rows_per_pass;
  gn_stats_kernel<<<grid, threads, groups * 2 * sizeof(float), stream>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<float*>(workspace), spatial,
      static_cast<int>(c), static_cast<int>(groups), static_cast<int>(rows_per_cta));
  VB_LAUNCH_CHECK();
  // apply: ~8 CTAs per SM, each with its own smem copy of the per-channel affine
  long long want2 = (8LL * vb_num_sms() + n - 1) / n;
  long long rpc = (spatial + want2 - 1) / want2;
  if (rpc < 4) rpc = 4;
  dim3 grid2(static_cast<unsigned>((spatial + rpc - 1) / rpc), static_cast<unsigned>(n));
  gn_apply_kernel<<<grid2, 256, 2 * c * sizeof(float), stream>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<const float*>(workspace),
      reinterpret_cast<const bf16*>(weight), reinterpret_cast<const bf16*>(bias),
      reinterpret_cast<bf16*>(out), spatial, static_cast<int>(c), static_cast<int>(groups), eps, act,
      static_cast<int>(rpc));
  VB_LAUNCH_CHECK();
  return VB_OK;
}
