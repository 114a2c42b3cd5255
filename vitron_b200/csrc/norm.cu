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

// d <= 512 and many rows (UNet / ViT token LayerNorms): one warp per FOUR rows. All eight 16-byte loads of a
// lane are issued before the first reduction, the four shuffle chains interleave, and the gain / bias
// vectors are fetched once per warp instead of once per row.
template <bool LAYER>
__global__ void __launch_bounds__(256)
rownorm_warp4_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w, const bf16* __restrict__ b,
                     bf16* __restrict__ out, long long ldo, long long rows, int d, float eps) {
  const long long row0 = (blockIdx.x * 8LL + (threadIdx.x >> 5)) * 4;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  const int c0 = lane * 8, c1 = (32 + lane) * 8;
  const bool h0 = c0 < d, h1 = c1 < d;
  uint4 u[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool ok = row0 + r < rows;
    const bf16* xr = x + (row0 + r) * ldx;
    u[r][0] = (ok && h0) ? __ldg(reinterpret_cast<const uint4*>(xr + c0)) : make_uint4(0u, 0u, 0u, 0u);
    u[r][1] = (ok && h1) ? __ldg(reinterpret_cast<const uint4*>(xr + c1)) : make_uint4(0u, 0u, 0u, 0u);
  }
  float wv[2][8], bv[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = i ? c1 : c0;
    const bool has = i ? h1 : h0;
    uint4 uw = make_uint4(0u, 0u, 0u, 0u), ub = make_uint4(0u, 0u, 0u, 0u);
    if (has) {
      uw = __ldg(reinterpret_cast<const uint4*>(w + c));
      if (LAYER && b != nullptr) ub = __ldg(reinterpret_cast<const uint4*>(b + c));
    }
    const uint32_t ww[4] = {uw.x, uw.y, uw.z, uw.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 wf = unpack_bf16(ww[j]), bf = unpack_bf16(bb[j]);
      wv[i][2 * j] = wf.x; wv[i][2 * j + 1] = wf.y; bv[i][2 * j] = bf.x; bv[i][2 * j + 1] = bf.y;
    }
  }
  float v[4][2][8], s[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t uu[4] = {u[r][i].x, u[r][i].y, u[r][i].z, u[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(uu[j]); v[r][i][2 * j] = f.x; v[r][i][2 * j + 1] = f.y; }
#pragma unroll
      for (int j = 0; j < 8; ++j) s[r] += LAYER ? v[r][i][j] : v[r][i][j] * v[r][i][j];  // absent pieces are zero
    }
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], sh);
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4];
  if (LAYER) {
    float q[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      mean[r] = s[r] / d;
      q[r] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (h0) { const float t = v[r][0][j] - mean[r]; q[r] += t * t; }
        if (h1) { const float t = v[r][1][j] - mean[r]; q[r] += t * t; }
      }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
      for (int r = 0; r < 4; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], sh);
#pragma unroll
    for (int r = 0; r < 4; ++r) rstd[r] = rsqrtf(q[r] / d + eps);
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) rstd[r] = rsqrtf(s[r] / d + eps);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (row0 + r >= rows) break;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!(i ? h1 : h0)) continue;
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = pack_bf16((v[r][i][2 * j] - mean[r]) * rstd[r] * wv[i][2 * j] + bv[i][2 * j],
                         (v[r][i][2 * j + 1] - mean[r]) * rstd[r] * wv[i][2 * j + 1] + bv[i][2 * j + 1]);
      *reinterpret_cast<uint4*>(out + (row0 + r) * ldo + (i ? c1 : c0)) = make_uint4(o[0], o[1], o[2], o[3]);
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
  if (d <= 512 && rows >= 1024)
    rownorm_warp4_kernel<LAYER><<<static_cast<unsigned>((rows + 31) / 32), 256, 0, st>>>(xp, ldx, wp, bp, op, ldo, rows, (int)d, eps);
  else if (d <= 1024 && rows >= 64)
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
__device__ __forceinline__ void gn_accumulate(const uint4 u, float (&s)[8], float (&q)[8]) {
  float2 f;
  f = unpack_bf16(u.x); s[0] += f.x; q[0] += f.x * f.x; s[1] += f.y; q[1] += f.y * f.y;
  f = unpack_bf16(u.y); s[2] += f.x; q[2] += f.x * f.x; s[3] += f.y; q[3] += f.y * f.y;
  f = unpack_bf16(u.z); s[4] += f.x; q[4] += f.x * f.x; s[5] += f.y; q[5] += f.y * f.y;
  f = unpack_bf16(u.w); s[6] += f.x; q[6] += f.x * f.x; s[7] += f.y; q[7] += f.y * f.y;
}

// fold the 8 per-channel partials of a thread into per-group shared accumulators
__device__ __forceinline__ void gn_fold(const float (&s)[8], const float (&q)[8], int ch0, int cpg, float* sacc) {
  int g_prev = ch0 / cpg;
  float as = 0.f, aq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (ch0 + j) / cpg;
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
}

// Shared-memory float atomics are CAS loops on this architecture (ATOMS.CAST.SPIN): hundreds of threads folding
// into a handful of group accumulators serialise for tens of microseconds. For cpg >= 8 an 8-channel vector
// touches at most two groups, so every thread publishes (sum, sq) for "its first group" and "the next one", and
// one warp per group then reduces the matching slots with shuffles: no atomics, deterministic order.
// part: [threads][4]; sacc: [ngroups][2]. All threads of the CTA must call this (contains __syncthreads()).
__device__ __forceinline__ void gn_block_group_sums(const float (&s)[8], const float (&q)[8], int my_vec, int vec_per_row,
                                                    int nthreads_active, int cpg, int ngroups, float* part, float* sacc) {
  const int tid = threadIdx.x;
  if (tid < nthreads_active) {
    const int ga = (my_vec * 8) / cpg;
    float sa = 0.f, qa = 0.f, sb = 0.f, qb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((my_vec * 8 + j) / cpg == ga) { sa += s[j]; qa += q[j]; }
      else { sb += s[j]; qb += q[j]; }
    }
    *reinterpret_cast<float4*>(part + tid * 4) = make_float4(sa, qa, sb, qb);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int rows_per_pass = nthreads_active / vec_per_row;
  for (int g = warp; g < ngroups && warp < nwarps; g += nwarps) {  // full warps only (blockDim need not be a multiple of 32)
    const int v_lo = (g * cpg) / 8, v_hi = ((g + 1) * cpg - 1) / 8;  // vectors overlapping group g
    const int nv = v_hi - v_lo + 1;
    float as = 0.f, aq = 0.f;
    for (int e = lane; e < rows_per_pass * nv; e += 32) {
      const int row = e / nv, v = v_lo + e % nv;
      const float4 pv = *reinterpret_cast<const float4*>(part + (row * vec_per_row + v) * 4);
      const int ga = (v * 8) / cpg;
      if (ga == g) { as += pv.x; aq += pv.y; }
      else if (ga == g - 1) { as += pv.z; aq += pv.w; }
    }
    as = warp_sum(as);
    aq = warp_sum(aq);
    if (lane == 0) { sacc[g * 2] = as; sacc[g * 2 + 1] = aq; }
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 gn_affine(const uint4 u, const float* sc, const float* sh, int act) {
  const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
  uint32_t oo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = unpack_bf16(uu[j]);
    float y0 = fmaf(f.x, sc[2 * j], sh[2 * j]);
    float y1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
    if (act == VB_ACT_SILU) { y0 = silu(y0); y1 = silu(y1); }
    else if (act == VB_ACT_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
    oo[j] = pack_bf16(y0, y1);
  }
  return make_uint4(oo[0], oo[1], oo[2], oo[3]);
}

// Single-pass GroupNorm for slabs that fit in shared memory: CTA (part, image) owns `cpart` channels (whole
// groups) of one image, pulls its [spatial x cpart] slab into smem while accumulating the group sums, then
// normalises out of smem: x is read from HBM once instead of twice and there is one launch instead of three.
__global__ void gn_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                bf16* __restrict__ out, int spatial, int c, int groups, float eps, int act, int cpart) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int cpg = c / groups;
  const int gpart = cpart / cpg;
  bf16* slab = reinterpret_cast<bf16*>(gsm);                                   // [spatial][cpart]
  float* ss = reinterpret_cast<float*>(gsm + static_cast<size_t>(spatial) * cpart * 2);  // scale[cpart], shift[cpart]
  float* sacc = ss + 2 * cpart;                                                // [gpart * 2]
  const int n = blockIdx.y, c0 = blockIdx.x * cpart;
  __shared__ __align__(16) float part[512 * 4];
  const int vec_per_row = cpart / 8;
  const int rows_per_pass = blockDim.x / vec_per_row;
  const int my_vec = threadIdx.x % vec_per_row, my_row = threadIdx.x / vec_per_row;
  const bf16* base = x + (static_cast<long long>(n) * spatial) * c + c0 + my_vec * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (my_row < rows_per_pass) {  // threads beyond vec_per_row * rows_per_pass idle (blockDim is exact: none)
    int r = my_row;
    for (; r + 7 * rows_per_pass < spatial; r += 8 * rows_per_pass) {  // eight independent 16-byte loads in flight
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r + k * rows_per_pass) * c));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        *reinterpret_cast<uint4*>(slab + static_cast<size_t>(r + k * rows_per_pass) * cpart + my_vec * 8) = u[k];
        gn_accumulate(u[k], s, q);
      }
    }
    for (; r < spatial; r += rows_per_pass) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * c));
      *reinterpret_cast<uint4*>(slab + static_cast<size_t>(r) * cpart + my_vec * 8) = u;
      gn_accumulate(u, s, q);
    }
  }
  gn_block_group_sums(s, q, my_vec, vec_per_row, vec_per_row * rows_per_pass, cpg, gpart, part, sacc);
  const float cnt = static_cast<float>(cpg) * static_cast<float>(spatial);
  for (int ch = threadIdx.x; ch < cpart; ch += blockDim.x) {
    const int g = ch / cpg;
    const float mean = sacc[g * 2] / cnt;
    const float rstd = rsqrtf(fmaxf(sacc[g * 2 + 1] / cnt - mean * mean, 0.f) + eps);
    const float scl = rstd * __bfloat162float(w[c0 + ch]);
    ss[ch] = scl;
    ss[cpart + ch] = __bfloat162float(b[c0 + ch]) - mean * scl;
  }
  __syncthreads();
  if (my_row < rows_per_pass) {
    float scl[8], shf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { scl[j] = ss[my_vec * 8 + j]; shf[j] = ss[cpart + my_vec * 8 + j]; }
    bf16* obase = out + (static_cast<long long>(n) * spatial) * c + c0 + my_vec * 8;
    for (int r = my_row; r < spatial; r += rows_per_pass) {  // each thread re-reads exactly what it staged
      const uint4 u = *reinterpret_cast<const uint4*>(slab + static_cast<size_t>(r) * cpart + my_vec * 8);
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(r) * c) = gn_affine(u, scl, shf, act);
    }
  }
}

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
  long long r = r0 + my_row;
  for (; r + 3LL * rows_per_pass < rend; r += 4LL * rows_per_pass) {  // four independent 16-byte loads in flight
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(base + (r + static_cast<long long>(k) * rows_per_pass) * c));
#pragma unroll
    for (int k = 0; k < 4; ++k) gn_accumulate(u[k], s, q);
  }
  for (; r < rend; r += rows_per_pass) gn_accumulate(__ldg(reinterpret_cast<const uint4*>(base + r * c)), s, q);
  __shared__ __align__(16) float part[1024 * 4];
  if (cpg >= 8) {
    gn_block_group_sums(s, q, my_vec, vec_per_row, vec_per_row * rows_per_pass, cpg, groups, part, sacc);
  } else {
    gn_fold(s, q, my_vec * 8, cpg, sacc);  // narrow groups: shared atomics (slow, rare)
  }
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
  long long i = threadIdx.x;
  for (; i + 3LL * blockDim.x < nvec; i += 4LL * blockDim.x) {  // four independent 16-byte loads in flight
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(x + base + (i + static_cast<long long>(k) * blockDim.x) * 8));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long ii = i + static_cast<long long>(k) * blockDim.x;
      const int cv = static_cast<int>(ii % vec_per_row) * 8;
      *reinterpret_cast<uint4*>(out + base + ii * 8) = gn_affine(u[k], ss + cv, ss + c + cv, act);
    }
  }
  for (; i < nvec; i += blockDim.x) {
    const int cv = static_cast<int>(i % vec_per_row) * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + base + i * 8));
    *reinterpret_cast<uint4*>(out + base + i * 8) = gn_affine(u, ss + cv, ss + c + cv, act);
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
  {
    // single-pass kernel when a CTA's [spatial x cpart] slab fits in shared memory; cpart = whole groups and
    // whole 16-byte vectors. Prefer slabs <= 48 KB (several CTAs per SM), else the narrowest part up to 200 KB.
    const int cpg = static_cast<int>(c / groups);
    int unit = cpg;
    while (unit % 8 != 0) unit += cpg;  // lcm(cpg, 8)
    int cpart = 0;
    if (cpg >= 8 && c % unit == 0 && spatial <= (1 << 20)) {
      for (int cand = unit; cand <= c; cand += unit) {
        if (c % cand != 0 || cand / 8 > 512) continue;
        const long long slab = spatial * cand * 2;
        if (cpart == 0 && slab <= 200 * 1024) cpart = cand;       // narrowest that fits at all
        if (slab <= 48 * 1024) cpart = cand;                      // widest small slab
      }
    }
    if (cpart > 0 && x != out) {
      const int vec_per_row = cpart / 8;
      const int rows_per_pass = 512 / vec_per_row;
      const int threads = vec_per_row * rows_per_pass;
      const size_t smem = static_cast<size_t>(spatial) * cpart * 2 + 2 * cpart * sizeof(float) + (cpart / cpg) * 2 * sizeof(float);
      static size_t smem_set = 0;
      if (smem > smem_set) {
        cudaError_t ea = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
        if (ea != cudaSuccess) { vb_set_last_error(ea); return VB_ERR_CUDA; }
        smem_set = 208 * 1024;
      }
      dim3 gridf(static_cast<unsigned>(c / cpart), static_cast<unsigned>(n));
      gn_fused_kernel<<<gridf, threads, smem, stream>>>(
          reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(weight), reinterpret_cast<const bf16*>(bias),
          reinterpret_cast<bf16*>(out), static_cast<int>(spatial), static_cast<int>(c), static_cast<int>(groups), eps, act,
          cpart);
      VB_LAUNCH_CHECK();
      return VB_OK;
    }
  }
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
