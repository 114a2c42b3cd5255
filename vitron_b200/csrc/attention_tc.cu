// vitron_b200 — flash attention on tcgen05 tensor cores with TMEM accumulators (sm_100a).
//
//   O = softmax(Q K^T * scale [causal, kv_len, bool mask]) V     bf16 in / out, fp32 math
//   head_dim 64 / 128 natively; 40 / 80 / 160 (GLIGEN, attention.py:192-257) run the 64 / 128 / 192-column
//   instantiation: the tensor maps carry the true head_dim, so the TMA unit zero-fills the padding columns of
//   Q / K / V in shared memory (they add 0 to QK^T and produce zero O columns that are never stored).
//   Boolean masks (SEEM masked cross-attention, utils/attn.py:296-316): uint8 [B|1, H|1, Sq, Skv], non-zero = masked
//   out, applied to S in registers; a row with every key masked yields zeros.
//
// One CTA = 128 query rows of one (batch, head); keys stream through in blocks of 64.
//   warp 0      TMA producer: Q once, K_j / V_j into 2-stage 128B-swizzled rings (4-D tensor maps over the
//               caller's strided [B, S, H, D] views, so fused QKV buffers are read in place)
//   warp 1      MMA issuer:   S_j = Q K_j^T  (M128 x N64 x K=D, both K-major)   -> TMEM S[j&1]
//                             O  += P_j V_j  (M128 x N=D x K64, V is MN-major)   -> TMEM O
//               QK_{j+1} is issued BEFORE PV_j, so the tensor pipe works on PV_j while the softmax warps
//               are already on block j+1 (S is double buffered in TMEM).
//   warps 2..5  softmax: thread == query row (tcgen05.ld 32x32b), so row max / sum need no shuffles;
//               P_j goes to smem (bf16, K-major, swizzled) as the A operand of PV_j; O stays in TMEM and is
//               rescaled lazily (only when the running max grows by more than 2^8), final 1/l in the epilogue.
// Two CTAs fit per SM (TMEM 256 columns, <= 113 KB smem each) so one CTA's softmax overlaps the other's MMAs.
//
// Replaces (when there is no boolean mask and D in {64, 128}) the mma.sync kernel of attention.cu for:
// LLaMA prefill, CLIP ViT, UNet spatial self/cross attention, SEEM pixel-decoder encoder, SEEM self-attn.
#include "common.cuh"
#include "vitron_b200.h"

namespace vb {

constexpr int TC_BM = 128;  // query rows per CTA
constexpr int TC_BN = 64;   // keys per block
constexpr int TC_THREADS = 192;

struct TcAttnParams {
  bf16* o;
  long long o_sb, o_ss, o_sh;
  int B, H, Sq, Skv;
  int D;                 // true head dim (<= the kernel's HD; the difference is zero padding)
  float scale_log2;
  int causal;
  const int32_t* kv_len;
  const uint8_t* mask;   // or null
  long long m_sb, m_sh, m_sq;
  int splits;            // > 1: the key blocks are divided over `splits` CTAs per query tile (few-query attention: SEEM's
  float* ws_o;           //      101 queries over 16384 keys would otherwise run on 8 CTAs); each writes its un-normalised
  float* ws_ml;          //      O (fp32) and (m, l) to the workspace, attn_split_merge_kernel combines them
};

// Bounded mbarrier wait: a protocol or descriptor bug must never hang the GPU. After 4 s without progress the
// waiter records (site id, block) in g_tc_watchdog and TRAPS: the launch fails with a sticky CUDA error that the
// next vb200_* call / stream synchronisation reports (VB_ERR_CUDA) — a timed-out attention never returns garbage
// as if it were a result. The abort flag only short-cuts the other waiters of the CTA until the trap lands.
__device__ unsigned int g_tc_watchdog[4];

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __noinline__ void tc_wait_slow(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, int site) {
  const unsigned long long t0 = gtime_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort_flag) return;
    if (gtime_ns() - t0 > 4000000000ull) {
      *abort_flag = 1;
      if (atomicCAS(&g_tc_watchdog[0], 0u, static_cast<unsigned int>(site)) == 0u) {
        g_tc_watchdog[1] = blockIdx.x | (blockIdx.y << 12) | (blockIdx.z << 22);
        g_tc_watchdog[2] = threadIdx.x;
      }
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, int site) {
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    if (mbar_try_wait(bar, parity)) return;
  tc_wait_slow(bar, parity, abort_flag, site);
}

// MN-major operand (rows = K index, 128-byte rows of 64 MN elements, 128B swizzle):
// SBO = stride between groups of 8 K rows, LBO = stride between 64-element MN atoms.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2, flush-to-zero: exp2(-inf) = 0, no denormal fix-up
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int HD>
__global__ void __launch_bounds__(TC_THREADS)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const TcAttnParams p) {
  pdl_trigger();
  pdl_wait();   // (the tcgen05 / mbarrier set-up below touches no global data, but the kernel's first TMA load follows at once)
  constexpr int KC = HD / 64;                   // 64-element chunks along the head dim
  constexpr int Q_BYTES = TC_BM * HD * 2;       // [KC][128][64]
  constexpr int K_BYTES = TC_BN * HD * 2;       // [KC][64][64]   K-major
  constexpr int V_BYTES = TC_BN * HD * 2;       // [KC][64 keys][64]   MN-major atoms
  constexpr int P_BYTES = TC_BM * TC_BN * 2;    // [128][64]   K-major
  constexpr uint32_t TMEM_COLS = (2 * TC_BN + HD <= 128) ? 128 : (2 * TC_BN + HD <= 256) ? 256 : 512;
  constexpr uint32_t S_COL = 0, O_COL = 2 * TC_BN;

  // No static shared memory in this kernel, so the dynamic window starts at the (1024-byte aligned) base of the
  // CTA's allocation; spending another KB on manual alignment would cost the second resident CTA at HD = 128.
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;          // 2 stages
  uint8_t* sV = sK + 2 * K_BYTES;      // 2 stages
  uint8_t* sP = sV + 2 * V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;             // 1
  uint64_t* k_full = bars + 1;         // 2
  uint64_t* k_empty = bars + 3;        // 2
  uint64_t* v_full = bars + 5;         // 2
  uint64_t* v_empty = bars + 7;        // 2
  uint64_t* s_full = bars + 9;         // 2
  uint64_t* p_full = bars + 11;        // 1 (count 4: one arrive per softmax warp)
  uint64_t* pv_done = bars + 12;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x / p.splits, sp = blockIdx.x - qb * p.splits, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * TC_BM;
  const int kv_len = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
  const int causal_off = p.Skv - p.Sq;
  int kv_end = kv_len;
  if (p.causal) kv_end = min(kv_len, q0 + TC_BM + causal_off);
  const int nblk_all = kv_end > 0 ? (kv_end + TC_BN - 1) / TC_BN : 0;
  // this CTA's share of the key blocks: [jb0, jb0 + nblk)
  const int per_split = (nblk_all + p.splits - 1) / p.splits;
  const int jb0 = sp * per_split;
  const int nblk = max(0, min(nblk_all, jb0 + per_split) - jb0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    *abort_flag = 0;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && nblk > 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int c = 0; c < KC; ++c) tma_load_4d(sQ + c * (TC_BM * 128), &tmap_q, q_full, c * 64, q0, h, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        tc_wait(&k_empty[st], ph ^ 1, abort_flag, 5);
        mbar_arrive_expect_tx(&k_full[st], K_BYTES);
#pragma unroll
        for (int c = 0; c < KC; ++c) tma_load_4d(sK + st * K_BYTES + c * (TC_BN * 128), &tmap_k, &k_full[st], c * 64, (jb0 + j) * TC_BN, h, b);
        tc_wait(&v_empty[st], ph ^ 1, abort_flag, 6);
        mbar_arrive_expect_tx(&v_full[st], V_BYTES);
#pragma unroll
        for (int c = 0; c < KC; ++c) tma_load_4d(sV + st * V_BYTES + c * (TC_BN * 128), &tmap_v, &v_full[st], c * 64, (jb0 + j) * TC_BN, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (nblk > 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(TC_BM, HD) | (1u << 16);  // B (= V) is MN-major
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        tc_wait(&k_full[st], (j >> 1) & 1, abort_flag, 2);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d_tmem = tmem_base + S_COL + st * TC_BN;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint32_t qa = smem_u32(sQ) + (ks / 4) * (TC_BM * 128) + (ks % 4) * 32;
            const uint32_t ka = smem_u32(sK) + st * K_BYTES + (ks / 4) * (TC_BN * 128) + (ks % 4) * 32;
            tc_mma_bf16(d_tmem, umma_desc_kmajor_sw128(qa), umma_desc_kmajor_sw128(ka), idesc_qk, ks > 0 ? 1u : 0u);
          }
          tc_commit(&k_empty[st]);
          tc_commit(&s_full[st]);
        }
        __syncwarp();
      };
      tc_wait(q_full, 0, abort_flag, 1);
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_qk(j + 1);  // S[(j+1)&1] was fully read before P_{j-1} was published
        const int st = j & 1;
        tc_wait(p_full, j & 1, abort_flag, 3);
        tc_wait(&v_full[st], (j >> 1) & 1, abort_flag, 4);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d_tmem = tmem_base + O_COL;
#pragma unroll
          for (int ks = 0; ks < TC_BN / 16; ++ks) {
            const uint32_t pa = smem_u32(sP) + ks * 32;                          // K-major: +16 keys = +32 B
            const uint32_t va = smem_u32(sV) + st * V_BYTES + ks * (16 * 128);   // MN-major: +16 key rows
            tc_mma_bf16(d_tmem, umma_desc_kmajor_sw128(pa), umma_desc_mnmajor_sw128(va, TC_BN * 128, 1024), idesc_pv,
                        (j > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(&v_empty[st]);
          tc_commit(pv_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================== softmax / correction / epilogue (thread = row)
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_used = -INFINITY;  // the max the stored exponentials are relative to
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      tc_wait(&s_full[st], (j >> 1) & 1, abort_flag, 7);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_addr + S_COL + st * TC_BN;
      const int kbase = (jb0 + j) * TC_BN;
      // CTA-uniform: does any row of this tile see a masked key in this block?
      const bool need_mask = (kbase + TC_BN > kv_end) || (p.causal && kbase + TC_BN - 1 > q0 + causal_off);
      // boolean mask bytes of this row / key block: fetched before the wait on the TMEM load
      uint4 mb[4];
      bool mvec = false;
      const uint8_t* mrow = nullptr;
      if (p.mask != nullptr && qrow < p.Sq) {
        mrow = p.mask + b * p.m_sb + h * p.m_sh + static_cast<long long>(qrow) * p.m_sq + kbase;
        mvec = kbase + TC_BN <= p.Skv && (reinterpret_cast<uintptr_t>(mrow) & 15) == 0;
        if (mvec) {
#pragma unroll
          for (int i = 0; i < 4; ++i) mb[i] = __ldg(reinterpret_cast<const uint4*>(mrow) + i);
        }
      }
      uint32_t sv[2][32];
      tmem_ld_32x32(s_addr, sv[0]);
      tmem_ld_32x32(s_addr + 32, sv[1]);
      tmem_ld_wait();
      if (need_mask) {
        const int lim = min(kv_end, p.causal ? qrow + causal_off + 1 : kv_end) - kbase;  // keys [0, lim) are live
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= lim) sv[c][i] = 0xff800000u;  // -inf
      }
      if (mrow != nullptr) {
        if (mvec) {
#pragma unroll
          for (int w = 0; w < 16; ++w) {
            const uint32_t word = w % 4 == 0 ? mb[w / 4].x : w % 4 == 1 ? mb[w / 4].y : w % 4 == 2 ? mb[w / 4].z : mb[w / 4].w;
#pragma unroll
            for (int bt = 0; bt < 4; ++bt)
              if ((word >> (8 * bt)) & 0xffu) sv[(w * 4 + bt) / 32][(w * 4 + bt) % 32] = 0xff800000u;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kbase + c * 32 + i < p.Skv && mrow[c * 32 + i]) sv[c][i] = 0xff800000u;
        }
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains: the softmax warps are
#pragma unroll                                                      // latency-, not issue-bound
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx[0] = fmaxf(mx[0], __uint_as_float(sv[c][i]));
          mx[1] = fmaxf(mx[1], __uint_as_float(sv[c][i + 1]));
          mx[2] = fmaxf(mx[2], __uint_as_float(sv[c][i + 2]));
          mx[3] = fmaxf(mx[3], __uint_as_float(sv[c][i + 3]));
        }
      const float mblk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      // lazy rescale: keep m_used unless the running max grew by more than 2^8 in the exp2 domain
      float corr = 1.f;
      bool rescale = false;
      if (mblk > m_used) {
        if (m_used == -INFINITY) {
          m_used = mblk;
        } else if ((mblk - m_used) * p.scale_log2 > 8.f) {
          corr = ex2_approx((m_used - mblk) * p.scale_log2);
          m_used = mblk;
          rescale = true;
        }
      }
      const float moff = (m_used == -INFINITY) ? 0.f : m_used * p.scale_log2;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[32];  // 64 bf16 probabilities of this row
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -moff));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c][i + 1]), p.scale_log2, -moff));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sv[c][i + 2]), p.scale_log2, -moff));
          const float p3 = ex2_approx(fmaf(__uint_as_float(sv[c][i + 3]), p.scale_log2, -moff));
          ps[0] += p0; ps[1] += p1; ps[2] += p2; ps[3] += p3;
          pk[c * 16 + i / 2] = pack_bf16(p0, p1);
          pk[c * 16 + i / 2 + 1] = pack_bf16(p2, p3);
        }
      }
      const float psum = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      // PV_{j-1} must have retired before sP is overwritten and before O is touched
      if (j > 0) {
        tc_wait(pv_done, (j - 1) & 1, abort_flag, 8);
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, rescale)) {  // warp-uniform TMEM traffic (aligned instructions)
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t ov[32];
            tmem_ld_32x32(tmem_base + lane_addr + O_COL + c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
            tmem_st_32x32(tmem_base + lane_addr + O_COL + c, ov);
          }
          tmem_st_wait();
        }
      }
      l = l * corr + psum;
      uint8_t* prow = sP + r * 128;
#pragma unroll
      for (int piece = 0; piece < 8; ++piece)
        *reinterpret_cast<uint4*>(prow + ((piece ^ (r & 7)) << 4)) =
            make_uint4(pk[piece * 4], pk[piece * 4 + 1], pk[piece * 4 + 2], pk[piece * 4 + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global (row per thread)
    if (nblk > 0) {
      tc_wait(pv_done, (nblk - 1) & 1, abort_flag, 9);
      tc_fence_after();
    }
    if (p.splits > 1) {
      // ---- partial result of this key range: un-normalised O, the max its exponentials are relative to (exp2 domain), l
      // (the TMEM loads are warp-aligned instructions: every lane executes them, only the stores are predicated)
      const long long slot = ((static_cast<long long>(b) * p.H + h) * p.splits + sp) * p.Sq + min(qrow, p.Sq - 1);
      if (qrow < p.Sq) {
        p.ws_ml[slot * 2] = (m_used == -INFINITY) ? -INFINITY : m_used * p.scale_log2;
        p.ws_ml[slot * 2 + 1] = l;
      }
      float* orow_ws = p.ws_o + slot * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += 32) {
        uint32_t ov[32];
        if (nblk > 0) {
          tmem_ld_32x32(tmem_base + lane_addr + O_COL + c, ov);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = 0;
        }
        if (qrow < p.Sq) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c + i < p.D)
              *reinterpret_cast<float4*>(orow_ws + c + i) = make_float4(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1]),
                                                                        __uint_as_float(ov[i + 2]), __uint_as_float(ov[i + 3]));
        }
      }
    } else {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.o + b * p.o_sb + h * p.o_sh + static_cast<long long>(qrow) * p.o_ss;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t ov[32];
      if (nblk > 0) {
        tmem_ld_32x32(tmem_base + lane_addr + O_COL + c, ov);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = 0;
      }
      if (qrow < p.Sq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          if (c + i < p.D)
          *reinterpret_cast<uint4*>(orow + c + i) =
              make_uint4(pack_bf16(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv),
                         pack_bf16(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv),
                         pack_bf16(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv),
                         pack_bf16(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv));
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// out[b, q, h, :] = sum_s O_s 2^(m_s - M) / sum_s l_s 2^(m_s - M): one thread per 4 head-dim columns of a (b, h, q) row
__global__ void attn_split_merge_kernel(const float* __restrict__ ws_o, const float* __restrict__ ws_ml, bf16* __restrict__ out,
                                        long long o_sb, long long o_ss, long long o_sh, int B, int H, int Sq, int D, int HD, int splits) {
  pdl_trigger();
  pdl_wait();
  const int vecs = D / 4;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(B) * H * Sq * vecs) return;
  const int v = static_cast<int>(idx % vecs);
  const long long row = idx / vecs;          // (b * H + h) * Sq + q
  const int q = static_cast<int>(row % Sq);
  const long long bh = row / Sq;
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, ws_ml[((bh * splits + s) * Sq + q) * 2]);
  float L = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (M != -INFINITY) {
    for (int s = 0; s < splits; ++s) {
      const long long slot = (bh * splits + s) * Sq + q;
      const float m = ws_ml[slot * 2];
      if (m == -INFINITY) continue;
      const float wgt = exp2f(m - M);
      L += ws_ml[slot * 2 + 1] * wgt;
      const float4 o = *reinterpret_cast<const float4*>(ws_o + slot * HD + v * 4);
      a0 += o.x * wgt; a1 += o.y * wgt; a2 += o.z * wgt; a3 += o.w * wgt;
    }
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  const int b = static_cast<int>(bh / H), h = static_cast<int>(bh % H);
  bf16* dst = out + b * o_sb + h * o_sh + static_cast<long long>(q) * o_ss + v * 4;
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(a0 * inv, a1 * inv), pack_bf16(a2 * inv, a3 * inv));
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled2 encode_fn() {
  static PFN_encodeTiled2 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled2>(ptr);
  return fn;
}

// [B, S, H, D] view with element strides (sb, ss, sh), D contiguous -> 4-D map (d, s, h, b), box {64, rows, 1, 1}
static int make_qkv_map(CUtensorMap* m, const void* base, long long B, long long S, long long H, long long D,
                        long long sb, long long ss, long long sh, int box_rows) {
  PFN_encodeTiled2 fn = encode_fn();
  if (!fn) return VB_ERR_DRIVER;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ss) * 2, static_cast<cuuint64_t>(sh) * 2, static_cast<cuuint64_t>(sb) * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VB_OK : VB_ERR_DRIVER;
}

template <int HD>
static int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const TcAttnParams& p,
                     cudaStream_t stream) {
  constexpr int smem = TC_BM * HD * 2 + 4 * TC_BN * HD * 2 + TC_BM * TC_BN * 2 + 128;
  static bool attr = false;
  auto kern = flash_attn_tc_kernel<HD>;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vb_set_last_error(e); return VB_ERR_CUDA; }
    // two CTAs per SM need (nearly) the whole 228 KB at HD = 128: ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  dim3 grid(((p.Sq + TC_BM - 1) / TC_BM) * p.splits, p.H, p.B);
  { cudaError_t le = vb_launch(kern, grid, dim3(TC_THREADS), smem, stream, tq, tk, tv, p); if (le != cudaSuccess) { vb_set_last_error(le); return VB_ERR_CUDA; } }
  if (p.splits > 1) {
    const long long items = static_cast<long long>(p.B) * p.H * p.Sq * (p.D / 4);
    cudaError_t le = vb_launch(attn_split_merge_kernel, dim3(static_cast<unsigned>((items + 255) / 256)), dim3(256), 0, stream,
                               static_cast<const float*>(p.ws_o), static_cast<const float*>(p.ws_ml), p.o, p.o_sb, p.o_ss, p.o_sh,
                               p.B, p.H, p.Sq, p.D, HD, p.splits);
    if (le != cudaSuccess) { vb_set_last_error(le); return VB_ERR_CUDA; }
  }
  VB_LAUNCH_CHECK();
  return VB_OK;
}

}  // namespace vb

using namespace vb;

// Diagnostics: {site id of the first timed-out wait (0 = none), packed block index, thread}; clears the record.
extern "C" int vb200_attention_watchdog(uint32_t* out3) {
  VB_CHECK_ARG(out3);
  unsigned int h[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpyFromSymbol(h, g_tc_watchdog, sizeof(h));
  if (e != cudaSuccess) { vb_set_last_error(e); return VB_ERR_CUDA; }
  out3[0] = h[0]; out3[1] = h[1]; out3[2] = h[2];
  if (h[0]) {
    unsigned int z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_tc_watchdog, z, sizeof(z));
  }
  return VB_OK;
}

// Resident CTAs per SM of the tcgen05 attention kernel (by registers / shared memory; TMEM allows 2).
extern "C" int vb200_attention_tc_occupancy(int head_dim) {
  int n = 0;
  cudaError_t e;
  if (head_dim == 64) {
    constexpr int smem = TC_BM * 64 * 2 + 4 * TC_BN * 64 * 2 + TC_BM * TC_BN * 2 + 128;
    cudaFuncSetAttribute(flash_attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, flash_attn_tc_kernel<64>, TC_THREADS, smem);
  } else if (head_dim == 128) {
    constexpr int smem = TC_BM * 128 * 2 + 4 * TC_BN * 128 * 2 + TC_BM * TC_BN * 2 + 128;
    cudaFuncSetAttribute(flash_attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, flash_attn_tc_kernel<128>, TC_THREADS, smem);
  } else {
    return VB_ERR_ARG;
  }
  if (e != cudaSuccess) { vb_set_last_error(e); return VB_ERR_CUDA; }
  return n;
}

static int tc_hd(int64_t head_dim) { return head_dim <= 64 ? 64 : head_dim <= 128 ? 128 : 192; }

// split-KV factor: only when the (query tile, head, batch) grid leaves most SMs idle and there are enough key blocks
static int tc_splits(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int causal) {
  if (causal) return 1;
  const long long ctas = ((Sq + TC_BM - 1) / TC_BM) * H * B;
  const long long nblk = (Skv + TC_BN - 1) / TC_BN;
  const int sms = vb_num_sms();
  if (ctas * 2 > sms || nblk < 8) return 1;
  long long s = (2LL * sms + ctas - 1) / ctas;
  if (s > nblk / 4) s = nblk / 4;
  if (s > 32) s = 32;
  return s < 2 ? 1 : static_cast<int>(s);
}

size_t vb_attention_tc_workspace(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t head_dim, int causal) {
  const int s = tc_splits(B, H, Sq, Skv, causal);
  if (s <= 1) return 0;
  return static_cast<size_t>(s) * B * H * Sq * (tc_hd(head_dim) + 2) * sizeof(float);
}

// Returns VB_ERR_UNSUPPORTED when the shape / strides are outside what this kernel handles; the caller
// (vb200_attention) then uses the mma.sync kernel.
int vb_attention_tc(const void* q, const void* k, const void* v, void* out, int64_t B, int64_t H, int64_t Sq,
                    int64_t Skv, int64_t head_dim, int64_t q_sb, int64_t q_ss, int64_t q_sh, int64_t k_sb,
                    int64_t k_ss, int64_t k_sh, int64_t v_sb, int64_t v_ss, int64_t v_sh, int64_t o_sb, int64_t o_ss,
                    int64_t o_sh, float scale, int causal, const int32_t* kv_len, const uint8_t* mask, int64_t m_sb,
                    int64_t m_sh, int64_t m_sq, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (head_dim != 40 && head_dim != 64 && head_dim != 80 && head_dim != 128 && head_dim != 160) return VB_ERR_UNSUPPORTED;
  if (Skv < 1 || Sq < 1) return VB_ERR_UNSUPPORTED;
  const int64_t st[12] = {q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh};
  for (int i = 0; i < 12; ++i)
    if (st[i] % 8 != 0 || st[i] < 0) return VB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return VB_ERR_UNSUPPORTED;
  if (q_ss == 0 || k_ss == 0 || v_ss == 0) return VB_ERR_UNSUPPORTED;
  // a tensor-map dim of extent > 1 needs a non-zero stride: broadcast operands stay on the mma.sync kernel
  if ((B > 1 && (q_sb == 0 || k_sb == 0 || v_sb == 0)) || (H > 1 && (q_sh == 0 || k_sh == 0 || v_sh == 0)))
    return VB_ERR_UNSUPPORTED;
  // extent-1 dims still need a legal (16-byte multiple, non-zero) stride value
  auto fix = [](int64_t stride, int64_t fallback) { return stride == 0 ? fallback : stride; };
  CUtensorMap tq, tk, tv;
  if (int r = make_qkv_map(&tq, q, B, Sq, H, head_dim, fix(q_sb, q_ss * Sq), q_ss, fix(q_sh, q_ss * Sq), TC_BM)) return r;
  if (int r = make_qkv_map(&tk, k, B, Skv, H, head_dim, fix(k_sb, k_ss * Skv), k_ss, fix(k_sh, k_ss * Skv), TC_BN)) return r;
  if (int r = make_qkv_map(&tv, v, B, Skv, H, head_dim, fix(v_sb, v_ss * Skv), v_ss, fix(v_sh, v_ss * Skv), TC_BN)) return r;
  TcAttnParams p;
  p.o = reinterpret_cast<bf16*>(out);
  p.o_sb = o_sb; p.o_ss = o_ss; p.o_sh = o_sh;
  p.B = static_cast<int>(B); p.H = static_cast<int>(H); p.Sq = static_cast<int>(Sq); p.Skv = static_cast<int>(Skv);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.kv_len = kv_len;
  p.D = static_cast<int>(head_dim);
  p.mask = mask;
  p.m_sb = m_sb; p.m_sh = m_sh; p.m_sq = m_sq;
  p.splits = 1;
  p.ws_o = nullptr;
  p.ws_ml = nullptr;
  const size_t need = vb_attention_tc_workspace(B, H, Sq, Skv, head_dim, causal);
  if (need > 0 && kv_len == nullptr && workspace != nullptr && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
      (head_dim % 4) == 0) {
    p.splits = tc_splits(B, H, Sq, Skv, causal);
    p.ws_o = reinterpret_cast<float*>(workspace);
    p.ws_ml = p.ws_o + static_cast<size_t>(p.splits) * B * H * Sq * tc_hd(head_dim);
  }
  if (head_dim <= 64) return launch_tc<64>(tq, tk, tv, p, stream);
  if (head_dim <= 128) return launch_tc<128>(tq, tk, tv, p, stream);
  return launch_tc<192>(tq, tk, tv, p, stream);
}
