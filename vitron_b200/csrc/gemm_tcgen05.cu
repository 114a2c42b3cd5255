// vitron_b200 — bf16 GEMM / implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[M, N] = epilogue( A[M, K] · B[N, K]^T )        (both operands K-major, fp32 accumulate)
//
// One persistent, warp-specialised kernel:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM)
//   warps 2..5  epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue -> global)
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//
// The A operand comes either from a plain 2-D row-major matrix or, for convolutions, straight
// from the NHWC activation tensor through a 4-D tensor map: every (tap, channel-chunk) k-step is
// one TMA box {64 ch, TW, TH, TN} fetched at the tap-shifted coordinate, out-of-bounds pixels are
// zero-filled by the TMA unit (= zero padding), so no im2col buffer ever exists.
//
// Covers (reference file:line in DESIGN.md): LLaMA q/k/v/o/gate/up/down/lm_head, CLIP ViT
// q/k/v/out/fc1/fc2, mm_projector, region MLP, every Linear / Conv2d(3x3,1x1) / Conv3d(3,1,1) of
// UNetSD_I2VGen, SEEM FPN convs + mask einsum, GLIGEN fuser linears.
#include <cstdlib>
#include "gemm_v2.cuh"

namespace vb {

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
  constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_STAGES = 2;
  constexpr int C_STAGE_BYTES = (BLOCK_N >= 32) ? 2 * BLOCK_M * 128 : 0;
  constexpr uint32_t TMEM_COLS = (ACC_STAGES * BLOCK_N <= 32)    ? 32
                                 : (ACC_STAGES * BLOCK_N <= 64)  ? 64
                                 : (ACC_STAGES * BLOCK_N <= 128) ? 128
                                 : (ACC_STAGES * BLOCK_N <= 256) ? 256
                                                                 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* cstage = smem + STAGES * STAGE_BYTES;  // 2 x 16 KB output staging for the TMA-store epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(cstage + C_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  volatile int* fin_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // BLOCK_N floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (p.c_box > 0) prefetch_tmap(&tmap_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
    }
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

  const int total_units = p.m_blocks * p.n_blocks * p.splits;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        TileCoord t = decode_work(p, unit);
        int k0, k1;
        split_range(p, t.split, k0, k1);
        int cw = 0, ch = 0, cn = 0;
        if (p.a_mode == 1) {
          int tiw = t.m_blk % p.tiles_w;
          int rest = t.m_blk / p.tiles_w;
          int tih = rest % p.tiles_h;
          int tin = rest / p.tiles_h;
          cw = tiw * p.tw * p.stride - p.pad_w;
          ch = tih * p.th * p.stride - p.pad_h;
          cn = tin * p.tn;
        }
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], p.a_box_bytes + B_BYTES);
          if (p.a_mode == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, t.m_blk * BLOCK_M);
          } else {
            int tap = kb / p.cin_chunks;
            int cc = kb - tap * p.cin_chunks;
            int dy = tap / p.kw, dx = tap - dy * p.kw;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], cc * BLOCK_K, cw + dx, ch + dy, cn);
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, t.n_blk * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      TileCoord t = decode_work(p, unit);
      int k0, k1;
      split_range(p, t.split, k0, k1);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = k0; kb < k1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle row: +2 in 16-byte units
            tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > k0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (kb == k1 - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (k1 <= k0 && lane == 0) tc_commit(&tmem_full[acc]);  // degenerate (never for K>0)
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue warps
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int ehalf = (warp - 2) >> 2;  // two warps per quadrant: they split the accumulator columns
    const int et = threadIdx.x - 64;    // 0 .. EPI_THREADS-1
    int acc = 0;
    uint32_t acc_phase = 0;
    int cbuf = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      TileCoord t = decode_work(p, unit);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int r = quad * 32 + lane;  // accumulator row owned by this thread
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quad * 32) << 16);

      // ---- output row of this accumulator row
      long long orow = -1;
      if (p.a_mode == 0) {
        long long g = static_cast<long long>(t.m_blk) * BLOCK_M + r;
        if (g < p.M) orow = g;
      } else {
        int tiw = t.m_blk % p.tiles_w;
        int rest = t.m_blk / p.tiles_w;
        int tih = rest % p.tiles_h;
        int tin = rest / p.tiles_h;
        int tx = r % p.tw;
        int r2 = r / p.tw;
        int ty = r2 % p.th;
        int tz = r2 / p.th;
        int x = tiw * p.tw + tx, y = tih * p.th + ty, n = tin * p.tn + tz;
        if (tz < p.tn && x < p.wo && y < p.ho && n < p.nb)
          orow = (static_cast<long long>(n) * p.ho + y) * p.wo + x;
      }
      const int col0 = t.n_blk * BLOCK_N;

      if (p.ws != nullptr) {
        // -------- raw fp32 partials to the workspace (split-K and/or swap-AB)
        float* wsp = p.ws + static_cast<long long>(t.split) * p.ws_split_stride;
#pragma unroll 1
        for (int c = ehalf * 16; c < BLOCK_N; c += 32) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c, v);
          tmem_ld_wait();
          if (orow >= 0) {
            if (p.swap) {
              // C[token = col, feature = orow]; a warp writes 32 consecutive features per token
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                int tok = col0 + c + j;
                if (tok < p.N) wsp[static_cast<long long>(tok) * p.ws_ld + orow] = __uint_as_float(v[j]);
              }
            } else {
              float* dst = wsp + orow * p.ws_ld + col0 + c;
              if (col0 + c + 16 <= p.N && (p.ws_ld & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(dst + j) =
                      make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (col0 + c + j < p.N) dst[j] = __uint_as_float(v[j]);
              }
            }
          }
        }
      } else {
        if constexpr (BLOCK_N >= 32) {
          if (p.c_box > 0) {
            // -------- fused epilogue -> 128B/64B-swizzled smem tile -> TMA store (coalesced, clipped by the
            // tensor map at the ragged edges; conv: one 4-D box per pixel tile, mirroring the A load).
            // The two warps of a TMEM lane quadrant take alternate 32-column accumulator chunks; each keeps
            // the tcgen05.ld and the residual loads of its next chunk in flight while it works on the current one.
            const bool g = p.glu != VB_GLU_NONE;
            const int bw = p.c_box;                       // output columns per box
            const int outs_per_acc = g ? 16 : 32;         // outputs produced by one 32-column accumulator chunk
            const int acc_per_box = bw / outs_per_acc;
            const int nboxes = (g ? BLOCK_N / 2 : BLOCK_N) / bw;
            constexpr int NCHUNKS = BLOCK_N / 32;
            const int row_bytes = bw * 2;
            const int swz_shift = bw == 64 ? 0 : 1, swz_mask = bw == 64 ? 7 : 3;
            const bf16* rb = nullptr;
            if (p.rowbias != nullptr && orow >= 0) rb = p.rowbias + (orow / p.rowbias_rows) * static_cast<long long>(p.N);
            const float rs = (p.rowscale != nullptr && orow >= 0) ? p.rowscale[orow] : 1.f;
            const int n_out_total = g ? (p.N >> 1) : p.N;
            const int ocol0 = g ? col0 / 2 : col0;
            // bias of this tile's accumulator columns -> smem (read back as broadcasts); visible after the
            // bar.sync that opens the first box, and not rewritten before every reader passed the last one
            for (int i = et; i < BLOCK_N; i += EPI_THREADS)
              sbias[i] = (p.bias != nullptr && col0 + i < p.N) ? __bfloat162float(p.bias[col0 + i]) : 0.f;
            const bf16* rrow = (p.residual != nullptr && orow >= 0) ? p.residual + orow * p.ldr + ocol0 : nullptr;
            const bool res_vec = rrow != nullptr && (p.ldr & 7) == 0;
            uint32_t v[32];
            uint4 rnext[4];
            auto issue = [&](int c) {
              tmem_ld_32x32(taddr + c * 32, v);
              if (res_vec && ocol0 + (c + 1) * outs_per_acc <= n_out_total) {
                const uint4* rp = reinterpret_cast<const uint4*>(rrow + c * outs_per_acc);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (q * 8 < outs_per_acc) rnext[q] = __ldg(rp + q);
              }
            };
            if (ehalf < NCHUNKS) issue(ehalf);
            // staging buffers alternate box by box; the store of box b overlaps the epilogue math of box b+1
#pragma unroll 1
            for (int bx = 0; bx < nboxes; ++bx) {
              uint8_t* cb = cstage + (cbuf & 1) * (BLOCK_M * 128);
              ++cbuf;
              if (et == 0) bulk_wait_read<1>();           // the store that last used this buffer has read it
              epi_bar_sync();
#pragma unroll 1
              for (int a = (ehalf + bx * acc_per_box) & 1; a < acc_per_box; a += 2) {
                const int ci = bx * acc_per_box + a;      // accumulator chunk (32 columns); ci % 2 == ehalf
                tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 bv = *reinterpret_cast<const float4*>(sbias + ci * 32 + j);
                  f[j] = fmaf(__uint_as_float(v[j]), rs, bv.x);
                  f[j + 1] = fmaf(__uint_as_float(v[j + 1]), rs, bv.y);
                  f[j + 2] = fmaf(__uint_as_float(v[j + 2]), rs, bv.z);
                  f[j + 3] = fmaf(__uint_as_float(v[j + 3]), rs, bv.w);
                }
                uint4 rcur[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
                if (ci + 2 < NCHUNKS) issue(ci + 2);
                const int gc = col0 + ci * 32;
                if (rb != nullptr) {
                  if (gc + 32 <= p.N && (p.N & 7) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const uint4 u = __ldg(reinterpret_cast<const uint4*>(rb + gc) + q);
                      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                      for (int w = 0; w < 4; ++w) {
                        const float2 r2 = unpack_bf16(uu[w]);
                        f[q * 8 + 2 * w] += r2.x;
                        f[q * 8 + 2 * w + 1] += r2.y;
                      }
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                      if (gc + j < p.N) f[j] += __bfloat162float(rb[gc + j]);
                  }
                }
                int nout = 32;
                if (g) {
                  if (p.glu == VB_GLU_SWIGLU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = silu(f[j]) * f[j + 16];
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = f[j] * gelu_erf_fast(f[j + 16]);
                  }
                  nout = 16;
                } else if (p.act == VB_ACT_GELU) {       // the switch hoisted out of the 32-wide loop
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = gelu_erf_fast(f[j]);
                } else if (p.act == VB_ACT_SILU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = silu(f[j]);
                } else if (p.act != VB_ACT_NONE) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
                }
                const int oc = ocol0 + ci * outs_per_acc;  // first output column of these values
                if (rrow != nullptr) {
                  if (res_vec && oc + nout <= n_out_total) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      if (q * 8 < nout) {
                        const uint32_t uu[4] = {rcur[q].x, rcur[q].y, rcur[q].z, rcur[q].w};
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                          const float2 r2 = unpack_bf16(uu[w]);
                          f[q * 8 + 2 * w] = fmaf(p.alpha, f[q * 8 + 2 * w], r2.x);
                          f[q * 8 + 2 * w + 1] = fmaf(p.alpha, f[q * 8 + 2 * w + 1], r2.y);
                        }
                      }
                    }
                  } else {
                    const bf16* rp = rrow + ci * outs_per_acc;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                      if (j < nout && oc + j < n_out_total) f[j] = __bfloat162float(rp[j]) + p.alpha * f[j];
                  }
                } else if (p.alpha != 1.0f) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
                }
                // 16-byte pieces into the swizzled staging row of this thread
                uint8_t* rowp = cb + r * row_bytes;
                const int piece0 = (a * outs_per_acc) >> 3;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  if (j < nout) {
                    const int piece = (piece0 + (j >> 3)) ^ ((r >> swz_shift) & swz_mask);
                    *reinterpret_cast<uint4*>(rowp + piece * 16) =
                        make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                   pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
                  }
                }
              }
              fence_proxy_async_smem();
              epi_bar_sync();
              if (et == 0) {
                const int cc = ocol0 + bx * bw;
                if (cc < n_out_total) {
                  if (p.a_mode == 0) {
                    tma_store_2d(&tmap_c, cb, cc, t.m_blk * BLOCK_M);
                  } else {
                    const int tiw = t.m_blk % p.tiles_w, rest = t.m_blk / p.tiles_w;
                    tma_store_4d(&tmap_c, cb, cc, tiw * p.tw, (rest % p.tiles_h) * p.th, (rest / p.tiles_h) * p.tn);
                  }
                }
                bulk_commit();
              }
            }
          }
        }
        if (BLOCK_N < 32 || p.c_box == 0) {
        // -------- fused epilogue straight to the output tensor
        const bf16* rb = nullptr;
        if (p.rowbias != nullptr && orow >= 0)
          rb = p.rowbias + (orow / p.rowbias_rows) * static_cast<long long>(p.N);
        constexpr int CH = (BLOCK_N >= 32) ? 32 : 16;
#pragma unroll 1
        for (int c = ehalf * CH; c < BLOCK_N; c += 2 * CH) {
          float f[CH];
          if constexpr (CH == 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          } else {
            uint32_t v[16];
            tmem_ld_32x16(taddr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          }
          if (orow < 0) continue;
          const int gc = col0 + c;  // first accumulator column of this chunk
          if (gc >= p.N) continue;
          if (p.rowscale != nullptr) {
            const float rs = p.rowscale[orow];
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] *= rs;
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (gc + j < p.N) f[j] += __bfloat162float(p.bias[gc + j]);
          }
          if (rb != nullptr) {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (gc + j < p.N) f[j] += __bfloat162float(rb[gc + j]);
          }
          int nout = CH, oc = gc;
          if (p.glu != VB_GLU_NONE) {
            // packed GLU: accumulator columns come in blocks of 32 = [16 x "a" | 16 x "b"]
            if constexpr (CH == 32) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a = f[j], b = f[j + 16];
                f[j] = (p.glu == VB_GLU_SWIGLU) ? silu(a) * b : a * gelu_erf(b);
              }
            }
            nout = 16;
            oc = gc >> 1;
          } else if (p.act != VB_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] = apply_act(f[j], p.act);
          }
          const int n_out_total = (p.glu != VB_GLU_NONE) ? (p.N >> 1) : p.N;
          if (p.residual != nullptr) {
            const bf16* rp = p.residual + orow * p.ldr + oc;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (j < nout && oc + j < n_out_total) f[j] = __bfloat162float(rp[j]) + p.alpha * f[j];
          } else if (p.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < CH; ++j) f[j] *= p.alpha;
          }
          if (p.out_fp32) {
            float* dst = reinterpret_cast<float*>(p.out) + orow * p.ldo + oc;
            if (oc + nout <= n_out_total && (p.ldo & 3) == 0) {
#pragma unroll
              for (int j = 0; j < CH; j += 4)
                if (j < nout)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (j < nout && oc + j < n_out_total) dst[j] = f[j];
            }
          } else {
            bf16* dst = reinterpret_cast<bf16*>(p.out) + orow * p.ldo + oc;
            if (oc + nout <= n_out_total && (p.ldo & 7) == 0) {
#pragma unroll
              for (int j = 0; j < CH; j += 8)
                if (j < nout)
                  *reinterpret_cast<uint4*>(dst + j) =
                      make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                 pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (j < nout && oc + j < n_out_total) dst[j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
        }
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }

      if (p.ws != nullptr && p.counters != nullptr) {
        // -------- split-K finalisation without a second kernel: the CTA that completes a tile last
        // sums the partials in split order (deterministic) and applies the fused epilogue.
        const int tile = unit / p.splits;
        __threadfence();
        epi_bar_sync();
        if (et == 0) {
          const int prev = atomicAdd(&p.counters[tile], 1);
          *fin_flag = (prev == p.splits - 1) ? 1 : 0;
        }
        epi_bar_sync();
        if (*fin_flag) {
          __threadfence();
          int r0, r1, c0, c1;  // tile extent in C orientation
          if (p.swap) { r0 = 0; r1 = p.rows_c; c0 = t.m_blk * BLOCK_M; c1 = min(p.cols_c, c0 + BLOCK_M); }
          else { r0 = t.m_blk * BLOCK_M; r1 = min(p.rows_c, r0 + BLOCK_M); c0 = col0; c1 = min(p.cols_c, c0 + BLOCK_N); }
          const bool g = p.glu != VB_GLU_NONE;
          const int oc0 = g ? c0 / 2 : c0, oc1 = g ? c1 / 2 : c1;
          const int chunks = (oc1 - oc0 + 7) / 8;
          const int items = (r1 - r0) * chunks;
          for (int it = et; it < items; it += EPI_THREADS) {
            const int row = r0 + it / chunks, oc = oc0 + (it % chunks) * 8;
            reduce_item(p.ws, p.splits, p.ws_split_stride, p.ws_ld, row, oc, p.cols_c, p.out, p.ldo, p.bias,
                        p.rowbias, p.rowbias_rows, p.residual, p.ldr, p.alpha, p.act, p.glu, p.out_fp32, p.rowscale);
          }
          if (et == 0) p.counters[tile] = 0;
        }
        epi_bar_sync();  // fin_flag is reused by the next unit
      }
    }
  }

  if (threadIdx.x == 64 && p.c_box > 0) bulk_wait<0>();  // staged tiles must be read out before smem goes away
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ split-K / swap reduce
// out[m, n'] = epilogue( sum_s ws[s, m, n] ), same epilogue semantics as the fused path.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits,
                                     long long split_stride, long long ws_ld, int rows, int ncols,
                                     void* out, long long ldo, const bf16* __restrict__ bias,
                                     const bf16* __restrict__ rowbias, int rowbias_rows,
                                     const bf16* __restrict__ residual, long long ldr, float alpha,
                                     int act, int glu, int out_fp32, const float* __restrict__ rowscale) {
  pdl_trigger();
  pdl_wait();
  const int n_out_total = glu != VB_GLU_NONE ? ncols / 2 : ncols;
  const int groups = (n_out_total + 7) / 8;
  long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * groups) return;
  reduce_item(ws, splits, split_stride, ws_ld, static_cast<int>(idx / groups), static_cast<int>(idx % groups) * 8,
              ncols, out, ldo, bias, rowbias, rowbias_rows, residual, ldr, alpha, act, glu, out_fp32, rowscale);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

static int make_tmap(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box,
                     const uint32_t* estr, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return VB_ERR_DRIVER;
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims,
                  strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VB_OK : VB_ERR_DRIVER;
}

template <int BN, int STAGES>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                      cudaStream_t stream) {
  constexpr int smem = STAGES * (BLOCK_M * BLOCK_K * 2 + BN * BLOCK_K * 2) + (BN >= 32 ? 2 * BLOCK_M * 128 : 0) + 1024 + 256 + 1024;
  static_assert(smem <= SMEM_LIMIT, "smem budget");
  static bool attr_set = false;
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { vb_set_last_error(e); return VB_ERR_CUDA; }
    attr_set = true;
  }
  int units = p.m_blocks * p.n_blocks * p.splits;
  int grid = units < vb_num_sms() ? units : vb_num_sms();
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, tc, p);
  VB_LAUNCH_CHECK();
  return VB_OK;
}

static int launch_gemm(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                       const GemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 256: return launch_cfg<256, 4>(ta, tb, tc, p, stream);
    case 160: return launch_cfg<160, 5>(ta, tb, tc, p, stream);
    case 128: return launch_cfg<128, 6>(ta, tb, tc, p, stream);
    case 64: return launch_cfg<64, 7>(ta, tb, tc, p, stream);
    case 32: return launch_cfg<32, 8>(ta, tb, tc, p, stream);
    case 16: return launch_cfg<16, 10>(ta, tb, tc, p, stream);
    default: return VB_ERR_ARG;
  }
}

static int pick_block_n(long long M, long long N, int glu) {
  // widest tile wins unless it leaves SMs idle or pads many columns:
  // score = tile efficiency x wave utilisation x column fill. 160 divides the UNet widths 320/640/1280.
  (void)glu;
  if (const char* force = getenv("VB200_FORCE_BN")) {  // tuning knob for tools/bn_sweep.py
    const int v = atoi(force);
    if (v == 256 || v == 160 || v == 128 || v == 64 || v == 32) return v;
  }
  const int sms = vb_num_sms();
  const long long mb = (M + BLOCK_M - 1) / BLOCK_M;
  const int cands[5] = {256, 160, 128, 64, 32};
  // relative per-tile efficiency measured on B200 (narrow tiles are L2->SMEM bandwidth bound)
  const float eff[5] = {1.0f, 0.80f, 0.66f, 0.40f, 0.22f};
  int best = 32;
  float best_score = -1.f;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    const long long nbk = (N + bn - 1) / bn;
    const long long tiles = mb * nbk;
    const long long waves = (tiles + sms - 1) / sms;
    const float util = static_cast<float>(tiles) / static_cast<float>(waves * sms);
    const float fill = static_cast<float>(N) / static_cast<float>(nbk * bn);  // padded columns
    const float score = eff[i] * util * fill;
    if (score > best_score) { best_score = score; best = bn; }
  }
  return best;
}

constexpr size_t GEMM_COUNTER_BYTES = 16384;  // 4096 tile counters at the head of the workspace

static int swap_splits(long long M, long long N, long long K) {
  // maximise (CTA wave utilisation) x (k-block balance across splits); fewer splits win ties
  const int sms = vb_num_sms();
  const int m_blocks = static_cast<int>((N + BLOCK_M - 1) / BLOCK_M);
  const int kblocks = static_cast<int>((K + BLOCK_K - 1) / BLOCK_K);
  int best = 1;
  float best_score = -1.f;
  for (int s = 1; s <= 32 && s <= kblocks; ++s) {
    if (static_cast<size_t>(s) * M * N * sizeof(float) > (static_cast<size_t>(96) << 20)) break;
    const long long units = static_cast<long long>(m_blocks) * s;
    const long long waves = (units + sms - 1) / sms;
    const float util = static_cast<float>(units) / static_cast<float>(waves * sms);
    const int per = (kblocks + s - 1) / s;
    const float bal = static_cast<float>(kblocks) / static_cast<float>(s * per);
    // every extra wave / split costs a fixed ramp (pipeline fill, partial write + reduction)
    const float score = util * bal - 0.01f * static_cast<float>(s);
    if (score > best_score + 1e-6f) { best_score = score; best = s; }
  }
  return best;
}

// output box width of the TMA-store epilogue: 64 (128B swizzle) or 32 (64B swizzle) columns, 0 = direct stores
static int c_box_for(int bn, int glu, int out_fp32, long long ldo, const void* out) {
  if (out_fp32 || bn < 32 || (ldo & 7) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return 0;
  const int out_w = glu != VB_GLU_NONE ? bn / 2 : bn;
  if (out_w % 64 == 0) return 64;
  if (out_w % 32 == 0) return 32;
  return 0;
}

static int validate_epi(const vb_epilogue* e, long long N) {
  if (e->glu != VB_GLU_NONE && (N % 32) != 0) return VB_ERR_ARG;
  if (e->act < VB_ACT_NONE || e->act > VB_ACT_SILU) return VB_ERR_ARG;
  if (e->glu < VB_GLU_NONE || e->glu > VB_GLU_GEGLU) return VB_ERR_ARG;
  return VB_OK;
}


// ------------------------------------------------------------------ v2 planning
static int g_gemm_impl = 0;  // 0: v2 whenever eligible, 1: generic kernel only (A/B measurements, parity tests of both)
static int g_gemm_resb = 1;  // resident-B variant: bit 0 = K <= 320 (5 k-blocks), bit 1 = also K <= 640 at BN = 128; 0 = off
                             // bit 2: force the CTA-pair (cta_group::2) variant wherever it applies; bit 3: never use it
static int g_gemm_dbg = 0;   // GemmParams.dbg of the v2 launches (measurement aid)
int resb_smem_bytes(int bn, int nkb);  // gemm_v2_resb.cu

// Tile width of the resident-B variant for M x N x (nkb k-blocks), or 0 when it does not apply: the widest slab that
// fits next to an A ring deep enough to cover the L2 latency (BN = 160 with 6 A stages at K <= 320; BN = 128 at K <= 640).
static int resb_tile(long long M, long long N, int nkb, int need) {
  if (!(g_gemm_resb & 3) || (need != 0 && need != F_RES && need != F_GLU && need != F_ACT)) return 0;
  if (nkb > 5 && !(g_gemm_resb & 2)) return 0;
  if (M < 4096) return 0;                       // few m-tiles: nothing to amortise the slab load over
  const int cands[3] = {160, 256, 128};
  int best = 0;
  float best_cost = 1e30f;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (resb_smem_bytes(bn, nkb) > SMEM_LIMIT) continue;
    const long long nb = (N + bn - 1) / bn;
    if (nb > vb_num_sms()) continue;
    const int groups = vb_num_sms() / static_cast<int>(nb);
    const long long mb = (M + BLOCK_M - 1) / BLOCK_M;
    const long long per_cta = (mb + groups - 1) / groups;     // m-tiles of the busiest CTA
    // per tile: MMA ~ bn / 256 units per k-block, A stream ~ 0.55 unit per k-block with a deep ring (BN 160), 0.9 with 3 stages
    const float mma = static_cast<float>(bn) / 256.f, astream = bn == 160 ? 0.55f : 0.9f;
    const float cost = static_cast<float>(per_cta) * nkb * (mma > astream ? mma : astream);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Tile width of the CTA-pair (cta_group::2, 256-row tile) variant for `mb` m-blocks, or 0 when it does not apply. Time per
// k-step of a pair in units of the 1-CTA 128x256x64 step: operand bytes per SM drop from 16 + BN/8 KB to 16 + BN/16 KB.
static int cl_tile(long long mb, long long N, int nkb, int need, float* t_out) {
  if (need != 0 && need != F_RES && need != F_GLU && need != F_ACT && need != F_RB && need != (F_RB | F_RES)) return 0;
  if (mb < 2) return 0;
  // measured (profiles/r02_gemm_pair_vs_1cta.jsonl): 3-8 % faster than the 1-CTA kernel from K = 4096 up (LLM prefill), equal
  // or slower on the short-K UNet shapes, whose time is tile hand-over + epilogue + output stores, not operand traffic
  if (nkb < 32 && !(g_gemm_resb & 4)) return 0;
  const int cands[3] = {256, 160, 128};
  const float tk[3] = {0.93f, 0.74f, 0.70f};
  int best = 0;
  float best_t = 1e30f;
  const int clusters = vb_num_sms() / 2;
  for (int i = 0; i < 3; ++i) {
    const long long nb = (N + cands[i] - 1) / cands[i];
    const long long units = ((mb + 1) / 2) * nb;
    const long long waves = (units + clusters - 1) / clusters;
    const float t = static_cast<float>(waves) * (nkb * tk[i]) + 2.0f * cands[i] / 256.f;
    if (t < best_t - 1e-4f) { best_t = t; best = cands[i]; }
  }
  if (t_out) *t_out = best_t;
  return best;
}

struct TilePlan {
  int bn, splits;
  float t;   // modelled time (units of one 1-CTA 128x256x64 k-step)
};

// Tile width and split-K factor for `mb` 128-row m-blocks x N columns x `kblocks` k-steps. Time model in units of one
// 128x256x64 k-step: per-unit main loop = k-steps x tk[bn] (narrow tiles are L2->smem bound: measured in round 1,
// profiles/r01_tile_width_sweep.jsonl), waves = ceil(units / SMs); a split costs its partial write + the finaliser.
static TilePlan plan_tiles(long long mb, long long N, int kblocks, bool allow_split, long long out_rows) {
  const int sms = vb_num_sms();
  const int cands[5] = {256, 160, 128, 64, 32};
  const float tk[5] = {1.0f, 0.78f, 0.76f, 0.625f, 0.57f};
  TilePlan best = {256, 1, 1e30f};
  float best_t = 1e30f;
  if (const char* force = getenv("VB200_FORCE_BN")) {
    const int v = atoi(force);
    if (v == 256 || v == 160 || v == 128 || v == 64 || v == 32) return {v, 1, 0.f};
  }
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    const long long nbk = (N + bn - 1) / bn;
    const long long tiles = mb * nbk;
    if (tiles > 4096 && allow_split) { /* counters cover 4096 tiles: no split for such grids (they fill the GPU anyway) */ }
    for (int sp = 1; sp <= 8; ++sp) {
      if (sp > 1) {
        if (!allow_split || tiles > 4096 || tiles * 10 >= static_cast<long long>(sms) * 7) break;
        if (kblocks / sp < 4) break;
        if (static_cast<unsigned long long>(sp) * out_rows * N * sizeof(float) > (64ull << 20)) break;
      }
      const long long units = tiles * sp;
      const long long waves = (units + sms - 1) / sms;
      const int kpu = (kblocks + sp - 1) / sp;
      float t = static_cast<float>(waves) * (kpu * tk[i] + (sp > 1 ? 1.5f : 0.f));
      t += 2.0f * bn / 256.f;
      // a split writes and re-reads sp fp32 copies of the output (~5 TB/s through L2; one unit = 0.27 us) and pays one
      // more launch (~2.5 us)
      if (sp > 1) t += static_cast<float>(sp) * out_rows * N * 8.0f / 5.0e6f / 0.27f + 9.0f;
      if (t < best_t - 1e-4f) { best_t = t; best = {bn, sp, t}; }
    }
  }
  return best;
}

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

// can the v2 kernel (256-bit epilogue accesses) serve this call?
static bool v2_eligible(const vb_epilogue* e, const void* out, long long ldo, long long N) {
  if (g_gemm_impl == 1) return false;
  if (!aligned32(out)) return false;
  if (e->out_fp32 ? (ldo % 8) != 0 : (ldo % 16) != 0) return false;
  if (e->bias && !aligned32(e->bias)) return false;
  if (e->rowbias && (!aligned32(e->rowbias) || (N % 16) != 0)) return false;
  if (e->residual && (!aligned32(e->residual) || (e->ldr % 16) != 0)) return false;
  if ((N % 32) != 0 && (N % 16) != 0) return false;  // ragged tail handled per element, rows must stay 32-byte aligned
  return true;
}

static int v2_need(const vb_epilogue* e, int splits) {
  if (splits > 1) return F_WS;
  int need = 0;
  if (e->rowscale) need |= F_RS;
  if (e->rowbias) need |= F_RB;
  if (e->residual || e->alpha != 1.0f) need |= F_RES;
  if (e->act != VB_ACT_NONE) need |= F_ACT;
  if (e->glu != VB_GLU_NONE) need |= F_GLU;
  if (e->out_fp32) need |= F_F32;
  return need;
}

static int launch_gemm_v2(int bn, int need, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          cudaStream_t stream) {
  int r;
  switch (bn) {
    case 256: r = launch_gemm_v2_256(need, ta, tb, p, stream); break;
    case 160: r = launch_gemm_v2_160(need, ta, tb, p, stream); break;
    case 128: r = launch_gemm_v2_128(need, ta, tb, p, stream); break;
    case 64: r = launch_gemm_v2_64(need, ta, tb, p, stream); break;
    case 32: r = launch_gemm_v2_32(need, ta, tb, p, stream); break;
    default: return VB_ERR_ARG;
  }
  if (r != VB_OK || !(need & F_WS)) return r;
  // split-K: sum the partials in split order + the fused epilogue, one thread per 8 outputs of a row
  const int n_out = p.glu != VB_GLU_NONE ? p.N / 2 : p.N;
  const long long items = static_cast<long long>(p.M) * ((n_out + 7) / 8);
  cudaError_t le = vb_launch(splitk_reduce_kernel, dim3(static_cast<unsigned>((items + 255) / 256)), dim3(256), 0, stream,
                             static_cast<const float*>(p.ws), p.splits, p.ws_split_stride, p.ws_ld, p.M, p.N, p.out, p.ldo, p.bias,
                             p.rowbias, p.rowbias_rows, p.residual, p.ldr, p.alpha, p.act, p.glu, p.out_fp32, p.rowscale);
  if (le != cudaSuccess) { vb_set_last_error(le); return VB_ERR_CUDA; }
  return VB_OK;
}

// conv output-pixel tile tw x th x tn (<= 128 accumulator rows): the shape that covers the output with the fewest
// tiles; ties go to the widest tile (longest contiguous runs for the TMA boxes)
static void conv_pixel_tile(int wo, int ho, long long nb, int& tw, int& th, int& tn) {
  tw = th = tn = 1;
  long long best = -1;
  for (int cw = wo < 128 ? wo : 128; cw >= 1; --cw) {
    const int hmax = 128 / cw < ho ? 128 / cw : ho;
    for (int ch = hmax; ch >= 1; --ch) {
      int cn = 128 / (cw * ch);
      if (cn > nb) cn = static_cast<int>(nb);
      const long long blocks = static_cast<long long>((wo + cw - 1) / cw) * ((ho + ch - 1) / ch) * ((nb + cn - 1) / cn);
      if (best < 0 || blocks < best) { best = blocks; tw = cw; th = ch; tn = cn; }
    }
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_set_gemm_debug(int resident_b, int dbg) {
  const int prev = g_gemm_resb | (g_gemm_dbg << 8);
  if (resident_b >= 0) g_gemm_resb = resident_b;
  if (dbg >= 0) g_gemm_dbg = dbg;
  return prev;
}

extern "C" int vb200_set_gemm_impl(int impl) {
  const int prev = g_gemm_impl;
  if (impl == 0 || impl == 1) g_gemm_impl = impl;
  return prev;
}

int vb_launch_gemv(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int64_t M,
                   int64_t N, int64_t K, const vb_epilogue* e, cudaStream_t stream);  // gemv.cu

// 17..64 rows: swap-AB + split-K (weights in the 128-row MMA slot, every SM streams a K slice) pays off only when the
// weight matrix is large enough to be an HBM stream (LLM decode at batch 17..64: 33-180 MB per GEMM). For the small
// matrices of the SEEM / GLIGEN token paths it measured 25.8 us per dependent GEMM (M = 64, N = K = 512) against 4.8 us on
// the persistent kernel with a mostly empty 128-row tile (profiles/r02_smallm_chain_latency.jsonl).
static bool use_swap(int64_t M, int64_t N, int64_t K) { return M <= 64 && N * K > (4ll << 20); }

extern "C" size_t vb200_gemm_bf16_workspace_size(int64_t M, int64_t N, int64_t K) {
  // M <= 16: weight-streaming kernel (gemv.cu), no workspace; 16 < M <= 64 and a long weight stream: swap-AB + split-K
  if (M <= 16 || N <= 0 || K <= 0) return 0;
  // [tile counters | fp32 partials]; the caller zero-fills it ONCE, the kernels leave the counters zeroed
  if (use_swap(M, N, K))
    return GEMM_COUNTER_BYTES + static_cast<size_t>(swap_splits(M, N, K)) * static_cast<size_t>(M) *
                                    static_cast<size_t>(N) * sizeof(float);
  // M > 64: split-K only when the tile set cannot fill the SMs (plan_tiles)
  const TilePlan pl = plan_tiles((M + BLOCK_M - 1) / BLOCK_M, N, static_cast<int>((K + BLOCK_K - 1) / BLOCK_K), true, M);
  if (pl.splits <= 1) return 0;
  return GEMM_COUNTER_BYTES + static_cast<size_t>(pl.splits) * static_cast<size_t>(M) * static_cast<size_t>(N) * sizeof(float);
}

extern "C" int vb200_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out,
                               int64_t ldo, int64_t M, int64_t N, int64_t K,
                               const vb_epilogue* epi, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
  VB_CHECK_ARG(A && W && out && epi);
  VB_CHECK_ARG(M > 0 && N > 0 && K > 0);
  VB_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && lda >= K && ldw >= K);
  VB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (int r = validate_epi(epi, N)) return r;
  if (M <= 16) return vb_launch_gemv(A, lda, W, ldw, out, ldo, M, N, K, epi, stream);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.num_k_blocks = static_cast<int>((K + BLOCK_K - 1) / BLOCK_K);
  p.a_mode = 0;
  p.a_box_bytes = BLOCK_M * BLOCK_K * 2;
  p.bias = reinterpret_cast<const bf16*>(epi->bias);
  p.rowbias = reinterpret_cast<const bf16*>(epi->rowbias);
  p.rowbias_rows = epi->rowbias_rows > 0 ? epi->rowbias_rows : 1;
  p.residual = reinterpret_cast<const bf16*>(epi->residual);
  p.ldr = epi->ldr;
  p.alpha = epi->alpha;
  p.act = epi->act;
  p.glu = epi->glu;
  p.out_fp32 = epi->out_fp32;
  p.rowscale = epi->rowscale;
  VB_CHECK_ARG(epi->rms_eps <= 0.f || epi->rowscale != nullptr);  // in-kernel rstd exists for M <= 16 only
  p.out = out;
  p.ldo = ldo;

  CUtensorMap ta, tb;
  const uint32_t estr2[2] = {1, 1};
  const bool swap = use_swap(M, N, K);  // small M, long weight stream: weights take the 128-row MMA slot
  if (swap) {
    // kernel-M = N (weight rows), kernel-N = M (tokens); partials -> workspace -> reduce kernel
    const int bn = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
    p.M = static_cast<int>(N);
    p.N = static_cast<int>(M);
    p.m_blocks = static_cast<int>((N + BLOCK_M - 1) / BLOCK_M);
    p.n_blocks = 1;
    const int splits = swap_splits(M, N, K);
    p.splits = splits;
    size_t need = GEMM_COUNTER_BYTES + static_cast<size_t>(splits) * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) return VB_ERR_WORKSPACE;
    if (static_cast<size_t>(p.m_blocks) * sizeof(int) > GEMM_COUNTER_BYTES) return VB_ERR_UNSUPPORTED;
    p.swap = 1;
    p.counters = reinterpret_cast<int*>(workspace);
    p.rows_c = static_cast<int>(M);
    p.cols_c = static_cast<int>(N);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + GEMM_COUNTER_BYTES);
    p.ws_ld = N;
    p.ws_split_stride = M * N;
    uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t sA[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t bA[2] = {BLOCK_K, BLOCK_M};
    if (int r = make_tmap(&ta, W, 2, dA, sA, bA, estr2)) return r;
    uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t sB[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(bn)};
    if (int r = make_tmap(&tb, A, 2, dB, sB, bB, estr2)) return r;
    return launch_gemm(bn, ta, tb, ta, p, stream);  // the last CTA of every tile finalises it (no TMA store)
  }

  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.m_blocks = static_cast<int>((M + BLOCK_M - 1) / BLOCK_M);
  if (v2_eligible(epi, out, ldo, N)) {
    p.dbg = g_gemm_dbg;
    if (const int rbn = resb_tile(M, N, p.num_k_blocks, v2_need(epi, 1))) {
      p.n_blocks = static_cast<int>((N + rbn - 1) / rbn);
      p.splits = 1;
      uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
      uint64_t sA[1] = {static_cast<uint64_t>(lda) * 2};
      uint32_t bA[2] = {BLOCK_K, BLOCK_M};
      if (int r = make_tmap(&ta, A, 2, dA, sA, bA, estr2)) return r;
      uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
      uint64_t sB[1] = {static_cast<uint64_t>(ldw) * 2};
      uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(rbn)};
      if (int r = make_tmap(&tb, W, 2, dB, sB, bB, estr2)) return r;
      const int r = launch_gemm_v2_resb(rbn, v2_need(epi, 1), ta, tb, p, stream);
      if (r != VB_ERR_UNSUPPORTED) return r;
    }
    TilePlan pl = plan_tiles(p.m_blocks, N, p.num_k_blocks, true, M);
    if (pl.splits > 1) {
      const size_t need_ws = GEMM_COUNTER_BYTES + static_cast<size_t>(pl.splits) * M * N * sizeof(float);
      if (workspace == nullptr || workspace_bytes < need_ws || !aligned32(workspace)) pl = plan_tiles(p.m_blocks, N, p.num_k_blocks, false, M);
    }
    // CTA pair (cta_group::2) when its modelled time beats the best 1-CTA plan
    float t_cl = 0.f;
    const int cbn = (g_gemm_resb & 8) ? 0 : cl_tile(p.m_blocks, N, p.num_k_blocks, v2_need(epi, 1), &t_cl);
    if (cbn && ((g_gemm_resb & 4) || t_cl < pl.t)) {
      p.n_blocks = static_cast<int>((N + cbn - 1) / cbn);
      p.splits = 1;
      uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
      uint64_t sA[1] = {static_cast<uint64_t>(lda) * 2};
      uint32_t bA[2] = {BLOCK_K, BLOCK_M};
      if (int r = make_tmap(&ta, A, 2, dA, sA, bA, estr2)) return r;
      uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
      uint64_t sB[1] = {static_cast<uint64_t>(ldw) * 2};
      uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(cbn / 2)};   // each CTA of the pair loads half of the B tile
      if (int r = make_tmap(&tb, W, 2, dB, sB, bB, estr2)) return r;
      const int r = launch_gemm_v2_cl(cbn, v2_need(epi, 1), ta, tb, p, stream);
      if (r != VB_ERR_UNSUPPORTED) return r;
    }
    p.n_blocks = static_cast<int>((N + pl.bn - 1) / pl.bn);
    p.splits = pl.splits;
    if (pl.splits > 1) {
      p.counters = reinterpret_cast<int*>(workspace);
      p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + GEMM_COUNTER_BYTES);
      p.ws_ld = N;
      p.ws_split_stride = M * N;
      p.vec_ok = (N % 8) == 0;
    }
    uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t sA[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t bA[2] = {BLOCK_K, BLOCK_M};
    if (int r = make_tmap(&ta, A, 2, dA, sA, bA, estr2)) return r;
    uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t sB[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(pl.bn)};
    if (int r = make_tmap(&tb, W, 2, dB, sB, bB, estr2)) return r;
    return launch_gemm_v2(pl.bn, v2_need(epi, pl.splits), ta, tb, p, stream);
  }
  int bn = pick_block_n(M, N, epi->glu);
  p.n_blocks = static_cast<int>((N + bn - 1) / bn);
  p.splits = 1;
  uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  uint64_t sA[1] = {static_cast<uint64_t>(lda) * 2};
  uint32_t bA[2] = {BLOCK_K, BLOCK_M};
  if (int r = make_tmap(&ta, A, 2, dA, sA, bA, estr2)) return r;
  uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
  uint64_t sB[1] = {static_cast<uint64_t>(ldw) * 2};
  uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(bn)};
  if (int r = make_tmap(&tb, W, 2, dB, sB, bB, estr2)) return r;
  CUtensorMap tc = ta;
  p.c_box = c_box_for(bn, epi->glu, epi->out_fp32, ldo, out);
  if (p.c_box > 0) {
    const long long n_out = epi->glu != VB_GLU_NONE ? N / 2 : N;
    uint64_t dC[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
    uint64_t sC[1] = {static_cast<uint64_t>(ldo) * 2};
    uint32_t bC[2] = {static_cast<uint32_t>(p.c_box), BLOCK_M};
    if (int r = make_tmap(&tc, out, 2, dC, sC, bC, estr2,
                          p.c_box == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
      return r;
  }
  return launch_gemm(bn, ta, tb, tc, p, stream);
}

static void conv_out_dims(int64_t h, int64_t w, int kh, int kw, int stride, int pad_h, int pad_w, int& ho, int& wo) {
  ho = static_cast<int>((h + 2 * pad_h - kh) / stride + 1);
  wo = static_cast<int>((w + 2 * pad_w - kw) / stride + 1);
}

extern "C" size_t vb200_conv_nhwc_workspace_size(int64_t nb, int64_t h, int64_t w, int64_t cin, int64_t cout, int kh,
                                                 int kw, int stride, int pad_h, int pad_w) {
  // split-K partials for convolutions whose pixel tiles cannot fill the SMs (0 for every other shape);
  // [tile counters | fp32 partials], zero-filled ONCE by the caller, left zeroed by the kernel
  if (nb <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || kh <= 0 || kw <= 0 || (stride != 1 && stride != 2)) return 0;
  int ho, wo;
  conv_out_dims(h, w, kh, kw, stride, pad_h, pad_w, ho, wo);
  if (ho <= 0 || wo <= 0) return 0;
  int tw, th, tn;
  conv_pixel_tile(wo, ho, nb, tw, th, tn);
  const long long mb = static_cast<long long>((wo + tw - 1) / tw) * ((ho + th - 1) / th) * ((nb + tn - 1) / tn);
  const int kblocks = kh * kw * static_cast<int>((cin + BLOCK_K - 1) / BLOCK_K);
  const long long rows = nb * ho * wo;
  const TilePlan pl = plan_tiles(mb, cout, kblocks, true, rows);
  if (pl.splits <= 1) return 0;
  return GEMM_COUNTER_BYTES + static_cast<size_t>(pl.splits) * rows * cout * sizeof(float);
}

extern "C" int vb200_conv_nhwc_bf16(const void* X, const void* Wt, void* out, int64_t nb,
                                    int64_t h, int64_t w, int64_t cin, int64_t cout, int kh,
                                    int kw, int stride, int pad_h, int pad_w,
                                    const vb_epilogue* epi, void* workspace, size_t workspace_bytes,
                                    cudaStream_t stream) {
  // X [nb, h, w, cin] bf16 NHWC; Wt [cout, kh*kw, cin_pad] bf16 with cin_pad = ceil64(cin)
  // (zero padded); out [nb, ho, wo, cout(/2 if GLU)].
  VB_CHECK_ARG(X && Wt && out && epi);
  VB_CHECK_ARG(nb > 0 && h > 0 && w > 0 && cin > 0 && cout > 0 && kh > 0 && kw > 0);
  VB_CHECK_ARG(stride == 1 || stride == 2);
  VB_CHECK_ARG((cin % 8) == 0);
  if (int r = validate_epi(epi, cout)) return r;
  int ho, wo;
  conv_out_dims(h, w, kh, kw, stride, pad_h, pad_w, ho, wo);
  VB_CHECK_ARG(ho > 0 && wo > 0);
  const int cin_chunks = static_cast<int>((cin + BLOCK_K - 1) / BLOCK_K);
  const int cin_pad = cin_chunks * BLOCK_K;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int tw, th, tn;
  conv_pixel_tile(wo, ho, nb, tw, th, tn);
  p.a_mode = 1;
  p.cin_chunks = cin_chunks;
  p.kw = kw;
  p.stride = stride;
  p.pad_h = pad_h;
  p.pad_w = pad_w;
  p.tw = tw; p.th = th; p.tn = tn;
  p.wo = wo; p.ho = ho; p.nb = static_cast<int>(nb);
  p.tiles_w = (wo + tw - 1) / tw;
  p.tiles_h = (ho + th - 1) / th;
  int tiles_n = static_cast<int>((nb + tn - 1) / tn);
  p.m_blocks = p.tiles_w * p.tiles_h * tiles_n;
  p.M = static_cast<int>(nb) * ho * wo;
  p.N = static_cast<int>(cout);
  p.num_k_blocks = kh * kw * cin_chunks;
  p.splits = 1;
  p.a_box_bytes = static_cast<uint32_t>(tw * th * tn) * BLOCK_K * 2;
  p.bias = reinterpret_cast<const bf16*>(epi->bias);
  p.rowbias = reinterpret_cast<const bf16*>(epi->rowbias);
  p.rowbias_rows = epi->rowbias_rows > 0 ? epi->rowbias_rows : 1;
  p.residual = reinterpret_cast<const bf16*>(epi->residual);
  p.ldr = epi->ldr;
  p.alpha = epi->alpha;
  p.act = epi->act;
  p.glu = epi->glu;
  p.out_fp32 = epi->out_fp32;
  p.out = out;
  p.ldo = epi->glu != VB_GLU_NONE ? cout / 2 : cout;

  const bool v2 = v2_eligible(epi, out, p.ldo, cout);
  int bn;
  float t_one = 1e30f;   // modelled time of the chosen 1-CTA plan
  if (v2) {
    TilePlan pl = plan_tiles(p.m_blocks, cout, p.num_k_blocks, true, p.M);
    if (pl.splits > 1) {
      const size_t need_ws = GEMM_COUNTER_BYTES + static_cast<size_t>(pl.splits) * p.M * cout * sizeof(float);
      if (workspace == nullptr || workspace_bytes < need_ws || !aligned32(workspace)) pl = plan_tiles(p.m_blocks, cout, p.num_k_blocks, false, p.M);
    }
    bn = pl.bn;
    t_one = pl.t;
    p.splits = pl.splits;
    if (pl.splits > 1) {
      p.counters = reinterpret_cast<int*>(workspace);
      p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + GEMM_COUNTER_BYTES);
      p.ws_ld = cout;
      p.ws_split_stride = static_cast<long long>(p.M) * cout;
      p.vec_ok = (cout % 8) == 0;
    }
  } else {
    bn = pick_block_n(static_cast<long long>(p.m_blocks) * BLOCK_M, cout, epi->glu);
  }
  // CTA pair (cta_group::2): two pixel tiles per 256-row MMA, each CTA loads its own tile + half of the weight tile
  int cl_bn = 0;
  if (v2 && !(g_gemm_resb & 8)) {
    float t_cl = 0.f;
    const int cbn = cl_tile(p.m_blocks, cout, p.num_k_blocks, v2_need(epi, 1), &t_cl);
    if (cbn && ((g_gemm_resb & 4) || t_cl < t_one)) {
      cl_bn = bn = cbn;
      p.splits = 1;
      p.counters = nullptr;
      p.ws = nullptr;
    }
  }
  p.n_blocks = static_cast<int>((cout + bn - 1) / bn);

  CUtensorMap ta, tb;
  uint64_t dA[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                    static_cast<uint64_t>(nb)};
  uint64_t sA[3] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(cin) * w * 2,
                    static_cast<uint64_t>(cin) * w * h * 2};
  // box extents are in traversed global elements: (tw-1)*stride+1 columns yield tw samples
  uint32_t bA[4] = {BLOCK_K, static_cast<uint32_t>((tw - 1) * stride + 1),
                    static_cast<uint32_t>((th - 1) * stride + 1), static_cast<uint32_t>(tn)};
  uint32_t eA[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
  if (int r = make_tmap(&ta, X, 4, dA, sA, bA, eA)) return r;
  uint64_t dB[2] = {static_cast<uint64_t>(kh) * kw * cin_pad, static_cast<uint64_t>(cout)};
  uint64_t sB[1] = {static_cast<uint64_t>(kh) * kw * cin_pad * 2};
  uint32_t bB[2] = {BLOCK_K, static_cast<uint32_t>(cl_bn ? bn / 2 : bn)};
  const uint32_t estr2[2] = {1, 1};
  if (int r = make_tmap(&tb, Wt, 2, dB, sB, bB, estr2)) return r;
  if (cl_bn) {
    p.dbg = g_gemm_dbg;
    return launch_gemm_v2_cl(bn, v2_need(epi, 1), ta, tb, p, stream);
  }
  if (v2) return launch_gemm_v2(bn, v2_need(epi, p.splits), ta, tb, p, stream);
  CUtensorMap tc = tb;
  p.c_box = c_box_for(bn, epi->glu, epi->out_fp32, p.ldo, out);
  if (p.c_box > 0) {
    uint64_t dC[4] = {static_cast<uint64_t>(p.ldo), static_cast<uint64_t>(wo), static_cast<uint64_t>(ho),
                      static_cast<uint64_t>(nb)};
    uint64_t sC[3] = {static_cast<uint64_t>(p.ldo) * 2, static_cast<uint64_t>(p.ldo) * wo * 2,
                      static_cast<uint64_t>(p.ldo) * wo * ho * 2};
    uint32_t bC[4] = {static_cast<uint32_t>(p.c_box), static_cast<uint32_t>(tw), static_cast<uint32_t>(th),
                      static_cast<uint32_t>(tn)};
    const uint32_t e4[4] = {1, 1, 1, 1};
    if (int r = make_tmap(&tc, out, 4, dC, sC, bC, e4,
                          p.c_box == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
      return r;
  }
  return launch_gemm(bn, ta, tb, tc, p, stream);
}
