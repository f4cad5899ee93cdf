// Implicit-GEMM convolution fed by TMA im2col descriptors (reference op K8).
//
//   D[M, N] = epilogue( A_im2col[M, K] * W[N, K]^T ),   M = B * GH * GW output positions, N = Cout,
//   K = KH * KW * C ordered (kh, kw, c) -- the layout of the weight matrix [Cout, KH, KW, Cin].
//
// The A operand is never materialised and no thread gathers it: the producer issues ONE
// cp.async.bulk.tensor.4d...im2col per k-block -- "128 consecutive output positions x 64 (or 32) channels of
// filter tap (kh, kw)" -- and the TMA unit walks the NHWC activation itself (bounding box = padding, traversal
// stride = convolution stride, out-of-image taps zero-filled), writing the 128-byte (64-byte) swizzled tile
// tcgen05.mma consumes.  SASS: UTMALDG.4D.IM2COL.
//
// The kernel is PERSISTENT: one CTA per SM loops over output tiles with the fp32 accumulator double-buffered
// in TMEM, so the epilogue of tile i (bias / ReLU / dReLU mask -> swizzled smem -> TMA store) overlaps the
// im2col loads and MMAs of tile i + 1.  Convolution GEMMs have tiny K (288 .. 576 here) and hundreds to
// thousands of M tiles: with one short-lived CTA per tile the fixed per-CTA latency (barrier setup, TMEM
// allocation, first-byte latency, epilogue, teardown) dominated the run time of the gather-based kernel.
//
//   forward:  S = X  [B, H, W, Cin],    grid = OH x OW, stride s, pad p
//   dgrad:    S = dZ [B, OH, OW, Cout], grid = H x W,   stride 1, pad KH - 1 - p, weights flipped / transposed
//             (conv_weight_flip_kernel), stride-1 convolutions only
#include "common.cuh"
#include "gemm.h"
#include "gemm_device.cuh"

namespace dk {

constexpr int kConvBlockM = 128;
constexpr int kConvThreads = 320;  // producer + MMA + 8 epilogue warps

template <int BN, int STAGES, int KBYTES>
struct ConvTmaSmem {
  static constexpr int kABytes = kConvBlockM * KBYTES;
  static constexpr int kBBytes = BN * KBYTES;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutStage = 8 * 4096;
  static constexpr int kMaskStage = 8 * 4096;
  static constexpr int kBiasStage = 8 * 1024;
  static constexpr int kTotal = STAGES * kStageBytes + kOutStage + kMaskStage + kBiasStage + 512 + 1024;
};

struct ConvGeom {
  int GH, GW;       // output position grid
  int mul, off;     // base = position * mul - off
  int KW, taps;     // filter width, KH * KW
  int c_chunks;     // C / channels per k-block
};

// K-major operand tile of [rows][KBYTES] bytes written by TMA with the KBYTES-byte swizzle
template <int KBYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * KBYTES) >> 4) << 32;          // stride between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(KBYTES == 128 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const void* tmap, int c, int w, int h, int n, int off_w,
                                                   int off_h, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(static_cast<uint16_t>(off_w)), "h"(static_cast<uint16_t>(off_h))
      : "memory");
}

template <int BN, int STAGES, int KBYTES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_m,
                const GemmEpilogue ep, const ConvGeom g, const int M, const int N, const int wres) {
  using S = ConvTmaSmem<BN, STAGES, KBYTES>;
  // wres: the layer has ONE tile along N and its whole weight matrix fits next to the pipeline, so every k-block's B
  // tile is loaded once per CTA (num_kb requests in total) instead of once per k-block of every tile -- a TMA request
  // costs ~190 cycles whatever its size, and with 9 taps x (A + B) requests per 128-pixel tile this kernel is
  // request-bound.  The stage pool then holds [resident weights | A-only stages].
  constexpr int kKElems = KBYTES / 2;      // bf16 channels per k-block
  constexpr int kMmaPerStage = kKElems / 16;
  constexpr uint32_t kTmemCols = 2 * BN;   // two accumulators
  constexpr uint32_t kIdesc = make_idesc(1u, kConvBlockM, BN);
  static_assert(BN == 64 || BN == 128, "conv_tma_kernel: BN in {64, 128}");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES * S::kStageBytes;
  uint8_t* mask_stage = out_stage + S::kOutStage;
  float* bias_stage = reinterpret_cast<float*>(mask_stage + S::kMaskStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(mask_stage + S::kMaskStage + S::kBiasStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* mask_bar = tmem_empty_bar + 2;        // [8 epilogue warps]
  uint64_t* w_bar = mask_bar + 8;                 // resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = g.taps * g.c_chunks;
  const int wres_bytes = wres ? num_kb * S::kBBytes : 0;
  const int a_stride = wres ? S::kABytes : S::kStageBytes;          // bytes between pipeline stages
  const int nst = wres ? min(STAGES, (STAGES * S::kStageBytes - wres_bytes) / S::kABytes) : STAGES;
  uint8_t* stages = smem + wres_bytes;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + kConvBlockM - 1) / kConvBlockM;
  const int num_tiles = n_tiles * m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    if (ep.tma_mask) tma_prefetch_desc(&tmap_m);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&mask_bar[i], 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  DK_PDL_WAIT();
  DK_PDL_TRIGGER();

  if (warp == 0) {
    // ------------------------------ TMA producer (im2col A, tiled B) ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (wres) {
        mbar_expect_tx(w_bar, wres_bytes);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem + kb * S::kBBytes, &tmap_b, kb * kKElems, 0, w_bar);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kConvBlockM, n0 = (tile % n_tiles) * BN;
        // first output position of the tile -> base pixel of its receptive field
        const int b = m0 / (g.GH * g.GW);
        const int rem = m0 - b * g.GH * g.GW;
        const int y = rem / g.GW, x = rem - y * g.GW;
        const int w0 = x * g.mul - g.off, h0 = y * g.mul - g.off;
        int tap = 0, cc = 0, kh = 0, kw = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stages + stage * a_stride;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], wres ? S::kABytes : S::kStageBytes);
          tma_load_im2col_4d(sa, &tmap_a, cc * kKElems, w0, h0, b, kw, kh, &full_bar[stage]);
          if (!wres) tma_load_2d(sb, &tmap_b, kb * kKElems, n0, &full_bar[stage]);
          if (++cc == g.c_chunks) {
            cc = 0;
            ++tap;
            if (++kw == g.KW) { kw = 0; ++kh; }
          }
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (wres) mbar_wait(w_bar, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(stages + stage * a_stride);
          const uint64_t adesc = make_kmajor_desc<KBYTES>(sa);
          const uint64_t bdesc = make_kmajor_desc<KBYTES>(wres ? smem_u32(smem + kb * S::kBBytes) : sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < kMmaPerStage; ++k) umma_f16(tmem_acc, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (8 warps, 64-column slices) ------------------------------
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int group = e >> 2;
    const uint32_t out_region = smem_u32(out_stage) + e * 4096;
    const uint32_t mask_region = smem_u32(mask_stage) + e * 4096;
    float* my_bias = bias_stage + e * 256;
    uint64_t* my_mask_bar = mask_bar + e;
    uint32_t mask_phase = 0;
    constexpr int kSlices = BN / 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles) * kConvBlockM, n0 = (tile % n_tiles) * BN;
      const int acc = it & 1;
      const int valid_slices = min(kSlices, (N - n0 + 63) / 64);
      int sl = ((it * kSlices) & 1) == group ? 0 : 1;
      if (ep.tma_mask && lane == 0 && sl < valid_slices) {
        mbar_expect_tx(my_mask_bar, 4096);
        tma_load_2d_addr(mask_region, &tmap_m, n0 + sl * 64, m0 + quarter * 32, my_mask_bar);
      }
      if (ep.bias != nullptr) {
#pragma unroll
        for (int j = lane; j < BN; j += 32) my_bias[j] = (n0 + j < N) ? __ldg(ep.bias + n0 + j) : 0.f;
        __syncwarp();
      }
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      bool arrived = false;
#pragma unroll 1
      for (; sl < valid_slices; sl += 2) {
        const int nc = n0 + sl * 64;
        float v[64];
        {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + sl * 64;
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(taddr, r0);
          tmem_ld_32x32b_x32(taddr + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[32 + j] = __uint_as_float(r1[j]);
          }
        }
        if (sl + 2 >= valid_slices) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          arrived = true;
        }
        if (ep.alpha != 1.f) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] *= ep.alpha;
        }
        if (ep.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(my_bias + sl * 64 + j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (ep.tma_mask) {
          mbar_wait(my_mask_bar, mask_phase);
          mask_phase ^= 1;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const uint4 q = ld_shared_v4(mask_region + sw128_off(lane, t));
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (!(__bfloat162float(h[u]) > 0.f)) v[t * 8 + u] = 0.f;
          }
          __syncwarp();
          if (lane == 0 && sl + 2 < valid_slices) {
            mbar_expect_tx(my_mask_bar, 4096);
            tma_load_2d_addr(mask_region, &tmap_m, nc + 128, m0 + quarter * 32, my_mask_bar);
          }
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t)
          st_shared_v4(out_region + sw128_off(lane, t), pack_bf16x2(v[8 * t], v[8 * t + 1]),
                       pack_bf16x2(v[8 * t + 2], v[8 * t + 3]), pack_bf16x2(v[8 * t + 4], v[8 * t + 5]),
                       pack_bf16x2(v[8 * t + 6], v[8 * t + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_addr(&tmap_d, out_region, nc, m0 + quarter * 32);
          tma_store_commit();
        }
      }
      if (!arrived) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------
// Weight gradient of a convolution with the im2col operand produced by TMA:
//     dW[co, (kh, kw, c)] += sum_m dZ[m, co] * X[b(m), y(m) s - p + kh, x(m) s - p + kw, c]
// i.e. D[M = Cout, N = taps * C] = A^T B with GEMM-K = the B * OH * OW output positions.  Per 64 positions the
// producer issues ONE tiled load of dZ (MN-major A) and one im2col load per (tap, 64-channel chunk) "unit"
// (MN-major B: [64 positions x chan channels], exactly what an im2col load writes); the MMA warp issues
// 4 tcgen05.mma per unit into that unit's TMEM column range, so the whole [Cout x units * chan] gradient block
// stays in TMEM while the CTA streams its share of the positions (split-K across the persistent grid), and
// is added into the fp32 gradient with TMA reduce-add at the end.  The bias gradient (column sums of dZ)
// comes from one more N = 16 MMA against a tile of ones.  No column matrix, no im2col / colsum kernels.
// ---------------------------------------------------------------------------------------------
struct WgradGeom {
  int GH, GW, mul, off;
  int KW, c_chunks;        // filter width, C / chan
  int unit0, units;        // (tap, channel-chunk) units handled by this launch: unit = tap * c_chunks + chunk
  int total_pb;            // 64-position blocks
  int Cout, Ktot;          // dW is [Cout, Ktot] fp32
};

template <int CHAN>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_d, const WgradGeom g, float* __restrict__ bias_grad,
                      const int stages, const int stage_bytes) {
  constexpr int kABytes = 128 * 128;               // dZ^T tile: two [64 positions x 64 co] boxes
  constexpr int kUnitBytes = 64 * CHAN * 2;        // one im2col box
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kIdesc = make_idesc(1u, 128, CHAN) | (1u << 15) | (1u << 16);   // A and B MN-major
  constexpr uint32_t kIdescBias = make_idesc(1u, 128, 16) | (1u << 15);
  constexpr uint32_t kStepA = 2048 >> 4;           // 16 positions x 128 B
  constexpr uint32_t kStepB = (16 * CHAN * 2) >> 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ones = smem + stages * stage_bytes;
  uint8_t* out_stage = ones + 2048;                // 4 epilogue warps x 4 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + 4 * 4096);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int per = (g.total_pb + gridDim.x - 1) / gridDim.x;
  const int pb_begin = blockIdx.x * per;
  const int pb_end = min(g.total_pb, pb_begin + per);
  const int num_pb = pb_end - pb_begin;
  const bool do_bias = bias_grad != nullptr;
  if (num_pb <= 0) return;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    st_shared_v4(smem_u32(ones) + t * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  DK_PDL_WAIT();
  DK_PDL_TRIGGER();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pb = pb_begin; pb < pb_end; ++pb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], kABytes + g.units * kUnitBytes);
        const int m0 = pb * 64;
        tma_load_2d(sa, &tmap_a, 0, m0, &full_bar[stage]);
        tma_load_2d(sa + 8192, &tmap_a, 64, m0, &full_bar[stage]);
        const int b = m0 / (g.GH * g.GW);
        const int rem = m0 - b * g.GH * g.GW;
        const int y = rem / g.GW, x = rem - y * g.GW;
        const int w0 = x * g.mul - g.off, h0 = y * g.mul - g.off;
        int unit = g.unit0;
        int tap = unit / g.c_chunks, cc = unit - tap * g.c_chunks;
        int kh = tap / g.KW, kw = tap - kh * g.KW;
        for (int u = 0; u < g.units; ++u) {
          tma_load_im2col_4d(sa + kABytes + u * kUnitBytes, &tmap_b, cc * CHAN, w0, h0, b, kw, kh, &full_bar[stage]);
          if (++cc == g.c_chunks) {
            cc = 0;
            if (++kw == g.KW) { kw = 0; ++kh; }
          }
        }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < num_pb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t adesc = make_smem_desc_sw128_mn(sa);
        const uint64_t odesc = make_smem_desc_sw128(smem_u32(ones));
        for (int u = 0; u < g.units; ++u) {
          // MN-major B tile of [64 positions][CHAN channels]: 8-position groups CHAN * 16 bytes apart
          uint64_t bdesc = 0;
          const uint32_t sb = sa + kABytes + u * kUnitBytes;
          bdesc |= static_cast<uint64_t>((sb & 0x3FFFF) >> 4);
          bdesc |= static_cast<uint64_t>(1) << 16;
          bdesc |= static_cast<uint64_t>((8 * CHAN * 2) >> 4) << 32;
          bdesc |= static_cast<uint64_t>(1) << 46;
          bdesc |= static_cast<uint64_t>(CHAN == 64 ? 2 : 4) << 61;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + u * CHAN, adesc + kStepA * k, bdesc + kStepB * k, kIdesc, (i | k) != 0);
        }
        if (do_bias) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 496, adesc + kStepA * k, odesc + 2 * k, kIdescBias, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (i == num_pb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ---- epilogue: TMEM -> swizzled smem -> TMA reduce-add into the fp32 gradient (rows = co) ----
    const int quarter = warp & 3;
    const int co = quarter * 32 + lane;
    const uint32_t region = smem_u32(out_stage) + (warp - 2) * 4096;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    if (quarter * 32 < g.Cout) {
      const int ncols = g.units * CHAN;
      const int n_base = g.unit0 * CHAN;   // units are contiguous in the (tap, c) column order
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + c, r);
        tmem_ld_wait();
        if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's reduce has read the staging slot
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) st_shared_v4(region + sw128_off(lane, t), r[4 * t], r[4 * t + 1], r[4 * t + 2], r[4 * t + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmap_d, region, n_base + c, quarter * 32);
          tma_store_commit();
        }
      }
      if (do_bias) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(trow + 496, r);
        tmem_ld_wait();
        if (co < g.Cout) atomicAdd(bias_grad + co, __uint_as_float(r[0]));
      }
      if (lane == 0) tma_store_wait_read<0>();
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeIm2col get_im2col_fn() {
  static PFN_encodeIm2col fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeIm2col>(p);
  }
  return fn;
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN, int STAGES, int KBYTES>
static int launch_conv_tma(const CUtensorMap* ta, const CUtensorMap* tb, const CUtensorMap* td, const CUtensorMap* tm,
                           GemmEpilogue ep, const ConvGeom& g, int M, int N, cudaStream_t stream) {
  using S = ConvTmaSmem<BN, STAGES, KBYTES>;
  auto kern = conv_tma_kernel<BN, STAGES, KBYTES>;
  static bool configured[64] = {};
  static int sm_count[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    DK_HOST_CHECK(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    configured[dev & 63] = true;
  }
  ep.tma_store = 1;
  ep.tma_mask = tm != nullptr ? 1 : 0;
  const int tiles = ((M + kConvBlockM - 1) / kConvBlockM) * ((N + BN - 1) / BN);
  const int grid = tiles < sm_count[dev & 63] ? tiles : sm_count[dev & 63];
  // resident weights: one tile along N, the weight matrix + at least two A stages fit the stage pool (DK_CONV_WRES=0: off)
  static int wres_env = -1;
  if (wres_env < 0) {
    const char* we = getenv("DK_CONV_WRES");
    wres_env = (we != nullptr && we[0] == '0') ? 0 : 1;
  }
  const int num_kb = g.taps * g.c_chunks;
  const int wres = (wres_env && N <= BN && num_kb * S::kBBytes + 2 * S::kABytes <= STAGES * S::kStageBytes) ? 1 : 0;
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, kConvThreads, S::kTotal, stream, *ta, *tb, *td, ep.tma_mask ? *tm : *ta, ep, g, M, N, wres));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace dk

extern "C" {

// 1 if this (C, div) geometry can run on the TMA-im2col kernel
int dk_conv_tma_supported(int C, int div, int N, int ldd, int d_fp32) {
  if (div != 1 || d_fp32 || (ldd % 8) != 0 || N < 8) return 0;
  return (C % 64 == 0 || C == 32) ? 1 : 0;
}

// im2col tensor map over the NHWC activation S [B, SH, SW, C] (bf16) for a GH x GW grid of output positions:
// base pixel of position (y, x) = (y * mul - off, x * mul - off); 128 positions x `chan` channels per load.
static int encode_im2col(void* out_tmap, const void* src, int B, int SH, int SW, int C, int GH, int GW, int mul, int off,
                         int chan, int pixels);

int dk_conv_tma_encode_a(void* out_tmap, const void* src, int B, int SH, int SW, int C, int GH, int GW, int mul, int off,
                         int chan) {
  return encode_im2col(out_tmap, src, B, SH, SW, C, GH, GW, mul, off, chan, dk::kConvBlockM);
}

static int encode_im2col(void* out_tmap, const void* src, int B, int SH, int SW, int C, int GH, int GW, int mul, int off,
                         int chan, int pixels) {
  auto fn = dk::get_im2col_fn();
  if (fn == nullptr) return -1;
  if ((reinterpret_cast<uintptr_t>(src) & 15) != 0 || (C * 2) % 16 != 0) return -2;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(SW), static_cast<cuuint64_t>(SH),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(SW) * C * 2,
                           static_cast<cuuint64_t>(SH) * SW * C * 2};
  // the box of BASE pixels: [-off, (G - 1) * mul - off] in each spatial dimension
  int lower[2] = {-off, -off};
  int upper[2] = {(GW - 1) * mul - off - (SW - 1), (GH - 1) * mul - off - (SH - 1)};
  cuuint32_t estride[4] = {1, static_cast<cuuint32_t>(mul), static_cast<cuuint32_t>(mul), 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim,
                  gstride, lower, upper, static_cast<cuuint32_t>(chan), static_cast<cuuint32_t>(pixels), estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, chan == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

// weight matrix [N, K] bf16 K-major, box [bn rows x chan elements], swizzle matching the A tile
int dk_conv_tma_encode_b(void* out_tmap, const void* W, long ldw, int N, int K, int bn, int chan) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return -1;
  auto fn = reinterpret_cast<dk::PFN_encodeTiled2>(p);
  if ((ldw * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0) return -2;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(N)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ldw) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(chan), static_cast<cuuint32_t>(bn)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W), gdim,
                  gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  chan == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

int dk_conv_tma_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, const void* tmap_m,
                       const DkGemmEpilogue* ep, int C, int GH, int GW, int KH, int KW, int mul, int off, int M, int N,
                       void* stream) {
  if (tmap_d == nullptr) return -3;
  if (ep->mask != nullptr && tmap_m == nullptr) return -5;
  const int chan = C % 64 == 0 ? 64 : 32;
  dk::ConvGeom g;
  g.GH = GH; g.GW = GW; g.mul = mul; g.off = off; g.KW = KW; g.taps = KH * KW; g.c_chunks = C / chan;
  const CUtensorMap* ta = reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap* tb = reinterpret_cast<const CUtensorMap*>(tmap_b);
  const CUtensorMap* td = reinterpret_cast<const CUtensorMap*>(tmap_d);
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(tmap_m);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bn = N <= 64 ? 64 : 128;
  if (chan == 64) {
    if (bn == 64) return dk::launch_conv_tma<64, 6, 128>(ta, tb, td, tm, *ep, g, M, N, st);
    return dk::launch_conv_tma<128, 4, 128>(ta, tb, td, tm, *ep, g, M, N, st);
  }
  if (bn == 64) return dk::launch_conv_tma<64, 8, 64>(ta, tb, td, tm, *ep, g, M, N, st);
  return dk::launch_conv_tma<128, 6, 64>(ta, tb, td, tm, *ep, g, M, N, st);
}

int dk_conv_tma_bn(int N) { return N <= 64 ? 64 : 128; }

// units (tap, channel chunk) one wgrad launch can hold: TMEM columns (496 + 16 for the bias gradient) and two
// pipeline stages of shared memory
int dk_conv_wgrad_tma_units(int C) {
  const int chan = C % 64 == 0 ? 64 : 32;
  const int by_tmem = 496 / chan;
  const int by_smem = (96 * 1024 - 16384) / (64 * chan * 2);
  return by_tmem < by_smem ? by_tmem : by_smem;
}

int dk_conv_wgrad_tma_supported(int C, int Cout, long lddz, long lddw) {
  return ((C % 64 == 0 || C == 32) && Cout <= 128 && Cout % 8 == 0 && lddz % 8 == 0 && lddw % 4 == 0) ? 1 : 0;
}

int dk_conv_wgrad_tma_encode(void* tmap_a, void* tmap_b, void* tmap_d, const void* src, int B, int SH, int SW, int C, int GH,
                             int GW, int KH, int KW, int stride, int pad, const void* dz, long lddz, float* dw, long lddw,
                             int Cout) {
  const int chan = C % 64 == 0 ? 64 : 32;
  int r = dk_tmap_encode_2d(tmap_a, dz, DK_BF16, static_cast<long>(B) * GH * GW, Cout, lddz, 64);
  if (r != 0) return r;
  r = encode_im2col(tmap_b, src, B, SH, SW, C, GH, GW, stride, pad, chan, 64);
  if (r != 0) return r;
  return dk_gemm_encode_output(tmap_d, dw, lddw, Cout, KH * KW * C, 1);
}

int dk_conv_wgrad_tma_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, int B, int C, int GH, int GW, int KH,
                             int KW, int stride, int pad, int Cout, int unit0, int units, float* bias_grad, void* stream) {
  const int chan = C % 64 == 0 ? 64 : 32;
  dk::WgradGeom g;
  g.GH = GH; g.GW = GW; g.mul = stride; g.off = pad; g.KW = KW; g.c_chunks = C / chan; g.unit0 = unit0; g.units = units;
  const long rows = static_cast<long>(B) * GH * GW;
  g.total_pb = static_cast<int>((rows + 63) / 64);
  g.Cout = Cout; g.Ktot = KH * KW * C;
  if (units < 1 || units > dk_conv_wgrad_tma_units(C) || unit0 + units > KH * KW * g.c_chunks) return -3;
  const int unit_bytes = 64 * chan * 2;
  const int stage_bytes = 16384 + units * unit_bytes;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return -4;
  const int smem = stages * stage_bytes + 2048 + 4 * 4096 + 256 + 1024;
  static int sm_count[64] = {};
  static bool configured[2][64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (sm_count[dev & 63] == 0) DK_HOST_CHECK(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const int which = chan == 64 ? 1 : 0;
  if (!configured[which][dev & 63]) {
    if (chan == 64)
      DK_HOST_CHECK(cudaFuncSetAttribute(dk::conv_wgrad_tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    else
      DK_HOST_CHECK(cudaFuncSetAttribute(dk::conv_wgrad_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[which][dev & 63] = true;
  }
  int grid = sm_count[dev & 63];
  if (grid > g.total_pb) grid = g.total_pb;
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  const CUtensorMap& td = *reinterpret_cast<const CUtensorMap*>(tmap_d);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (chan == 64)
    DK_HOST_CHECK(DK_LAUNCH(dk::conv_wgrad_tma_kernel<64>, grid, 192, smem, st, ta, tb, td, g, bias_grad, stages, stage_bytes));
  else
    DK_HOST_CHECK(DK_LAUNCH(dk::conv_wgrad_tma_kernel<32>, grid, 192, smem, st, ta, tb, td, g, bias_grad, stages, stage_bytes));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

// one-shot variant (tests): dW (zeroed by the caller) += wgrad, bias_grad (zeroed, may be NULL) += colsum(dZ)
int dk_conv_wgrad_tma(const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                      const void* dz, long lddz, float* dw, long lddw, int Cout, float* bias_grad, void* stream) {
  alignas(64) CUtensorMap ta, tb, td;
  int r = dk_conv_wgrad_tma_encode(&ta, &tb, &td, src, B, SH, SW, C, GH, GW, KH, KW, stride, pad, dz, lddz, dw, lddw, Cout);
  if (r != 0) return r;
  const int chan = C % 64 == 0 ? 64 : 32;
  const int total_units = KH * KW * (C / chan), per = dk_conv_wgrad_tma_units(C);
  for (int u0 = 0; u0 < total_units; u0 += per) {
    const int n = total_units - u0 < per ? total_units - u0 : per;
    r = dk_conv_wgrad_tma_launch(&ta, &tb, &td, B, C, GH, GW, KH, KW, stride, pad, Cout, u0, n, u0 == 0 ? bias_grad : nullptr, stream);
    if (r != 0) return r;
  }
  return 0;
}



// one-shot variant (tests): encodes every tensor map first
int dk_conv_tma(const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off,
                const void* Wmat, long ldw, const DkGemmEpilogue* ep, int M, int N, void* stream) {
  alignas(64) CUtensorMap ta, tb, td, tm;
  const int chan = C % 64 == 0 ? 64 : 32;
  int r = dk_conv_tma_encode_a(&ta, src, B, SH, SW, C, GH, GW, mul, off, chan);
  if (r != 0) return r;
  r = dk_conv_tma_encode_b(&tb, Wmat, ldw, N, KH * KW * C, dk_conv_tma_bn(N), chan);
  if (r != 0) return r;
  if (ep->d == nullptr || dk_gemm_encode_output(&td, ep->d, ep->ldd, M, N, 0) != 0) return -4;
  const bool has_m = ep->mask != nullptr && dk_gemm_encode_output(&tm, ep->mask, ep->ld_mask, M, N, 0) == 0;
  return dk_conv_tma_launch(&ta, &tb, &td, has_m ? &tm : nullptr, ep, C, GH, GW, KH, KW, mul, off, M, N, stream);
}

}  // extern "C"
