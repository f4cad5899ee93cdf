// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )      (both operands K-major, "TN" GEMM)
//
// One CTA computes one 128 x BN output tile:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B, STAGES-deep ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulator in TMEM)
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> fused bias / ReLU / dReLU-mask ->
//                 row-major store and optional transposed bf16 copy)
// smem full/empty mbarriers form the TMA<->MMA pipeline, one more mbarrier hands the finished
// accumulator to the epilogue warps.  Out-of-range rows/columns/K are zero-filled by TMA, so any
// M, N, K are accepted as long as the global row strides are multiples of 16 bytes.
//
// Dense-layer mapping (reference ops K6/K7):
//   forward  Y  = act(X W^T + b)   : A = X [B,in],     B = W  [out,in]
//   dgrad    dX = (dY W) * relu'   : A = dY [B,out],   B = W^T[in,out]
//   wgrad    dW = dY^T X           : A = dY^T[out,B],  B = X^T[in,B]
// The transposed copies are produced by the epilogue of the kernel that wrote the tensor.
#include "common.cuh"
#include "gemm.h"
#include "gemm_device.cuh"

namespace dk {

constexpr int kBlockM = 128;
constexpr int kGemmThreads = 192;

template <int BN, int STAGES, bool TF32>
struct GemmSmem {
  static constexpr int kBlockKBytes = 128;                    // one swizzle row
  static constexpr int kABytes = kBlockM * kBlockKBytes;      // 16 KB
  static constexpr int kBBytes = BN * kBlockKBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = STAGES * kStageBytes + kBarrierBytes + 1024;  // + align slack
};

// Epilogue of one 128 x BN accumulator tile (called by the 4 epilogue warps); see the kernel comment.
template <int BN>
__device__ __forceinline__ void gemm_epilogue(const CUtensorMap& tmap_d, const CUtensorMap& tmap_m,
                                              const GemmEpilogue& ep, const int M, const int N, const int m0,
                                              const int n0, const int warp, const int lane, const uint32_t tmem_base,
                                              uint8_t* smem, uint64_t* tmem_full_bar, uint64_t* mask_bar) {
  const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
  const int m = m0 + quarter * 32 + lane;        // output row owned by this thread
  constexpr int kChunk = BN < 32 ? 16 : 32;
  const bool tma_out = ep.tma_store != 0;
  const int esize = ep.d_fp32 ? 4 : 2;
  // staging regions inside the (idle after the main loop) pipeline buffers
  const uint32_t out_region = smem_u32(smem) + quarter * (32 * BN * esize);
  const uint32_t mask_region = smem_u32(smem) + 4 * (32 * BN * esize) + quarter * (32 * BN * 2);
  mbar_wait(tmem_full_bar, 0);
  tcgen05_fence_after();
  if (ep.tma_mask) {
    // dReLU / dropout mask tile (bf16, same shape as the output) via TMA into swizzled smem
    if (lane == 0) {
      const int ngroups = (BN + 63) / 64;
      mbar_expect_tx(&mask_bar[quarter], ngroups * 4096);
      for (int g = 0; g < ngroups; ++g)
        tma_load_2d_addr(mask_region + g * 4096, &tmap_m, n0 + g * 64, m0 + quarter * 32, &mask_bar[quarter]);
    }
    mbar_wait(&mask_bar[quarter], 0);
  }
  const bool row_ok = m < M;
  const float bias_m = (ep.bias != nullptr && ep.bias_along_m && row_ok) ? ep.bias[m] : 0.f;
  uint32_t drop_salt = 0, drop_thr = 0;
  float keep_scale = 1.f;
  if (ep.drop_p > 0.f) {
    drop_salt = ep.drop_seed + (ep.step != nullptr ? static_cast<uint32_t>(*ep.step) : 0u) * 0x85EBCA77u;
    drop_thr = static_cast<uint32_t>(ep.drop_p * 256.f + 0.5f);          // 8-bit resolution
    keep_scale = 256.f / (256.f - static_cast<float>(drop_thr));
  }
#pragma unroll 1
  for (int c = 0; c < BN; c += kChunk) {
    const int nc = n0 + c;
    if (nc >= N) break;  // warp-uniform
    float v[kChunk];
    {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c;
      if constexpr (kChunk == 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      }
    }
    const bool full_chunk = nc + kChunk <= N;
    // ---- fused elementwise epilogue ----
    if (ep.alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < kChunk; ++j) v[j] *= ep.alpha;
    }
    if (ep.bias != nullptr) {
      if (ep.bias_along_m) {
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] += bias_m;
      } else if (full_chunk && ((reinterpret_cast<uintptr_t>(ep.bias + nc) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < kChunk; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + nc + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
          if (nc + j < N) v[j] += __ldg(ep.bias + nc + j);
      }
    }
    if (ep.act == 1) {
#pragma unroll
      for (int j = 0; j < kChunk; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (ep.drop_p > 0.f) {
      // inverted dropout (reference op K10): one counter-based hash per 4 elements, 8 bits each
#pragma unroll
      for (int j = 0; j < kChunk; j += 4) {
        uint32_t h = (static_cast<uint32_t>(m) * static_cast<uint32_t>(N) + static_cast<uint32_t>(nc + j)) * 0x9E3779B1u ^ drop_salt;
        h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          v[j + t] = (((h >> (8 * t)) & 0xFFu) < drop_thr) ? 0.f : v[j + t] * keep_scale;
      }
    }
    if (ep.tma_mask) {
      // chunk c covers 64 bytes of the 128-byte mask row: 16-byte pieces (c % 64) / 8 + 0..3
      const int g = c >> 6, j0 = (c & 63) >> 3;
#pragma unroll
      for (int t = 0; t < kChunk / 8; ++t) {
        const uint4 q = ld_shared_v4(mask_region + g * 4096 + sw128_off(lane, j0 + t));
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (!(__bfloat162float(h[u]) > 0.f)) v[t * 8 + u] = 0.f;
      }
    } else if (ep.mask != nullptr && row_ok) {
      const __nv_bfloat16* mrow = ep.mask + static_cast<size_t>(m) * ep.ld_mask + nc;
      if (full_chunk && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < kChunk; j += 8) {
          const uint4 q = *reinterpret_cast<const uint4*>(mrow + j);
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (!(__bfloat162float(h[t]) > 0.f)) v[j + t] = 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
          if (nc + j < N && !(__bfloat162float(mrow[j]) > 0.f)) v[j] = 0.f;
      }
    }
    if (tma_out) {
      // ---- stage in swizzled smem, one TMA store per completed 128-byte column group ----
      if (ep.d_fp32) {
        const int g = c / 32;                      // kChunk == 32 here (host guarantees BN >= 32)
#pragma unroll
        for (int t = 0; t < kChunk / 4; ++t)
          st_shared_v4(out_region + g * 4096 + sw128_off(lane, t), __float_as_uint(v[4 * t]),
                       __float_as_uint(v[4 * t + 1]), __float_as_uint(v[4 * t + 2]), __float_as_uint(v[4 * t + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (gridDim.z > 1 || ep.accumulate)
            tma_reduce_add_2d(&tmap_d, out_region + g * 4096, nc, m0 + quarter * 32);
          else
            tma_store_2d_addr(&tmap_d, out_region + g * 4096, nc, m0 + quarter * 32);
        }
      } else {
        const int g = c >> 6, j0 = (c & 63) >> 3;
#pragma unroll
        for (int t = 0; t < kChunk / 8; ++t)
          st_shared_v4(out_region + g * 4096 + sw128_off(lane, j0 + t), pack_bf16x2(v[8 * t], v[8 * t + 1]),
                       pack_bf16x2(v[8 * t + 2], v[8 * t + 3]), pack_bf16x2(v[8 * t + 4], v[8 * t + 5]),
                       pack_bf16x2(v[8 * t + 6], v[8 * t + 7]));
        const bool group_done = ((c & 63) + kChunk >= 64) || (nc + kChunk >= N) || (c + kChunk >= BN);
        if (group_done) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d_addr(&tmap_d, out_region + g * 4096, n0 + g * 64, m0 + quarter * 32);
        }
      }
      continue;
    }
    // ---- direct row-major store (small / unaligned outputs) ----
    if (ep.d != nullptr && row_ok) {
      if (ep.d_fp32) {
        float* drow = reinterpret_cast<float*>(ep.d) + static_cast<size_t>(m) * ep.ldd + nc;
        if (gridDim.z > 1) {
#pragma unroll
          for (int j = 0; j < kChunk; ++j)
            if (nc + j < N) atomicAdd(drow + j, v[j]);
        } else if (full_chunk && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < kChunk; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (ep.accumulate) {
              const float4 p = *reinterpret_cast<const float4*>(drow + j);
              o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *reinterpret_cast<float4*>(drow + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kChunk; ++j)
            if (nc + j < N) drow[j] = ep.accumulate ? drow[j] + v[j] : v[j];
        }
      } else {
        __nv_bfloat16* drow =
            reinterpret_cast<__nv_bfloat16*>(ep.d) + static_cast<size_t>(m) * ep.ldd + nc;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < kChunk; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(v[j], v[j + 1]);
            o.y = pack_bf16x2(v[j + 2], v[j + 3]);
            o.z = pack_bf16x2(v[j + 4], v[j + 5]);
            o.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(drow + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < kChunk; ++j)
            if (nc + j < N) drow[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
    // ---- transposed bf16 copy: DT[n][m]; lanes are consecutive m -> coalesced ----
    if (ep.dt != nullptr && row_ok) {
#pragma unroll
      for (int j = 0; j < kChunk; ++j)
        if (nc + j < N)
          ep.dt[static_cast<size_t>(nc + j) * ep.lddt + m] = __float2bfloat16_rn(v[j]);
    }
  }
  if (tma_out && lane == 0) {
    tma_store_commit();
    tma_store_wait_read<0>();  // smem must stay valid until the TMA engine has read it
  }
}

// Epilogue of the bn = 16 forward GEMM with the classifier head fused in (see DkGemmEpilogue::head_*): the
// activation slice [128 x 16] of this CTA is written out, its contribution to the logits is red.add'ed to a
// [128 x 16] fp32 scratch, the (<= 32, co-resident) CTAs of the grid meet on a counter, and every CTA then computes
// softmax / dZ for its rows and its own 16 columns of dH = alpha (dZ W3) * (H > 0).  Replaces the stand-alone head
// kernel (one launch + one dependent-kernel latency per step in the small-batch regime).
__device__ __forceinline__ void gemm_epilogue_head(const GemmEpilogue& ep, const int M, const int m_end, const int N,
                                                   const int m0, const int n0, const int warp, const int lane,
                                                   const uint32_t tmem_base, uint64_t* tmem_full_bar) {
  // This code runs once per launch: loops over the classes are kept rolled (instruction-cache misses cost more
  // than the loop overhead) and the per-row class vector lives in shared memory, one conflict-free column per thread.
  __shared__ __align__(16) float s_w3[16 * 16];  // W3[c][n0 + j] of this CTA's column slice (zero padded)
  __shared__ float s_g[16][128];                 // [class][epilogue thread]: logits, then dZ
  __shared__ float s_b3[16];
  const int quarter = warp & 3;
  const int tid = (warp - 2) * 32 + lane;        // 0..127 over the four epilogue warps
  const int m = m0 + quarter * 32 + lane;
  const bool row_ok = m < m_end;      // (M = the whole mini-batch: loss / gradient scale; m_end = end of this CTA's M tile)
  const int C = ep.head_c;
  for (int idx = tid; idx < 256; idx += 128) {
    const int c = idx >> 4, jj = idx & 15;
    s_w3[idx] = (c < C && n0 + jj < N) ? __bfloat162float(ep.head_w[static_cast<size_t>(c) * ep.head_ldw + n0 + jj]) : 0.f;
  }
  if (tid < 16) s_b3[tid] = (ep.head_bias != nullptr && tid < C) ? __ldg(ep.head_bias + tid) : 0.f;
  unsigned long long* const trw = (blockIdx.x == 0 && warp == 4 && lane == 0) ? ep.trace : nullptr;
  const int label = row_ok ? __ldg(ep.head_labels + m) : 0;
  float bias_v[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr && n0 + 4 * q4 + 4 <= N) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + 4 * q4));
    bias_v[4 * q4] = b4.x; bias_v[4 * q4 + 1] = b4.y; bias_v[4 * q4 + 2] = b4.z; bias_v[4 * q4 + 3] = b4.w;
  }
  // head_sync[0]: arrivals, never reset (launch L of the grid completes it to (L + 1) * target); head_sync[1] = L, the
  // number of finished launches (written by one thread at the end of each).  The logits scratch is double-buffered
  // by launch parity: this launch accumulates into acc[L & 1] while CTA 0 clears acc[(L + 1) & 1] -- whose last
  // readers belonged to launch L - 1 -- so nothing has to be cleaned up on the way out.
  const unsigned launch = *reinterpret_cast<volatile const unsigned*>(ep.head_sync + 1);
  float* const acc = ep.head_acc + (launch & 1u) * (128 * 16);
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    float4* nxt = reinterpret_cast<float4*>(ep.head_acc + ((launch + 1u) & 1u) * (128 * 16));
    for (int i4 = tid; i4 < 128 * 16 / 4; i4 += 128) nxt[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint32_t drop_salt = 0, drop_thr = 0;
  float keep_scale = 1.f;
  if (ep.drop_p > 0.f) {
    drop_salt = ep.drop_seed + (ep.step != nullptr ? static_cast<uint32_t>(*ep.step) : 0u) * 0x85EBCA77u;
    drop_thr = static_cast<uint32_t>(ep.drop_p * 256.f + 0.5f);
    keep_scale = 256.f / (256.f - static_cast<float>(drop_thr));
  }
  int slot = ep.head_step != nullptr ? (*ep.head_step - 1) : 0;
  named_bar_sync(1, 128);
  mbar_wait(tmem_full_bar, 0);
  tcgen05_fence_after();
  float v[16];
  {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16), r);
    tmem_ld_wait();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(r[jj]) + bias_v[jj];
  }
  // ---- the producing layer's own epilogue: (bias,) ReLU, inverted dropout (same hash as gemm_epilogue) ----
  if (ep.act == 1) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) v[jj] = fmaxf(v[jj], 0.f);
  }
  if (ep.drop_p > 0.f) {
#pragma unroll
    for (int jj = 0; jj < 16; jj += 4) {
      uint32_t h = (static_cast<uint32_t>(m) * static_cast<uint32_t>(N) + static_cast<uint32_t>(n0 + jj)) * 0x9E3779B1u ^ drop_salt;
      h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
#pragma unroll
      for (int t = 0; t < 4; ++t) v[jj + t] = (((h >> (8 * t)) & 0xFFu) < drop_thr) ? 0.f : v[jj + t] * keep_scale;
    }
  }
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) v[jj] = (n0 + jj < N) ? __bfloat162float(__float2bfloat16_rn(v[jj])) : 0.f;  // what the next layer reads
  if (row_ok) {
    if (ep.d != nullptr) {
      __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(ep.d) + static_cast<size_t>(m) * ep.ldd + n0;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (n0 + 8 * h8 + 8 <= N)
          *reinterpret_cast<uint4*>(drow + 8 * h8) =
              make_uint4(pack_bf16x2(v[8 * h8], v[8 * h8 + 1]), pack_bf16x2(v[8 * h8 + 2], v[8 * h8 + 3]),
                         pack_bf16x2(v[8 * h8 + 4], v[8 * h8 + 5]), pack_bf16x2(v[8 * h8 + 6], v[8 * h8 + 7]));
      }
    }
    // ---- this slice's share of the logits: four classes per 16-byte vector reduction ----
#pragma unroll 1
    for (int c4 = 0; c4 < C; c4 += 4) {
      float s4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4* wr = reinterpret_cast<const float4*>(s_w3 + (c4 + u) * 16);   // rows >= C of s_w3 are zero
        float a = 0.f;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 w4 = wr[q4];
          a = fmaf(v[4 * q4], w4.x, fmaf(v[4 * q4 + 1], w4.y, fmaf(v[4 * q4 + 2], w4.z, fmaf(v[4 * q4 + 3], w4.w, a))));
        }
        s4[u] = a;
      }
      asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(acc + m * 16 + c4), "f"(s4[0]),
                   "f"(s4[1]), "f"(s4[2]), "f"(s4[3])
                   : "memory");
    }
  }
  // ---- rendezvous of the grid's epilogue warps (warps without live rows only arrive) ----
  trace_stamp(trw, 8);
  __syncwarp();
  const unsigned target = (launch + 1u) * (gridDim.x * gridDim.y * 4u);
  if (lane == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ep.head_sync) : "memory");   // cumulative over the warp's reds
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ep.head_sync) : "memory");
    } while (static_cast<int>(seen - target) < 0);
  }
  __syncwarp();
  trace_stamp(trw, 9);
  // ---- softmax cross-entropy of this row, dZ, and this CTA's 16 columns of dH ----
  float row_loss = 0.f, correct = 0.f;
  if (row_ok) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      float4 a;
      asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                   : "l"(acc + m * 16 + 4 * q4)
                   : "memory");
      s_g[4 * q4][tid] = a.x; s_g[4 * q4 + 1][tid] = a.y; s_g[4 * q4 + 2][tid] = a.z; s_g[4 * q4 + 3][tid] = a.w;
    }
    trace_stamp(trw, 10);
    float mx = -INFINITY, zl = 0.f;
    int amax = 0;
    // modest unrolling: these loops are single-warp dependent chains (latency bound), but full unrolling of
    // run-once code costs more in instruction-cache misses than it saves
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float z = s_g[c][tid] + s_b3[c];
      s_g[c][tid] = z;
      if (z > mx) { mx = z; amax = c; }
      if (c == label) zl = z;
    }
    float se = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float e = __expf(s_g[c][tid] - mx);
      s_g[c][tid] = e;
      se += e;
    }
    const float lse = __logf(se) + mx, inv_se = __fdividef(1.f, se), inv_b = 1.f / static_cast<float>(M);
    row_loss = lse - zl;
    correct = amax == label ? 1.f : 0.f;
    trace_stamp(trw, 11);
    // the bf16-rounded gradient is what the weight-gradient GEMM sees: use the same value for dH
#pragma unroll 4
    for (int c = 0; c < 16; ++c)
      s_g[c][tid] = c < C ? __bfloat162float(__float2bfloat16_rn((s_g[c][tid] * inv_se - (c == label ? 1.f : 0.f)) * inv_b)) : 0.f;
    float d[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) d[jj] = 0.f;
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float gc = s_g[c][tid];
      const float4* wr = reinterpret_cast<const float4*>(s_w3 + c * 16);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 w4 = wr[q4];
        d[4 * q4] = fmaf(gc, w4.x, d[4 * q4]);
        d[4 * q4 + 1] = fmaf(gc, w4.y, d[4 * q4 + 1]);
        d[4 * q4 + 2] = fmaf(gc, w4.z, d[4 * q4 + 2]);
        d[4 * q4 + 3] = fmaf(gc, w4.w, d[4 * q4 + 3]);
      }
    }
    if (ep.head_dh != nullptr) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) d[jj] = (ep.act != 1 || v[jj] > 0.f) ? d[jj] * ep.head_alpha : 0.f;
      __nv_bfloat16* hrow = ep.head_dh + static_cast<size_t>(m) * ep.head_lddh + n0;
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        if (n0 + 8 * h8 + 8 <= N)
          *reinterpret_cast<uint4*>(hrow + 8 * h8) =
              make_uint4(pack_bf16x2(d[8 * h8], d[8 * h8 + 1]), pack_bf16x2(d[8 * h8 + 2], d[8 * h8 + 3]),
                         pack_bf16x2(d[8 * h8 + 4], d[8 * h8 + 5]), pack_bf16x2(d[8 * h8 + 6], d[8 * h8 + 7]));
      }
    }
    if (blockIdx.x == 0 && ep.head_dz != nullptr) {
      __nv_bfloat16* zrow = ep.head_dz + static_cast<size_t>(m) * ep.head_ldz;
#pragma unroll 1
      for (int h8 = 0; h8 < 2; ++h8) {
        if (8 * h8 + 8 <= ep.head_ldz)
          *reinterpret_cast<uint4*>(zrow + 8 * h8) =
              make_uint4(pack_bf16x2(s_g[8 * h8][tid], s_g[8 * h8 + 1][tid]), pack_bf16x2(s_g[8 * h8 + 2][tid], s_g[8 * h8 + 3][tid]),
                         pack_bf16x2(s_g[8 * h8 + 4][tid], s_g[8 * h8 + 5][tid]), pack_bf16x2(s_g[8 * h8 + 6][tid], s_g[8 * h8 + 7][tid]));
      }
    }
  }
  trace_stamp(trw, 12);
  if (blockIdx.x == 0 && ep.head_hist != nullptr) {
    const float l = warp_sum(row_loss), cr = warp_sum(correct);
    if (lane == 0 && m0 + quarter * 32 < m_end) {
      if (slot < 0) slot = 0;
      if (ep.head_hist_slots > 0) slot %= ep.head_hist_slots;
      const float inv_b = 1.f / static_cast<float>(M);
      atomicAdd(ep.head_hist + 2 * slot, l * inv_b);
      atomicAdd(ep.head_hist + 2 * slot + 1, cr * inv_b);
    }
  }
  // ---- one thread publishes the launch count for the next launch (kernel boundary orders it) ----
  if (blockIdx.x == 0 && blockIdx.y == 0 && warp == 2 && lane == 0) ep.head_sync[1] = launch + 1u;
  trace_stamp(trw, 13);
}

// AMN / BMN: the operand is MN-major in global memory, i.e. stored as [K, M] (resp. [K, N])
// row-major with the M (N) index contiguous.  TMA then loads [64 K-rows x 64 MN-elements] boxes
// (one 128-byte swizzle row per K index) and the UMMA descriptor walks K in 8-row atoms
// (SBO = 1024 B) and MN in 64-element chunks (LBO = 8192 B).
//
// Epilogue: each of the 4 epilogue warps owns 32 accumulator rows.  It reads them from TMEM in
// 32-column chunks, applies the fused elementwise work, stages the result in the (now idle)
// pipeline shared memory in the 128-byte-swizzled layout and hands 32 x 128-byte boxes to the TMA
// store engine, so global writes are full coalesced lines and ragged M / N edges are clipped by
// the tensor map.  The dReLU / dropout mask tile is fetched the same way (TMA load).  With
// split-K (gridDim.z > 1) the store is a TMA add-reduction into the fp32 output.
template <int BN, int STAGES, bool TF32, bool AMN, bool BMN>
__global__ void __launch_bounds__(kGemmThreads, (BN <= 128 ? 2 : 1))
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_m,
               const GemmEpilogue ep, const int M, const int N, const int K, const int kb_per_split,
               const int a_box_rows, const int cluster, const int tile_m, const int kch) {
  using S = GemmSmem<BN, STAGES, TF32>;
  // cluster > 1 (DK_GEMM_MCAST_A): the CTAs of a cluster are neighbours along N and read the SAME A tile -- each loads
  // a_box_rows / cluster rows of every k-block and multicasts them into all CTAs' stages; a stage is free again when
  // every CTA of the cluster has consumed it (empty barriers count `cluster` commits, arrived by multicast).
  const uint32_t crank = cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = static_cast<uint16_t>((1u << cluster) - 1u);
  // bytes one k-block of the A operand brings in (DK_GEMM_SHORT_A: a short single M tile uses a short box)
  const int a_stage_bytes = AMN ? S::kABytes : a_box_rows * 128;
  constexpr int kBlockK = TF32 ? 32 : 64;   // elements per 128-byte swizzle row
  constexpr int kUmmaK = TF32 ? 8 : 16;     // 32 bytes of K per tcgen05.mma
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kIdesc = make_idesc(TF32 ? 2u : 1u, kBlockM, BN) | (AMN ? (1u << 15) : 0u) |
                              (BMN ? (1u << 16) : 0u);
  static_assert(!(TF32 && (AMN || BMN)), "MN-major operands are implemented for 16-bit types only");
  static_assert(!BMN || BN >= 64, "MN-major B needs BN >= 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;  // [4], one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4);

  unsigned long long* const tr = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 ? ep.trace : nullptr;
  if (threadIdx.x == 0) trace_stamp(tr, 0);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * tile_m;          // tile_m < 128 (DK_GEMM_TILE_ROWS): short M tiles, rows past it are not this CTA's
  const int m_end = min(M, m0 + tile_m);
  const int total_kb = (K + kBlockK - 1) / kBlockK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;  // uniform for the CTA (never with a cluster: split-K is off there)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (ep.tma_store || kch > 1) tma_prefetch_desc(&tmap_d);
    if (ep.tma_mask || kch > 1) tma_prefetch_desc(&tmap_m);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], cluster > 1 ? cluster : 1);
    }
    mbar_init(tmem_full_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&mask_bar[q], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (cluster > 1) cluster_sync_all();   // every CTA's barriers exist before anybody multicasts into them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(tr, 1);
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail
  // of the previous kernel; operands / bias / mask produced by it are only touched after this point
  DK_PDL_WAIT();
  DK_PDL_TRIGGER();
  if (threadIdx.x == 0) trace_stamp(tr, 2);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int rows_per = cluster > 1 ? a_box_rows / cluster : 0;
      if (kch > 1) {
        // ---- several k-blocks per request (DK_GEMM_KCH): tmap_d / tmap_m are the 3-D views of A / B ----
        const int a_chunk = a_stage_bytes, b_chunk = S::kBBytes, stage_bytes = kch * (a_chunk + b_chunk);
        const int nst = min(STAGES, STAGES * S::kStageBytes / stage_bytes);
        const int groups = (total_kb + kch - 1) / kch;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kch * a_chunk;
          const int nch = min(kch, total_kb - g * kch);
          mbar_expect_tx(&full_bar[stage], nch * (a_chunk + b_chunk));
          if ((g + 1) * kch * kBlockK <= K) {          // every chunk of the group is fully inside K: one request each
            tma_load_3d(sa, &tmap_d, 0, m0, g * kch, &full_bar[stage]);
            tma_load_3d(sb, &tmap_m, 0, n0, g * kch, &full_bar[stage]);
          } else {                                      // K tail: per-chunk 2-D requests (out-of-range columns zero-filled)
            for (int c = 0; c < nch; ++c) {
              tma_load_2d(sa + c * a_chunk, &tmap_a, (g * kch + c) * kBlockK, m0, &full_bar[stage]);
              tma_load_2d(sb + c * b_chunk, &tmap_b, (g * kch + c) * kBlockK, n0, &full_bar[stage]);
            }
          }
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      } else
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[stage], a_stage_bytes + S::kBBytes);
        if constexpr (AMN) {
#pragma unroll
          for (int c = 0; c < kBlockM / 64; ++c)
            tma_load_2d(sa + c * 8192, &tmap_a, m0 + c * 64, kb * kBlockK, &full_bar[stage]);
        } else if (cluster > 1) {
          // this CTA's slice of the shared A k-block (whole 8-row swizzle atoms), delivered to every CTA of the cluster
          tma_load_2d_mcast(sa + crank * rows_per * 128, &tmap_a, kb * kBlockK, m0 + static_cast<int>(crank) * rows_per,
                            &full_bar[stage], cmask);
        } else {
          tma_load_2d(sa, &tmap_a, kb * kBlockK, m0, &full_bar[stage]);
        }
        if constexpr (BMN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sb + c * 8192, &tmap_b, n0 + c * 64, kb * kBlockK, &full_bar[stage]);
        } else {
          tma_load_2d(sb, &tmap_b, kb * kBlockK, n0, &full_bar[stage]);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    if (kch > 1) {
      const int a_chunk = a_stage_bytes, b_chunk = S::kBBytes, stage_bytes = kch * (a_chunk + b_chunk);
      const int nst = min(STAGES, STAGES * S::kStageBytes / stage_bytes);
      const int groups = (total_kb + kch - 1) / kch;
      for (int g = 0; g < groups; ++g) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (g == 0 && lane == 0) trace_stamp(tr, 3);
        if (g == groups - 1 && lane == 0) trace_stamp(tr, 4);
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kch * a_chunk;
          const int nch = min(kch, total_kb - g * kch);
          for (int c = 0; c < nch; ++c) {
            const uint64_t adesc = make_smem_desc_sw128(sa + c * a_chunk);
            const uint64_t bdesc = make_smem_desc_sw128(sb + c * b_chunk);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, kIdesc, (g | c | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (g == groups - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (kb == 0 && lane == 0) trace_stamp(tr, 3);
      if (kb == num_kb - 1 && lane == 0) trace_stamp(tr, 4);
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t adesc = AMN ? make_smem_desc_sw128_mn(sa) : make_smem_desc_sw128(sa);
        const uint64_t bdesc = BMN ? make_smem_desc_sw128_mn(sb) : make_smem_desc_sw128(sb);
        // per-MMA advance along K in (addr >> 4) units: K-major = 32 bytes inside the swizzled
        // row; MN-major = 16 K-rows = two 1024-byte atoms
        constexpr uint32_t kAStep = AMN ? (2048 >> 4) : 2;
        constexpr uint32_t kBStep = BMN ? (2048 >> 4) : 2;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          if constexpr (TF32)
            umma_tf32(tmem_base, adesc + kAStep * k, bdesc + kBStep * k, kIdesc, (kb | k) != 0);
          else
            umma_f16(tmem_base, adesc + kAStep * k, bdesc + kBStep * k, kIdesc, (kb | k) != 0);
        }
        if (cluster > 1) umma_commit_mcast(&empty_bar[stage], cmask);   // ... in every CTA that multicasts into it
        else umma_commit(&empty_bar[stage]);         // frees the smem slot once the MMAs retire
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    if (warp == 2 && lane == 0) {
      mbar_wait(tmem_full_bar, 0);
      trace_stamp(tr, 5);
    }
    bool head_done = false;
    if constexpr (BN == 16 && !TF32 && !AMN && !BMN) {
      if (ep.head_w != nullptr) {
        gemm_epilogue_head(ep, M, m_end, N, m0, n0, warp, lane, tmem_base, tmem_full_bar);
        head_done = true;
      }
    }
    if (!head_done)
      gemm_epilogue<BN>(tmap_d, tmap_m, ep, m_end, N, m0, n0, warp, lane, tmem_base, smem, tmem_full_bar, mask_bar);
    tcgen05_fence_before();
    if (warp == 2 && lane == 0) trace_stamp(tr, 6);
  }

  __syncthreads();
  if (cluster > 1) cluster_sync_all();   // nobody leaves while a peer may still arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
    if (lane == 0) trace_stamp(tr, 7);
  }
}


// ---------------------------------------------------------------------------------------------
// Implicit-GEMM convolution: D[M, N] = epilogue(A_gather[M, K] * B[N, K]^T) where the A operand is
// never materialised -- row m = (b, y, x) of a (GH x GW) position grid, column k = (kh, kw, c), and
//     A[m, k] = S[b, (y * mul - off + kh) / div, (x * mul - off + kw) / div, c]      (0 outside / if not divisible)
// is gathered from the NHWC activation S straight into the 128-byte-swizzled shared-memory tile the
// UMMA descriptor expects (16-byte chunks of 8 channels; C % 8 == 0).  The four epilogue warps are
// idle during the main loop of the non-persistent kernel, so they ARE the gather producers: each
// thread owns one of the 128 tile rows, issues the eight 16-byte loads of k-block i + 1 before it
// stores k-block i (software pipelining over registers), then fence.proxy.async + one mbarrier
// arrive per warp.  B (the weights, K-major) still arrives by TMA.
//   forward:  S = X [B, H, W, Cin],   grid = OH x OW, mul = stride, off = pad,          div = 1
//   dgrad:    S = dZ [B, OH, OW, Cout], grid = H x W,  mul = 1,      off = KH - 1 - pad, div = stride,
//             B = the weights flipped / transposed to [Cin, (kh', kw', cout)]
// This replaces im2col + GEMM (forward reads the activation instead of a 9x larger column matrix)
// and GEMM + col2im (dgrad writes the input gradient once, with the producer's dReLU mask fused).
// ---------------------------------------------------------------------------------------------
struct ConvGather {
  const __nv_bfloat16* src;
  int SH, SW, C;      // source image
  int GH, GW;         // position grid of the GEMM rows
  int KH, KW;
  int mul, off, div;
};

// LDGSTS = true: the gather is issued as cp.async (16 bytes, zero-fill for padding) straight into the
// swizzled tile and completes on the stage's mbarrier (cp.async.mbarrier.arrive.noinc), so up to STAGES
// k-blocks of loads are in flight per thread without holding them in registers; the MMA thread then
// crosses to the async proxy with fence.proxy.async before issuing tcgen05.mma.
template <int BN, int STAGES, bool LDGSTS>
__global__ void __launch_bounds__(kGemmThreads, 2)
conv_gemm_kernel(const ConvGather g, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_m,
                 const GemmEpilogue ep, const int M, const int N, const int K) {
  using S = GemmSmem<BN, STAGES, false>;
  constexpr int kBlockK = 64, kUmmaK = 16;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kIdesc = make_idesc(1u, kBlockM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBlockM;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_b);
    if (ep.tma_store) tma_prefetch_desc(&tmap_d);
    if (ep.tma_mask) tma_prefetch_desc(&tmap_m);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + 128);  // TMA (expect_tx arrive) + every gather thread
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&mask_bar[q], 1);
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
    // ------------------------------ TMA producer (weights) ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sb = smem + stage * S::kStageBytes + S::kABytes;
        mbar_expect_tx(&full_bar[stage], S::kBBytes);
        tma_load_2d(sb, &tmap_b, kb * kBlockK, n0, &full_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        if constexpr (LDGSTS) fence_proxy_async_smem();  // cp.async wrote through the generic proxy
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t adesc = make_smem_desc_sw128(sa);
        const uint64_t bdesc = make_smem_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) != 0);
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ------------------------------ gather producers, then epilogue ----------------------------
    const int r = (warp - 2) * 32 + lane;  // tile row owned by this thread
    const int m = m0 + r;
    const bool row_ok = m < M;
    int b = 0, y = 0, x = 0;
    if (row_ok) {
      b = m / (g.GH * g.GW);
      const int rem = m - b * g.GH * g.GW;
      y = rem / g.GW;
      x = rem - y * g.GW;
    }
    const int ty0 = y * g.mul - g.off, tx0 = x * g.mul - g.off;
    const __nv_bfloat16* img = g.src + static_cast<size_t>(b) * g.SH * g.SW * g.C;
    auto gather = [&](int kb, uint4* v) {
      int k = kb * kBlockK;
      int tap = k / g.C;
      int c = k - tap * g.C;
      int kh = tap / g.KW;
      int kw = tap - kh * g.KW;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = make_uint4(0, 0, 0, 0);
        if (row_ok && k < K) {
          int ty = ty0 + kh, tx = tx0 + kw;
          bool ok = ty >= 0 && tx >= 0;
          if (g.div > 1) {
            ok = ok && (ty % g.div == 0) && (tx % g.div == 0);
            ty /= g.div;
            tx /= g.div;
          }
          if (ok && ty < g.SH && tx < g.SW)
            v[j] = __ldg(reinterpret_cast<const uint4*>(img + (static_cast<size_t>(ty) * g.SW + tx) * g.C + c));
        }
        k += 8;
        c += 8;
        if (c >= g.C) {
          c = 0;
          if (++kw == g.KW) { kw = 0; ++kh; }
        }
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (LDGSTS) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        int k = kb * kBlockK;
        int tap = k / g.C;
        int c = k - tap * g.C;
        int kh = tap / g.KW;
        int kw = tap - kh * g.KW;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat16* src = g.src;  // any valid address when nothing is read (src-size 0 = zero fill)
          uint32_t nbytes = 0;
          if (row_ok && k < K) {
            int ty = ty0 + kh, tx = tx0 + kw;
            bool ok = ty >= 0 && tx >= 0;
            if (g.div > 1) {
              ok = ok && (ty % g.div == 0) && (tx % g.div == 0);
              ty /= g.div;
              tx /= g.div;
            }
            if (ok && ty < g.SH && tx < g.SW) {
              src = img + (static_cast<size_t>(ty) * g.SW + tx) * g.C + c;
              nbytes = 16;
            }
          }
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa + sw128_off(r, j)), "l"(src), "r"(nbytes)
                       : "memory");
          k += 8;
          c += 8;
          if (c >= g.C) {
            c = 0;
            if (++kw == g.KW) { kw = 0; ++kh; }
          }
        }
        // arrive on the stage barrier when this thread's copies have landed (counted in the init value)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else {
    uint4 cur[8], nxt[8];
    gather(0, cur);
    for (int kb = 0; kb < num_kb; ++kb) {
      if (kb + 1 < num_kb) gather(kb + 1, nxt);  // loads in flight while we wait for the slot
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(sa + sw128_off(r, j), cur[j].x, cur[j].y, cur[j].z, cur[j].w);
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[stage]);
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    }
    gemm_epilogue<BN>(tmap_d, tmap_m, ep, M, N, m0, n0, warp, lane, tmem_base, smem, tmem_full_bar, mask_bar);
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES, bool LDGSTS>
static int launch_conv_gemm(const ConvGather& g, const CUtensorMap* tb, const CUtensorMap* td, const CUtensorMap* tm,
                            GemmEpilogue ep, int M, int N, int K, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES, false>;
  auto kern = conv_gemm_kernel<BN, STAGES, LDGSTS>;
  static bool configured[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured[dev & 63] = true;
  }
  const int esize = ep.d_fp32 ? 4 : 2;
  const bool fits = 4 * 32 * BN * esize + (tm != nullptr ? 4 * 32 * BN * 2 : 0) <= STAGES * S::kStageBytes;
  ep.tma_store = (td != nullptr && fits && ep.dt == nullptr) ? 1 : 0;
  ep.tma_mask = (tm != nullptr && ep.tma_store) ? 1 : 0;
  dim3 grid((N + BN - 1) / BN, (M + kBlockM - 1) / kBlockM, 1);
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, kGemmThreads, S::kTotal, stream, g, *tb, ep.tma_store ? *td : *tb,
                          ep.tma_mask ? *tm : *tb, ep, M, N, K));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL (compiled, reachable only with DK_IMPLICIT_WGRAD=1; not yet validated on hardware --
// the round's GPU budget ended before its first run): implicit-GEMM weight gradient
//     dW[co, (kh, kw, c)] += sum_m dZ[m, co] * X[b(m), y(m) * s - p + kh, x(m) * s - p + kw, c]
// i.e. D[M = Cout, N = KH KW Cin] = A^T B with both operands MN-major and the GEMM-K dimension = the
// B OH OW output positions (split-K over gridDim.z, fp32 TMA add-reduction like the dense wgrad).
// A = dZ [rows, Cout] arrives by TMA ([64 co x 64 rows] boxes); the B tile -- 64 positions x BN
// columns of the never-materialised column matrix -- is gathered with the validated cp.async zero-fill
// pattern of conv_gemm_kernel into the MN-major swizzled layout ([64 k-rows][128 B] per 64-column chunk,
// chunks 8192 B apart).  With it the column matrix (and im2col) disappears from training entirely.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_wgrad_kernel(const ConvGather g, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_d, const GemmEpilogue ep, const int M, const int N,
                  const int K, const int kb_per_split) {
  using S = GemmSmem<BN, STAGES, false>;
  constexpr int kBlockK = 64, kUmmaK = 16;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kIdesc = make_idesc(1u, kBlockM, BN) | (1u << 15) | (1u << 16);  // A and B MN-major
  static_assert(BN % 64 == 0, "MN-major B tiles come in 64-column chunks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBlockM;
  const int total_kb = (K + kBlockK - 1) / kBlockK;  // K = number of output positions
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    if (ep.tma_store) tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + 128);  // TMA (expect_tx arrive) + every gather thread
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&mask_bar[q], 1);
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
    // ------------------------------ TMA producer (dZ, MN-major A) ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        mbar_expect_tx(&full_bar[stage], S::kABytes);
#pragma unroll
        for (int c = 0; c < kBlockM / 64; ++c)
          tma_load_2d(sa + c * 8192, &tmap_a, m0 + c * 64, kb * kBlockK, &full_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        fence_proxy_async_smem();  // the gathered tile was written through the generic proxy (cp.async)
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t adesc = make_smem_desc_sw128_mn(sa);
        const uint64_t bdesc = make_smem_desc_sw128_mn(sb);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_f16(tmem_base, adesc + (2048 >> 4) * k, bdesc + (2048 >> 4) * k, kIdesc, (kb | k) != 0);
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ------------------------------ gather producers (B), then epilogue ----------------------------
    const int t = (warp - 2) * 32 + lane;   // 0..127
    const int kr = t & 63;                  // GEMM-k row (output position) inside the k-block
    constexpr int kUnits = BN / 8;          // 16-byte units per row
    const int u_begin = (t >> 6) * (kUnits / 2), u_end = u_begin + kUnits / 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t sb = smem_u32(smem + stage * S::kStageBytes + S::kABytes);
      const int m = kb * kBlockK + kr;      // output position
      const bool row_ok = m < K;
      int b = 0, y = 0, x = 0;
      if (row_ok) {
        b = m / (g.GH * g.GW);
        const int rem = m - b * g.GH * g.GW;
        y = rem / g.GW;
        x = rem - y * g.GW;
      }
      const int ty0 = y * g.mul - g.off, tx0 = x * g.mul - g.off;
      const __nv_bfloat16* img = g.src + static_cast<size_t>(b) * g.SH * g.SW * g.C;
      int n = n0 + u_begin * 8;             // column of the (virtual) column matrix = (tap, c)
      int tap = n / g.C;
      int c = n - tap * g.C;
      int kh = tap / g.KW;
      int kw = tap - kh * g.KW;
      for (int u = u_begin; u < u_end; ++u) {
        const __nv_bfloat16* src = g.src;
        uint32_t nbytes = 0;
        if (row_ok && n < N) {
          const int ty = ty0 + kh, tx = tx0 + kw;
          if (ty >= 0 && tx >= 0 && ty < g.SH && tx < g.SW) {
            src = img + (static_cast<size_t>(ty) * g.SW + tx) * g.C + c;
            nbytes = 16;
          }
        }
        // unit u -> 64-column chunk u / 8 (8192 B apart), 16-byte slot (u % 8) of row kr, 128-byte swizzle
        const uint32_t dst = sb + (u >> 3) * 8192 + kr * 128 + (((u & 7) ^ (kr & 7)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        n += 8;
        c += 8;
        if (c >= g.C) {
          c = 0;
          if (++kw == g.KW) { kw = 0; ++kh; }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[stage])) : "memory");
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    gemm_epilogue<BN>(tmap_d, tmap_d, ep, M, N, m0, n0, warp, lane, tmem_base, smem, tmem_full_bar, mask_bar);
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES>
static int launch_conv_wgrad(const ConvGather& g, const CUtensorMap* ta, const CUtensorMap* td, GemmEpilogue ep, int M,
                             int N, int K, int splits, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES, false>;
  auto kern = conv_wgrad_kernel<BN, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured[dev & 63] = true;
  }
  const int total_kb = (K + 63) / 64;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;
  if (!ep.d_fp32 || td == nullptr) return -6;  // split-K accumulates with the fp32 TMA add-reduction
  ep.tma_store = 1;
  ep.tma_mask = 0;
  dim3 grid((N + BN - 1) / BN, (M + kBlockM - 1) / kBlockM, splits);
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, kGemmThreads, S::kTotal, stream, g, *ta, *td, ep, M, N, K, kb_per_split));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

// dgrad weights: Wd[c, (kh', kw', co)] = W[co, (KH - 1 - kh', KW - 1 - kw', c)]   (both bf16, K-major)
__global__ void __launch_bounds__(256)
conv_weight_flip_kernel(const __nv_bfloat16* __restrict__ w, int ldw, __nv_bfloat16* __restrict__ wd, int ldwd, int Cout,
                        int Cin, int KH, int KW) {
  DK_PDL_ENTER();
  const int total = Cin * KH * KW * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int co = t % Cout; t /= Cout;
    const int kw = t % KW; t /= KW;
    const int kh = t % KH; t /= KH;
    const int c = t;
    wd[static_cast<size_t>(c) * ldwd + (kh * KW + kw) * Cout + co] =
        w[static_cast<size_t>(co) * ldw + ((KH - 1 - kh) * KW + (KW - 1 - kw)) * Cin + c];
  }
}


// ---------------------------------------------------------------------------------------------
// Pull fused into the first GEMM (reference op K2): the forward GEMM of the first layer
// whose B operand -- the layer's weights -- is read by TMA **straight from the parameter server's
// HBM** (the peer-mapped center variable, fp32, consumed as tf32 by tcgen05.mma).  The CTAs of the
// first M-tile row also persist every landed weight tile: while the MMA consumes it from shared
// memory, the otherwise idle epilogue warps copy it to the local fp32 master W, the W1 snapshot
// and the bf16 shadow.  The weights cross NVLink exactly once per N-tile column group and the
// transfer is hidden behind the tensor-core work of the same kernel; no separate pull pass, no
// NCCL call and no host socket is involved.
//   A = X  [M = batch, K = in]  fp32 (local)      B = C_w [N = out, K = in] fp32 (peer HBM)
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_pull_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const GemmEpilogue ep, const int M, const int N,
                 const int K, float* __restrict__ w_local, float* __restrict__ w1_local,
                 __nv_bfloat16* __restrict__ wb_local, const int ldw) {
  using S = GemmSmem<BN, STAGES, true>;
  constexpr int kBlockK = 32, kUmmaK = 8;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kIdesc = make_idesc(2u, kBlockM, BN);
  static_assert(BN == 128 || BN == 64, "persist mapping assumes 32 or 16 weight rows per epilogue warp");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBlockM;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const bool persist = blockIdx.y == 0;  // one CTA per weight-tile column writes the local copies

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (ep.tma_store) tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], persist ? 5 : 1);  // MMA commit (+ the 4 persisting warps)
    }
    mbar_init(tmem_full_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&mask_bar[q], 1);
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
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[stage], S::kStageBytes);
        tma_load_2d(sa, &tmap_a, kb * kBlockK, m0, &full_bar[stage]);  // local activations
        tma_load_2d(sb, &tmap_b, kb * kBlockK, n0, &full_bar[stage]);  // weights from the PS GPU over NVLink
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t adesc = make_smem_desc_sw128(sa);
        const uint64_t bdesc = make_smem_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) != 0);
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    if (persist) {
      // ---- persist the pulled weight tiles while the tensor core consumes them ----
      const int quarter = warp & 3;
      constexpr int kRowsPerWarp = BN / 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sb = smem_u32(smem + stage * S::kStageBytes) + S::kABytes;
        const int k = kb * kBlockK + ((lane & 7) << 2);  // this lane's 4 consecutive K elements
#pragma unroll
        for (int it = 0; it < kRowsPerWarp / 4; ++it) {
          const int r = quarter * kRowsPerWarp + it * 4 + (lane >> 3);  // weight row inside the tile
          const int n = n0 + r;
          const uint4 q = ld_shared_v4(sb + sw128_off(r, lane & 7));
          if (n < N && k + 3 < K) {
            const size_t off = static_cast<size_t>(n) * ldw + k;
            *reinterpret_cast<uint4*>(w_local + off) = q;
            *reinterpret_cast<uint4*>(w1_local + off) = q;
            uint2 h;
            h.x = pack_bf16x2(__uint_as_float(q.x), __uint_as_float(q.y));
            h.y = pack_bf16x2(__uint_as_float(q.z), __uint_as_float(q.w));
            *reinterpret_cast<uint2*>(wb_local + off) = h;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    gemm_epilogue<BN>(tmap_d, tmap_d, ep, M, N, m0, n0, warp, lane, tmem_base, smem, tmem_full_bar, mask_bar);
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

struct GemmLaunch {
  const CUtensorMap* ta;
  const CUtensorMap* tb;
  const CUtensorMap* td;  // may be nullptr
  const CUtensorMap* tm;  // may be nullptr
  GemmEpilogue ep;
  int M, N, K, splits;
  int a_box_rows = kBlockM;
  int cluster = 1;   // DK_GEMM_MCAST_A: CTAs per cluster sharing (and multicasting) the A tile
  int tile_m = kBlockM;   // DK_GEMM_TILE_ROWS: height of the M tile (short tiles: several CTAs along M for M < 128)
  int kch = 0;            // DK_GEMM_KCH: k-blocks per TMA request (td / tm then carry the 3-D operand views)
  cudaStream_t stream;
};

template <int BN, int STAGES, bool TF32, bool AMN = false, bool BMN = false>
static int launch_gemm(const GemmLaunch& L) {
  const CUtensorMap& ta = *L.ta;
  const CUtensorMap& tb = *L.tb;
  GemmEpilogue ep = L.ep;
  const int M = L.M, N = L.N, K = L.K;
  cudaStream_t stream = L.stream;
  using S = GemmSmem<BN, STAGES, TF32>;
  auto kern = gemm_tn_kernel<BN, STAGES, TF32, AMN, BMN>;
  static bool configured[64] = {};  // function attributes are per device
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured[dev & 63] = true;
  }
  constexpr int kBlockK = TF32 ? 32 : 64;
  const int total_kb = (K + kBlockK - 1) / kBlockK;
  int splits = L.splits < 1 ? 1 : L.splits;
  if (splits > total_kb) splits = total_kb;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;
  // the TMA-store epilogue stages 4 x 32 x BN outputs (+ the mask tile) in the pipeline buffers
  const int esize = ep.d_fp32 ? 4 : 2;
  const bool fits = 4 * 32 * BN * esize + (L.tm != nullptr ? 4 * 32 * BN * 2 : 0) <= STAGES * S::kStageBytes;
  const bool wide = ep.d_fp32 ? BN >= 32 : BN >= 64;
  ep.tma_store = (L.td != nullptr && fits && wide && ep.dt == nullptr) ? 1 : 0;
  ep.tma_mask = (L.tm != nullptr && ep.tma_store && BN >= 64) ? 1 : 0;
  if (splits > 1 && !ep.d_fp32) return -6;
  dim3 grid((N + BN - 1) / BN, (M + L.tile_m - 1) / L.tile_m, splits);
  if (L.tile_m != kBlockM && (ep.tma_store || ep.tma_mask || AMN || splits > 1)) return -8;   // short tiles: direct-store epilogue only
  const int kch = L.kch > 1 ? L.kch : 0;
  if (kch) {
    // stages of kch k-blocks must fit the kernel's shared-memory pool at least twice; td / tm are operand views here
    if (TF32 || AMN || BMN || BN > 32 || splits > 1 || L.cluster > 1 || L.td == nullptr || L.tm == nullptr ||
        2 * kch * (L.a_box_rows * 128 + S::kBBytes) > STAGES * S::kStageBytes)
      return -8;
    ep.tma_store = 0;
    ep.tma_mask = 0;
  }
  if (L.cluster > 1) {
    if (AMN || splits > 1 || grid.y != 1) return -8;
    grid.x = (grid.x + L.cluster - 1) / L.cluster * L.cluster;   // whole clusters; the extra CTAs' tiles are clipped
    DK_HOST_CHECK(launch_kernel_cluster(kern, grid, dim3(kGemmThreads), S::kTotal, stream, static_cast<unsigned>(L.cluster), ta, tb,
                                        ep.tma_store ? *L.td : ta, ep.tma_mask ? *L.tm : ta, ep, M, N, K, kb_per_split,
                                        L.a_box_rows, L.cluster, L.tile_m, 0));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, kGemmThreads, S::kTotal, stream, ta, tb, (ep.tma_store || kch) ? *L.td : ta,
                          (ep.tma_mask || kch) ? *L.tm : ta, ep, M, N, K, kb_per_split, L.a_box_rows, 1, L.tile_m, kch));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Persistent variant for the bf16-output GEMMs of the forward / dgrad chain (short K, epilogue
// heavy): one CTA per SM loops over output tiles; the fp32 accumulator is DOUBLE-BUFFERED in TMEM
// (2 x BN columns) so the epilogue warps drain tile i (tcgen05.ld -> bias / ReLU / dropout / mask
// -> swizzled smem slice -> TMA store, 64 columns at a time through a two-slot staging ring)
// while the MMA warp already accumulates tile i+1 and the TMA producer prefetches its operands.
// A is K-major; B is K-major (forward) or MN-major (dgrad).  256-wide tiles raise the arithmetic
// intensity against L2 to 85 FLOP/B.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct PersistentSmem {
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutStage = 8 * 4096;    // 8 epilogue warps x one 32-row x 128-byte slot
  static constexpr int kMaskStage = 8 * 4096;
  static constexpr int kBiasStage = 8 * 1024;   // per-warp copy of the tile's bias slice (<= 256 floats)
  static constexpr int kBarrierBytes = 512;
  static constexpr int kTotal = STAGES * kStageBytes + kOutStage + kMaskStage + kBiasStage + kBarrierBytes + 1024;
};

constexpr int kPersistentThreads = 320;  // producer + MMA + 8 epilogue warps

template <int BN, int STAGES, bool BMN>
__global__ void __launch_bounds__(kPersistentThreads, 1)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_m,
                       const GemmEpilogue ep, const int M, const int N, const int K) {
  using S = PersistentSmem<BN, STAGES>;
  constexpr int kBlockK = 64, kUmmaK = 16;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulators
  constexpr uint32_t kIdesc = make_idesc(1u, kBlockM, BN) | (BMN ? (1u << 16) : 0u);
  static_assert(BN == 64 || BN == 128 || BN == 256, "persistent kernel: BN in {64, 128, 256}");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_stage = smem + STAGES * S::kStageBytes;
  uint8_t* mask_stage = out_stage + S::kOutStage;
  float* bias_stage = reinterpret_cast<float*>(mask_stage + S::kMaskStage);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(mask_stage + S::kMaskStage + S::kBiasStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* mask_bar = tmem_empty_bar + 2;        // [8 epilogue warps]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int n_tiles = (N + BN - 1) / BN;
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
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
      mbar_init(&tmem_empty_bar[a], 8);  // one arrival per epilogue warp
    }
    for (int i = 0; i < 8; ++i) mbar_init(&mask_bar[i], 1);
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
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kBlockM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_2d(sa, &tmap_a, kb * kBlockK, m0, &full_bar[stage]);
          if constexpr (BMN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * 8192, &tmap_b, n0 + c * 64, kb * kBlockK, &full_bar[stage]);
          } else {
            tma_load_2d(sb, &tmap_b, kb * kBlockK, n0, &full_bar[stage]);
          }
          if (++stage == STAGES) {
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = BMN ? make_smem_desc_sw128_mn(sb) : make_smem_desc_sw128(sb);
          constexpr uint32_t kBStep = BMN ? (2048 >> 4) : 2;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_f16(tmem_acc, adesc + 2 * k, bdesc + kBStep * k, kIdesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    // 8 warps: warp e = warp - 2; TMEM lane quarter = warp & 3; the two warps of a quarter
    // alternate over the 64-column slices (group = e >> 2 takes slices of matching parity).
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int group = e >> 2;
    const uint32_t out_region = smem_u32(out_stage) + e * 4096;
    const uint32_t mask_region = smem_u32(mask_stage) + e * 4096;
    float* my_bias = bias_stage + e * 256;
    uint64_t* my_mask_bar = mask_bar + e;
    uint32_t mask_phase = 0;
    uint32_t drop_salt = 0, drop_thr = 0;
    float keep_scale = 1.f;
    if (ep.drop_p > 0.f) {
      drop_salt = ep.drop_seed + (ep.step != nullptr ? static_cast<uint32_t>(*ep.step) : 0u) * 0x85EBCA77u;
      drop_thr = static_cast<uint32_t>(ep.drop_p * 256.f + 0.5f);
      keep_scale = 256.f / (256.f - static_cast<float>(drop_thr));
    }
    constexpr int kSlices = BN / 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles) * kBlockM, n0 = (tile % n_tiles) * BN;
      const int acc = it & 1;
      const int m = m0 + quarter * 32 + lane;
      const int valid_slices = min(kSlices, (N - n0 + 63) / 64);
      // first slice of this tile handled by this warp
      int sl = ((it * kSlices) & 1) == group ? 0 : 1;
      if (ep.tma_mask && lane == 0 && sl < valid_slices) {
        mbar_expect_tx(my_mask_bar, 4096);
        tma_load_2d_addr(mask_region, &tmap_m, n0 + sl * 64, m0 + quarter * 32, my_mask_bar);
      }
      if (ep.bias != nullptr) {
        // stage this tile's bias slice once (broadcast reads from smem in the slice loop)
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
          // this warp's last read of the accumulator: hand it back to the MMA warp
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
        if (ep.drop_p > 0.f) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            uint32_t h = (static_cast<uint32_t>(m) * static_cast<uint32_t>(N) + static_cast<uint32_t>(nc + j)) * 0x9E3779B1u ^ drop_salt;
            h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
#pragma unroll
            for (int t = 0; t < 4; ++t)
              v[j + t] = (((h >> (8 * t)) & 0xFFu) < drop_thr) ? 0.f : v[j + t] * keep_scale;
          }
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
          if (lane == 0 && sl + 2 < valid_slices) {  // prefetch this warp's next mask slice
            mbar_expect_tx(my_mask_bar, 4096);
            tma_load_2d_addr(mask_region, &tmap_m, nc + 128, m0 + quarter * 32, my_mask_bar);
          }
        }
        // the previous TMA store out of this warp's slot must have finished reading it
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
      if (!arrived) {  // no slice of this tile for this warp: still release the accumulator
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

template <int BN, int STAGES, bool BMN>
static int launch_persistent(const GemmLaunch& L) {
  using S = PersistentSmem<BN, STAGES>;
  auto kern = gemm_persistent_kernel<BN, STAGES, BMN>;
  static bool configured[64] = {};
  static int sm_count[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    DK_HOST_CHECK(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    configured[dev & 63] = true;
  }
  GemmEpilogue ep = L.ep;
  ep.tma_store = 1;
  ep.tma_mask = L.tm != nullptr ? 1 : 0;
  const int tiles = ((L.M + kBlockM - 1) / kBlockM) * ((L.N + BN - 1) / BN);
  const int grid = tiles < sm_count[dev & 63] ? tiles : sm_count[dev & 63];
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, kPersistentThreads, S::kTotal, L.stream, *L.ta, *L.tb, *L.td,
                          ep.tma_mask ? *L.tm : *L.ta, ep, L.M, L.N, L.K));
  return 0;
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs on the SMs of one TPC compute a 256 x BN tile with ONE
// tcgen05.mma stream issued by the leader.  Each CTA stages its own 128 rows of A and only HALF of
// the B tile (BN/2 rows), so the shared-memory traffic per CTA (TMA writes + UMMA operand reads)
// drops from 16+32 KB to 16+16 KB per k-block at BN = 256 and the L2 -> SM traffic for B halves.
// TMA loads of both CTAs credit the leader's full barrier; tcgen05.commit multicasts the
// "slot free" / "accumulator ready" arrivals to both CTAs.  K-major bf16 operands.
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct PairSmem {
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = (BN / 2) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = STAGES * kStageBytes + kBarrierBytes + 1024;
};

template <int BN, int STAGES, bool AMN, bool BMN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_m,
                 const GemmEpilogue ep, const int M, const int N, const int K, const int kb_per_split) {
  using S = PairSmem<BN, STAGES>;
  constexpr int kBlockK = 64, kUmmaK = 16;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kIdesc = make_idesc(1u, 2 * kBlockM, BN) | (AMN ? (1u << 15) : 0u) |
                              (BMN ? (1u << 16) : 0u);  // M = 256 across the pair
  static_assert(!BMN || BN >= 128, "MN-major B halves need BN / 2 >= 64");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* mask_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n0 = blockIdx.y * BN;
  const int m0 = pair * (2 * kBlockM) + static_cast<int>(rank) * kBlockM;
  const int nb0 = n0 + static_cast<int>(rank) * (BN / 2);  // this CTA's half of the B tile
  const int total_kb = (K + kBlockK - 1) / kBlockK;
  const int kb_begin = blockIdx.z * kb_per_split;            // split-K slice (uniform for the pair)
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (ep.tma_store) tma_prefetch_desc(&tmap_d);
    if (ep.tma_mask) tma_prefetch_desc(&tmap_m);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&mask_bar[q], 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrival
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  DK_PDL_WAIT();
  DK_PDL_TRIGGER();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * S::kStageBytes);  // bytes of BOTH CTAs
        if constexpr (AMN) {
#pragma unroll
          for (int c = 0; c < kBlockM / 64; ++c)
            tma_load_2d_2cta(sa + c * 8192, &tmap_a, m0 + c * 64, kb * kBlockK, &full_bar[stage]);
        } else {
          tma_load_2d_2cta(sa, &tmap_a, kb * kBlockK, m0, &full_bar[stage]);
        }
        if constexpr (BMN) {
#pragma unroll
          for (int c = 0; c < BN / 2 / 64; ++c)
            tma_load_2d_2cta(sb + c * 8192, &tmap_b, nb0 + c * 64, kb * kBlockK, &full_bar[stage]);
        } else {
          tma_load_2d_2cta(sb, &tmap_b, kb * kBlockK, nb0, &full_bar[stage]);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
          const uint64_t adesc = AMN ? make_smem_desc_sw128_mn(sa) : make_smem_desc_sw128(sa);
          const uint64_t bdesc = BMN ? make_smem_desc_sw128_mn(sb) : make_smem_desc_sw128(sb);
          constexpr uint32_t kAStep = AMN ? (2048 >> 4) : 2;
          constexpr uint32_t kBStep = BMN ? (2048 >> 4) : 2;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_f16_2cta(tmem_base, adesc + kAStep * k, bdesc + kBStep * k, kIdesc, (kb | k) != 0);
          umma_commit_2cta(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit_2cta(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    gemm_epilogue<BN>(tmap_d, tmap_m, ep, M, N, m0, n0, warp, lane, tmem_base, smem, tmem_full_bar, mask_bar);
    tcgen05_fence_before();
  }

  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or arrivals are in flight
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

template <int BN, int STAGES, bool AMN = false, bool BMN = false>
static int launch_pair(const GemmLaunch& L) {
  using S = PairSmem<BN, STAGES>;
  auto kern = gemm_pair_kernel<BN, STAGES, AMN, BMN>;
  static bool configured[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured[dev & 63] = true;
  }
  GemmEpilogue ep = L.ep;
  const int esize = ep.d_fp32 ? 4 : 2;
  const bool fits = 4 * 32 * BN * esize + (L.tm != nullptr ? 4 * 32 * BN * 2 : 0) <= STAGES * S::kStageBytes;
  ep.tma_store = (L.td != nullptr && fits && ep.dt == nullptr) ? 1 : 0;
  ep.tma_mask = (L.tm != nullptr && ep.tma_store) ? 1 : 0;
  const int pairs = (L.M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int total_kb = (L.K + 63) / 64;
  int splits = L.splits < 1 ? 1 : L.splits;
  if (splits > total_kb) splits = total_kb;
  const int kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + kb_per_split - 1) / kb_per_split;
  if (splits > 1 && !ep.d_fp32) return -6;
  dim3 grid(2 * pairs, (L.N + BN - 1) / BN, splits);
  DK_HOST_CHECK(launch_kernel_cluster(kern, grid, dim3(kGemmThreads), S::kTotal, L.stream, 2u, *L.ta, *L.tb,
                                      ep.tma_store ? *L.td : *L.ta, ep.tma_mask ? *L.tm : *L.ta, ep, L.M, L.N, L.K,
                                      kb_per_split));
  return 0;
}

}  // namespace dk

extern "C" {

int dk_tmap_encode_2d(void* out_tmap, const void* base, int dtype, long rows, long cols, long ld,
                      int box_rows) {
  auto fn = dk::get_encode_fn();
  if (fn == nullptr) return -1;
  const int esize = dtype == DK_F32 ? 4 : 2;
  const int box_cols = 128 / esize;
  if ((ld * esize) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return -2;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esize};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap),
                  dtype == DK_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

// Rows of the TMA box of a K-major A operand: a single short M tile (M < 128, the small-batch regime) is
// loaded with a box of just ceil8(M) rows -- half the shared-memory fill of a 128-row box whose rows past M
// would only be zero-filled; the accumulator rows past M are never stored.
// [rows, K] row-major bf16 operand viewed as [64 elements, rows, K / 64 full chunks]: box = [64, box_rows, kch]
int dk_tmap_encode_kchunks(void* out_tmap, const void* base, long rows, long K, long ld, int box_rows, int kch) {
  auto fn = dk::get_encode_fn();
  if (fn == nullptr) return -1;
  const long full = K / 64;
  if (full < 1 || kch < 1 || kch > 16 || box_rows < 1 || box_rows > 256) return -3;
  if ((ld * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return -2;
  cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(full)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(kch)};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim,
                  gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -4;
}

int dk_gemm_a_box_rows(int M) {
  if (M >= dk::kBlockM) return dk::kBlockM;
  return (M + 7) / 8 * 8;
}

// DK_GEMM_MCAST_A: the A box is split into whole 8-row swizzle atoms over the CTAs of a cluster (at most 8)
int dk_gemm_mcast_cluster(int M) {
  const int atoms = dk_gemm_a_box_rows(M) / 8;
  for (int cl = 8; cl >= 2; --cl)
    if (atoms % cl == 0) return cl;
  return 1;
}

int dk_gemm_mcast_box_rows(int M) { return dk_gemm_a_box_rows(M) / dk_gemm_mcast_cluster(M); }

int dk_gemm_pick_bn(int N) {
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return 128;
}

// Tile width that also looks at M: narrow tiles when 128-wide ones cannot fill the 148 SMs.
int dk_gemm_pick_bn2(int M, int N) {
  const int bn = dk_gemm_pick_bn(N);
  if (bn < 128) return bn;
  const long tiles128 = static_cast<long>((M + 127) / 128) * ((N + 127) / 128);
  if (tiles128 < 120 && N > 64) return 64;
  // measured (profiles/): for the short-K forward / dgrad GEMMs two co-resident 128-wide CTAs per
  // SM (epilogue of one overlapping the main loop of the other) beat one 256-wide CTA
  return 128;
}

// wgrad-style GEMMs (fp32 output, split-K): wide tiles, the split factor restores the parallelism
int dk_gemm_pick_bn_splitk(int M, int N, int K) {
  if (N >= 512 && K >= 2048) return 256;
  if (N > 64) return 128;
  return dk_gemm_pick_bn(N) < 64 ? 64 : dk_gemm_pick_bn(N);
}

// Launch with pre-encoded tensor maps.  flags: DK_GEMM_TF32 | DK_GEMM_A_MN | DK_GEMM_B_MN.
// K-major operand maps are encoded with box_rows = 128 (A) / bn (B) over a [rows, K] matrix;
// MN-major operand maps with box_rows = 64 over the [K, rows] matrix.  tmap_d / tmap_m (optional)
// describe the output / mask matrices [M, N] with box_rows = 32 for the TMA-store epilogue.
int dk_gemm_tn_launch2(const void* tmap_a, const void* tmap_b, const void* tmap_d, const void* tmap_m,
                       const DkGemmEpilogue* ep, int M, int N, int K, int bn, int flags, int splits, void* stream) {
  dk::GemmLaunch L;
  L.ta = reinterpret_cast<const CUtensorMap*>(tmap_a);
  L.tb = reinterpret_cast<const CUtensorMap*>(tmap_b);
  L.td = reinterpret_cast<const CUtensorMap*>(tmap_d);
  L.tm = reinterpret_cast<const CUtensorMap*>(tmap_m);
  L.ep = *ep;
  L.M = M; L.N = N; L.K = K; L.splits = splits;
  L.stream = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0 || K <= 0) return -3;
  const bool tf32 = flags & DK_GEMM_TF32, amn = flags & DK_GEMM_A_MN, bmn = flags & DK_GEMM_B_MN;
  if (ep->head_w != nullptr) {
    // fused classifier head: plain bn = 16 kernel, one M tile, a grid small enough to be co-resident
    if (bn != 16 || M > dk::kBlockM || splits > 1 || (flags & ~(DK_GEMM_SHORT_A | DK_GEMM_MCAST_A | 0xFFF00)) != 0 ||
        ((N + 15) / 16) * (DK_GEMM_TILE_ROWS_OF(flags) ? (M + DK_GEMM_TILE_ROWS_OF(flags) - 1) / DK_GEMM_TILE_ROWS_OF(flags) : 1) > 64 || N % 8 != 0 ||
        ep->head_c > 16 || ep->head_c < 1 || ep->head_labels == nullptr || ep->head_acc == nullptr ||
        ep->head_sync == nullptr || ep->d_fp32 || ep->dt != nullptr || ep->mask != nullptr || ep->bias_along_m ||
        (ep->ldd % 8) != 0 || (ep->head_lddh % 8) != 0 || (ep->head_ldz % 8) != 0)
      return -9;
    L.ep.tma_store = 0;
    L.ep.tma_mask = 0;
    if (DK_GEMM_KCH_OF(flags) <= 1) {   // (with DK_GEMM_KCH the two extra maps are operand views, not output / mask maps)
      L.td = nullptr;
      L.tm = nullptr;
    }
  }
  L.kch = DK_GEMM_KCH_OF(flags);
  if (L.kch > 1 && (flags & (DK_GEMM_PAIR | DK_GEMM_PERSISTENT | DK_GEMM_A_MN | DK_GEMM_B_MN | DK_GEMM_TF32 | DK_GEMM_MCAST_A))) return -8;
  if (flags & DK_GEMM_SHORT_A) {
    if ((flags & (DK_GEMM_PAIR | DK_GEMM_PERSISTENT | DK_GEMM_A_MN)) || M > dk::kBlockM) return -8;  // plain kernel, M <= 128
    L.a_box_rows = dk_gemm_a_box_rows(M);
    const int tr = DK_GEMM_TILE_ROWS_OF(flags);
    if (tr != 0) {
      if (tr % 8 != 0 || tr >= dk::kBlockM || (flags & DK_GEMM_MCAST_A)) return -8;
      L.a_box_rows = tr;
      L.tile_m = tr;
    }
    if (flags & DK_GEMM_MCAST_A) {
      L.cluster = dk_gemm_mcast_cluster(M);
      if (splits > 1 || tf32 || bmn) return -8;
    }
  } else if (flags & DK_GEMM_MCAST_A) {
    return -8;
  }
  if ((flags & DK_GEMM_PAIR) && !tf32) {
    // K-major B tensor maps must have been encoded with box_rows = bn / 2 (each CTA loads half the tile)
    if (!amn && !bmn) {
      if (bn == 256) return dk::launch_pair<256, 6>(L);
      if (bn == 128) return dk::launch_pair<128, 8>(L);
    } else if (amn && bmn) {
      if (bn == 256) return dk::launch_pair<256, 6, true, true>(L);
      if (bn == 128) return dk::launch_pair<128, 8, true, true>(L);
    } else if (!amn && bmn) {
      if (bn == 256) return dk::launch_pair<256, 6, false, true>(L);
      if (bn == 128) return dk::launch_pair<128, 8, false, true>(L);
    }
    return -4;
  }
  if ((flags & DK_GEMM_PERSISTENT) && !tf32 && !amn && L.td != nullptr && !ep->d_fp32 && ep->dt == nullptr &&
      splits <= 1 && (ep->mask == nullptr || L.tm != nullptr) && !ep->bias_along_m) {
    if (bn == 256) return bmn ? dk::launch_persistent<256, 3, true>(L) : dk::launch_persistent<256, 3, false>(L);
    if (bn == 128) return bmn ? dk::launch_persistent<128, 4, true>(L) : dk::launch_persistent<128, 4, false>(L);
    if (bn == 64) return bmn ? dk::launch_persistent<64, 6, true>(L) : dk::launch_persistent<64, 6, false>(L);
  }
  if (tf32) {
    if (amn || bmn) return -5;
    switch (bn) {
      case 16: return dk::launch_gemm<16, 4, true>(L);
      case 32: return dk::launch_gemm<32, 4, true>(L);
      case 64: return dk::launch_gemm<64, 4, true>(L);
      case 128: return dk::launch_gemm<128, 3, true>(L);
      default: return -4;
    }
  }
  if (!amn && !bmn) {
    // One M tile (the reference's batch sizes): the K loop is a chain of L2 round trips, so the ring is made
    // deep enough to have (nearly) every k-block in flight at once; the CTA then owns the SM's shared memory.
    if (M <= dk::kBlockM && K > 6 * 64) {
      if (bn == 16) return dk::launch_gemm<16, 12, false>(L);
      if (bn == 32) return dk::launch_gemm<32, 11, false>(L);
    }
    switch (bn) {
      case 16: return dk::launch_gemm<16, 6, false>(L);
      case 32: return dk::launch_gemm<32, 6, false>(L);
      case 64: return dk::launch_gemm<64, 4, false>(L);
      case 128: return dk::launch_gemm<128, 3, false>(L);
      case 256: return dk::launch_gemm<256, 4, false>(L);
      default: return -4;
    }
  }
  if (!amn && bmn) {
    switch (bn) {
      case 64: return dk::launch_gemm<64, 4, false, false, true>(L);
      case 128: return dk::launch_gemm<128, 3, false, false, true>(L);
      case 256: return dk::launch_gemm<256, 4, false, false, true>(L);
      default: return -4;
    }
  }
  if (amn && bmn) {
    switch (bn) {
      case 64: return dk::launch_gemm<64, 4, false, true, true>(L);
      case 128: return dk::launch_gemm<128, 3, false, true, true>(L);
      case 256: return dk::launch_gemm<256, 4, false, true, true>(L);
      default: return -4;
    }
  }
  switch (bn) {  // A MN-major, B K-major
    case 16: return dk::launch_gemm<16, 6, false, true, false>(L);
    case 32: return dk::launch_gemm<32, 6, false, true, false>(L);
    case 64: return dk::launch_gemm<64, 4, false, true, false>(L);
    case 128: return dk::launch_gemm<128, 3, false, true, false>(L);
    default: return -4;
  }
}

int dk_gemm_tn_launch(const void* tmap_a, const void* tmap_b, const DkGemmEpilogue* ep, int M, int N,
                      int K, int bn, int flags, void* stream) {
  return dk_gemm_tn_launch2(tmap_a, tmap_b, nullptr, nullptr, ep, M, N, K, bn, flags, 1, stream);
}

// Output / mask tensor maps for the TMA-store epilogue; returns 0 on success, non-zero if the
// matrix cannot be described (unaligned base / stride) in which case the direct-store path is used.
int dk_gemm_encode_output(void* tmap_d, const void* D, long ldd, int M, int N, int d_fp32) {
  return dk_tmap_encode_2d(tmap_d, D, d_fp32 ? DK_F32 : DK_BF16, M, N, ldd, 32);
}

// Split-K factor that fills the GPU (2 CTAs / SM) without making the slices too short.
// split factor for the CTA-pair kernel: 74 pairs fill the 148 SMs
int dk_gemm_pick_splits_pair(int M, int N, int K, int bn) {
  const int tiles = ((M + 255) / 256) * ((N + bn - 1) / bn);
  const int total_kb = (K + 63) / 64;
  int splits = 74 / tiles;
  if (splits > total_kb / 4) splits = total_kb / 4;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;
  return splits;
}

int dk_gemm_pick_splits(int M, int N, int K, int bn, int tf32) {
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  const int total_kb = (K + (tf32 ? 32 : 64) - 1) / (tf32 ? 32 : 64);
  const int slots = bn > 128 ? 148 : 2 * 148;  // resident CTAs (wide tiles run 1 CTA / SM)
  int splits = slots / tiles;  // floor: tiles x splits must fit in ONE wave (a 2nd, nearly empty wave doubles the time)
  if (splits > total_kb / 4) splits = total_kb / 4;
  if (splits < 1) splits = 1;
  if (splits > 32) splits = 32;
  return splits;
}

int dk_gemm_encode_operands(void* tmap_a, void* tmap_b, const void* A, long lda, const void* B, long ldb,
                            int M, int N, int K, int bn, int flags) {
  const int dt = (flags & DK_GEMM_TF32) ? DK_F32 : DK_BF16;
  int r = (flags & DK_GEMM_A_MN) ? dk_tmap_encode_2d(tmap_a, A, dt, K, M, lda, 64)
                                 : dk_tmap_encode_2d(tmap_a, A, dt, M, K, lda,
                                                     DK_GEMM_TILE_ROWS_OF(flags) ? DK_GEMM_TILE_ROWS_OF(flags)
                                                     : (flags & DK_GEMM_MCAST_A) ? dk_gemm_mcast_box_rows(M)
                                                     : (flags & DK_GEMM_SHORT_A) ? dk_gemm_a_box_rows(M) : dk::kBlockM);
  if (r != 0) return r;
  return (flags & DK_GEMM_B_MN) ? dk_tmap_encode_2d(tmap_b, B, dt, K, N, ldb, 64)
                                : dk_tmap_encode_2d(tmap_b, B, dt, N, K, ldb, (flags & DK_GEMM_PAIR) ? bn / 2 : bn);
}


// Fused pull + forward GEMM of the first layer (see gemm_pull_kernel).  `center_w` is the layer's
// weight block inside the peer-mapped center variable; w / w1 / wb are the local copies to refresh.
int dk_gemm_pull_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, const DkGemmEpilogue* ep_in, int M,
                        int N, int K, float* w_local, float* w1_local, void* wb_local, int ldw, void* stream) {
  using S = dk::GemmSmem<128, 3, true>;
  auto kern = dk::gemm_pull_kernel<128, 3>;
  static bool configured[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured[dev & 63] = true;
  }
  dk::GemmEpilogue ep = *ep_in;
  ep.tma_store = tmap_d != nullptr ? 1 : 0;
  ep.tma_mask = 0;
  ep.mask = nullptr;
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  const CUtensorMap& td = tmap_d != nullptr ? *reinterpret_cast<const CUtensorMap*>(tmap_d) : ta;
  dim3 grid((N + 127) / 128, (M + dk::kBlockM - 1) / dk::kBlockM, 1);
  DK_HOST_CHECK(DK_LAUNCH(kern, grid, dk::kGemmThreads, S::kTotal, stream, ta, tb, td, ep, M, N, K, w_local, w1_local,
                          reinterpret_cast<__nv_bfloat16*>(wb_local), ldw));
  return 0;
}

// one-shot variant (tests): encodes the tensor maps first
int dk_gemm_pull(const float* X, long ldx, const float* center_w, long ldc, const DkGemmEpilogue* ep, int M, int N, int K,
                 float* w_local, float* w1_local, void* wb_local, void* stream) {
  alignas(64) CUtensorMap ta, tb, td;
  int r = dk_tmap_encode_2d(&ta, X, DK_F32, M, K, ldx, dk::kBlockM);
  if (r != 0) return r;
  r = dk_tmap_encode_2d(&tb, center_w, DK_F32, N, K, ldc, 128);
  if (r != 0) return r;
  const bool has_d = ep->d != nullptr && dk_gemm_encode_output(&td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0 &&
                     (ep->d_fp32 || true);
  return dk_gemm_pull_launch(&ta, &tb, has_d ? &td : nullptr, ep, M, N, K, w_local, w1_local, wb_local, (int)ldc, stream);
}

// Convenience one-shot entry: encodes the tensor maps, then launches (splits: 0 = auto for fp32
// outputs without bias / activation, otherwise 1).
int dk_gemm_tn_ex(const void* A, long lda, const void* B, long ldb, const DkGemmEpilogue* ep, int M, int N,
                  int K, int flags, int bn, int splits, void* stream) {
  alignas(64) CUtensorMap ta, tb, td, tm;
  if (bn <= 0) bn = dk_gemm_pick_bn(N);
  if ((flags & DK_GEMM_B_MN) && bn < 64) bn = 64;
  int r = dk_gemm_encode_operands(&ta, &tb, A, lda, B, ldb, M, N, K, bn, flags);
  if (r != 0) return r;
  const bool has_d = ep->d != nullptr && dk_gemm_encode_output(&td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0;
  const bool has_m = ep->mask != nullptr && dk_gemm_encode_output(&tm, ep->mask, ep->ld_mask, M, N, 0) == 0;
  if (splits <= 0) splits = 1;
  return dk_gemm_tn_launch2(&ta, &tb, has_d ? &td : nullptr, has_m ? &tm : nullptr, ep, M, N, K, bn, flags, splits,
                            stream);
}

int dk_gemm_tn(const void* A, long lda, const void* B, long ldb, const DkGemmEpilogue* ep, int M,
               int N, int K, int flags, int bn, void* stream) {
  return dk_gemm_tn_ex(A, lda, B, ldb, ep, M, N, K, flags, bn, 1, stream);
}

// Implicit-GEMM convolution launch (see conv_gemm_kernel).  tmap_b: weights [N, K] K-major encoded with
// box_rows = bn; tmap_d / tmap_m: output / mask [M, N] (optional) as for dk_gemm_tn_launch2.
// A-gather implementation of the implicit-GEMM convolution: 0 = register-pipelined ld.global + st.shared
// (default), 1 = cp.async ring completing on the stage mbarrier.  mode < 0 only queries; the initial value
// comes from DK_CONV_LDGSTS.
int dk_conv_gather_mode(int mode) {
  static int current = -1;
  if (current < 0) {
    const char* e = getenv("DK_CONV_LDGSTS");
    current = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (mode >= 0) current = mode ? 1 : 0;
  return current;
}

int dk_conv_gemm_launch(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off,
                        int div, const void* tmap_b, const void* tmap_d, const void* tmap_m, const DkGemmEpilogue* ep,
                        int M, int N, int K, int bn, void* stream) {
  if (C % 8 != 0 || K != KH * KW * C || M <= 0 || N <= 0) return -3;
  dk::ConvGather g;
  g.src = reinterpret_cast<const __nv_bfloat16*>(src);
  g.SH = SH; g.SW = SW; g.C = C; g.GH = GH; g.GW = GW; g.KH = KH; g.KW = KW; g.mul = mul; g.off = off; g.div = div;
  const CUtensorMap* tb = reinterpret_cast<const CUtensorMap*>(tmap_b);
  const CUtensorMap* td = reinterpret_cast<const CUtensorMap*>(tmap_d);
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(tmap_m);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dk_conv_gather_mode(-1) == 1) {
    if (bn == 64) return dk::launch_conv_gemm<64, 4, true>(g, tb, td, tm, *ep, M, N, K, st);
    if (bn == 128) return dk::launch_conv_gemm<128, 3, true>(g, tb, td, tm, *ep, M, N, K, st);
    return -4;
  }
  if (bn == 64) return dk::launch_conv_gemm<64, 4, false>(g, tb, td, tm, *ep, M, N, K, st);
  if (bn == 128) return dk::launch_conv_gemm<128, 3, false>(g, tb, td, tm, *ep, M, N, K, st);
  return -4;
}

int dk_conv_pick_bn(int N) { return N <= 64 ? 64 : 128; }

// one-shot variant (tests): encodes the tensor maps first
int dk_conv_gemm(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off, int div,
                 const void* Bmat, long ldb, const DkGemmEpilogue* ep, int M, int N, int K, void* stream) {
  alignas(64) CUtensorMap tb, td, tm;
  const int bn = dk_conv_pick_bn(N);
  int r = dk_tmap_encode_2d(&tb, Bmat, DK_BF16, N, K, ldb, bn);
  if (r != 0) return r;
  const bool has_d = ep->d != nullptr && dk_gemm_encode_output(&td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0;
  const bool has_m = ep->mask != nullptr && dk_gemm_encode_output(&tm, ep->mask, ep->ld_mask, M, N, 0) == 0;
  return dk_conv_gemm_launch(src, SH, SW, C, GH, GW, KH, KW, mul, off, div, &tb, has_d ? &td : nullptr,
                             has_m ? &tm : nullptr, ep, M, N, K, bn, stream);
}

// EXPERIMENTAL implicit wgrad (see conv_wgrad_kernel): dW [Cout, KH KW Cin] fp32 += dZ^T * gather(X).
// tmap_a: dZ stored [rows, Cout] encoded as an MN-major A operand (dk_tmap_encode_2d(.., rows = K_gemm, cols = Cout, box 64));
// tmap_d: the fp32 gradient matrix [Cout, KH KW Cin] (must be zeroed by the caller: split-K adds into it).
int dk_conv_wgrad_launch(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                         const void* tmap_a, const void* tmap_d, const DkGemmEpilogue* ep, int Cout, int rows, int splits,
                         void* stream) {
  if (C % 8 != 0 || Cout <= 0 || rows <= 0) return -3;
  dk::ConvGather g;
  g.src = reinterpret_cast<const __nv_bfloat16*>(src);
  g.SH = SH; g.SW = SW; g.C = C; g.GH = GH; g.GW = GW; g.KH = KH; g.KW = KW; g.mul = stride; g.off = pad; g.div = 1;
  return dk::launch_conv_wgrad<128, 4>(g, reinterpret_cast<const CUtensorMap*>(tmap_a),
                                       reinterpret_cast<const CUtensorMap*>(tmap_d), *ep, Cout, KH * KW * C, rows, splits,
                                       reinterpret_cast<cudaStream_t>(stream));
}

// one-shot variant (tests)
int dk_conv_wgrad(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                  const void* dz, long lddz, float* dw, long lddw, int Cout, int rows, int splits, void* stream) {
  alignas(64) CUtensorMap ta, td;
  int r = dk_tmap_encode_2d(&ta, dz, DK_BF16, rows, Cout, lddz, 64);
  if (r != 0) return r;
  r = dk_gemm_encode_output(&td, dw, lddw, Cout, KH * KW * C, 1);
  if (r != 0) return r;
  DkGemmEpilogue ep = {};
  ep.alpha = 1.f;
  ep.d = dw;
  ep.ldd = static_cast<int>(lddw);
  ep.d_fp32 = 1;
  return dk_conv_wgrad_launch(src, SH, SW, C, GH, GW, KH, KW, stride, pad, &ta, &td, &ep, Cout, rows, splits, stream);
}

int dk_conv_weight_flip(const void* w, int ldw, void* wd, int ldwd, int Cout, int Cin, int KH, int KW, void* stream) {
  const int total = Cin * KH * KW * Cout;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  DK_HOST_CHECK(DK_LAUNCH(dk::conv_weight_flip_kernel, blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream),
                          reinterpret_cast<const __nv_bfloat16*>(w), ldw, reinterpret_cast<__nv_bfloat16*>(wd), ldwd, Cout,
                          Cin, KH, KW));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"

