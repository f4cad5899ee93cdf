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

// AMN / BMN: the operand is MN-major in global memory, i.e. stored as [K, M] (resp. [K, N])
// row-major with the M (N) index contiguous.  TMA then loads [64 K-rows x 64 MN-elements] boxes
// (one 128-byte swizzle row per K index) and the UMMA descriptor walks K in 8-row atoms
// (SBO = 1024 B) and MN in 64-element chunks (LBO = 8192 B).
template <int BN, int STAGES, bool TF32, bool AMN, bool BMN>
__global__ void __launch_bounds__(kGemmThreads, (BN <= 128 ? 2 : 1))
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmEpilogue ep, const int M, const int N, const int K) {
  using S = GemmSmem<BN, STAGES, TF32>;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kBlockM;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
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
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[stage], S::kStageBytes);
        if constexpr (AMN) {
#pragma unroll
          for (int c = 0; c < kBlockM / 64; ++c)
            tma_load_2d(sa + c * 8192, &tmap_a, m0 + c * 64, kb * kBlockK, &full_bar[stage]);
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
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
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
        umma_commit(&empty_bar[stage]);              // frees the smem slot once the MMAs retire
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
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int m = m0 + quarter * 32 + lane;        // output row owned by this thread
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const bool row_ok = m < M;
    const float bias_m = (ep.bias != nullptr && ep.bias_along_m && row_ok) ? ep.bias[m] : 0.f;
    constexpr int kChunk = BN < 32 ? 16 : 32;
#pragma unroll 1
    for (int c = 0; c < BN; c += kChunk) {
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
      const int nc = n0 + c;
      if (nc >= N) break;  // warp-uniform
      const bool full_chunk = nc + kChunk <= N;
      // ---- fused elementwise epilogue ----
#pragma unroll
      for (int j = 0; j < kChunk; ++j) {
        float x = v[j] * ep.alpha;
        if (ep.bias != nullptr) {
          if (ep.bias_along_m) x += bias_m;
          else if (nc + j < N) x += __ldg(ep.bias + nc + j);
        }
        if (ep.act == 1) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      if (ep.drop_p > 0.f) {
        // inverted dropout (reference op K10): counter-based hash of (seed, step, element index)
        const uint32_t salt = ep.drop_seed + (ep.step != nullptr ? static_cast<uint32_t>(*ep.step) : 0u) * 0x85EBCA77u;
        const float keep_scale = 1.f / (1.f - ep.drop_p);
#pragma unroll
        for (int j = 0; j < kChunk; ++j) {
          uint32_t h = (static_cast<uint32_t>(m) * static_cast<uint32_t>(N) + static_cast<uint32_t>(nc + j)) * 0x9E3779B1u ^ salt;
          h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
          v[j] = (static_cast<float>(h) * 2.3283064365386963e-10f < ep.drop_p) ? 0.f : v[j] * keep_scale;
        }
      }
      if (ep.mask != nullptr && row_ok) {
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
      // ---- row-major store ----
      if (ep.d != nullptr && row_ok) {
        if (ep.d_fp32) {
          float* drow = reinterpret_cast<float*>(ep.d) + static_cast<size_t>(m) * ep.ldd + nc;
          if (full_chunk && ((reinterpret_cast<uintptr_t>(drow) & 15) == 0)) {
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

template <int BN, int STAGES, bool TF32, bool AMN = false, bool BMN = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M,
                       int N, int K, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES, TF32>;
  auto kern = gemm_tn_kernel<BN, STAGES, TF32, AMN, BMN>;
  static bool configured = false;
  if (!configured) {
    DK_HOST_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + kBlockM - 1) / kBlockM);
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(ta, tb, ep, M, N, K);
  DK_HOST_CHECK(cudaGetLastError());
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

int dk_gemm_pick_bn(int N) {
  if (N <= 16) return 16;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  return 128;
}

// Launch with pre-encoded tensor maps.  flags: DK_GEMM_TF32 | DK_GEMM_A_MN | DK_GEMM_B_MN.
// K-major operand maps are encoded with box_rows = 128 (A) / bn (B) over a [rows, K] matrix;
// MN-major operand maps with box_rows = 64 over the [K, rows] matrix.
int dk_gemm_tn_launch(const void* tmap_a, const void* tmap_b, const DkGemmEpilogue* ep, int M, int N,
                      int K, int bn, int flags, void* stream) {
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0 || K <= 0) return -3;
  const bool tf32 = flags & DK_GEMM_TF32, amn = flags & DK_GEMM_A_MN, bmn = flags & DK_GEMM_B_MN;
  if (tf32) {
    if (amn || bmn) return -5;
    switch (bn) {
      case 16: return dk::launch_gemm<16, 4, true>(ta, tb, *ep, M, N, K, st);
      case 32: return dk::launch_gemm<32, 4, true>(ta, tb, *ep, M, N, K, st);
      case 64: return dk::launch_gemm<64, 4, true>(ta, tb, *ep, M, N, K, st);
      case 128: return dk::launch_gemm<128, 3, true>(ta, tb, *ep, M, N, K, st);
      default: return -4;
    }
  }
  if (!amn && !bmn) {
    switch (bn) {
      case 16: return dk::launch_gemm<16, 6, false>(ta, tb, *ep, M, N, K, st);
      case 32: return dk::launch_gemm<32, 6, false>(ta, tb, *ep, M, N, K, st);
      case 64: return dk::launch_gemm<64, 6, false>(ta, tb, *ep, M, N, K, st);
      case 128: return dk::launch_gemm<128, 3, false>(ta, tb, *ep, M, N, K, st);
      case 256: return dk::launch_gemm<256, 4, false>(ta, tb, *ep, M, N, K, st);
      default: return -4;
    }
  }
  if (!amn && bmn) {
    switch (bn) {
      case 64: return dk::launch_gemm<64, 6, false, false, true>(ta, tb, *ep, M, N, K, st);
      case 128: return dk::launch_gemm<128, 3, false, false, true>(ta, tb, *ep, M, N, K, st);
      case 256: return dk::launch_gemm<256, 4, false, false, true>(ta, tb, *ep, M, N, K, st);
      default: return -4;
    }
  }
  if (amn && bmn) {
    switch (bn) {
      case 64: return dk::launch_gemm<64, 6, false, true, true>(ta, tb, *ep, M, N, K, st);
      case 128: return dk::launch_gemm<128, 3, false, true, true>(ta, tb, *ep, M, N, K, st);
      case 256: return dk::launch_gemm<256, 4, false, true, true>(ta, tb, *ep, M, N, K, st);
      default: return -4;
    }
  }
  switch (bn) {  // A MN-major, B K-major
    case 16: return dk::launch_gemm<16, 6, false, true, false>(ta, tb, *ep, M, N, K, st);
    case 32: return dk::launch_gemm<32, 6, false, true, false>(ta, tb, *ep, M, N, K, st);
    case 64: return dk::launch_gemm<64, 6, false, true, false>(ta, tb, *ep, M, N, K, st);
    case 128: return dk::launch_gemm<128, 3, false, true, false>(ta, tb, *ep, M, N, K, st);
    default: return -4;
  }
}

int dk_gemm_encode_operands(void* tmap_a, void* tmap_b, const void* A, long lda, const void* B, long ldb,
                            int M, int N, int K, int bn, int flags) {
  const int dt = (flags & DK_GEMM_TF32) ? DK_F32 : DK_BF16;
  int r = (flags & DK_GEMM_A_MN) ? dk_tmap_encode_2d(tmap_a, A, dt, K, M, lda, 64)
                                 : dk_tmap_encode_2d(tmap_a, A, dt, M, K, lda, dk::kBlockM);
  if (r != 0) return r;
  return (flags & DK_GEMM_B_MN) ? dk_tmap_encode_2d(tmap_b, B, dt, K, N, ldb, 64)
                                : dk_tmap_encode_2d(tmap_b, B, dt, N, K, ldb, bn);
}

// Convenience one-shot entry: encodes both tensor maps, then launches.
int dk_gemm_tn(const void* A, long lda, const void* B, long ldb, const DkGemmEpilogue* ep, int M,
               int N, int K, int flags, int bn, void* stream) {
  alignas(64) CUtensorMap ta, tb;
  if (bn <= 0) bn = dk_gemm_pick_bn(N);
  if ((flags & DK_GEMM_B_MN) && bn < 64) bn = 64;
  int r = dk_gemm_encode_operands(&ta, &tb, A, lda, B, ldb, M, N, K, bn, flags);
  if (r != 0) return r;
  return dk_gemm_tn_launch(&ta, &tb, ep, M, N, K, bn, flags, stream);
}

}  // extern "C"
