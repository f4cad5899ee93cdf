// Fused dense backward-update kernel for the small-batch (reference-regime) training step.
//
// The reference trains at 4-64 rows per mini-batch (examples/mnist_analysis.ipynb:387,
// example_1_analysis.ipynb:434).  At that size a step is launch- and latency-bound, so everything
// that follows the input-gradient chain is ONE kernel, for ALL dense layers of the model at once:
//
//   per 128 x 64 tile of a layer's kernel [n_out, k_in]:
//     dW  = dZ^T X            tcgen05.mma, both operands MN-major straight from their row-major
//                             storage by TMA (GEMM-K = the mini-batch), fp32 accumulator in TMEM
//     db  = colsum(dZ)        one extra N = 16 tcgen05.mma of the same A tile against a tile of ones
//                             (second TMEM accumulator) -- no column-sum kernel
//     W, m, v <- optimizer    applied by the epilogue straight from TMEM: the gradient never goes
//                             to memory, there is no gradient buffer, memset or optimizer launch
//     Wb <- bf16(W)           the shadow the next forward GEMM reads
//   on the step that closes a communication window, additionally (same epilogue, no extra launch):
//     r = (W - W1) * s ; old = atom.add.sys(center, r) ; W = W1 = old + r        (ADAG / DOWNPOUR / DynSGD)
//     E = a (W - C) ; W -= E ; red.add.sys(center, E)                             (AEASGD / EAMSGD)
//   with `center` in the parameter server's HBM (peer-mapped over NVLink): the worker -> PS push, its
//   scale and the PS-side update, and the PS -> worker pull are fused into the weight-gradient GEMM
//   of the backward pass, tile by tile (reference: distkeras/workers.py:327-342 does this with
//   get_weights + numpy + pickle + TCP; parameter_servers.py:276-285 applies it under a mutex).
//
// Replaces, per step: 3 column-sum kernels, 3 wgrad GEMMs, the gradient memset, the optimizer kernel
// and (once per window) the exchange kernel of the wide-batch path.
#include "dense_fused.h"

#include <string.h>

#include <new>

#include "common.cuh"
#include "gemm.h"
#include "gemm_device.cuh"
#include "optim.cuh"
#include "ps.h"

namespace dk {

constexpr int kBwdBlockM = 128;
constexpr int kBwdBN = 64;
constexpr int kBwdThreads = 320;  // producer, MMA issuer, 8 epilogue warps
constexpr int kBwdStages = 2;  // GEMM-K = the mini-batch: one or two 64-row k-blocks in the compact regime

// TMA descriptors of a layer's parameter state (device memory, written once at plan time): the fp32 master,
// optimizer state and last-pulled-center tiles are LOADED as [128 x 32] boxes while the MMA runs, and the
// updated tiles (and the bf16 shadow) are STORED back as [32-row] boxes -- the optimizer never issues a
// global load or store of its own.
enum { kStW = 0, kStS0 = 1, kStS1 = 2, kStW1 = 3, kStWb = 4 };
struct StateMaps {
  alignas(64) CUtensorMap ld[4];  // W, s0, s1, W1: fp32 [n_out, k_in], box 32 cols x 128 rows, SWIZZLE_128B
  alignas(64) CUtensorMap st[5];  // W, s0, s1, W1: box 32 x 32; Wb: bf16 box 64 cols x 32 rows (kept for reference / tools)
  alignas(64) CUtensorMap wb128;  // Wb: bf16 box 64 cols x 128 rows: the whole shadow tile in one store
};

struct BwdLayerDev {
  alignas(64) CUtensorMap ta;  // dZ [batch, n_out] as an MN-major A operand (64 x 64 boxes)
  alignas(64) CUtensorMap tb;  // X  [batch, k_in]  as an MN-major B operand
  const StateMaps* maps;       // nullptr for layers on the element-wise path
  long w_off, b_off;
  __nv_bfloat16* wb_pad;
  int ldwb_pad;
  int n_out, k_in, tiles_n, tile_begin, vec;
};

struct BwdUpdateDev {
  BwdLayerDev layer[DK_BWD_MAX_LAYERS];
  float* shard_center[DK_BWD_MAX_SHARDS];
  float* w;
  float* s0;
  float* s1;
  float* w1;
  __nv_bfloat16* wb;
  int* step;
  unsigned* done_counter;
  const float* scale_dev;
  unsigned* ctrl;
  unsigned* last_update;
  unsigned long long* trace;
  long shard_per;
  OptimArgs opt;
  int nlayers, batch, step_inc, comm_mode, nshards, worker;
  float comm_scale, alpha;
};

struct BwdRecord {
  BwdUpdateDev dev;
  const void* x[DK_BWD_MAX_LAYERS];
  long ldx[DK_BWD_MAX_LAYERS];
  StateMaps* maps_dev;  // device array [nlayers]
  int total_tiles;
};

__device__ __forceinline__ float* center_of(const BwdUpdateDev& p, long idx) {
  if (p.nshards <= 1) return p.shard_center[0] + idx;
  const long s = idx / p.shard_per;
  return p.shard_center[s] + (idx - s * p.shard_per);
}

__device__ __forceinline__ float4 atom_add_v4_sys_f(float* addr, float4 v) {
  float4 o;
  asm volatile("atom.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4], {%5, %6, %7, %8};"
               : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
               : "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
  return o;
}

__device__ __forceinline__ float atom_add_sys_f(float* addr, float v) {
  float o;
  asm volatile("atom.relaxed.sys.global.add.f32 %0, [%1], %2;" : "=f"(o) : "l"(addr), "f"(v) : "memory");
  return o;
}

__device__ __forceinline__ void red_add_v4_sys_f(float* addr, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void red_add_sys_f(float* addr, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float4 ld_sys_v4_f(const float* addr) {
  float4 o;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
               : "l"(addr)
               : "memory");
  return o;
}

__device__ __forceinline__ float ld_sys_f(const float* addr) {
  float o;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(o) : "l"(addr) : "memory");
  return o;
}

// one parameter: optimizer rule (+ bf16 shadow when no exchange follows)
template <int KIND>
__device__ __forceinline__ void update_scalar(const BwdUpdateDev& p, long idx, float g, float lr, const OptimCorr& corr,
                                              bool write_shadow) {
  float w = p.w[idx];
  float s0 = p.s0 != nullptr ? p.s0[idx] : 0.f;
  float s1 = p.s1 != nullptr ? p.s1[idx] : 0.f;
  optim_update<KIND>(w, g, s0, s1, lr, p.opt, corr);
  p.w[idx] = w;
  if (p.s0 != nullptr) p.s0[idx] = s0;
  if (p.s1 != nullptr) p.s1[idx] = s1;
  if (write_shadow) p.wb[idx] = __float2bfloat16_rn(w);
}

// one parameter: window-boundary exchange with the parameter server
__device__ __forceinline__ float exchange_scalar(const BwdUpdateDev& p, long idx, float s) {
  float w = p.w[idx];
  if (p.comm_mode == DK_COMM_EXCHANGE) {
    const float r = (w - p.w1[idx]) * s;
    w = atom_add_sys_f(center_of(p, idx), r) + r;
    p.w1[idx] = w;
  } else {
    const float c = ld_sys_f(center_of(p, idx));
    const float e = p.alpha * (w - c);
    w -= e;
    red_add_sys_f(center_of(p, idx), e);
  }
  p.w[idx] = w;
  p.wb[idx] = __float2bfloat16_rn(w);
  return w;
}

// Shared-memory plan (after the 1024-byte alignment):
//   [0, kStagesBytes)        TMA <-> MMA pipeline (A 16 KB + B 8 KB per stage); idle once the accumulator is
//                            complete, the first 32 KB are then reused as the 8 epilogue warps' 4 KB gradient slots
//   ones (2 KB)              bf16 1.0 tile: B operand of the bias-gradient MMA
//   state tiles              W, s0, s1, W1: two [128 rows x 128 B] swizzled half-tiles each (32 KB per array)
//   Wb tile (16 KB)          [128 rows x 128 B] bf16
constexpr int kBwdStagesBytes = kBwdStages * (kBwdBlockM * 128 + kBwdBN * 128);
constexpr int kBwdTileBytes = 2 * kBwdBlockM * 128;  // one fp32 state array of a 128 x 64 tile
constexpr int kBwdSmemUse = kBwdStagesBytes + 2048 + 4 * kBwdTileBytes + kBwdBlockM * 128 + 256;


// Epilogue of one (quarter, half) = 32 rows x 32 columns slice of the accumulator tile, executed by one warp
// with lane = row (the TMEM-native mapping).  The parameter state of the slice is already in shared memory
// (TMA, 128-byte swizzle: lane = row accesses are bank-conflict free); the loop over the slice's eight 16-byte
// column groups is NOT unrolled -- these kernels run once per step, so straight-line code is paid for in
// instruction-cache misses.  On a window boundary the slice is then exchanged with the parameter server with
// lane = column-group (coalesced 128-byte runs per row over NVLink), all 8 requests of a lane in flight.
template <int KIND>
__device__ __forceinline__ void bwd_epilogue(const BwdUpdateDev& p, const BwdLayerDev& ly, const int m0, const int n0,
                                             const int quarter, const int half, const int lane, const uint32_t tmem_base,
                                             const uint32_t gslot, const uint32_t tiles, const uint32_t wb_tile,
                                             uint64_t* state_bar, const bool has_bias, const float lr, const OptimCorr& corr,
                                             const float cs, float bw, float bs0, float bs1) {
  constexpr bool kS0 = KIND != DK_OPT_SGD;
  constexpr bool kS1 = KIND == DK_OPT_ADAM || KIND == DK_OPT_ADADELTA || KIND == DK_OPT_ADAMAX || KIND == DK_OPT_NADAM;
  const bool comm = p.comm_mode != DK_COMM_NONE;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
  const int nbase = n0 + half * 32;                  // first column of this warp's slice
  const int mbase = m0 + quarter * 32;               // first row
  const bool slice_live = nbase < ly.k_in;
  if (slice_live) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(trow + half * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4)                   // row `lane`, 16-byte chunk c4 -> swizzled slot
      st_shared_v4(gslot + lane * 128 + ((c4 ^ (lane & 7)) << 4), r[4 * c4], r[4 * c4 + 1], r[4 * c4 + 2], r[4 * c4 + 3]);
    __syncwarp();
  }
  unsigned long long* const tr = (blockIdx.x == 0 && quarter == 2 && half == 0 && lane == 0) ? p.trace : nullptr;
  trace_stamp(tr, 8);
  if (ly.maps != nullptr) {
    // =============================== TMA-fed path (k_in % 4 == 0) ===============================
    const uint32_t rowoff = (quarter * 32 + lane) * 128;          // this lane's row inside a half-tile
    const uint32_t tW = tiles + half * (kBwdBlockM * 128);
    const uint32_t tS0 = tW + kBwdTileBytes, tS1 = tW + 2 * kBwdTileBytes, tW1 = tW + 3 * kBwdTileBytes;
    if (slice_live) {
      mbar_wait(state_bar, 0);                                     // the state tiles have landed
      if (blockIdx.x == 0 && quarter == 2 && half == 0 && lane == 0) trace_stamp(p.trace, 5);
#pragma unroll 1
      for (int jj = 0; jj < 4; ++jj) {                             // two 16-byte column groups per trip
        uint32_t packed[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int j = 2 * jj + h2;
          const uint32_t off = rowoff + ((j ^ (lane & 7)) << 4);
          const uint4 gq = ld_shared_v4(gslot + lane * 128 + ((j ^ (lane & 7)) << 4));
          uint4 wq = ld_shared_v4(tW + off);
          uint4 aq = kS0 ? ld_shared_v4(tS0 + off) : make_uint4(0, 0, 0, 0);
          uint4 bq = kS1 ? ld_shared_v4(tS1 + off) : make_uint4(0, 0, 0, 0);
          float w[4] = {__uint_as_float(wq.x), __uint_as_float(wq.y), __uint_as_float(wq.z), __uint_as_float(wq.w)};
          float a[4] = {__uint_as_float(aq.x), __uint_as_float(aq.y), __uint_as_float(aq.z), __uint_as_float(aq.w)};
          float b[4] = {__uint_as_float(bq.x), __uint_as_float(bq.y), __uint_as_float(bq.z), __uint_as_float(bq.w)};
          const float g[4] = {__uint_as_float(gq.x), __uint_as_float(gq.y), __uint_as_float(gq.z), __uint_as_float(gq.w)};
#pragma unroll
          for (int e = 0; e < 4; ++e) optim_update<KIND>(w[e], g[e], a[e], b[e], lr, p.opt, corr);
          st_shared_v4(tW + off, __float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3]));
          if constexpr (kS0)
            st_shared_v4(tS0 + off, __float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
          if constexpr (kS1)
            st_shared_v4(tS1 + off, __float_as_uint(b[0]), __float_as_uint(b[1]), __float_as_uint(b[2]), __float_as_uint(b[3]));
          packed[2 * h2] = pack_bf16x2(w[0], w[1]);
          packed[2 * h2 + 1] = pack_bf16x2(w[2], w[3]);
        }
        // 8 bf16 = 16-byte chunk (4 half + jj) of this row of the shadow tile
        st_shared_v4(wb_tile + rowoff + (((4 * half + jj) ^ (lane & 7)) << 4), packed[0], packed[1], packed[2], packed[3]);
      }
      trace_stamp(tr, 9);
      if (comm) {
        // ---- window boundary: push the window's displacement, adopt the center ----
        __syncwarp();
        const int sub = lane >> 3, c4 = lane & 7;     // lane = 16-byte column group, 4 rows per warp access
        const int n = nbase + 4 * c4;
        const bool col_ok = n < ly.k_in;
        float4 x[8], y[8];
        bool ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = quarter * 32 + 4 * i + sub;
          ok[i] = col_ok && (m0 + row) < ly.n_out;
          const uint32_t off = row * 128 + ((c4 ^ (row & 7)) << 4);
          const uint4 q = ld_shared_v4(tW + off);     // updated W
          x[i] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
          y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.comm_mode == DK_COMM_EXCHANGE) {
            const uint4 b = ld_shared_v4(tW1 + off);  // last pulled center
            x[i] = make_float4((x[i].x - __uint_as_float(b.x)) * cs, (x[i].y - __uint_as_float(b.y)) * cs,
                               (x[i].z - __uint_as_float(b.z)) * cs, (x[i].w - __uint_as_float(b.w)) * cs);  // residual
          }
        }
        if (p.comm_mode == DK_COMM_EXCHANGE) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (ok[i]) y[i] = atom_add_v4_sys_f(center_of(p, ly.w_off + static_cast<long>(mbase + 4 * i + sub) * ly.k_in + n), x[i]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y[i].x += x[i].x; y[i].y += x[i].y; y[i].z += x[i].z; y[i].w += x[i].w;   // new center = new W = new W1
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (ok[i]) y[i] = ld_sys_v4_f(center_of(p, ly.w_off + static_cast<long>(mbase + 4 * i + sub) * ly.k_in + n));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!ok[i]) continue;
            const float4 e = make_float4(p.alpha * (x[i].x - y[i].x), p.alpha * (x[i].y - y[i].y), p.alpha * (x[i].z - y[i].z),
                                         p.alpha * (x[i].w - y[i].w));
            red_add_v4_sys_f(center_of(p, ly.w_off + static_cast<long>(mbase + 4 * i + sub) * ly.k_in + n), e);
            y[i] = make_float4(x[i].x - e.x, x[i].y - e.y, x[i].z - e.z, x[i].w - e.w);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = quarter * 32 + 4 * i + sub;
          const uint32_t off = row * 128 + ((c4 ^ (row & 7)) << 4);
          st_shared_v4(tW + off, __float_as_uint(y[i].x), __float_as_uint(y[i].y), __float_as_uint(y[i].z), __float_as_uint(y[i].w));
          if (p.comm_mode == DK_COMM_EXCHANGE)
            st_shared_v4(tW1 + off, __float_as_uint(y[i].x), __float_as_uint(y[i].y), __float_as_uint(y[i].z),
                         __float_as_uint(y[i].w));
          // bf16 shadow: 4 values = 8 bytes at byte (64 half + 8 c4) of the row: chunk (4 half + c4 / 2), half (c4 & 1)
          const uint32_t wboff = row * 128 + (((4 * half + (c4 >> 1)) ^ (row & 7)) << 4) + ((c4 & 1) << 3);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(wb_tile + wboff), "r"(pack_bf16x2(y[i].x, y[i].y)),
                       "r"(pack_bf16x2(y[i].z, y[i].w))
                       : "memory");
        }
      }
      // ---- write the half-tile back: ONE TMA store per array for the four warps of this half (a request costs the
      //      TMA unit ~190 cycles whatever its size: 3 + 3 + 1 stores per tile instead of 8 x 3 + 4), out of the
      //      swizzled tiles through the same [128 x 32] maps that loaded them (clipped at the matrix edge) ----
      fence_proxy_async_smem();
      named_bar_sync(5 + half, 128);
      if (quarter == 0 && lane == 0) {
        tma_store_2d_addr(&ly.maps->ld[kStW], tW, nbase, m0);
        if constexpr (kS0) tma_store_2d_addr(&ly.maps->ld[kStS0], tS0, nbase, m0);
        if constexpr (kS1) tma_store_2d_addr(&ly.maps->ld[kStS1], tS1, nbase, m0);
        if (p.comm_mode == DK_COMM_EXCHANGE) tma_store_2d_addr(&ly.maps->ld[kStW1], tW1, nbase, m0);
      }
      trace_stamp(tr, 10);
    }
    // all eight warps have filled the [128 x 128 B] bf16 shadow tile: one store
    named_bar_sync(7, 256);
    if (quarter == 0 && half == 0 && lane == 0 && n0 < ly.k_in) tma_store_2d_addr(&ly.maps->wb128, wb_tile, n0, m0);
    if (ly.wb_pad != nullptr && half == 1) {
      // the TMA store above went to the 8-padded shadow; the flat shadow (rows 8-byte aligned: k_in % 4 == 0) gets the
      // same 32 x 64 slice through plain stores -- half a warp per row, 8 bytes per lane, 128 contiguous bytes per row
      const int c16 = (lane & 15) >> 1, h8 = lane & 1;
      const int ncol = n0 + 8 * c16 + 4 * h8;
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += 2) {
        const int row = quarter * 32 + rr + (lane >> 4);
        const int m = m0 + row;
        if (m < ly.n_out && ncol < ly.k_in) {
          uint32_t lo, hi;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(wb_tile + row * 128 + ((c16 ^ (row & 7)) << 4) + (h8 << 3)));
          *reinterpret_cast<uint2*>(p.wb + ly.w_off + static_cast<long>(m) * ly.k_in + ncol) = make_uint2(lo, hi);
        }
      }
    }
    trace_stamp(tr, 11);
    if (lane == 0) {
      tma_store_commit();
      tma_store_wait_read<0>();
    }
    trace_stamp(tr, 12);
  } else if (slice_live) {
    // ====== element-wise path (k_in % 4 != 0, e.g. the 30-feature Higgs input layer): lane = column group ======
    const int sub = lane >> 3, c4 = lane & 7;
    const int n = nbase + 4 * c4;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + sub;
      const int m = mbase + row;
      if (m >= ly.n_out) continue;
      const uint4 q = ld_shared_v4(gslot + row * 128 + ((c4 ^ (row & 7)) << 4));
      const float gq[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        if (n + j >= ly.k_in) break;
        const long idx = ly.w_off + static_cast<long>(m) * ly.k_in + n + j;
        update_scalar<KIND>(p, idx, gq[j], lr, corr, !comm);
        float wn = p.w[idx];
        if (comm) wn = exchange_scalar(p, idx, cs);
        if (ly.wb_pad != nullptr) ly.wb_pad[static_cast<long>(m) * ly.ldwb_pad + n + j] = __float2bfloat16_rn(wn);
      }
    }
  }
  // ---- bias: its gradient is column 0 of the ones-tile accumulator; lane = row, contiguous in memory ----
  if (has_bias && half == 0) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(trow + kBwdBN, r);
    tmem_ld_wait();
    const int m = mbase + lane;
    if (m < ly.n_out) {
      const long idx = ly.b_off + m;                 // bw / bs0 / bs1 were loaded before the accumulator wait
      optim_update<KIND>(bw, __uint_as_float(r[0]), bs0, bs1, lr, p.opt, corr);
      p.w[idx] = bw;
      if constexpr (kS0) p.s0[idx] = bs0;
      if constexpr (kS1) p.s1[idx] = bs1;
      if (comm) exchange_scalar(p, idx, cs);
      else p.wb[idx] = __float2bfloat16_rn(bw);
    }
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
dense_bwd_update_kernel(const __grid_constant__ BwdUpdateDev p) {
  constexpr int BN = kBwdBN;
  constexpr int kABytes = kBwdBlockM * 128;
  constexpr int kBBytes = BN * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 128;  // BN accumulator columns + 16 bias-gradient columns, power of two
  constexpr uint32_t kIdesc = make_idesc(1u, kBwdBlockM, BN) | (1u << 15) | (1u << 16);  // A and B MN-major
  constexpr uint32_t kIdescBias = make_idesc(1u, kBwdBlockM, 16) | (1u << 15);          // A MN-major, ones K-major

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ones = smem + kBwdStagesBytes;            // 16 rows x 128 B of bf16 1.0 (layout-invariant)
  uint8_t* tiles = ones + 2048;                      // W, s0, s1, W1 state tiles
  uint8_t* wb_tile = tiles + 4 * kBwdTileBytes;      // bf16 shadow tile
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wb_tile + kBwdBlockM * 128);
  uint64_t* empty_bar = full_bar + kBwdStages;
  uint64_t* tmem_full_bar = empty_bar + kBwdStages;
  uint64_t* state_bar = tmem_full_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(state_bar + 1);

  unsigned long long* const tr = blockIdx.x == 0 ? p.trace : nullptr;
  if (threadIdx.x == 0) trace_stamp(tr, 0);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int L = 0;
  while (L + 1 < p.nlayers && static_cast<int>(blockIdx.x) >= p.layer[L + 1].tile_begin) ++L;
  const BwdLayerDev& ly = p.layer[L];
  const int local = static_cast<int>(blockIdx.x) - ly.tile_begin;
  const int m0 = (local / ly.tiles_n) * kBwdBlockM;
  const int n0 = (local % ly.tiles_n) * BN;
  const bool has_bias = ly.b_off >= 0 && n0 == 0;
  const int num_kb = (p.batch + 63) / 64;
  const int kind = p.opt.kind;
  const bool use_s0 = kind != DK_OPT_SGD;
  const bool use_s1 = kind == DK_OPT_ADAM || kind == DK_OPT_ADADELTA || kind == DK_OPT_ADAMAX || kind == DK_OPT_NADAM;
  const bool use_w1 = p.comm_mode == DK_COMM_EXCHANGE;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ly.ta);
    tma_prefetch_desc(&ly.tb);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(state_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 2 && warp < 6) {  // 128 threads x 16 B = the ones tile
    const int t = threadIdx.x - 64;
    st_shared_v4(smem_u32(ones) + t * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_stamp(tr, 1);
  DK_PDL_WAIT();
  DK_PDL_TRIGGER();
  if (threadIdx.x == 0) trace_stamp(tr, 2);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
#pragma unroll
        for (int c = 0; c < kBwdBlockM / 64; ++c) tma_load_2d(sa + c * 8192, &ly.ta, m0 + c * 64, kb * 64, &full_bar[stage]);
#pragma unroll
        for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &ly.tb, n0 + c * 64, kb * 64, &full_bar[stage]);
        if (kb == 0 && ly.maps != nullptr) {
          // the parameter state of this tile: in flight while the gradient GEMM runs
          const int narr = 1 + (use_s0 ? 1 : 0) + (use_s1 ? 1 : 0) + (use_w1 ? 1 : 0);
          mbar_expect_tx(state_bar, narr * kBwdTileBytes);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const bool want = a == kStW || (a == kStS0 && use_s0) || (a == kStS1 && use_s1) || (a == kStW1 && use_w1);
            if (!want) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              tma_load_2d(tiles + a * kBwdTileBytes + h * (kBwdBlockM * 128), &ly.maps->ld[a], n0 + 32 * h, m0, state_bar);
          }
        }
        if (++stage == kBwdStages) {
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
      if (kb == 0 && lane == 0) trace_stamp(tr, 3);
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
        const uint64_t adesc = make_smem_desc_sw128_mn(sa);
        const uint64_t bdesc = make_smem_desc_sw128_mn(sa + kABytes);
        const uint64_t odesc = make_smem_desc_sw128(smem_u32(ones));
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 64 batch rows per k-block, 16 per MMA: two 1024-byte atoms each
          umma_f16(tmem_base, adesc + (2048 >> 4) * k, bdesc + (2048 >> 4) * k, kIdesc, (kb | k) != 0);
          if (has_bias) umma_f16(tmem_base + BN, adesc + (2048 >> 4) * k, odesc + 2 * k, kIdescBias, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
      if (++stage == kBwdStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue: optimizer (+ exchange) ------------------------------
    const int e = warp - 2;            // 0..7
    const int quarter = warp & 3;      // TMEM lane quarter this warp may read
    const int half = e >> 2;           // which 32 of the 64 accumulator columns
    const uint32_t gslot = smem_u32(smem) + e * 4096;   // pipeline buffers are idle once the accumulator is complete
    const uint32_t tl = smem_u32(tiles), wbt = smem_u32(wb_tile);
    // everything that only depends on the previous step is fetched while the operands / state tiles are in
    // flight: step number -> learning rate and bias correction, DynSGD scale, this lane's bias parameter
    const int t = max(*p.step, 1);
    float lr;
    OptimCorr corr;
    optim_prelude(p.opt, t, lr, corr);
    const float cs = p.comm_mode != DK_COMM_NONE
                         ? (p.scale_dev != nullptr ? p.comm_scale * __ldg(p.scale_dev) : p.comm_scale) : 0.f;
    float bw = 0.f, bs0 = 0.f, bs1 = 0.f;
    if (has_bias && half == 0 && m0 + quarter * 32 + lane < ly.n_out) {
      const long idx = ly.b_off + m0 + quarter * 32 + lane;
      bw = p.w[idx];
      if (use_s0) bs0 = p.s0[idx];
      if (use_s1) bs1 = p.s1[idx];
    }
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    if (warp == 2 && lane == 0) trace_stamp(tr, 4);
    switch (kind) {
      case DK_OPT_SGD: bwd_epilogue<DK_OPT_SGD>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_MOMENTUM: bwd_epilogue<DK_OPT_MOMENTUM>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_ADAGRAD: bwd_epilogue<DK_OPT_ADAGRAD>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_RMSPROP: bwd_epilogue<DK_OPT_RMSPROP>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_ADAM: bwd_epilogue<DK_OPT_ADAM>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_ADADELTA: bwd_epilogue<DK_OPT_ADADELTA>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      case DK_OPT_NADAM: bwd_epilogue<DK_OPT_NADAM>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
      default: bwd_epilogue<DK_OPT_ADAMAX>(p, ly, m0, n0, quarter, half, lane, tmem_base, gslot, tl, wbt, state_bar, has_bias, lr, corr, cs, bw, bs0, bs1); break;
    }
    tcgen05_fence_before();
    if (warp == 2 && lane == 0) trace_stamp(tr, 6);
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) trace_stamp(tr, 7);
  if (threadIdx.x == 0) {
    // last CTA of the grid: advance the step counter (every CTA read it before arriving here) and
    // publish the commit in the parameter server's control block
    __threadfence();
    const unsigned done = atomicAdd(p.done_counter, 1u);
    if (done == gridDim.x - 1) {
      *p.done_counter = 0u;
      if (p.step_inc) *p.step += 1;
      if (p.comm_mode != DK_COMM_NONE && p.ctrl != nullptr) {
        if (p.comm_mode == DK_COMM_EXCHANGE) {
          unsigned v = 0;
          if (p.scale_dev == nullptr)
            asm volatile("atom.relaxed.sys.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(p.ctrl + DK_CTRL_NUM_UPDATES) : "memory");
          else  // DynSGD: the ticket kernel already counted this commit
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ctrl + DK_CTRL_NUM_UPDATES) : "memory");
          if (p.last_update != nullptr) *p.last_update = p.scale_dev == nullptr ? v + 1 : v;
        } else {
          asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.ctrl + DK_CTRL_NUM_UPDATES) : "memory");
        }
        asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.ctrl + DK_CTRL_HEARTBEAT + p.worker) : "memory");
      }
    }
  }
}

constexpr int kBwdSmem = kBwdSmemUse + 1024;  // + alignment slack

}  // namespace dk

using namespace dk;

extern "C" {

long dk_bwd_update_record_bytes() { return static_cast<long>(sizeof(BwdRecord)); }

long dk_bwd_update_desc_bytes() { return static_cast<long>(sizeof(DkBwdUpdateDesc)); }

int dk_bwd_update_prepare(void* record, const DkBwdUpdateDesc* d) {
  if (d->nlayers < 1 || d->nlayers > DK_BWD_MAX_LAYERS || d->nshards < 1 || d->nshards > DK_BWD_MAX_SHARDS) return -1;
  if (d->batch < 1 || d->w == nullptr || d->wb == nullptr || d->step == nullptr || d->done_counter == nullptr) return -2;
  if (d->comm_mode == DK_COMM_EXCHANGE && d->w1 == nullptr) return -3;
  BwdRecord* rec = new (record) BwdRecord();
  BwdUpdateDev& p = rec->dev;
  memset(&p, 0, sizeof(p));
  int tiles = 0;
  for (int l = 0; l < d->nlayers; ++l) {
    const DkBwdLayerDesc& s = d->layer[l];
    BwdLayerDev& ly = p.layer[l];
    int r = dk_tmap_encode_2d(&ly.ta, s.dz, DK_BF16, d->batch, s.n_out, s.lddz, 64);
    if (r != 0) return r;
    rec->x[l] = s.x;
    rec->ldx[l] = s.ldx;
    if (s.x != nullptr) {
      r = dk_tmap_encode_2d(&ly.tb, s.x, DK_BF16, d->batch, s.k_in, s.ldx, 64);
      if (r != 0) return r;
    }
    ly.maps = nullptr;
    ly.w_off = s.w_off;
    ly.b_off = s.b_off;
    ly.wb_pad = reinterpret_cast<__nv_bfloat16*>(s.wb_pad);
    ly.ldwb_pad = static_cast<int>(s.ldwb_pad);
    ly.n_out = s.n_out;
    ly.k_in = s.k_in;
    ly.tiles_n = (s.k_in + kBwdBN - 1) / kBwdBN;
    ly.tile_begin = tiles;
    ly.vec = (s.w_off % 4 == 0 && s.k_in % 4 == 0 && (s.wb_pad == nullptr || s.ldwb_pad % 4 == 0)) ? 1 : 0;
    tiles += ((s.n_out + kBwdBlockM - 1) / kBwdBlockM) * ly.tiles_n;
  }
  rec->total_tiles = tiles;
  // parameter-state tensor maps (device memory): layers whose rows are 16-byte aligned take the TMA-fed path
  rec->maps_dev = nullptr;
  {
    StateMaps host[DK_BWD_MAX_LAYERS];
    memset(host, 0, sizeof(host));
    bool any = false;
    for (int l = 0; l < d->nlayers; ++l) {
      const DkBwdLayerDesc& sd = d->layer[l];
      BwdLayerDev& ly = p.layer[l];
      // fp32 rows must be 16-byte aligned (k_in % 4 == 0).  The bf16 shadow tile leaves by TMA store into the flat
      // shadow when its rows are 16-byte aligned too (k_in % 8 == 0), otherwise into the 8-padded shadow the GEMMs
      // read (the flat copy is then written with plain 8-byte stores by the kernel).
      const bool flat_rows_ok = sd.k_in % 8 == 0 && sd.wb_pad == nullptr;
      const bool pad_rows_ok = sd.k_in % 8 != 0 && sd.wb_pad != nullptr && sd.ldwb_pad % 8 == 0;
      if (!ly.vec || sd.k_in % 4 != 0 || sd.w_off % 8 != 0 || !(flat_rows_ok || pad_rows_ok)) continue;
      float* arrs[4] = {d->w, d->s0, d->s1, d->w1};
      bool ok = true;
      for (int a = 0; a < 4 && ok; ++a) {
        if (arrs[a] == nullptr) continue;
        ok = dk_tmap_encode_2d(&host[l].ld[a], arrs[a] + sd.w_off, DK_F32, sd.n_out, sd.k_in, sd.k_in, kBwdBlockM) == 0 &&
             dk_tmap_encode_2d(&host[l].st[a], arrs[a] + sd.w_off, DK_F32, sd.n_out, sd.k_in, sd.k_in, 32) == 0;
      }
      ok = ok && (flat_rows_ok
                      ? dk_tmap_encode_2d(&host[l].st[kStWb], reinterpret_cast<__nv_bfloat16*>(d->wb) + sd.w_off, DK_BF16,
                                          sd.n_out, sd.k_in, sd.k_in, 32)
                      : dk_tmap_encode_2d(&host[l].st[kStWb], sd.wb_pad, DK_BF16, sd.n_out, sd.k_in, sd.ldwb_pad, 32)) == 0;
      ok = ok && (flat_rows_ok
                      ? dk_tmap_encode_2d(&host[l].wb128, reinterpret_cast<__nv_bfloat16*>(d->wb) + sd.w_off, DK_BF16,
                                          sd.n_out, sd.k_in, sd.k_in, kBwdBlockM)
                      : dk_tmap_encode_2d(&host[l].wb128, sd.wb_pad, DK_BF16, sd.n_out, sd.k_in, sd.ldwb_pad, kBwdBlockM)) == 0;
      if (ok) {
        any = true;
        ly.maps = reinterpret_cast<const StateMaps*>(static_cast<uintptr_t>(l + 1));  // patched to the device address below
      }
    }
    if (any) {
      DK_HOST_CHECK(cudaMalloc(reinterpret_cast<void**>(&rec->maps_dev), sizeof(StateMaps) * DK_BWD_MAX_LAYERS));
      DK_HOST_CHECK(cudaMemcpy(rec->maps_dev, host, sizeof(StateMaps) * DK_BWD_MAX_LAYERS, cudaMemcpyHostToDevice));
      for (int l = 0; l < d->nlayers; ++l)
        if (p.layer[l].maps != nullptr) p.layer[l].maps = rec->maps_dev + l;
    }
  }
  for (int s = 0; s < d->nshards; ++s) p.shard_center[s] = d->shard_center[s];
  p.w = d->w; p.s0 = d->s0; p.s1 = d->s1; p.w1 = d->w1;
  p.wb = reinterpret_cast<__nv_bfloat16*>(d->wb);
  p.step = d->step; p.done_counter = d->done_counter; p.scale_dev = d->scale_dev;
  p.ctrl = d->ctrl; p.last_update = d->last_update; p.trace = d->trace; p.shard_per = d->shard_per > 0 ? d->shard_per : 1;
  memset(&p.opt, 0, sizeof(p.opt));
  p.opt.kind = d->opt_kind; p.opt.lr = d->lr; p.opt.p0 = d->p0; p.opt.p1 = d->p1; p.opt.eps = d->eps;
  p.opt.decay = d->decay; p.opt.nesterov = d->nesterov; p.opt.grad_scale = 1.f;
  p.nlayers = d->nlayers; p.batch = d->batch; p.step_inc = d->step_inc; p.comm_mode = d->comm_mode;
  p.nshards = d->nshards; p.worker = d->worker; p.comm_scale = d->comm_scale; p.alpha = d->alpha;
  return 0;
}

void dk_bwd_update_release(void* record) {
  BwdRecord* rec = reinterpret_cast<BwdRecord*>(record);
  if (rec->maps_dev != nullptr) cudaFree(rec->maps_dev);
  rec->maps_dev = nullptr;
}

int dk_bwd_update_set_input(void* record, int layer, const void* x) {
  BwdRecord* rec = reinterpret_cast<BwdRecord*>(record);
  if (layer < 0 || layer >= rec->dev.nlayers) return -1;
  rec->x[layer] = x;
  return dk_tmap_encode_2d(&rec->dev.layer[layer].tb, x, DK_BF16, rec->dev.batch, rec->dev.layer[layer].k_in, rec->ldx[layer],
                           64);
}

int dk_bwd_update_launch(const void* record, void* stream) {
  const BwdRecord* rec = reinterpret_cast<const BwdRecord*>(record);
  static bool configured[64] = {};
  int dev = 0;
  DK_HOST_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    DK_HOST_CHECK(cudaFuncSetAttribute(dense_bwd_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    configured[dev & 63] = true;
  }
  for (int l = 0; l < rec->dev.nlayers; ++l)
    if (rec->x[l] == nullptr) return -4;
  DK_HOST_CHECK(DK_LAUNCH(dense_bwd_update_kernel, rec->total_tiles, kBwdThreads, kBwdSmem,
                          reinterpret_cast<cudaStream_t>(stream), rec->dev));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_bwd_update(const DkBwdUpdateDesc* desc, void* stream) {
  alignas(64) static thread_local unsigned char storage[sizeof(BwdRecord) + 64];
  void* rec = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(storage) + 63) & ~static_cast<uintptr_t>(63));
  int r = dk_bwd_update_prepare(rec, desc);
  if (r != 0) return r;
  r = dk_bwd_update_launch(rec, stream);
  cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));  // one-shot helper: the maps die with this call
  dk_bwd_update_release(rec);
  return r;
}

}  // extern "C"
