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
constexpr int kBwdThreads = 192;
constexpr int kBwdStages = 3;

struct BwdLayerDev {
  alignas(64) CUtensorMap ta;  // dZ [batch, n_out] as an MN-major A operand (64 x 64 boxes)
  alignas(64) CUtensorMap tb;  // X  [batch, k_in]  as an MN-major B operand
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
  long shard_per;
  OptimArgs opt;
  int nlayers, batch, step_inc, comm_mode, nshards, worker;
  float comm_scale, alpha;
};

struct BwdRecord {
  BwdUpdateDev dev;
  const void* x[DK_BWD_MAX_LAYERS];
  long ldx[DK_BWD_MAX_LAYERS];
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
__device__ __forceinline__ void update_scalar(const BwdUpdateDev& p, long idx, float g, float lr, float corr,
                                              bool write_shadow) {
  float w = p.w[idx];
  float s0 = p.s0 != nullptr ? p.s0[idx] : 0.f;
  float s1 = p.s1 != nullptr ? p.s1[idx] : 0.f;
  optim_update_rt(p.opt.kind, w, g, s0, s1, lr, p.opt, corr);
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
  uint8_t* ones = smem + kBwdStages * kStageBytes;  // 16 rows x 128 B of bf16 1.0 (layout-invariant)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + 2048);
  uint64_t* empty_bar = full_bar + kBwdStages;
  uint64_t* tmem_full_bar = empty_bar + kBwdStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ly.ta);
    tma_prefetch_desc(&ly.tb);
    for (int s = 0; s < kBwdStages; ++s) {
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
  if (warp >= 2) {  // 128 threads x 16 B = the ones tile
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
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;  // output unit (row of the kernel) owned by this thread
    const bool row_ok = m < ly.n_out;
    const int t = max(*p.step, 1);
    float lr, corr;
    optim_prelude(p.opt, t, lr, corr);
    const bool comm = p.comm_mode != DK_COMM_NONE;
    const float cs = comm ? (p.scale_dev != nullptr ? p.comm_scale * __ldg(p.scale_dev) : p.comm_scale) : 0.f;
    const bool has_s0 = p.s0 != nullptr, has_s1 = p.s1 != nullptr;
    const long row_idx = ly.w_off + static_cast<long>(m) * ly.k_in;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    // ---- pass 1: gradient (TMEM) -> optimizer rule -> W / state (/ shadow) ----
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      const int n = n0 + c;
      if (n >= ly.k_in) break;  // warp-uniform
      uint32_t r[16];
      tmem_ld_32x32b_x16(trow + c, r);
      tmem_ld_wait();
      if (!row_ok) continue;
      const long idx = row_idx + n;
      if (ly.vec && n + 16 <= ly.k_in) {
        float4 wv[4], av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          wv[q] = *reinterpret_cast<const float4*>(p.w + idx + 4 * q);
          av[q] = has_s0 ? *reinterpret_cast<const float4*>(p.s0 + idx + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
          bv[q] = has_s1 ? *reinterpret_cast<const float4*>(p.s1 + idx + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          optim_update_rt(p.opt.kind, wv[q].x, __uint_as_float(r[4 * q]), av[q].x, bv[q].x, lr, p.opt, corr);
          optim_update_rt(p.opt.kind, wv[q].y, __uint_as_float(r[4 * q + 1]), av[q].y, bv[q].y, lr, p.opt, corr);
          optim_update_rt(p.opt.kind, wv[q].z, __uint_as_float(r[4 * q + 2]), av[q].z, bv[q].z, lr, p.opt, corr);
          optim_update_rt(p.opt.kind, wv[q].w, __uint_as_float(r[4 * q + 3]), av[q].w, bv[q].w, lr, p.opt, corr);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float4*>(p.w + idx + 4 * q) = wv[q];
          if (has_s0) *reinterpret_cast<float4*>(p.s0 + idx + 4 * q) = av[q];
          if (has_s1) *reinterpret_cast<float4*>(p.s1 + idx + 4 * q) = bv[q];
        }
        if (!comm) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint2 o;
            o.x = pack_bf16x2(wv[q].x, wv[q].y);
            o.y = pack_bf16x2(wv[q].z, wv[q].w);
            *reinterpret_cast<uint2*>(p.wb + idx + 4 * q) = o;
            if (ly.wb_pad != nullptr)
              *reinterpret_cast<uint2*>(ly.wb_pad + static_cast<long>(m) * ly.ldwb_pad + n + 4 * q) = o;
          }
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {
          if (n + j >= ly.k_in) break;
          update_scalar(p, idx + j, __uint_as_float(r[j]), lr, corr, !comm);
          if (!comm && ly.wb_pad != nullptr)
            ly.wb_pad[static_cast<long>(m) * ly.ldwb_pad + n + j] = __float2bfloat16_rn(p.w[idx + j]);
        }
      }
    }
    if (has_bias) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(trow + BN, r);
      tmem_ld_wait();
      if (row_ok) {
        update_scalar(p, ly.b_off + m, __uint_as_float(r[0]), lr, corr, !comm);
        if (comm) exchange_scalar(p, ly.b_off + m, cs);
      }
    }
    // ---- pass 2 (window boundary): push the window's displacement, adopt the center ----
    if (comm && row_ok) {
      const int ncols = min(BN, ly.k_in - n0);
      const long idx0 = row_idx + n0;
      if (ly.vec && ncols == BN) {
        // all 16 float4 of this row slice: every NVLink request is issued before the first result is used
        float4 x[BN / 4], y[BN / 4];
        if (p.comm_mode == DK_COMM_EXCHANGE) {
#pragma unroll
          for (int q = 0; q < BN / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(p.w + idx0 + 4 * q);
            const float4 b = *reinterpret_cast<const float4*>(p.w1 + idx0 + 4 * q);
            x[q] = make_float4((a.x - b.x) * cs, (a.y - b.y) * cs, (a.z - b.z) * cs, (a.w - b.w) * cs);
          }
#pragma unroll
          for (int q = 0; q < BN / 4; ++q) y[q] = atom_add_v4_sys_f(center_of(p, idx0 + 4 * q), x[q]);
#pragma unroll
          for (int q = 0; q < BN / 4; ++q) {
            y[q].x += x[q].x; y[q].y += x[q].y; y[q].z += x[q].z; y[q].w += x[q].w;
            *reinterpret_cast<float4*>(p.w1 + idx0 + 4 * q) = y[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < BN / 4; ++q) x[q] = ld_sys_v4_f(center_of(p, idx0 + 4 * q));
#pragma unroll
          for (int q = 0; q < BN / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(p.w + idx0 + 4 * q);
            const float4 e = make_float4(p.alpha * (a.x - x[q].x), p.alpha * (a.y - x[q].y), p.alpha * (a.z - x[q].z),
                                         p.alpha * (a.w - x[q].w));
            y[q] = make_float4(a.x - e.x, a.y - e.y, a.z - e.z, a.w - e.w);
            red_add_v4_sys_f(center_of(p, idx0 + 4 * q), e);
          }
        }
#pragma unroll
        for (int q = 0; q < BN / 4; ++q) {
          *reinterpret_cast<float4*>(p.w + idx0 + 4 * q) = y[q];
          uint2 o;
          o.x = pack_bf16x2(y[q].x, y[q].y);
          o.y = pack_bf16x2(y[q].z, y[q].w);
          *reinterpret_cast<uint2*>(p.wb + idx0 + 4 * q) = o;
          if (ly.wb_pad != nullptr)
            *reinterpret_cast<uint2*>(ly.wb_pad + static_cast<long>(m) * ly.ldwb_pad + n0 + 4 * q) = o;
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < ncols; ++j) {
          const float wn = exchange_scalar(p, idx0 + j, cs);
          if (ly.wb_pad != nullptr) ly.wb_pad[static_cast<long>(m) * ly.ldwb_pad + n0 + j] = __float2bfloat16_rn(wn);
        }
      }
    }
    tcgen05_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
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

constexpr int kBwdSmem = kBwdStages * (kBwdBlockM * 128 + kBwdBN * 128) + 2048 + 256 + 1024;

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
  for (int s = 0; s < d->nshards; ++s) p.shard_center[s] = d->shard_center[s];
  p.w = d->w; p.s0 = d->s0; p.s1 = d->s1; p.w1 = d->w1;
  p.wb = reinterpret_cast<__nv_bfloat16*>(d->wb);
  p.step = d->step; p.done_counter = d->done_counter; p.scale_dev = d->scale_dev;
  p.ctrl = d->ctrl; p.last_update = d->last_update; p.shard_per = d->shard_per > 0 ? d->shard_per : 1;
  memset(&p.opt, 0, sizeof(p.opt));
  p.opt.kind = d->opt_kind; p.opt.lr = d->lr; p.opt.p0 = d->p0; p.opt.p1 = d->p1; p.opt.eps = d->eps;
  p.opt.decay = d->decay; p.opt.nesterov = d->nesterov; p.opt.grad_scale = 1.f;
  p.nlayers = d->nlayers; p.batch = d->batch; p.step_inc = d->step_inc; p.comm_mode = d->comm_mode;
  p.nshards = d->nshards; p.worker = d->worker; p.comm_scale = d->comm_scale; p.alpha = d->alpha;
  return 0;
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
  return dk_bwd_update_launch(rec, stream);
}

}  // extern "C"
