// Parameter-server kernels over peer-mapped HBM (NVLink 5 / NVSwitch).
//
// The center variable is one flat fp32 buffer living in the PS GPU's HBM and mapped into every
// worker's address space (CUDA IPC).  The reference's commit / pull protocol
// (distkeras/workers.py:224-240, parameter_servers.py:232-255: pickle + TCP + mutex + numpy add) is
// replaced by kernels that issue system-scope reductions / loads straight to that buffer:
//
//   commit    C += s * (W - W1)                    red.global.add.v4.f32 (.sys)     [K1]
//   pull      W = W1 = C (+ bf16 shadow)           ld.global (peer)                 [K2, copy form]
//   exchange  commit + pull in ONE NVLink round trip via fetching atomics:
//             old = atom.add(C, r); W = W1 = old + r                                [K1+K2]
//   elastic   E = a (W - C); W -= E; C += E        (AEASGD / EAMSGD, workers.py:402-407)  [K3]
//   damped    C += r / (inv_lr (C - C_stale)^2 + 1)  (Experimental PS, parameter_servers.py:372-386)
//   ticket    staleness counter / num_updates (DynSGD, parameter_servers.py:336-354)       [K5]
//   lock      ticket lock for the "strict" (serialised, reference-faithful) mode
//   average   in-place mean over P replicas through peer loads (AveragingTrainer)          [K13]
#include "common.cuh"
#include "ps.h"

namespace dk {

constexpr int kPsThreads = 256;
constexpr int kPsUnroll = 4;

__device__ __forceinline__ void red_add_v4_sys(float* addr, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void red_add_sys(float* addr, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ float4 atom_add_v4_sys(float* addr, float4 v) {
  float4 o;
  asm volatile("atom.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4], {%5, %6, %7, %8};"
               : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
               : "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
  return o;
}

__device__ __forceinline__ float atom_add_sys(float* addr, float v) {
  float o;
  asm volatile("atom.relaxed.sys.global.add.f32 %0, [%1], %2;" : "=f"(o) : "l"(addr), "f"(v) : "memory");
  return o;
}

// Peer loads: relaxed system-scope so they are never served from a stale local L1 line.
__device__ __forceinline__ float4 ld_sys_v4(const float* addr) {
  float4 o;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
               : "l"(addr)
               : "memory");
  return o;
}

__device__ __forceinline__ float ld_sys(const float* addr) {
  float o;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(o) : "l"(addr) : "memory");
  return o;
}

__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v) {
  uint2 o;
  o.x = pack_bf16x2(v.x, v.y);
  o.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = o;
}

// Flat traversal in three phases per thread: kPsUnroll float4 slots each.  Phase 0 issues every local
// load (prep), phase 1 issues every remote operation (NVLink load / atomic) and only phase 2 consumes
// the results, so a thread keeps kPsUnroll independent 16-byte NVLink requests in flight instead of
// serialising kPsUnroll round trips (a 1 M-parameter exchange is latency-bound: 4 MB is ~4.4 us at link
// rate, one round trip is ~2 us).  A scalar tail covers n % 4.
struct NoRegs {};

template <typename Prep, typename Issue, typename Finish, typename ScalarOp>
__device__ __forceinline__ void flat_pipeline(long n, Prep prep, Issue issue, Finish finish, ScalarOp scalar_op) {
  const long n4 = n >> 2;
  const long tile = static_cast<long>(blockDim.x) * kPsUnroll;
  for (long base = static_cast<long>(blockIdx.x) * tile + threadIdx.x; base < n4;
       base += static_cast<long>(gridDim.x) * tile) {
    decltype(prep(0L)) a[kPsUnroll];
    decltype(issue(0L, a[0])) b[kPsUnroll];
#pragma unroll
    for (int u = 0; u < kPsUnroll; ++u) {
      const long i = base + static_cast<long>(u) * blockDim.x;
      if (i < n4) a[u] = prep(i << 2);
    }
#pragma unroll
    for (int u = 0; u < kPsUnroll; ++u) {
      const long i = base + static_cast<long>(u) * blockDim.x;
      if (i < n4) b[u] = issue(i << 2, a[u]);
    }
#pragma unroll
    for (int u = 0; u < kPsUnroll; ++u) {
      const long i = base + static_cast<long>(u) * blockDim.x;
      if (i < n4) finish(i << 2, a[u], b[u]);
    }
  }
  const long tail0 = n4 << 2;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = tail0 + static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    scalar_op(i);
}

__device__ __forceinline__ float load_scale(const float* scale_dev, float scale) {
  return scale_dev != nullptr ? scale * __ldg(scale_dev) : scale;
}

__device__ __forceinline__ float4 residual4(const float* w, const float* w1, long i, float s) {
  const float4 a = *reinterpret_cast<const float4*>(w + i);
  const float4 b = *reinterpret_cast<const float4*>(w1 + i);
  return make_float4((a.x - b.x) * s, (a.y - b.y) * s, (a.z - b.z) * s, (a.w - b.w) * s);
}

// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPsThreads)
ps_commit_kernel(float* __restrict__ center, const float* __restrict__ w, const float* __restrict__ w1,
                 long n, float scale, const float* __restrict__ scale_dev, unsigned* ctrl, int worker,
                 unsigned iteration) {
  DK_PDL_ENTER();
  const float s = load_scale(scale_dev, scale);
  flat_pipeline(
      n, [&](long i) { return residual4(w, w1, i, s); },
      [&](long i, const float4& r) {
        red_add_v4_sys(center + i, r);
        return NoRegs{};
      },
      [&](long, const float4&, const NoRegs&) {},
      [&](long i) { red_add_sys(center + i, (w[i] - w1[i]) * s); });
  if (ctrl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    if (scale_dev == nullptr)  // DynSGD already bumped the counter in its ticket kernel
      asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_HEARTBEAT + worker) : "memory");  // commits by this worker (graph-replay safe liveness counter)
  }
}

__global__ void __launch_bounds__(kPsThreads)
ps_pull_kernel(const float* __restrict__ center, float* __restrict__ w, float* __restrict__ w1,
               __nv_bfloat16* __restrict__ wb, long n, const unsigned* ctrl, unsigned* last_update) {
  DK_PDL_ENTER();
  flat_pipeline(
      n, [&](long) { return NoRegs{}; }, [&](long i, const NoRegs&) { return ld_sys_v4(center + i); },
      [&](long i, const NoRegs&, const float4& c) {
        *reinterpret_cast<float4*>(w + i) = c;
        if (w1 != nullptr) *reinterpret_cast<float4*>(w1 + i) = c;
        if (wb != nullptr) st_bf16x4(wb + i, c);
      },
      [&](long i) {
        const float c = ld_sys(center + i);
        w[i] = c;
        if (w1 != nullptr) w1[i] = c;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(c);
      });
  if (ctrl != nullptr && last_update != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    *last_update = v;
  }
}

// commit + pull fused through returning atomics: one NVLink round trip per 16 bytes.
__global__ void __launch_bounds__(kPsThreads)
ps_exchange_kernel(float* __restrict__ center, float* __restrict__ w, float* __restrict__ w1,
                   __nv_bfloat16* __restrict__ wb, long n, float scale,
                   const float* __restrict__ scale_dev, unsigned* ctrl, int worker,
                   unsigned iteration, unsigned* last_update) {
  DK_PDL_ENTER();
  const float s = load_scale(scale_dev, scale);
  flat_pipeline(
      n, [&](long i) { return residual4(w, w1, i, s); },
      [&](long i, const float4& r) { return atom_add_v4_sys(center + i, r); },
      [&](long i, const float4& r, const float4& o) {
        const float4 c = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
        *reinterpret_cast<float4*>(w + i) = c;
        *reinterpret_cast<float4*>(w1 + i) = c;
        if (wb != nullptr) st_bf16x4(wb + i, c);
      },
      [&](long i) {
        const float r = (w[i] - w1[i]) * s;
        const float c = atom_add_sys(center + i, r) + r;
        w[i] = c;
        w1[i] = c;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(c);
      });
  if (ctrl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned v = 0;
    if (scale_dev == nullptr)
      asm volatile("atom.relaxed.sys.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    else
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    if (last_update != nullptr) *last_update = scale_dev == nullptr ? v + 1 : v;
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_HEARTBEAT + worker) : "memory");  // commits by this worker (graph-replay safe liveness counter)
  }
}

// Elastic averaging step (AEASGD / EAMSGD): pull + local move + commit in one pass.
__global__ void __launch_bounds__(kPsThreads)
ps_elastic_kernel(float* __restrict__ center, float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                  long n, float alpha, unsigned* ctrl, int worker, unsigned iteration) {
  DK_PDL_ENTER();
  flat_pipeline(
      n, [&](long i) { return *reinterpret_cast<const float4*>(w + i); },
      [&](long i, const float4&) { return ld_sys_v4(center + i); },
      [&](long i, const float4& x0, const float4& c) {
        float4 x = x0;
        const float4 e = make_float4(alpha * (x.x - c.x), alpha * (x.y - c.y), alpha * (x.z - c.z),
                                     alpha * (x.w - c.w));
        x.x -= e.x; x.y -= e.y; x.z -= e.z; x.w -= e.w;
        *reinterpret_cast<float4*>(w + i) = x;
        if (wb != nullptr) st_bf16x4(wb + i, x);
        red_add_v4_sys(center + i, e);
      },
      [&](long i) {
        const float c = ld_sys(center + i);
        const float e = alpha * (w[i] - c);
        const float x = w[i] - e;
        w[i] = x;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(x);
        red_add_sys(center + i, e);
      });
  if (ctrl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_HEARTBEAT + worker) : "memory");  // commits by this worker (graph-replay safe liveness counter)
  }
}

// Experimental PS: per-element staleness damping, then the worker adopts the new center.
struct Float4x2 {
  float4 a, b;
};

__global__ void __launch_bounds__(kPsThreads)
ps_damped_exchange_kernel(float* __restrict__ center, float* __restrict__ w, float* __restrict__ w1,
                          __nv_bfloat16* __restrict__ wb, long n, float scale, float inv_lr,
                          unsigned* ctrl, int worker, unsigned iteration) {
  DK_PDL_ENTER();
  // w1 doubles as the stale center variable (the worker's last pulled copy, workers.py:553-563).
  // Two NVLink round trips per element are inherent here (the damping needs the CURRENT center before
  // the update can be formed); the kPsUnroll slots of a thread overlap theirs.
  flat_pipeline(
      n,
      [&](long i) {
        Float4x2 t;
        t.a = *reinterpret_cast<const float4*>(w + i);
        t.b = *reinterpret_cast<const float4*>(w1 + i);
        return t;
      },
      [&](long i, const Float4x2&) { return ld_sys_v4(center + i); },
      [&](long i, const Float4x2& t, const float4& c) {
        const float4 a = t.a, b = t.b;
        float4 r = make_float4((a.x - b.x) * scale, (a.y - b.y) * scale, (a.z - b.z) * scale,
                               (a.w - b.w) * scale);
        const float dx = c.x - b.x, dy = c.y - b.y, dz = c.z - b.z, dw = c.w - b.w;
        r.x = __fdividef(r.x, inv_lr * dx * dx + 1.f);
        r.y = __fdividef(r.y, inv_lr * dy * dy + 1.f);
        r.z = __fdividef(r.z, inv_lr * dz * dz + 1.f);
        r.w = __fdividef(r.w, inv_lr * dw * dw + 1.f);
        const float4 o = atom_add_v4_sys(center + i, r);
        const float4 nc = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
        *reinterpret_cast<float4*>(w + i) = nc;
        *reinterpret_cast<float4*>(w1 + i) = nc;
        if (wb != nullptr) st_bf16x4(wb + i, nc);
      },
      [&](long i) {
        const float c = ld_sys(center + i);
        float r = (w[i] - w1[i]) * scale;
        const float d = c - w1[i];
        r = __fdividef(r, inv_lr * d * d + 1.f);
        const float nc = atom_add_sys(center + i, r) + r;
        w[i] = nc;
        w1[i] = nc;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(nc);
      });
  if (ctrl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_HEARTBEAT + worker) : "memory");  // commits by this worker (graph-replay safe liveness counter)
  }
}

// DynSGD ticket: num_updates += 1; scale = 1 / (num_updates_before - last_update + 1).
__global__ void ps_ticket_kernel(unsigned* ctrl, const unsigned* last_update, float* scale_out) {
  DK_PDL_ENTER();
  unsigned old;
  asm volatile("atom.relaxed.sys.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
  const unsigned last = *last_update;
  const unsigned staleness = (old >= last ? old - last : 0u) + 1u;
  *scale_out = 1.f / static_cast<float>(staleness);
  if (ctrl != nullptr) {
    // staleness histogram (observability): bucket = min(staleness, 31)
    const unsigned b = staleness < 31u ? staleness : 31u;
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_STALENESS_HIST + b) : "memory");
  }
}

// Device-side rendezvous of `workers` ranks (synchronous EASGD): instance r of the barrier completes when the
// monotonic arrival counter reaches (r + 1) * workers.  Release on arrive / acquire on the spin make every rank's
// earlier center updates visible to every rank's later center reads.  A rank that waits longer than the timeout
// (a peer died) or sees the stop flag marks the rendezvous broken and moves on instead of hanging the GPU.
__global__ void ps_barrier_kernel(unsigned* ctrl, unsigned workers, unsigned* round, unsigned* broken,
                                  unsigned long long timeout_ns) {
  DK_PDL_ENTER();
  const unsigned r = *round;
  *round = r + 1u;
  if (*broken != 0u) return;
  const unsigned target = (r + 1u) * workers;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_BARRIER) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned seen, stop;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctrl + DK_CTRL_BARRIER) : "memory");
    if (static_cast<int>(seen - target) >= 0) break;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(stop) : "l"(ctrl + DK_CTRL_STOP) : "memory");
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (stop != 0u || t1 - t0 > timeout_ns) {
      *broken = 1u;
      break;
    }
    __nanosleep(100);
  }
}

// Synchronous EASGD, phase 1: elastic difference against the (quiescent) center; the worker moves, the center is
// only read.  Phase 2 (after everybody has read): the center absorbs the stored difference.
__global__ void __launch_bounds__(kPsThreads)
ps_easgd_read_kernel(const float* __restrict__ center, float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                     float* __restrict__ e_out, long n, float alpha) {
  DK_PDL_ENTER();
  flat_pipeline(
      n, [&](long i) { return *reinterpret_cast<const float4*>(w + i); },
      [&](long i, const float4&) { return ld_sys_v4(center + i); },
      [&](long i, const float4& x0, const float4& c) {
        float4 x = x0;
        const float4 e = make_float4(alpha * (x.x - c.x), alpha * (x.y - c.y), alpha * (x.z - c.z),
                                     alpha * (x.w - c.w));
        x.x -= e.x; x.y -= e.y; x.z -= e.z; x.w -= e.w;
        *reinterpret_cast<float4*>(w + i) = x;
        *reinterpret_cast<float4*>(e_out + i) = e;
        if (wb != nullptr) st_bf16x4(wb + i, x);
      },
      [&](long i) {
        const float c = ld_sys(center + i);
        const float e = alpha * (w[i] - c);
        const float x = w[i] - e;
        w[i] = x;
        e_out[i] = e;
        if (wb != nullptr) wb[i] = __float2bfloat16_rn(x);
      });
}

__global__ void __launch_bounds__(kPsThreads)
ps_easgd_add_kernel(float* __restrict__ center, const float* __restrict__ e, long n, unsigned* ctrl, int worker) {
  DK_PDL_ENTER();
  flat_pipeline(
      n, [&](long i) { return *reinterpret_cast<const float4*>(e + i); },
      [&](long, const float4&) { return NoRegs{}; },
      [&](long i, const float4& v, const NoRegs&) { red_add_v4_sys(center + i, v); },
      [&](long i) { red_add_sys(center + i, e[i]); });
  if (ctrl != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_NUM_UPDATES) : "memory");
    asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(ctrl + DK_CTRL_HEARTBEAT + worker) : "memory");
  }
}

// Host-visible fetch-add on a control word (dynamic shard queue: workers claim the next data
// partition; the replacement for Spark's task scheduler + `parallelism_factor` over-partitioning).
__global__ void ps_fetch_add_kernel(unsigned* word, unsigned inc, unsigned* out) {
  DK_PDL_ENTER();
  unsigned old;
  asm volatile("atom.relaxed.sys.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(word), "r"(inc) : "memory");
  *out = old;
}

// Ticket lock (strict mode): serialises whole commit(+pull) sequences like the reference's mutex.
__global__ void ps_lock_acquire_kernel(unsigned* ctrl, unsigned* my_ticket) {
  DK_PDL_ENTER();
  unsigned t;
  asm volatile("atom.relaxed.sys.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ctrl + DK_CTRL_LOCK_NEXT) : "memory");
  *my_ticket = t;
  unsigned serving;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(serving) : "l"(ctrl + DK_CTRL_LOCK_SERVING) : "memory");
    if (serving != t) {
      __nanosleep(200);
      // the holder died with the lock (rank loss): after 10 s take it over instead of spinning forever -- the
      // release below publishes ticket + 1, which also un-sticks everybody queued behind this waiter
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) break;
    }
  } while (serving != t);
}

__global__ void ps_lock_release_kernel(unsigned* ctrl, const unsigned* my_ticket) {
  DK_PDL_ENTER();
  const unsigned t = *my_ticket + 1u;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ctrl + DK_CTRL_LOCK_SERVING), "r"(t) : "memory");
}

// In-kernel all-reduce-mean over P replicas: this rank reduces slice [lo, hi) by loading the slice
// from every peer, then stores the mean back to every peer (reduce-scatter + all-gather in one
// kernel, no NCCL).  The caller brackets it with barriers.
struct PeerPtrs {
  float* p[DK_MAX_PEERS];
};

__global__ void __launch_bounds__(kPsThreads)
ps_average_kernel(PeerPtrs peers, int num_peers, long lo, long hi, float inv) {
  DK_PDL_ENTER();
  const long n = hi - lo;
  flat_pipeline(
      n, [&](long) { return NoRegs{}; },
      [&](long i, const NoRegs&) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < num_peers; ++r) {
          const float4 v = ld_sys_v4(peers.p[r] + lo + i);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        return acc;
      },
      [&](long i, const NoRegs&, const float4& sum) {
        const float4 acc = make_float4(sum.x * inv, sum.y * inv, sum.z * inv, sum.w * inv);
#pragma unroll 4
        for (int r = 0; r < num_peers; ++r) *reinterpret_cast<float4*>(peers.p[r] + lo + i) = acc;
      },
      [&](long i) {
        float acc = 0.f;
        for (int r = 0; r < num_peers; ++r) acc += ld_sys(peers.p[r] + lo + i);
        acc *= inv;
        for (int r = 0; r < num_peers; ++r) peers.p[r][lo + i] = acc;
      });
}

// Plain device copy/zero helpers used by benchmarks (peer read / write bandwidth).
__global__ void __launch_bounds__(kPsThreads)
ps_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, long n) {
  DK_PDL_ENTER();
  flat_pipeline(
      n, [&](long) { return NoRegs{}; }, [&](long i, const NoRegs&) { return ld_sys_v4(src + i); },
      [&](long i, const NoRegs&, const float4& v) { *reinterpret_cast<float4*>(dst + i) = v; },
      [&](long i) { dst[i] = ld_sys(src + i); });
}

static inline int ps_grid(long n) {
  long blocks = ((n >> 2) + kPsThreads * kPsUnroll - 1) / (kPsThreads * kPsUnroll);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return static_cast<int>(blocks);
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_ps_commit(float* center, const float* w, const float* w1, long n, float scale,
                 const float* scale_dev, unsigned* ctrl, int worker, unsigned iteration, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_commit_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, center, w, w1, n, scale, scale_dev,
                                                                       ctrl, worker, iteration));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_pull(const float* center, float* w, float* w1, void* wb, long n, const unsigned* ctrl,
               unsigned* last_update, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_pull_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, 
      center, w, w1, reinterpret_cast<__nv_bfloat16*>(wb), n, ctrl, last_update));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_exchange(float* center, float* w, float* w1, void* wb, long n, float scale,
                   const float* scale_dev, unsigned* ctrl, int worker, unsigned iteration,
                   unsigned* last_update, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_exchange_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, 
      center, w, w1, reinterpret_cast<__nv_bfloat16*>(wb), n, scale, scale_dev, ctrl, worker, iteration,
      last_update));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_elastic(float* center, float* w, void* wb, long n, float alpha, unsigned* ctrl, int worker,
                  unsigned iteration, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_elastic_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, 
      center, w, reinterpret_cast<__nv_bfloat16*>(wb), n, alpha, ctrl, worker, iteration));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_damped_exchange(float* center, float* w, float* w1, void* wb, long n, float scale,
                          float inv_lr, unsigned* ctrl, int worker, unsigned iteration, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_damped_exchange_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, 
      center, w, w1, reinterpret_cast<__nv_bfloat16*>(wb), n, scale, inv_lr, ctrl, worker, iteration));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_ticket(unsigned* ctrl, const unsigned* last_update, float* scale_out, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_ticket_kernel, 1, 1, 0, (cudaStream_t)stream, ctrl, last_update, scale_out));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_fetch_add(unsigned* word, unsigned inc, unsigned* out, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_fetch_add_kernel, 1, 1, 0, (cudaStream_t)stream, word, inc, out));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_barrier(unsigned* ctrl, int workers, unsigned* round, unsigned* broken, int timeout_ms, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_barrier_kernel, 1, 1, 0, (cudaStream_t)stream, ctrl, static_cast<unsigned>(workers), round,
                          broken, static_cast<unsigned long long>(timeout_ms) * 1000000ull));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_easgd_read(const float* center, float* w, void* wb, float* e, long n, float alpha, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_easgd_read_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, center, w,
                          reinterpret_cast<__nv_bfloat16*>(wb), e, n, alpha));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_easgd_add(float* center, const float* e, long n, unsigned* ctrl, int worker, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_easgd_add_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, center, e, n, ctrl, worker));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_lock_acquire(unsigned* ctrl, unsigned* my_ticket, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_lock_acquire_kernel, 1, 1, 0, (cudaStream_t)stream, ctrl, my_ticket));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_lock_release(unsigned* ctrl, const unsigned* my_ticket, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_lock_release_kernel, 1, 1, 0, (cudaStream_t)stream, ctrl, my_ticket));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_average(float* const* peer_ptrs, int num_peers, long lo, long hi, void* stream) {
  if (num_peers > DK_MAX_PEERS) return -1;
  PeerPtrs pp;
  for (int i = 0; i < num_peers; ++i) pp.p[i] = peer_ptrs[i];
  // slice bounds must keep float4 alignment
  if ((lo & 3) != 0) return -2;
  DK_HOST_CHECK(DK_LAUNCH(ps_average_kernel, ps_grid(hi - lo), kPsThreads, 0, (cudaStream_t)stream, pp, num_peers, lo, hi,
                                                                              1.f / num_peers));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_ps_copy(float* dst, const float* src, long n, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(ps_copy_kernel, ps_grid(n), kPsThreads, 0, (cudaStream_t)stream, dst, src, n));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
