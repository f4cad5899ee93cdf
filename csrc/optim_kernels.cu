// Fused flat-buffer worker optimizers (reference op K12: the Keras optimizer graph that
// `model.train_on_batch` runs, distkeras/workers.py:103-119 compiles it from `worker_optimizer`).
//
// One launch updates EVERY parameter of the replica: the model keeps a single flat fp32 master
// buffer (the `get_weights()` analogue), a flat fp32 gradient buffer and flat optimizer state, and
// the kernel also emits the flat bf16 shadow that feeds the tcgen05 GEMMs.  The step number is
// read from a device counter so the launch can be replayed from a CUDA graph.
#include "common.cuh"
#include "kernels.h"
#include "optim.cuh"

namespace dk {

template <int KIND>
__global__ void __launch_bounds__(256) optim_kernel(const OptimArgs a) {
  DK_PDL_ENTER();
  constexpr bool kS0 = KIND != DK_OPT_SGD;
  constexpr bool kS1 = KIND == DK_OPT_ADAM || KIND == DK_OPT_ADADELTA || KIND == DK_OPT_ADAMAX || KIND == DK_OPT_NADAM;
  const int t = a.step != nullptr ? max(*a.step, 1) : 1;
  float lr;
  OptimCorr corr;
  optim_prelude(a, t, lr, corr);

  const long n4 = a.n >> 2;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 w = reinterpret_cast<float4*>(a.w)[i];
    float4 g = reinterpret_cast<const float4*>(a.g)[i];
    g.x *= a.grad_scale; g.y *= a.grad_scale; g.z *= a.grad_scale; g.w *= a.grad_scale;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if constexpr (kS0) s0 = reinterpret_cast<float4*>(a.s0)[i];
    if constexpr (kS1) s1 = reinterpret_cast<float4*>(a.s1)[i];
    optim_update<KIND>(w.x, g.x, s0.x, s1.x, lr, a, corr);
    optim_update<KIND>(w.y, g.y, s0.y, s1.y, lr, a, corr);
    optim_update<KIND>(w.z, g.z, s0.z, s1.z, lr, a, corr);
    optim_update<KIND>(w.w, g.w, s0.w, s1.w, lr, a, corr);
    reinterpret_cast<float4*>(a.w)[i] = w;
    if constexpr (kS0) reinterpret_cast<float4*>(a.s0)[i] = s0;
    if constexpr (kS1) reinterpret_cast<float4*>(a.s1)[i] = s1;
    if (a.wb != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(w.x, w.y);
      o.y = pack_bf16x2(w.z, w.w);
      reinterpret_cast<uint2*>(a.wb)[i] = o;
    }
  }
  for (long i = (n4 << 2) + static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
       i += stride) {
    float w = a.w[i], g = a.g[i] * a.grad_scale, s0 = 0.f, s1 = 0.f;
    if constexpr (kS0) s0 = a.s0[i];
    if constexpr (kS1) s1 = a.s1[i];
    optim_update<KIND>(w, g, s0, s1, lr, a, corr);
    a.w[i] = w;
    if constexpr (kS0) a.s0[i] = s0;
    if constexpr (kS1) a.s1[i] = s1;
    if (a.wb != nullptr) a.wb[i] = __float2bfloat16_rn(w);
  }
}

// EAMSGD momentum algebra (reference K4, workers.py:447-457) on flat buffers:
//   pre : r_t = mu * r ; W_copy = W ; W += r_t
//   post: g = W_after - (W_copy + r_t) ; r = r_t - eta * g ; W = W_copy - r
__global__ void __launch_bounds__(256)
eamsgd_pre_kernel(float* __restrict__ w, float* __restrict__ r, float* __restrict__ wcopy,
                  __nv_bfloat16* __restrict__ wb, long n, float mu) {
  DK_PDL_ENTER();
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float rt = mu * r[i];
    r[i] = rt;
    const float x = w[i];
    wcopy[i] = x;
    const float y = x + rt;
    w[i] = y;
    if (wb != nullptr) wb[i] = __float2bfloat16_rn(y);
  }
}

__global__ void __launch_bounds__(256)
eamsgd_post_kernel(float* __restrict__ w, float* __restrict__ r, const float* __restrict__ wcopy,
                   __nv_bfloat16* __restrict__ wb, long n, float eta) {
  DK_PDL_ENTER();
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float rt = r[i];
    const float g = w[i] - (wcopy[i] + rt);
    const float rn = rt - eta * g;
    r[i] = rn;
    const float y = wcopy[i] - rn;
    w[i] = y;
    if (wb != nullptr) wb[i] = __float2bfloat16_rn(y);
  }
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long n) {
  DK_PDL_ENTER();
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

static inline int flat_grid(long n, int per_thread) {
  long b = (n / per_thread + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return static_cast<int>(b);
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_optim_step(int kind, float* w, const float* g, float* s0, float* s1, void* wb, long n, float lr,
                  float p0, float p1, float eps, float decay, int nesterov, const int* step,
                  float grad_scale, void* stream) {
  OptimArgs a;
  a.w = w; a.g = g; a.s0 = s0; a.s1 = s1; a.wb = reinterpret_cast<__nv_bfloat16*>(wb); a.n = n;
  a.kind = kind; a.lr = lr; a.p0 = p0; a.p1 = p1; a.eps = eps; a.decay = decay;
  a.nesterov = nesterov; a.step = step; a.grad_scale = grad_scale;
  const int grid = flat_grid(n, 4);
  cudaStream_t st = (cudaStream_t)stream;
  switch (kind) {
    case DK_OPT_SGD: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_SGD>, grid, 256, 0, st, a)); break;
    case DK_OPT_MOMENTUM: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_MOMENTUM>, grid, 256, 0, st, a)); break;
    case DK_OPT_ADAGRAD: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_ADAGRAD>, grid, 256, 0, st, a)); break;
    case DK_OPT_RMSPROP: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_RMSPROP>, grid, 256, 0, st, a)); break;
    case DK_OPT_ADAM: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_ADAM>, grid, 256, 0, st, a)); break;
    case DK_OPT_ADADELTA: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_ADADELTA>, grid, 256, 0, st, a)); break;
    case DK_OPT_ADAMAX: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_ADAMAX>, grid, 256, 0, st, a)); break;
    case DK_OPT_NADAM: DK_HOST_CHECK(DK_LAUNCH(optim_kernel<DK_OPT_NADAM>, grid, 256, 0, st, a)); break;
    default: return -1;
  }
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_eamsgd_pre(float* w, float* r, float* wcopy, void* wb, long n, float mu, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(eamsgd_pre_kernel, flat_grid(n, 1), 256, 0, (cudaStream_t)stream, 
      w, r, wcopy, reinterpret_cast<__nv_bfloat16*>(wb), n, mu));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_eamsgd_post(float* w, float* r, const float* wcopy, void* wb, long n, float eta, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(eamsgd_post_kernel, flat_grid(n, 1), 256, 0, (cudaStream_t)stream, 
      w, r, wcopy, reinterpret_cast<__nv_bfloat16*>(wb), n, eta));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_cast_bf16(const float* src, void* dst, long n, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(cast_bf16_kernel, flat_grid(n, 1), 256, 0, (cudaStream_t)stream, 
      src, reinterpret_cast<__nv_bfloat16*>(dst), n));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
