// Update rules of the worker optimizers (reference op K12), shared by the flat optimizer kernel
// (optim_kernels.cu) and the fused wgrad + optimizer epilogue (dense_fused.cu).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace dk {

struct OptimArgs {
  float* w;
  const float* g;
  float* s0;  // momentum / accumulator / adam m / adadelta acc
  float* s1;  // adam v / adadelta delta_acc / adamax u
  __nv_bfloat16* wb;
  long n;
  int kind;
  float lr, p0, p1, eps, decay;
  int nesterov;
  const int* step;  // device step counter (>= 1)
  float grad_scale;
};

// MUFU.SQRT (1 ulp-ish): the IEEE-rounded sqrtf costs ~10 instructions and a slow-path branch per element
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int KIND>
__device__ __forceinline__ void optim_update(float& w, float g, float& s0, float& s1, float lr,
                                             const OptimArgs& a, float corr) {
  if constexpr (KIND == DK_OPT_SGD) {
    w -= lr * g;
  } else if constexpr (KIND == DK_OPT_MOMENTUM) {
    const float v = a.p0 * s0 - lr * g;
    s0 = v;
    w += a.nesterov ? a.p0 * v - lr * g : v;
  } else if constexpr (KIND == DK_OPT_ADAGRAD) {
    s0 += g * g;
    w -= lr * __fdividef(g, fast_sqrt(s0) + a.eps);
  } else if constexpr (KIND == DK_OPT_RMSPROP) {
    s0 = a.p0 * s0 + (1.f - a.p0) * g * g;
    w -= lr * __fdividef(g, fast_sqrt(s0) + a.eps);
  } else if constexpr (KIND == DK_OPT_ADAM) {
    s0 = a.p0 * s0 + (1.f - a.p0) * g;
    s1 = a.p1 * s1 + (1.f - a.p1) * g * g;
    w -= lr * corr * __fdividef(s0, fast_sqrt(s1) + a.eps);
  } else if constexpr (KIND == DK_OPT_ADADELTA) {
    s0 = a.p0 * s0 + (1.f - a.p0) * g * g;
    const float upd = g * fast_sqrt(__fdividef(s1 + a.eps, s0 + a.eps));
    w -= lr * upd;
    s1 = a.p0 * s1 + (1.f - a.p0) * upd * upd;
  } else if constexpr (KIND == DK_OPT_ADAMAX) {
    s0 = a.p0 * s0 + (1.f - a.p0) * g;
    s1 = fmaxf(a.p1 * s1, fabsf(g));
    w -= lr * corr * __fdividef(s0, s1 + a.eps);
  }
}


// learning rate after decay and the bias-correction factor of step t (>= 1)
__device__ __forceinline__ void optim_prelude(const OptimArgs& a, int t, float& lr, float& corr) {
  lr = a.lr;
  if (a.decay > 0.f) lr = lr / (1.f + a.decay * static_cast<float>(t - 1));
  corr = 1.f;
  if (a.kind == DK_OPT_ADAM)
    corr = sqrtf(1.f - powf(a.p1, static_cast<float>(t))) / (1.f - powf(a.p0, static_cast<float>(t)));
  if (a.kind == DK_OPT_ADAMAX) corr = 1.f / (1.f - powf(a.p0, static_cast<float>(t)));
}

// run-time dispatch (warp-uniform branch) for kernels that are not templated on the rule
__device__ __forceinline__ void optim_update_rt(int kind, float& w, float g, float& s0, float& s1, float lr,
                                                const OptimArgs& a, float corr) {
  switch (kind) {
    case DK_OPT_SGD: optim_update<DK_OPT_SGD>(w, g, s0, s1, lr, a, corr); break;
    case DK_OPT_MOMENTUM: optim_update<DK_OPT_MOMENTUM>(w, g, s0, s1, lr, a, corr); break;
    case DK_OPT_ADAGRAD: optim_update<DK_OPT_ADAGRAD>(w, g, s0, s1, lr, a, corr); break;
    case DK_OPT_RMSPROP: optim_update<DK_OPT_RMSPROP>(w, g, s0, s1, lr, a, corr); break;
    case DK_OPT_ADAM: optim_update<DK_OPT_ADAM>(w, g, s0, s1, lr, a, corr); break;
    case DK_OPT_ADADELTA: optim_update<DK_OPT_ADADELTA>(w, g, s0, s1, lr, a, corr); break;
    default: optim_update<DK_OPT_ADAMAX>(w, g, s0, s1, lr, a, corr); break;
  }
}

}  // namespace dk
