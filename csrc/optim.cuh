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

// step-dependent scalars of the bias-corrected rules (computed once per thread from the step number t)
struct OptimCorr {
  float c0;  // Adam: sqrt(1 - b2^t) / (1 - b1^t); Adamax: 1 / (1 - b1^t); Nadam: 1 / (1 - b2^t)
  float c1;  // Nadam: b1 / (1 - b1^(t+1))
  float c2;  // Nadam: (1 - b1) / (1 - b1^t)
};

// MUFU.SQRT (1 ulp-ish): the IEEE-rounded sqrtf costs ~10 instructions and a slow-path branch per element
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int KIND>
__device__ __forceinline__ void optim_update(float& w, float g, float& s0, float& s1, float lr,
                                             const OptimArgs& a, const OptimCorr& k) {
  const float corr = k.c0;
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
  } else if constexpr (KIND == DK_OPT_NADAM) {
    s0 = a.p0 * s0 + (1.f - a.p0) * g;
    s1 = a.p1 * s1 + (1.f - a.p1) * g * g;
    const float mhat = k.c1 * s0 + k.c2 * g;
    w -= lr * __fdividef(mhat, fast_sqrt(s1 * k.c0) + a.eps);
  }
}


// learning rate after decay and the bias-correction scalars of step t (>= 1)
__device__ __forceinline__ void optim_prelude(const OptimArgs& a, int t, float& lr, OptimCorr& k) {
  lr = a.lr;
  if (a.decay > 0.f) lr = __fdividef(lr, 1.f + a.decay * static_cast<float>(t - 1));
  k.c0 = 1.f; k.c1 = 0.f; k.c2 = 0.f;
  const float tf = static_cast<float>(t);
  if (a.kind == DK_OPT_ADAM) k.c0 = __fdividef(sqrtf(1.f - __powf(a.p1, tf)), 1.f - __powf(a.p0, tf));
  if (a.kind == DK_OPT_ADAMAX) k.c0 = __fdividef(1.f, 1.f - __powf(a.p0, tf));
  if (a.kind == DK_OPT_NADAM) {
    k.c0 = __fdividef(1.f, 1.f - __powf(a.p1, tf));
    k.c1 = __fdividef(a.p0, 1.f - __powf(a.p0, tf + 1.f));
    k.c2 = __fdividef(1.f - a.p0, 1.f - __powf(a.p0, tf));
  }
}

}  // namespace dk
