// Fused loss kernels (reference op K11: Keras `categorical_crossentropy` on a softmax output plus
// the `accuracy` metric, i.e. everything `train_on_batch` returns in workers.py:199-202).
// One launch computes row softmax, the loss, the gradient (p - y)/B in bf16 (row-major and
// transposed, ready to be tcgen05 GEMM operands) and the batch accuracy; loss and accuracy are
// accumulated straight into the device-resident history record for the current step.
#include "common.cuh"
#include "kernels.h"

namespace dk {

__global__ void __launch_bounds__(256)
softmax_xent_kernel(const float* __restrict__ logits, int ldl, const int* __restrict__ labels,
                    const float* __restrict__ labels_dense, int B, int C,
                    __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dzt, int ldzt,
                    float* __restrict__ probs, float* __restrict__ hist, const int* __restrict__ step,
                    int hist_slots) {
  DK_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_b = 1.f / static_cast<float>(B);
  float loss_acc = 0.f, correct_acc = 0.f;
  for (int row = blockIdx.x * warps_per_block + warp_in_block; row < B;
       row += gridDim.x * warps_per_block) {
    const float* z = logits + static_cast<size_t>(row) * ldl;
    // pass 1: max + argmax
    float mx = -INFINITY;
    int amax = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = z[c];
      if (v > mx) { mx = v; amax = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
    }
    // pass 2: sum exp
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
    se = warp_sum(se);
    const float lse = __logf(se) + mx;
    const float inv_se = 1.f / se;
    // target
    int label = -1;
    float ysum = 1.f;
    if (labels_dense == nullptr) {
      label = labels[row];
    } else {
      // dense targets: label = argmax(y) for the accuracy metric; loss = -sum y log p
      const float* y = labels_dense + static_cast<size_t>(row) * C;
      float ym = -INFINITY; int ya = 0; float ys = 0.f;
      for (int c = lane; c < C; c += 32) { const float v = y[c]; ys += v; if (v > ym) { ym = v; ya = c; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, ym, o);
        const int oa = __shfl_xor_sync(0xffffffffu, ya, o);
        if (om > ym || (om == ym && oa < ya)) { ym = om; ya = oa; }
      }
      ysum = warp_sum(ys);
      label = ya;
    }
    // pass 3: gradient + loss
    float row_loss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(z[c] - mx) * inv_se;
      float y;
      if (labels_dense == nullptr) y = (c == label) ? 1.f : 0.f;
      else y = labels_dense[static_cast<size_t>(row) * C + c];
      row_loss += y * (lse - z[c]);
      const float g = (p * ysum - y) * inv_b;
      if (probs != nullptr) probs[static_cast<size_t>(row) * C + c] = p;
      if (dz != nullptr) dz[static_cast<size_t>(row) * ldz + c] = __float2bfloat16_rn(g);
      if (dzt != nullptr) dzt[static_cast<size_t>(c) * ldzt + row] = __float2bfloat16_rn(g);
    }
    row_loss = warp_sum(row_loss);
    if (lane == 0) {
      loss_acc += row_loss;
      correct_acc += (amax == label) ? 1.f : 0.f;
    }
  }
  // block reduce -> one atomic pair per block
  __shared__ float s_loss[8], s_corr[8];
  if (lane == 0) { s_loss[warp_in_block] = loss_acc; s_corr[warp_in_block] = correct_acc; }
  __syncthreads();
  if (threadIdx.x == 0 && hist != nullptr) {
    float l = 0.f, c = 0.f;
    for (int i = 0; i < warps_per_block; ++i) { l += s_loss[i]; c += s_corr[i]; }
    int slot = step != nullptr ? (*step - 1) : 0;
    if (slot < 0) slot = 0;
    if (hist_slots > 0) slot %= hist_slots;
    atomicAdd(hist + 2 * slot, l * inv_b);
    atomicAdd(hist + 2 * slot + 1, c * inv_b);
  }
}


// Small-C specialisation (MNIST / CIFAR: C = 10, Higgs: C = 2): one thread per row, so the
// transposed gradient store is coalesced across the warp.
template <int MAXC>
__global__ void __launch_bounds__(256)
softmax_xent_small_kernel(const float* __restrict__ logits, int ldl, const int* __restrict__ labels,
                          const float* __restrict__ labels_dense, int B, int C,
                          __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dzt,
                          int ldzt, float* __restrict__ probs, float* __restrict__ hist,
                          const int* __restrict__ step, int hist_slots) {
  DK_PDL_ENTER();
  const float inv_b = 1.f / static_cast<float>(B);
  float loss_acc = 0.f, correct_acc = 0.f;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < B; row += gridDim.x * blockDim.x) {
    const float* z = logits + static_cast<size_t>(row) * ldl;
    float v[MAXC];
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      v[c] = c < C ? z[c] : -INFINITY;
      if (v[c] > mx) { mx = v[c]; amax = c; }
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      v[c] = c < C ? __expf(v[c] - mx) : 0.f;
      se += v[c];
    }
    const float lse = __logf(se) + mx;
    const float inv_se = 1.f / se;
    int label = 0;
    float ysum = 1.f;
    const float* y = nullptr;
    if (labels_dense == nullptr) {
      label = labels[row];
    } else {
      y = labels_dense + static_cast<size_t>(row) * C;
      float ym = -INFINITY;
      ysum = 0.f;
      for (int c = 0; c < C; ++c) { const float t = y[c]; ysum += t; if (t > ym) { ym = t; label = c; } }
    }
    float row_loss = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const float p = v[c] * inv_se;
        const float t = y != nullptr ? y[c] : (c == label ? 1.f : 0.f);
        row_loss += t * (lse - z[c]);
        const float g = (p * ysum - t) * inv_b;
        if (probs != nullptr) probs[static_cast<size_t>(row) * C + c] = p;
        if (dz != nullptr) dz[static_cast<size_t>(row) * ldz + c] = __float2bfloat16_rn(g);
        if (dzt != nullptr) dzt[static_cast<size_t>(c) * ldzt + row] = __float2bfloat16_rn(g);
      }
    }
    loss_acc += row_loss;
    correct_acc += (amax == label) ? 1.f : 0.f;
  }
  loss_acc = warp_sum(loss_acc);
  correct_acc = warp_sum(correct_acc);
  __shared__ float s_loss[8], s_corr[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_loss[w] = loss_acc; s_corr[w] = correct_acc; }
  __syncthreads();
  if (threadIdx.x == 0 && hist != nullptr) {
    float l = 0.f, c = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { l += s_loss[i]; c += s_corr[i]; }
    int slot = step != nullptr ? (*step - 1) : 0;
    if (slot < 0) slot = 0;
    if (hist_slots > 0) slot %= hist_slots;
    atomicAdd(hist + 2 * slot, l * inv_b);
    atomicAdd(hist + 2 * slot + 1, c * inv_b);
  }
}

// mean-squared-error / binary cross-entropy (on probabilities) for non-softmax heads
__global__ void __launch_bounds__(256)
elementwise_loss_kernel(int kind, const float* __restrict__ out, const float* __restrict__ target, int B,
                        int C, __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dzt,
                        int ldzt, float* __restrict__ hist, const int* __restrict__ step, int hist_slots) {
  DK_PDL_ENTER();
  const long n = static_cast<long>(B) * C;
  const float inv = 1.f / static_cast<float>(n);
  float loss = 0.f, corr = 0.f;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / C), c = static_cast<int>(i % C);
    const float o = out[i], t = target[i];
    float g;
    if (kind == DK_LOSS_MSE) {
      const float d = o - t;
      loss += d * d;
      g = 2.f * d * inv;
    } else {
      const float p = fminf(fmaxf(o, 1e-7f), 1.f - 1e-7f);
      loss += -(t * __logf(p) + (1.f - t) * __logf(1.f - p));
      g = (p - t) / (p * (1.f - p)) * inv;
    }
    corr += ((o > 0.5f) == (t > 0.5f)) ? 1.f : 0.f;
    if (dz != nullptr) dz[static_cast<size_t>(r) * ldz + c] = __float2bfloat16_rn(g);
    if (dzt != nullptr) dzt[static_cast<size_t>(c) * ldzt + r] = __float2bfloat16_rn(g);
  }
  loss = warp_sum(loss);
  corr = warp_sum(corr);
  __shared__ float s_loss[8], s_corr[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_loss[w] = loss; s_corr[w] = corr; }
  __syncthreads();
  if (threadIdx.x == 0 && hist != nullptr) {
    float l = 0.f, c = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { l += s_loss[i]; c += s_corr[i]; }
    int slot = step != nullptr ? (*step - 1) : 0;
    if (slot < 0) slot = 0;
    if (hist_slots > 0) slot %= hist_slots;
    atomicAdd(hist + 2 * slot, l * inv);
    atomicAdd(hist + 2 * slot + 1, c * inv);
  }
}

// LabelIndexTransformer rule (distkeras/transformers.py:321-332) + AccuracyEvaluator count
// (evaluators.py:42-48): first index whose activation >= threshold, else arg-max; counts matches.
__global__ void __launch_bounds__(256)
label_index_kernel(const float* __restrict__ probs, int B, int C, float threshold, int default_index,
                   int* __restrict__ out_index, const int* __restrict__ labels, int* correct_count) {
  DK_PDL_ENTER();
  int local = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const float* p = probs + static_cast<size_t>(r) * C;
    int idx = -1;
    float best = -INFINITY;
    int arg = default_index;
    for (int c = 0; c < C; ++c) {
      const float v = p[c];
      if (idx < 0 && v >= threshold) idx = c;
      if (v > best) { best = v; arg = c; }
    }
    if (idx < 0) idx = arg;
    if (out_index != nullptr) out_index[r] = idx;
    if (labels != nullptr && labels[r] == idx) ++local;
  }
  if (correct_count != nullptr) {
    float s = warp_sum(static_cast<float>(local));
    if ((threadIdx.x & 31) == 0 && s > 0.f) atomicAdd(correct_count, static_cast<int>(s));
  }
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_softmax_xent(const float* logits, int ldl, const int* labels, const float* labels_dense, int B,
                    int C, void* dz, int ldz, void* dzt, int ldzt, float* probs, float* hist,
                    const int* step, int hist_slots, void* stream) {
  if (C <= 16) {
    int blocks = (B + 127) / 128;
    DK_HOST_CHECK(DK_LAUNCH(softmax_xent_small_kernel<16>, blocks, 128, 0, (cudaStream_t)stream, 
        logits, ldl, labels, labels_dense, B, C, reinterpret_cast<__nv_bfloat16*>(dz), ldz,
        reinterpret_cast<__nv_bfloat16*>(dzt), ldzt, probs, hist, step, hist_slots));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  int blocks = (B + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  DK_HOST_CHECK(DK_LAUNCH(softmax_xent_kernel, blocks, 256, 0, (cudaStream_t)stream, 
      logits, ldl, labels, labels_dense, B, C, reinterpret_cast<__nv_bfloat16*>(dz), ldz,
      reinterpret_cast<__nv_bfloat16*>(dzt), ldzt, probs, hist, step, hist_slots));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_elementwise_loss(int kind, const float* out, const float* target, int B, int C, void* dz, int ldz,
                        void* dzt, int ldzt, float* hist, const int* step, int hist_slots, void* stream) {
  long n = static_cast<long>(B) * C;
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  DK_HOST_CHECK(DK_LAUNCH(elementwise_loss_kernel, blocks, 256, 0, (cudaStream_t)stream, 
      kind, out, target, B, C, reinterpret_cast<__nv_bfloat16*>(dz), ldz,
      reinterpret_cast<__nv_bfloat16*>(dzt), ldzt, hist, step, hist_slots));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_label_index(const float* probs, int B, int C, float threshold, int default_index, int* out_index,
                   const int* labels, int* correct_count, void* stream) {
  int blocks = (B + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  DK_HOST_CHECK(DK_LAUNCH(label_index_kernel, blocks, 256, 0, (cudaStream_t)stream, probs, B, C, threshold, default_index,
                                                               out_index, labels, correct_count));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
