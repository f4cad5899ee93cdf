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

// ------------------------------------------------------------------------------------------
// Fused classifier head (training): ONE launch computes, for a dense softmax layer with C <= 16
// classes on top of H [B, K] (bf16),
//     logits = H W^T + b ; p = softmax(logits) ; loss / accuracy into the history ring ;
//     dZ = (p - y) / B  (bf16, [B, ldz], pad columns zeroed -- consumed by the wgrad GEMM / bias sum) ;
//     dH = alpha * (dZ W), zeroed where H <= 0 when `use_mask` (dReLU / dropout of the producer).
// It replaces a skinny forward GEMM (N = 10), the loss kernel and a skinny dgrad GEMM (K = 16): ~130
// MFLOP at batch 16384, i.e. nothing for the tensor cores and three launch latencies on the critical
// path.  Eight lanes share a row (each owns every 8th 16-byte chunk of it); W lives in shared memory
// as fp32; per-row reductions are three xor-shuffles per class.
// ------------------------------------------------------------------------------------------
constexpr int kHeadMaxC = 16;

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    f[2 * u] = __uint_as_float(w[u] << 16);
    f[2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(256)
dense_softmax_head_kernel(const __nv_bfloat16* __restrict__ H, int ldh, const __nv_bfloat16* __restrict__ Wb, int ldw,
                          const float* __restrict__ bias, const int* __restrict__ labels,
                          const float* __restrict__ labels_dense, int B, int C, int K,
                          __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dH, int lddh,
                          float alpha, int use_mask, float* __restrict__ hist, const int* __restrict__ step,
                          int hist_slots) {
  DK_PDL_ENTER();
  extern __shared__ float s_w[];  // [C][Kp] fp32
  const int chunks = K >> 3;      // K % 8 == 0
  const int Kp = K + 4;           // +4 floats: rows of W start on different banks
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) {
    const int c = i / K, k = i - c * K;
    s_w[c * Kp + k] = __bfloat162float(Wb[static_cast<size_t>(c) * ldw + k]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  const float inv_b = 1.f / static_cast<float>(B);
  float loss_acc = 0.f, correct_acc = 0.f;
  const int rows_per_iter = gridDim.x * (blockDim.x >> 5) * 4;
  for (int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 4; row0 < B; row0 += rows_per_iter) {
    const int row = row0 + grp;
    const bool valid = row < B;
    const __nv_bfloat16* hrow = H + static_cast<size_t>(valid ? row : 0) * ldh;
    float acc[kHeadMaxC];
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c) acc[c] = 0.f;
    for (int ch = sub; ch < chunks; ch += 8) {
      float h[8];
      unpack8(*reinterpret_cast<const uint4*>(hrow + (ch << 3)), h);
#pragma unroll
      for (int c = 0; c < kHeadMaxC; ++c) {
        if (c < C) {
          const float* w = s_w + c * Kp + (ch << 3);
          const float4 w0 = *reinterpret_cast<const float4*>(w), w1 = *reinterpret_cast<const float4*>(w + 4);
          acc[c] += h[0] * w0.x + h[1] * w0.y + h[2] * w0.z + h[3] * w0.w + h[4] * w1.x + h[5] * w1.y + h[6] * w1.z +
                    h[7] * w1.w;
        }
      }
    }
    // reduce over the 8 lanes of the row (every lane ends up with the full logits)
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c) {
      if (c < C) {
        float v = acc[c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        acc[c] = v + (bias != nullptr ? bias[c] : 0.f);
      }
    }
    float mx = -INFINITY;
    int amax = 0;
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c)
      if (c < C && acc[c] > mx) { mx = acc[c]; amax = c; }
    float e[kHeadMaxC], se = 0.f;
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c) {
      e[c] = c < C ? __expf(acc[c] - mx) : 0.f;
      se += e[c];
    }
    const float lse = __logf(se) + mx, inv_se = 1.f / se;
    int label = 0;
    float ysum = 1.f;
    const float* y = nullptr;
    if (valid) {
      if (labels_dense == nullptr) {
        label = labels[row];
      } else {
        y = labels_dense + static_cast<size_t>(row) * C;
        float ym = -INFINITY;
        ysum = 0.f;
        for (int c = 0; c < C; ++c) { const float t = y[c]; ysum += t; if (t > ym) { ym = t; label = c; } }
      }
    }
    float g[kHeadMaxC], row_loss = 0.f;
#pragma unroll
    for (int c = 0; c < kHeadMaxC; ++c) {
      g[c] = 0.f;
      if (c < C) {
        const float t = y != nullptr ? y[c] : (c == label ? 1.f : 0.f);
        row_loss += t * (lse - acc[c]);
        // the bf16-rounded gradient is what the wgrad GEMM sees: use the same value for dH
        g[c] = __bfloat162float(__float2bfloat16_rn((e[c] * inv_se * ysum - t) * inv_b));
      }
    }
    if (valid && sub == 0) {
      loss_acc += row_loss;
      correct_acc += (amax == label) ? 1.f : 0.f;
    }
    if (valid && dz != nullptr) {
      // lane `sub` stores columns sub and sub + 8 (pad columns up to ldz are written as zeros)
#pragma unroll
      for (int c = 0; c < kHeadMaxC; ++c)
        if ((c & 7) == sub && c < ldz) dz[static_cast<size_t>(row) * ldz + c] = __float2bfloat16_rn(g[c]);
    }
    if (valid && dH != nullptr) {
      for (int ch = sub; ch < chunks; ch += 8) {
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < kHeadMaxC; ++c) {
          if (c < C) {
            const float* w = s_w + c * Kp + (ch << 3);
            const float4 w0 = *reinterpret_cast<const float4*>(w), w1 = *reinterpret_cast<const float4*>(w + 4);
            d[0] += g[c] * w0.x; d[1] += g[c] * w0.y; d[2] += g[c] * w0.z; d[3] += g[c] * w0.w;
            d[4] += g[c] * w1.x; d[5] += g[c] * w1.y; d[6] += g[c] * w1.z; d[7] += g[c] * w1.w;
          }
        }
        if (use_mask) {
          float h[8];
          unpack8(*reinterpret_cast<const uint4*>(hrow + (ch << 3)), h);
#pragma unroll
          for (int u = 0; u < 8; ++u) d[u] = h[u] > 0.f ? d[u] * alpha : 0.f;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) d[u] *= alpha;
        }
        uint4 o;
        o.x = pack_bf16x2(d[0], d[1]); o.y = pack_bf16x2(d[2], d[3]);
        o.z = pack_bf16x2(d[4], d[5]); o.w = pack_bf16x2(d[6], d[7]);
        *reinterpret_cast<uint4*>(dH + static_cast<size_t>(row) * lddh + (ch << 3)) = o;
      }
    }
  }
  loss_acc = warp_sum(loss_acc);
  correct_acc = warp_sum(correct_acc);
  __shared__ float s_loss[8], s_corr[8];
  if (lane == 0) { s_loss[warp] = loss_acc; s_corr[warp] = correct_acc; }
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

// Fused classifier head; returns -7 when the shape is outside what the kernel covers (the caller
// then uses GEMM + dk_softmax_xent + GEMM): C <= 16, K % 8 == 0, ldh / lddh % 8 == 0, W in <= 48 KB smem.
int dk_dense_softmax_head(const void* H, int ldh, const void* Wb, int ldw, const float* bias, const int* labels,
                          const float* labels_dense, int B, int C, int K, void* dz, int ldz, void* dH, int lddh,
                          float alpha, int use_mask, float* hist, const int* step, int hist_slots, void* stream) {
  const size_t smem = static_cast<size_t>(C) * (K + 4) * sizeof(float);
  if (C < 1 || C > kHeadMaxC || K % 8 != 0 || ldh % 8 != 0 || (dH != nullptr && lddh % 8 != 0) || ldz > kHeadMaxC ||
      ldz < C || smem > 48 * 1024)
    return -7;
  int blocks = (B + 31) / 32;
  if (blocks > 148 * 4) blocks = 148 * 4;
  DK_HOST_CHECK(DK_LAUNCH(dense_softmax_head_kernel, blocks, 256, smem, (cudaStream_t)stream,
      reinterpret_cast<const __nv_bfloat16*>(H), ldh, reinterpret_cast<const __nv_bfloat16*>(Wb), ldw, bias, labels,
      labels_dense, B, C, K, reinterpret_cast<__nv_bfloat16*>(dz), ldz, reinterpret_cast<__nv_bfloat16*>(dH), lddh,
      alpha, use_mask, hist, step, hist_slots));
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
