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
// path.  Layout of the work: a warp owns 8 rows per pass; 8 lanes share a row (lane `sub` owns the
// 16-byte chunks sub, sub + 8, ... of it) and every lane works on TWO rows (r and r + 4) so each
// shared-memory read of W feeds 16 FMAs.  W is staged once per block as fp32 in a chunk-major layout
// [chunk][class][8] whose chunk stride (8 C + 4 words) makes the 8 lanes of an LDS.128 phase hit
// disjoint banks (measured: the naive [class][K] layout ran 6.5-way conflicted and smem-bound).
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

// MAXC: compile-time bound on the class count (2 / 10 / 16); NCH > 0: every lane keeps its (at most
// NCH) chunks of both rows in registers -- all loads are issued back to back and the dReLU mask needs
// no second read; NCH == 0 is the generic loop for K > 256.
// R: rows handled by each 8-lane group per pass (2 = a shared-memory read of W feeds 16 FMAs: best when the
// batch is large; 1 = half the serial work per lane, twice the warps: best in the latency-bound small-batch regime).
template <int MAXC, int NCH, int R>
__global__ void __launch_bounds__(128)
dense_softmax_head_kernel(const __nv_bfloat16* __restrict__ H, int ldh, const __nv_bfloat16* __restrict__ Wb, int ldw,
                          const float* __restrict__ bias, const int* __restrict__ labels,
                          const float* __restrict__ labels_dense, int B, int C, int K,
                          __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dH, int lddh,
                          float alpha, int use_mask, float* __restrict__ hist, const int* __restrict__ step,
                          int hist_slots) {
  DK_PDL_ENTER();
  extern __shared__ __align__(16) float s_w[];  // [chunks][C * 8 + 4]
  const int chunks = K >> 3;                    // K % 8 == 0
  const int SC = C * 8 + 4;
#pragma unroll 4
  for (int i = threadIdx.x; i < C * chunks; i += blockDim.x) {  // one 16-byte load per 8 weights
    const int c = i / chunks, ch = i - c * chunks;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(Wb + static_cast<size_t>(c) * ldw + (ch << 3)), f);
    float* dst = s_w + ch * SC + c * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane & 7, grp = lane >> 3;
  const float inv_b = 1.f / static_cast<float>(B);
  float loss_acc = 0.f, correct_acc = 0.f;
  const int rows_per_pass = gridDim.x * nwarps * 4 * R;
  for (int row0 = (blockIdx.x * nwarps + warp) * 4 * R; row0 < B; row0 += rows_per_pass) {
    int row[R];
    bool valid[R];
    const __nv_bfloat16* hrow[R];
    float acc[R][MAXC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row[r] = row0 + 4 * r + grp;
      valid[r] = row[r] < B;
      hrow[r] = H + static_cast<size_t>(valid[r] ? row[r] : 0) * ldh;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) acc[r][c] = 0.f;
    }
    uint4 hq[R][NCH > 0 ? NCH : 1];
    if constexpr (NCH > 0) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int ch = sub + 8 * j;
#pragma unroll
        for (int r = 0; r < R; ++r)
          hq[r][j] = ch < chunks ? *reinterpret_cast<const uint4*>(hrow[r] + (ch << 3)) : make_uint4(0, 0, 0, 0);
      }
    }
    auto fwd_chunk = [&](int ch, const uint4* q) {
      float hv[R][8];
#pragma unroll
      for (int r = 0; r < R; ++r) unpack8(q[r], hv[r]);
      const float* wch = s_w + ch * SC;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float4 w0 = *reinterpret_cast<const float4*>(wch + c * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wch + c * 8 + 4);
#pragma unroll
          for (int r = 0; r < R; ++r)
            acc[r][c] += hv[r][0] * w0.x + hv[r][1] * w0.y + hv[r][2] * w0.z + hv[r][3] * w0.w + hv[r][4] * w1.x +
                         hv[r][5] * w1.y + hv[r][6] * w1.z + hv[r][7] * w1.w;
        }
      }
    };
    if constexpr (NCH > 0) {
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        if (sub + 8 * j < chunks) {
          uint4 q[R];
#pragma unroll
          for (int r = 0; r < R; ++r) q[r] = hq[r][j];
          fwd_chunk(sub + 8 * j, q);
        }
    } else {
#pragma unroll 4  // independent loads of four chunks in flight per lane (the accumulation is the only dependency)
      for (int ch = sub; ch < chunks; ch += 8) {
        uint4 q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) q[r] = *reinterpret_cast<const uint4*>(hrow[r] + (ch << 3));
        fwd_chunk(ch, q);
      }
    }
    // reduce over the 8 lanes of a row (every lane ends up with the full logits), then softmax / loss
    float g[R][MAXC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float mx = -INFINITY;
      int amax = 0;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          float v = acc[r][c];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += bias != nullptr ? bias[c] : 0.f;
          acc[r][c] = v;
          if (v > mx) { mx = v; amax = c; }
        }
      }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        g[r][c] = c < C ? __expf(acc[r][c] - mx) : 0.f;
        se += g[r][c];
      }
      const float lse = __logf(se) + mx, inv_se = 1.f / se;
      int label = 0;
      float ysum = 1.f;
      const float* y = nullptr;
      if (valid[r]) {
        if (labels_dense == nullptr) {
          label = labels[row[r]];
        } else {
          y = labels_dense + static_cast<size_t>(row[r]) * C;
          float ym = -INFINITY;
          ysum = 0.f;
          for (int c = 0; c < C; ++c) { const float t = y[c]; ysum += t; if (t > ym) { ym = t; label = c; } }
        }
      }
      float row_loss = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float t = y != nullptr ? y[c] : (c == label ? 1.f : 0.f);
          row_loss += t * (lse - acc[r][c]);
          // the bf16-rounded gradient is what the wgrad GEMM sees: use the same value for dH
          g[r][c] = __bfloat162float(__float2bfloat16_rn((g[r][c] * inv_se * ysum - t) * inv_b));
        }
      }
      if (valid[r] && sub == 0) {
        loss_acc += row_loss;
        correct_acc += (amax == label) ? 1.f : 0.f;
      }
      if (valid[r] && dz != nullptr) {
        // lane `sub` stores the column pair (2 sub, 2 sub + 1) as one 4-byte word: the 8 lanes of a
        // row write 32 contiguous bytes (pad columns up to ldz are zeros); ldz is even
        uint32_t word = 0;
#pragma unroll
        for (int c = 0; c < MAXC; c += 2)
          if ((c >> 1) == sub) word = pack_bf16x2(g[r][c], c + 1 < MAXC ? g[r][c + 1] : 0.f);
        if (2 * sub < ldz) *reinterpret_cast<uint32_t*>(dz + static_cast<size_t>(row[r]) * ldz + 2 * sub) = word;
      }
    }
    auto bwd_chunk = [&](int ch, const uint4* q) {
      float d[R][8];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) d[r][u] = 0.f;
      const float* wch = s_w + ch * SC;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float4 w0 = *reinterpret_cast<const float4*>(wch + c * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wch + c * 8 + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) d[r][u] += g[r][c] * wv[u];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (use_mask) {
          float h[8];
          unpack8(q[r], h);
#pragma unroll
          for (int u = 0; u < 8; ++u) d[r][u] = h[u] > 0.f ? d[r][u] * alpha : 0.f;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) d[r][u] *= alpha;
        }
        if (valid[r]) {
          uint4 o;
          o.x = pack_bf16x2(d[r][0], d[r][1]); o.y = pack_bf16x2(d[r][2], d[r][3]);
          o.z = pack_bf16x2(d[r][4], d[r][5]); o.w = pack_bf16x2(d[r][6], d[r][7]);
          *reinterpret_cast<uint4*>(dH + static_cast<size_t>(row[r]) * lddh + (ch << 3)) = o;
        }
      }
    };
    if (dH != nullptr) {
      if constexpr (NCH > 0) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
          if (sub + 8 * j < chunks) {
            uint4 q[R];
#pragma unroll
            for (int r = 0; r < R; ++r) q[r] = hq[r][j];
            bwd_chunk(sub + 8 * j, q);
          }
      } else {
#pragma unroll 4
        for (int ch = sub; ch < chunks; ch += 8) {
          uint4 q[R];
#pragma unroll
          for (int r = 0; r < R; ++r) q[r] = *reinterpret_cast<const uint4*>(hrow[r] + (ch << 3));
          bwd_chunk(ch, q);
        }
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
    for (int i = 0; i < nwarps; ++i) { l += s_loss[i]; c += s_corr[i]; }
    int slot = step != nullptr ? (*step - 1) : 0;
    if (slot < 0) slot = 0;
    if (hist_slots > 0) slot %= hist_slots;
    atomicAdd(hist + 2 * slot, l * inv_b);
    atomicAdd(hist + 2 * slot + 1, c * inv_b);
  }
}


// ------------------------------------------------------------------------------------------
// Small-batch variant of the fused head (B <= 512, the reference's batch sizes): the step is a chain of
// dependent kernels there, so this one is built for latency, not throughput.  A block owns 8 rows, 16
// lanes per row.  ONE round of global loads brings the whole weight matrix (bf16, C x K), the block's 8
// activation rows, bias and labels into shared memory; every lane then owns the bf16 pairs lr, lr + 16,
// ... of its row: logits by 10 short FMA chains + a 4-step half-warp reduction, softmax / loss /
// accuracy redundantly per lane, dZ as packed pairs, and dH = alpha (dZ W) * (H > 0) for the lane's own
// pairs (coalesced 4-byte stores).  No loop is unrolled more than NW (<= 8) times: the kernel runs once
// per step and straight-line code is paid for in instruction-cache misses.
// ------------------------------------------------------------------------------------------
template <int MAXC>
__global__ void __launch_bounds__(128)
dense_softmax_head_small_kernel(const __nv_bfloat16* __restrict__ H, int ldh, const __nv_bfloat16* __restrict__ Wb, int ldw,
                                const float* __restrict__ bias, const int* __restrict__ labels,
                                const float* __restrict__ labels_dense, int B, int C, int K,
                                __nv_bfloat16* __restrict__ dz, int ldz, __nv_bfloat16* __restrict__ dH, int lddh,
                                float alpha, int use_mask, float* __restrict__ hist, const int* __restrict__ step,
                                int hist_slots) {
  DK_PDL_ENTER();
  extern __shared__ __align__(16) uint32_t s_head[];
  const int kw = K >> 1;                      // bf16 pairs per row (K % 8 == 0)
  uint32_t* sW = s_head;                      // [C][kw]
  uint32_t* sH = sW + C * kw;                 // [8][kw]
  float* sB = reinterpret_cast<float*>(sH + 8 * kw);  // [MAXC] bias, then [8] labels
  int* sL = reinterpret_cast<int*>(sB + MAXC);
  __shared__ float s_loss[4], s_corr[4];
  const int tid = threadIdx.x;
  const int row_in = tid >> 4, lr = tid & 15;
  const int row0 = blockIdx.x * 8;
  const int chunks = K >> 3;                  // 16-byte chunks per row
  // ---- one round of global loads ----
  for (int i = tid; i < C * chunks; i += 128) {
    const int c = i / chunks, ch = i - c * chunks;
    reinterpret_cast<uint4*>(sW + c * kw)[ch] = *reinterpret_cast<const uint4*>(Wb + static_cast<size_t>(c) * ldw + (ch << 3));
  }
  for (int i = tid; i < 8 * chunks; i += 128) {
    const int r = i / chunks, ch = i - r * chunks;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (row0 + r < B) q = *reinterpret_cast<const uint4*>(H + static_cast<size_t>(row0 + r) * ldh + (ch << 3));
    reinterpret_cast<uint4*>(sH + r * kw)[ch] = q;
  }
  if (tid < MAXC) sB[tid] = (bias != nullptr && tid < C) ? bias[tid] : 0.f;
  if (tid >= 64 && tid < 72) {
    const int r = tid - 64;
    int lab = 0;
    if (row0 + r < B && labels_dense == nullptr) lab = labels[row0 + r];
    sL[r] = lab;
  }
  __syncthreads();
  const int row = row0 + row_in;
  const bool valid = row < B;
  const uint32_t* hrow = sH + row_in * kw;
  // ---- logits: the lane's pairs against every class ----
  float acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
#pragma unroll 2
  for (int w = lr; w < kw; w += 16) {
    const uint32_t hp = hrow[w];
    const float h0 = __uint_as_float(hp << 16), h1 = __uint_as_float(hp & 0xffff0000u);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {
        const uint32_t wp = sW[c * kw + w];
        acc[c] = fmaf(h0, __uint_as_float(wp << 16), fmaf(h1, __uint_as_float(wp & 0xffff0000u), acc[c]));
      }
    }
  }
  float mx = -INFINITY;
  int amax = 0;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      float v = acc[c];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += sB[c];
      acc[c] = v;
      if (v > mx) { mx = v; amax = c; }
    }
  }
  // ---- softmax, loss, accuracy, dZ ----
  float g[MAXC];
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    g[c] = c < C ? __expf(acc[c] - mx) : 0.f;
    se += g[c];
  }
  const float lse = __logf(se) + mx, inv_se = __fdividef(1.f, se), inv_b = 1.f / static_cast<float>(B);
  int label = sL[row_in];
  float ysum = 1.f;
  const float* y = nullptr;
  if (valid && labels_dense != nullptr) {
    y = labels_dense + static_cast<size_t>(row) * C;
    float ym = -INFINITY;
    ysum = 0.f;
    for (int c = 0; c < C; ++c) { const float t = y[c]; ysum += t; if (t > ym) { ym = t; label = c; } }
  }
  float row_loss = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < C) {
      const float t = y != nullptr ? y[c] : (c == label ? 1.f : 0.f);
      row_loss += t * (lse - acc[c]);
      // the bf16-rounded gradient is what the weight-gradient GEMM sees: use the same value for dH
      g[c] = __bfloat162float(__float2bfloat16_rn((g[c] * inv_se * ysum - t) * inv_b));
    }
  }
  if (valid && dz != nullptr && 2 * lr < ldz) {
    uint32_t word = 0;
#pragma unroll
    for (int c = 0; c < MAXC; c += 2)
      if ((c >> 1) == lr) word = pack_bf16x2(g[c], c + 1 < MAXC ? g[c + 1] : 0.f);
    *reinterpret_cast<uint32_t*>(dz + static_cast<size_t>(row) * ldz + 2 * lr) = word;
  }
  // ---- dH = alpha (dZ W) * (H > 0): the lane's own pairs ----
  if (dH != nullptr && valid) {
#pragma unroll 2
    for (int w = lr; w < kw; w += 16) {
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const uint32_t wp = sW[c * kw + w];
          d0 = fmaf(g[c], __uint_as_float(wp << 16), d0);
          d1 = fmaf(g[c], __uint_as_float(wp & 0xffff0000u), d1);
        }
      }
      if (use_mask) {
        const uint32_t hp = hrow[w];
        d0 = __uint_as_float(hp << 16) > 0.f ? d0 * alpha : 0.f;
        d1 = __uint_as_float(hp & 0xffff0000u) > 0.f ? d1 * alpha : 0.f;
      } else {
        d0 *= alpha;
        d1 *= alpha;
      }
      *reinterpret_cast<uint32_t*>(dH + static_cast<size_t>(row) * lddh + 2 * w) = pack_bf16x2(d0, d1);
    }
  }
  // ---- history record: one atomic pair per block ----
  float l = (valid && lr == 0) ? row_loss : 0.f, cr = (valid && lr == 0 && amax == label) ? 1.f : 0.f;
  l = warp_sum(l);
  cr = warp_sum(cr);
  if ((tid & 31) == 0) { s_loss[tid >> 5] = l; s_corr[tid >> 5] = cr; }
  __syncthreads();
  if (tid == 0 && hist != nullptr) {
    int slot = step != nullptr ? (*step - 1) : 0;
    if (slot < 0) slot = 0;
    if (hist_slots > 0) slot %= hist_slots;
    atomicAdd(hist + 2 * slot, (s_loss[0] + s_loss[1] + s_loss[2] + s_loss[3]) * inv_b);
    atomicAdd(hist + 2 * slot + 1, (s_corr[0] + s_corr[1] + s_corr[2] + s_corr[3]) * inv_b);
  }
}

template <int MAXC>
static int launch_head_small(const void* H, int ldh, const void* Wb, int ldw, const float* bias, const int* labels,
                             const float* labels_dense, int B, int C, int K, void* dz, int ldz, void* dH, int lddh,
                             float alpha, int use_mask, float* hist, const int* step, int hist_slots, void* stream) {
  const size_t smem = static_cast<size_t>(C + 8) * (K / 2) * 4 + MAXC * 4 + 8 * 4;
  DK_HOST_CHECK(DK_LAUNCH((dense_softmax_head_small_kernel<MAXC>), (B + 7) / 8, 128, smem, (cudaStream_t)stream,
      reinterpret_cast<const __nv_bfloat16*>(H), ldh, reinterpret_cast<const __nv_bfloat16*>(Wb), ldw, bias, labels,
      labels_dense, B, C, K, reinterpret_cast<__nv_bfloat16*>(dz), ldz, reinterpret_cast<__nv_bfloat16*>(dH), lddh,
      alpha, use_mask, hist, step, hist_slots));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

// host-side dispatch of the fused head (class-count / chunk-count specialisations)
template <int MAXC>
static int launch_head(const void* H, int ldh, const void* Wb, int ldw, const float* bias, const int* labels,
                       const float* labels_dense, int B, int C, int K, void* dz, int ldz, void* dH, int lddh,
                       float alpha, int use_mask, float* hist, const int* step, int hist_slots, void* stream,
                       size_t smem) {
  int blocks = (B + 31) / 32;                // 4 warps x 8 rows per pass
  if (blocks > 148 * 4) blocks = 148 * 4;    // resident blocks stage W once and walk their rows
  const int per_lane = (K / 8 + 7) / 8;      // 16-byte chunks each of the 8 lanes of a row owns
#define DK_HEAD_LAUNCH(NCH)                                                                                          \
  DK_HOST_CHECK(DK_LAUNCH((dense_softmax_head_kernel<MAXC, NCH, 2>), blocks, 128, smem, (cudaStream_t)stream,        \
      reinterpret_cast<const __nv_bfloat16*>(H), ldh, reinterpret_cast<const __nv_bfloat16*>(Wb), ldw, bias, labels, \
      labels_dense, B, C, K, reinterpret_cast<__nv_bfloat16*>(dz), ldz, reinterpret_cast<__nv_bfloat16*>(dH), lddh,  \
      alpha, use_mask, hist, step, hist_slots))
  if (per_lane <= 2) { DK_HEAD_LAUNCH(2); }
  else if (per_lane <= 4) { DK_HEAD_LAUNCH(4); }
  else { DK_HEAD_LAUNCH(0); }
#undef DK_HEAD_LAUNCH
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
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
// (evaluators.py:42-48): first index whose activation >= threshold, else the arg-max over the positive
// entries, else default_index; counts matches.
__global__ void __launch_bounds__(256)
label_index_kernel(const float* __restrict__ probs, int B, int C, float threshold, int default_index,
                   int* __restrict__ out_index, const int* __restrict__ labels, int* correct_count) {
  DK_PDL_ENTER();
  int local = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const float* p = probs + static_cast<size_t>(r) * C;
    int idx = -1;
    float best = 0.f;  // the reference's running maximum starts at 0: no positive entry -> default_index
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
// then uses GEMM + dk_softmax_xent + GEMM): C <= 16, K % 8 == 0, ld % 8 == 0, W in <= 48 KB smem.
int dk_dense_softmax_head(const void* H, int ldh, const void* Wb, int ldw, const float* bias, const int* labels,
                          const float* labels_dense, int B, int C, int K, void* dz, int ldz, void* dH, int lddh,
                          float alpha, int use_mask, float* hist, const int* step, int hist_slots, void* stream) {
  const size_t smem = static_cast<size_t>(K / 8) * (C * 8 + 4) * sizeof(float);
  if (C < 1 || C > kHeadMaxC || K % 8 != 0 || ldh % 8 != 0 || (dH != nullptr && lddh % 8 != 0) || ldz > kHeadMaxC ||
      ldz < C || ldz % 2 != 0 || ldw % 8 != 0 || smem > 48 * 1024)
    return -7;
  if (B <= 512 && static_cast<size_t>(C + 8) * (K / 2) * 4 <= 40 * 1024) {  // latency-bound regime
    if (C <= 2)
      return launch_head_small<2>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask,
                                  hist, step, hist_slots, stream);
    if (C <= 10)
      return launch_head_small<10>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask,
                                   hist, step, hist_slots, stream);
    return launch_head_small<16>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask,
                                 hist, step, hist_slots, stream);
  }
  if (C <= 2)
    return launch_head<2>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask, hist,
                          step, hist_slots, stream, smem);
  if (C <= 10)
    return launch_head<10>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask,
                           hist, step, hist_slots, stream, smem);
  return launch_head<16>(H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, alpha, use_mask, hist,
                         step, hist_slots, stream, smem);
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
