// Layout / elementwise kernels around the tcgen05 GEMMs: fused input stage (dtype cast +
// MinMaxTransformer affine + transposed copy), bf16 transposes, bias-gradient row sums, im2col /
// col2im (reference op K8 lowered to GEMM), 2x2 max-pool (K9) and residual helpers.
#include "common.cuh"
#include "kernels.h"

namespace dk {

// ------------------------------------------------------------------------------------------
// input stage
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<uint8_t>(const uint8_t* p) { return static_cast<float>(*p); }
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(256)
input_stage_kernel(const T* __restrict__ x, int B, int F, float scale, float shift,
                   __nv_bfloat16* __restrict__ xb, int ldx, __nv_bfloat16* __restrict__ xt, int ldxt,
                   int* step_counter) {
  DK_PDL_ENTER();
  __shared__ __nv_bfloat16 tile[32][33];
  const int f0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r, f = f0 + tx;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (b < B && f < F) {
      v = __float2bfloat16_rn(load_as_float<T>(x + static_cast<size_t>(b) * F + f) * scale + shift);
      if (xb != nullptr) xb[static_cast<size_t>(b) * ldx + f] = v;
    }
    tile[r][tx] = v;
  }
  if (xt != nullptr) {
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int f = f0 + r, b = b0 + tx;
      if (f < F && b < B) xt[static_cast<size_t>(f) * ldxt + b] = tile[tx][r];
    }
  }
  if (step_counter != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    *step_counter += 1;
}

// Vectorised input stage (no transposed copy): 8 features per thread, one 16-byte bf16 store.
template <typename T>
__global__ void __launch_bounds__(256)
input_stage_vec_kernel(const T* __restrict__ x, long rows, int F, float scale, float shift,
                       __nv_bfloat16* __restrict__ xb, int ldx, float* __restrict__ xf, int ldxf, int* step_counter) {
  DK_PDL_ENTER();
  const int f8 = F >> 3;  // F % 8 == 0 guaranteed by the launcher
  const long total = rows * f8;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / f8;
    const int c = static_cast<int>(i - r * f8) << 3;
    const T* src = x + r * F + c;
    float v[8];
    if constexpr (sizeof(T) == 1) {
      const uint2 q = *reinterpret_cast<const uint2*>(src);
      const uint8_t* b = reinterpret_cast<const uint8_t*>(&q);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = static_cast<float>(b[t]);
    } else if constexpr (sizeof(T) == 4) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 b = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      const uint4 q = *reinterpret_cast<const uint4*>(src);
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = __bfloat162float(h[t]);
    }
    uint4 o;
    o.x = pack_bf16x2(v[0] * scale + shift, v[1] * scale + shift);
    o.y = pack_bf16x2(v[2] * scale + shift, v[3] * scale + shift);
    o.z = pack_bf16x2(v[4] * scale + shift, v[5] * scale + shift);
    o.w = pack_bf16x2(v[6] * scale + shift, v[7] * scale + shift);
    *reinterpret_cast<uint4*>(xb + r * ldx + c) = o;
    if (xf != nullptr) {  // fp32 copy: A operand of the tf32 pull-fused first-layer GEMM
      float* dst = xf + r * ldxf + c;
      *reinterpret_cast<float4*>(dst) = make_float4(v[0] * scale + shift, v[1] * scale + shift, v[2] * scale + shift,
                                                    v[3] * scale + shift);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * scale + shift, v[5] * scale + shift,
                                                        v[6] * scale + shift, v[7] * scale + shift);
    }
  }
  if (step_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1;
}

__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols, int lds,
                      __nv_bfloat16* __restrict__ dst, int ldd) {
  DK_PDL_ENTER();
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8)
    if (r0 + r < rows && c0 + tx < cols) tile[r][tx] = src[static_cast<size_t>(r0 + r) * lds + c0 + tx];
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8)
    if (c0 + r < cols && r0 + tx < rows) dst[static_cast<size_t>(c0 + r) * ldd + r0 + tx] = tile[tx][r];
}

// out[r] = scale * sum_c src[r, c]; one warp per row, 16-byte loads when aligned.
__global__ void __launch_bounds__(256)
rowsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols, int lds,
                   float* __restrict__ out, float scale) {
  DK_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* p = src + static_cast<size_t>(row) * lds;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const int c8 = cols >> 3;
    for (int i = lane; i < c8; i += 32) {
      const uint4 q = reinterpret_cast<const uint4*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __bfloat1622float2(h[t]);
        acc += f.x + f.y;
      }
    }
    for (int c = (c8 << 3) + lane; c < cols; c += 32) acc += __bfloat162float(p[c]);
  } else {
    for (int c = lane; c < cols; c += 32) acc += __bfloat162float(p[c]);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc * scale;
}

// out[c] += scale * sum_r src[r, c]  (bias gradient from dZ [B, N]).  Grid = (column blocks of 64,
// row splits); each warp reads 128-byte row segments (bf16x2 per lane), the block reduces through
// shared memory and issues one fp32 atomicAdd per column.  `out` must be zeroed beforehand.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols, int lds,
                   float* __restrict__ out, float scale, int rows_per_block) {
  DK_PDL_ENTER();
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f;
  if (c + 1 < cols && (lds & 1) == 0) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + static_cast<size_t>(r) * lds + c);
      const float2 f = __bfloat1622float2(v);
      a0 += f.x;
      a1 += f.y;
    }
  } else if (c < cols) {
    for (int r = r0 + warp; r < r1; r += 8) {
      a0 += __bfloat162float(src[static_cast<size_t>(r) * lds + c]);
      if (c + 1 < cols) a1 += __bfloat162float(src[static_cast<size_t>(r) * lds + c + 1]);
    }
  }
  red[warp][lane * 2] = a0;
  red[warp][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const int col = blockIdx.x * 64 + threadIdx.x;
    if (col < cols) atomicAdd(out + col, t * scale);
  }
}

// ------------------------------------------------------------------------------------------
// convolution lowering (NHWC, bf16): col[(b, oh, ow), (kh, kw, c)]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int KH, int KW,
              int stride, int pad, int OH, int OW, __nv_bfloat16* __restrict__ col, int ldcol) {
  DK_PDL_ENTER();
  // one thread per (row, kh, kw, c8-chunk); channels are innermost -> contiguous copies
  const int cvec = (C % 8 == 0) ? 8 : 1;
  const int cchunks = C / cvec;
  const long total = static_cast<long>(B) * OH * OW * KH * KW * cchunks;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int cc = static_cast<int>(t % cchunks); t /= cchunks;
    const int kw = static_cast<int>(t % KW); t /= KW;
    const int kh = static_cast<int>(t % KH); t /= KH;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH); t /= OH;
    const int b = static_cast<int>(t);
    const int ih = oh * stride - pad + kh, iw = ow * stride - pad + kw;
    const long row = (static_cast<long>(b) * OH + oh) * OW + ow;
    __nv_bfloat16* dst = col + row * ldcol + (kh * KW + kw) * C + cc * cvec;
    const bool inside = ih >= 0 && ih < H && iw >= 0 && iw < W;
    const __nv_bfloat16* srcp = x + ((static_cast<long>(b) * H + ih) * W + iw) * C + cc * cvec;
    if (cvec == 8) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (inside) v = *reinterpret_cast<const uint4*>(srcp);
      *reinterpret_cast<uint4*>(dst) = v;
    } else {
      *dst = inside ? *srcp : __float2bfloat16_rn(0.f);
    }
  }
}

// Small / odd channel counts (C = 1 or 3 in the first layer): one thread assembles 8 consecutive
// elements of a col row (scalar gathers that hit L1) and writes them as ONE 16-byte store; the
// padding columns K..ldcol-1 are written as zeros.  ldcol % 8 == 0.
__global__ void __launch_bounds__(256)
im2col_gather8_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int KH, int KW,
                      int stride, int pad, int OH, int OW, __nv_bfloat16* __restrict__ col, int ldcol) {
  DK_PDL_ENTER();
  const int K = KH * KW * C;
  const int chunks = ldcol >> 3;
  const long total = static_cast<long>(B) * OH * OW * chunks;
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const long row = i / chunks;
    long t = row;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH); t /= OH;
    const int b = static_cast<int>(t);
    const int ih0 = oh * stride - pad, iw0 = ow * stride - pad;
    int k = ch << 3;
    int c = k % C, tap = k / C;
    int kw = tap % KW, kh = tap / KW;
    unsigned short v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ih = ih0 + kh, iw = iw0 + kw;
      const bool ok = (k + e) < K && ih >= 0 && ih < H && iw >= 0 && iw < W;
      v[e] = ok ? __ldg(xs + ((static_cast<long>(b) * H + ih) * W + iw) * C + c) : static_cast<unsigned short>(0);
      if (++c == C) {
        c = 0;
        if (++kw == KW) { kw = 0; ++kh; }
      }
    }
    uint4 o;
    o.x = v[0] | (static_cast<uint32_t>(v[1]) << 16); o.y = v[2] | (static_cast<uint32_t>(v[3]) << 16);
    o.z = v[4] | (static_cast<uint32_t>(v[5]) << 16); o.w = v[6] | (static_cast<uint32_t>(v[7]) << 16);
    *reinterpret_cast<uint4*>(col + row * ldcol + (ch << 3)) = o;
  }
}

// dx[b, ih, iw, c] = sum over the (kh, kw) windows that cover it (gather form: no atomics)
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ col, int ldcol, int B, int H, int W, int C, int KH,
              int KW, int stride, int pad, int OH, int OW, __nv_bfloat16* __restrict__ dx,
              const __nv_bfloat16* __restrict__ mask) {
  DK_PDL_ENTER();
  const long total = static_cast<long>(B) * H * W * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % C); t /= C;
    const int iw = static_cast<int>(t % W); t /= W;
    const int ih = static_cast<int>(t % H); t /= H;
    const int b = static_cast<int>(t);
    float acc = 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int ohs = ih + pad - kh;
      if (ohs < 0 || ohs % stride != 0) continue;
      const int oh = ohs / stride;
      if (oh >= OH) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int ows = iw + pad - kw;
        if (ows < 0 || ows % stride != 0) continue;
        const int ow = ows / stride;
        if (ow >= OW) continue;
        const long row = (static_cast<long>(b) * OH + oh) * OW + ow;
        acc += __bfloat162float(col[row * ldcol + (kh * KW + kw) * C + c]);
      }
    }
    if (mask != nullptr && !(__bfloat162float(mask[i]) > 0.f)) acc = 0.f;  // fused dReLU of the producer
    dx[i] = __float2bfloat16_rn(acc);
  }
}

// 8-channel vector form of the gather (C % 8 == 0): one 16-byte load per covering window
__global__ void __launch_bounds__(256)
col2im_vec8_kernel(const __nv_bfloat16* __restrict__ col, int ldcol, int B, int H, int W, int C, int KH,
                   int KW, int stride, int pad, int OH, int OW, __nv_bfloat16* __restrict__ dx,
                   const __nv_bfloat16* __restrict__ mask) {
  DK_PDL_ENTER();
  const int c8n = C >> 3;
  const long total = static_cast<long>(B) * H * W * c8n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % c8n) << 3; t /= c8n;
    const int iw = static_cast<int>(t % W); t /= W;
    const int ih = static_cast<int>(t % H); t /= H;
    const int b = static_cast<int>(t);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < KH; ++kh) {
      const int ohs = ih + pad - kh;
      if (ohs < 0 || ohs % stride != 0) continue;
      const int oh = ohs / stride;
      if (oh >= OH) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int ows = iw + pad - kw;
        if (ows < 0 || ows % stride != 0) continue;
        const int ow = ows / stride;
        if (ow >= OW) continue;
        const long row = (static_cast<long>(b) * OH + oh) * OW + ow;
        const uint4 q = *reinterpret_cast<const uint4*>(col + row * ldcol + (kh * KW + kw) * C + c);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __bfloat1622float2(h[u]);
          acc[2 * u] += f.x;
          acc[2 * u + 1] += f.y;
        }
      }
    }
    if (mask != nullptr) {  // fused dReLU of the layer that produced this activation
      const uint4 mq = *reinterpret_cast<const uint4*>(mask + (i << 3));
      const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(&mq);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (!(__bfloat162float(mh[u]) > 0.f)) acc[u] = 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dx + (i << 3)) = o;
  }
}

// 8-channel vector forms of the pooling kernels (C % 8 == 0)
__global__ void __launch_bounds__(256)
maxpool_fwd_vec8_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int k, int stride,
                        int OH, int OW, __nv_bfloat16* __restrict__ y) {
  DK_PDL_ENTER();
  const int c8n = C >> 3;
  const long total = static_cast<long>(B) * OH * OW * c8n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % c8n) << 3; t /= c8n;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH); t /= OH;
    const int b = static_cast<int>(t);
    float m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = -INFINITY;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int ih = oh * stride + kh, iw = ow * stride + kw;
        if (ih < H && iw < W) {
          const uint4 q = *reinterpret_cast<const uint4*>(x + ((static_cast<long>(b) * H + ih) * W + iw) * C + c);
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
          for (int u = 0; u < 8; ++u) m[u] = fmaxf(m[u], __bfloat162float(h[u]));
        }
      }
    uint4 o;
    o.x = pack_bf16x2(m[0], m[1]); o.y = pack_bf16x2(m[2], m[3]);
    o.z = pack_bf16x2(m[4], m[5]); o.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(y + (i << 3)) = o;
  }
}

// one thread per OUTPUT window and 8 channels: writes the whole k x k input-gradient window
// (non-overlapping windows), first-max position receives the gradient
__global__ void __launch_bounds__(256)
maxpool_bwd_vec8_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int B, int H,
                        int W, int C, int k, int OH, int OW, __nv_bfloat16* __restrict__ dx, int relu) {
  DK_PDL_ENTER();
  const int c8n = C >> 3;
  const long total = static_cast<long>(B) * OH * OW * c8n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % c8n) << 3; t /= c8n;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH); t /= OH;
    const int b = static_cast<int>(t);
    float m[8];
    int arg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { m[u] = -INFINITY; arg[u] = 0; }
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const uint4 q = *reinterpret_cast<const uint4*>(
            x + ((static_cast<long>(b) * H + oh * k + kh) * W + ow * k + kw) * C + c);
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float v = __bfloat162float(h[u]);
          if (v > m[u]) { m[u] = v; arg[u] = kh * k + kw; }
        }
      }
    const uint4 gq = *reinterpret_cast<const uint4*>(dy + (i << 3));
    const __nv_bfloat16* g = reinterpret_cast<const __nv_bfloat16*>(&gq);
    // relu: x is a post-ReLU activation and dx feeds that ReLU's backward -> the mask (x > 0) is
    // applied here (the window max decides: a zero max means every input of the window was clipped)
    if (relu) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (!(m[u] > 0.f)) arg[u] = -1;
    }
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        __nv_bfloat16 o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = (arg[u] == kh * k + kw) ? g[u] : __float2bfloat16_rn(0.f);
        *reinterpret_cast<uint4*>(dx + ((static_cast<long>(b) * H + oh * k + kh) * W + ow * k + kw) * C + c) =
            *reinterpret_cast<const uint4*>(o);
      }
    // rows / columns past the last full window (H or W not a multiple of k) receive no gradient
    const uint4 z = make_uint4(0, 0, 0, 0);
    const bool tail_w = ow == OW - 1 && OW * k < W, tail_h = oh == OH - 1 && OH * k < H;
    if (tail_w)
      for (int kh = 0; kh < k; ++kh)
        for (int jw = OW * k; jw < W; ++jw)
          *reinterpret_cast<uint4*>(dx + ((static_cast<long>(b) * H + oh * k + kh) * W + jw) * C + c) = z;
    if (tail_h)
      for (int jh = OH * k; jh < H; ++jh)
        for (int kw = 0; kw < k; ++kw)
          *reinterpret_cast<uint4*>(dx + ((static_cast<long>(b) * H + jh) * W + ow * k + kw) * C + c) = z;
    if (tail_w && tail_h)
      for (int jh = OH * k; jh < H; ++jh)
        for (int jw = OW * k; jw < W; ++jw)
          *reinterpret_cast<uint4*>(dx + ((static_cast<long>(b) * H + jh) * W + jw) * C + c) = z;
  }
}

__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int k, int stride,
                   int OH, int OW, __nv_bfloat16* __restrict__ y) {
  DK_PDL_ENTER();
  const long total = static_cast<long>(B) * OH * OW * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % C); t /= C;
    const int ow = static_cast<int>(t % OW); t /= OW;
    const int oh = static_cast<int>(t % OH); t /= OH;
    const int b = static_cast<int>(t);
    float m = -INFINITY;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int ih = oh * stride + kh, iw = ow * stride + kw;
        if (ih < H && iw < W)
          m = fmaxf(m, __bfloat162float(x[((static_cast<long>(b) * H + ih) * W + iw) * C + c]));
      }
    y[i] = __float2bfloat16_rn(m);
  }
}

// dx = dy routed to the first arg-max position of each window (non-overlapping windows: k == stride)
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                   const __nv_bfloat16* __restrict__ dy, int B, int H, int W, int C, int k, int stride,
                   int OH, int OW, __nv_bfloat16* __restrict__ dx, int relu) {
  DK_PDL_ENTER();
  const long total = static_cast<long>(B) * H * W * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    long t = i;
    const int c = static_cast<int>(t % C); t /= C;
    const int iw = static_cast<int>(t % W); t /= W;
    const int ih = static_cast<int>(t % H); t /= H;
    const int b = static_cast<int>(t);
    const int oh = ih / stride, ow = iw / stride;
    float g = 0.f;
    if (oh < OH && ow < OW && ih - oh * stride < k && iw - ow * stride < k) {
      const long o = ((static_cast<long>(b) * OH + oh) * OW + ow) * C + c;
      const float yv = __bfloat162float(y[o]);
      if (__bfloat162float(x[i]) == yv) {
        // first-match tie break: only the earliest window position equal to the max gets the grad
        int fh = -1, fw = -1;
        for (int kh = 0; kh < k; ++kh)
          for (int kw = 0; kw < k; ++kw) {
            const int jh = oh * stride + kh, jw = ow * stride + kw;
            if (fh < 0 && jh < H && jw < W &&
                __bfloat162float(x[((static_cast<long>(b) * H + jh) * W + jw) * C + c]) == yv) {
              fh = jh;
              fw = jw;
            }
          }
        if (fh == ih && fw == iw && (!relu || yv > 0.f)) g = __bfloat162float(dy[o]);
      }
    }
    dx[i] = __float2bfloat16_rn(g);
  }
}

__global__ void __launch_bounds__(256)
relu_mask_kernel(__nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ act, long n) {
  DK_PDL_ENTER();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    if (!(__bfloat162float(act[i]) > 0.f)) dy[i] = __float2bfloat16_rn(0.f);
}

__global__ void __launch_bounds__(256)
add_bf16_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ a,
                const __nv_bfloat16* __restrict__ b, long n, int relu) {
  DK_PDL_ENTER();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float v = __bfloat162float(a[i]) + __bfloat162float(b[i]);
    if (relu) v = fmaxf(v, 0.f);
    dst[i] = __float2bfloat16_rn(v);
  }
}


// ------------------------------------------------------------------------------------------
// BatchNormalization over the channel (last) axis of a [rows, C] bf16 matrix, training mode
// (not in the reference's models; needed by ResNet-18, BASELINE config 5).
//   stats    : per-channel sum / sum of squares (atomics into a zeroed [2, C] fp32 buffer)
//   finalize : mean, invstd (saved for backward) + moving-average update in the fp32 master
//   apply    : y = (x - mean) * invstd * gamma + beta  (+ ReLU)
//   bwd_reduce: sum(dy'), sum(dy' * xhat) with dy' = dy masked by the ReLU output
//   bwd_apply : dx = gamma * invstd * (dy' - sum(dy')/n - xhat * sum(dy' xhat)/n)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long rows, int C, float* __restrict__ sums,
                int rows_per_block) {
  DK_PDL_ENTER();
  __shared__ float red[2][8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const long r0 = static_cast<long>(blockIdx.y) * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (c + 1 < C) {
    for (long r = r0 + warp; r < r1; r += 8) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * C + c));
      s0 += f.x; s1 += f.y; q0 += f.x * f.x; q1 += f.y * f.y;
    }
  }
  red[0][warp][lane * 2] = s0; red[0][warp][lane * 2 + 1] = s1;
  red[1][warp][lane * 2] = q0; red[1][warp][lane * 2 + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, col = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[which][w][col];
    if (blockIdx.x * 64 + col < C) atomicAdd(sums + which * C + blockIdx.x * 64 + col, t);
  }
}

__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ sums, long rows, int C, float eps, float momentum,
                   float* __restrict__ saved_mean, float* __restrict__ saved_invstd,
                   float* __restrict__ moving_mean, float* __restrict__ moving_var) {
  DK_PDL_ENTER();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_n = 1.f / static_cast<float>(rows);
  const float mean = sums[c] * inv_n;
  const float var = fmaxf(sums[C + c] * inv_n - mean * mean, 0.f);
  saved_mean[c] = mean;
  saved_invstd[c] = rsqrtf(var + eps);
  if (moving_mean != nullptr) {
    moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * mean;
    moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * var;
  }
}

// mode 0: training (saved batch statistics); mode 1: inference (moving statistics, mean/var given)
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long rows, int C, const float* __restrict__ mean,
                const float* __restrict__ invstd_or_var, const float* __restrict__ gamma,
                const float* __restrict__ beta, int relu, int inference, float eps,
                __nv_bfloat16* __restrict__ y) {
  DK_PDL_ENTER();
  const int c8n = C >> 3;
  const long total = rows * c8n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8n) << 3;
    const uint4 q = *reinterpret_cast<const uint4*>(x + (i << 3));
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float is = inference ? rsqrtf(invstd_or_var[c + u] + eps) : invstd_or_var[c + u];
      float v = (__bfloat162float(h[u]) - mean[c + u]) * is * gamma[c + u] + beta[c + u];
      o[u] = relu ? fmaxf(v, 0.f) : v;
    }
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(y + (i << 3)) = w;
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const __nv_bfloat16* __restrict__ y_relu, long rows, int C,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     float* __restrict__ sums, int rows_per_block) {
  DK_PDL_ENTER();
  __shared__ float red[2][8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const long r0 = static_cast<long>(blockIdx.y) * rows_per_block;
  const long r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  if (c + 1 < C) {
    const float m0 = mean[c], m1 = mean[c + 1], i0 = invstd[c], i1 = invstd[c + 1];
    for (long r = r0 + warp; r < r1; r += 8) {
      float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dy + r * C + c));
      if (y_relu != nullptr) {
        const float2 yy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(y_relu + r * C + c));
        if (!(yy.x > 0.f)) g.x = 0.f;
        if (!(yy.y > 0.f)) g.y = 0.f;
      }
      const float2 xv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * C + c));
      s0 += g.x; s1 += g.y;
      q0 += g.x * (xv.x - m0) * i0; q1 += g.y * (xv.y - m1) * i1;
    }
  }
  red[0][warp][lane * 2] = s0; red[0][warp][lane * 2 + 1] = s1;
  red[1][warp][lane * 2] = q0; red[1][warp][lane * 2 + 1] = q1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, col = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[which][w][col];
    if (blockIdx.x * 64 + col < C) atomicAdd(sums + which * C + blockIdx.x * 64 + col, t);
  }
}

// dx, and the parameter gradients dgamma = sum(dy' xhat), dbeta = sum(dy') (written by block 0)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ y_relu, long rows, int C,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ sums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx) {
  DK_PDL_ENTER();
  const int c8n = C >> 3;
  const long total = rows * c8n;
  const float inv_n = 1.f / static_cast<float>(rows);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = sums[c];
      dgamma[c] = sums[C + c];
    }
  }
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8n) << 3;
    const uint4 qg = *reinterpret_cast<const uint4*>(dy + (i << 3));
    const uint4 qx = *reinterpret_cast<const uint4*>(x + (i << 3));
    uint4 qy = make_uint4(0, 0, 0, 0);
    if (y_relu != nullptr) qy = *reinterpret_cast<const uint4*>(y_relu + (i << 3));
    const __nv_bfloat16* hg = reinterpret_cast<const __nv_bfloat16*>(&qg);
    const __nv_bfloat16* hx = reinterpret_cast<const __nv_bfloat16*>(&qx);
    const __nv_bfloat16* hy = reinterpret_cast<const __nv_bfloat16*>(&qy);
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float g = __bfloat162float(hg[u]);
      if (y_relu != nullptr && !(__bfloat162float(hy[u]) > 0.f)) g = 0.f;
      const float xhat = (__bfloat162float(hx[u]) - mean[c + u]) * invstd[c + u];
      o[u] = gamma[c + u] * invstd[c + u] * (g - sums[c + u] * inv_n - xhat * sums[C + c + u] * inv_n);
    }
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(dx + (i << 3)) = w;
  }
}

// global average pooling over the spatial positions: x [B, P, C] -> y [B, C] (and its backward)
__global__ void __launch_bounds__(256)
gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, int B, int P, int C, __nv_bfloat16* __restrict__ y) {
  DK_PDL_ENTER();
  const long total = static_cast<long>(B) * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long b = i / C;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += __bfloat162float(x[(b * P + p) * C + c]);
    y[i] = __float2bfloat16_rn(acc / static_cast<float>(P));
  }
}

__global__ void __launch_bounds__(256)
gap_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int B, int P, int C, __nv_bfloat16* __restrict__ dx) {
  DK_PDL_ENTER();
  const long total = static_cast<long>(B) * P * C;
  const float inv = 1.f / static_cast<float>(P);
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long b = i / (static_cast<long>(P) * C);
    dx[i] = __float2bfloat16_rn(__bfloat162float(dy[b * C + c]) * inv);
  }
}

static inline int ew_grid(long n) {
  long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return static_cast<int>(b);
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_input_stage(const void* x, int in_dtype, int B, int F, float scale, float shift, void* xb,
                   int ldx, void* xt, int ldxt, int* step_counter, void* xf, int ldxf, void* stream) {
  dim3 grid((F + 31) / 32, (B + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* xbp = reinterpret_cast<__nv_bfloat16*>(xb);
  __nv_bfloat16* xtp = reinterpret_cast<__nv_bfloat16*>(xt);
  const int in_size = in_dtype == DK_IN_U8 ? 1 : (in_dtype == DK_IN_F32 ? 4 : 2);
  if (xf != nullptr && !(xt == nullptr && xb != nullptr && F % 8 == 0 && ldx % 8 == 0 && ldxf % 4 == 0)) return -7;
  if (xt == nullptr && xb != nullptr && F % 8 == 0 && ldx % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(x) % (8 * in_size > 16 ? 16 : 8 * in_size)) == 0) {
    long total = static_cast<long>(B) * (F / 8);
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const int g = static_cast<int>(blocks < 1 ? 1 : blocks);
    if (in_dtype == DK_IN_U8)
      DK_HOST_CHECK(DK_LAUNCH(input_stage_vec_kernel<uint8_t>, g, 256, 0, st, reinterpret_cast<const uint8_t*>(x), B, F, scale, shift,
                                                         xbp, ldx, reinterpret_cast<float*>(xf), ldxf, step_counter));
    else if (in_dtype == DK_IN_F32)
      DK_HOST_CHECK(DK_LAUNCH(input_stage_vec_kernel<float>, g, 256, 0, st, reinterpret_cast<const float*>(x), B, F, scale, shift, xbp,
                                                       ldx, reinterpret_cast<float*>(xf), ldxf, step_counter));
    else
      DK_HOST_CHECK(DK_LAUNCH(input_stage_vec_kernel<__nv_bfloat16>, g, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(x), B, F, scale,
                                                               shift, xbp, ldx, reinterpret_cast<float*>(xf), ldxf, step_counter));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  if (in_dtype == DK_IN_U8)
    DK_HOST_CHECK(DK_LAUNCH(input_stage_kernel<uint8_t>, grid, 256, 0, st, reinterpret_cast<const uint8_t*>(x), B, F, scale,
                                                      shift, xbp, ldx, xtp, ldxt, step_counter));
  else if (in_dtype == DK_IN_F32)
    DK_HOST_CHECK(DK_LAUNCH(input_stage_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float*>(x), B, F, scale,
                                                    shift, xbp, ldx, xtp, ldxt, step_counter));
  else
    DK_HOST_CHECK(DK_LAUNCH(input_stage_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(x),
                                                            B, F, scale, shift, xbp, ldx, xtp, ldxt,
                                                            step_counter));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_transpose_bf16(const void* src, int rows, int cols, int lds, void* dst, int ldd, void* stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  DK_HOST_CHECK(DK_LAUNCH(transpose_bf16_kernel, grid, 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(src), rows, cols, lds, reinterpret_cast<__nv_bfloat16*>(dst),
      ldd));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_rowsum_bf16(const void* src, int rows, int cols, int lds, float* out, float scale, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(rowsum_bf16_kernel, (rows + 7) / 8, 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(src), rows, cols, lds, out, scale));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_colsum_bf16(const void* src, int rows, int cols, int lds, float* out, float scale, void* stream) {
  const int colblocks = (cols + 63) / 64;
  int splits = (rows + 255) / 256;
  int cap = (148 * 8) / colblocks;  // fill the GPU even when the matrix is narrow
  if (cap < 64) cap = 64;
  if (splits > cap) splits = cap;
  const int rpb = (rows + splits - 1) / splits;
  dim3 grid((cols + 63) / 64, splits);
  DK_HOST_CHECK(DK_LAUNCH(colsum_bf16_kernel, grid, 256, 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(src), rows,
                                                            cols, lds, out, scale, rpb));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_im2col(const void* x, int B, int H, int W, int C, int KH, int KW, int stride, int pad, int OH,
              int OW, void* col, int ldcol, void* stream) {
  if (C % 8 != 0 && ldcol % 8 == 0 && (reinterpret_cast<uintptr_t>(col) & 15) == 0) {
    const long total8 = static_cast<long>(B) * OH * OW * (ldcol / 8);
    DK_HOST_CHECK(DK_LAUNCH(im2col_gather8_kernel, ew_grid(total8), 256, 0, (cudaStream_t)stream,
        reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, C, KH, KW, stride, pad, OH, OW,
        reinterpret_cast<__nv_bfloat16*>(col), ldcol));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  const int cvec = (C % 8 == 0) ? 8 : 1;
  const long total = static_cast<long>(B) * OH * OW * KH * KW * (C / cvec);
  DK_HOST_CHECK(DK_LAUNCH(im2col_kernel, ew_grid(total), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, C, KH, KW, stride, pad, OH, OW,
      reinterpret_cast<__nv_bfloat16*>(col), ldcol));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_col2im(const void* col, int ldcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
              int OH, int OW, void* dx, void* stream) {
  return dk_col2im_ex(col, ldcol, B, H, W, C, KH, KW, stride, pad, OH, OW, dx, nullptr, stream);
}

// mask (optional, [B*H*W, C] bf16): dx is zeroed where mask <= 0 (dReLU of the producing layer)
int dk_col2im_ex(const void* col, int ldcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
                 int OH, int OW, void* dx, const void* mask, void* stream) {
  if (C % 8 == 0 && ldcol % 8 == 0) {
    const long total8 = static_cast<long>(B) * H * W * (C / 8);
    DK_HOST_CHECK(DK_LAUNCH(col2im_vec8_kernel, ew_grid(total8), 256, 0, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(col), ldcol, B, H, W, C, KH, KW, stride, pad, OH, OW,
        reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(mask)));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  const long total = static_cast<long>(B) * H * W * C;
  DK_HOST_CHECK(DK_LAUNCH(col2im_kernel, ew_grid(total), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(col), ldcol, B, H, W, C, KH, KW, stride, pad, OH, OW,
      reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(mask)));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_maxpool_fwd(const void* x, int B, int H, int W, int C, int k, int stride, void* y, void* stream) {
  const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
  if (C % 8 == 0) {
    const long total8 = static_cast<long>(B) * OH * OW * (C / 8);
    DK_HOST_CHECK(DK_LAUNCH(maxpool_fwd_vec8_kernel, ew_grid(total8), 256, 0, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, C, k, stride, OH, OW, reinterpret_cast<__nv_bfloat16*>(y)));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  const long total = static_cast<long>(B) * OH * OW * C;
  DK_HOST_CHECK(DK_LAUNCH(maxpool_fwd_kernel, ew_grid(total), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, C, k, stride, OH, OW,
      reinterpret_cast<__nv_bfloat16*>(y)));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_maxpool_bwd(const void* x, const void* y, const void* dy, int B, int H, int W, int C, int k,
                   int stride, void* dx, void* stream) {
  return dk_maxpool_bwd_ex(x, y, dy, B, H, W, C, k, stride, dx, 0, stream);
}

// relu != 0: additionally apply the (x > 0) mask of the ReLU that produced x (fused dReLU)
int dk_maxpool_bwd_ex(const void* x, const void* y, const void* dy, int B, int H, int W, int C, int k,
                      int stride, void* dx, int relu, void* stream) {
  const int OH = (H - k) / stride + 1, OW = (W - k) / stride + 1;
  if (C % 8 == 0 && k == stride) {
    const long total8 = static_cast<long>(B) * OH * OW * (C / 8);
    DK_HOST_CHECK(DK_LAUNCH(maxpool_bwd_vec8_kernel, ew_grid(total8), 256, 0, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), B, H, W, C, k, OH, OW,
        reinterpret_cast<__nv_bfloat16*>(dx), relu));
    DK_HOST_CHECK(cudaGetLastError());
    return 0;
  }
  const long total = static_cast<long>(B) * H * W * C;
  DK_HOST_CHECK(DK_LAUNCH(maxpool_bwd_kernel, ew_grid(total), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(y),
      reinterpret_cast<const __nv_bfloat16*>(dy), B, H, W, C, k, stride, OH, OW,
      reinterpret_cast<__nv_bfloat16*>(dx), relu));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}


static inline void stat_grid(long rows, int C, dim3* grid, int* rpb) {
  const int colblocks = (C + 63) / 64;
  long splits = (rows + 255) / 256;
  long cap = (148 * 8) / colblocks;
  if (cap < 64) cap = 64;
  if (splits > cap) splits = cap;
  if (splits < 1) splits = 1;
  *rpb = static_cast<int>((rows + splits - 1) / splits);
  *grid = dim3(colblocks, static_cast<unsigned>(splits));
}

// training forward: stats (sums must be zeroed) -> finalize -> apply, three launches
int dk_bn_forward(const void* x, long rows, int C, float* sums, float* saved_mean, float* saved_invstd,
                  float* moving_mean, float* moving_var, const float* gamma, const float* beta, float eps,
                  float momentum, int relu, void* y, void* stream) {
  if (C % 8 != 0) return -1;
  dim3 grid; int rpb;
  stat_grid(rows, C, &grid, &rpb);
  DK_HOST_CHECK(DK_LAUNCH(bn_stats_kernel, grid, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), rows, C,
                          sums, rpb));
  DK_HOST_CHECK(DK_LAUNCH(bn_finalize_kernel, (C + 255) / 256, 256, 0, stream, (const float*)sums, rows, C, eps,
                          momentum, saved_mean, saved_invstd, moving_mean, moving_var));
  DK_HOST_CHECK(DK_LAUNCH(bn_apply_kernel, ew_grid(rows * (C / 8)), 256, 0, stream,
                          reinterpret_cast<const __nv_bfloat16*>(x), rows, C, (const float*)saved_mean,
                          (const float*)saved_invstd, gamma, beta, relu, 0, eps, reinterpret_cast<__nv_bfloat16*>(y)));
  return 0;
}

int dk_bn_inference(const void* x, long rows, int C, const float* moving_mean, const float* moving_var,
                    const float* gamma, const float* beta, float eps, int relu, void* y, void* stream) {
  if (C % 8 != 0) return -1;
  DK_HOST_CHECK(DK_LAUNCH(bn_apply_kernel, ew_grid(rows * (C / 8)), 256, 0, stream,
                          reinterpret_cast<const __nv_bfloat16*>(x), rows, C, moving_mean, moving_var, gamma, beta,
                          relu, 1, eps, reinterpret_cast<__nv_bfloat16*>(y)));
  return 0;
}

// backward: reduce (sums zeroed) -> apply; y_relu (optional) is the ReLU output used as mask
int dk_bn_backward(const void* dy, const void* x, const void* y_relu, long rows, int C, const float* saved_mean,
                   const float* saved_invstd, const float* gamma, float* sums, float* dgamma, float* dbeta, void* dx,
                   void* stream) {
  if (C % 8 != 0) return -1;
  dim3 grid; int rpb;
  stat_grid(rows, C, &grid, &rpb);
  DK_HOST_CHECK(DK_LAUNCH(bn_bwd_reduce_kernel, grid, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(dy),
                          reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(y_relu),
                          rows, C, saved_mean, saved_invstd, sums, rpb));
  DK_HOST_CHECK(DK_LAUNCH(bn_bwd_apply_kernel, ew_grid(rows * (C / 8)), 256, 0, stream,
                          reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x),
                          reinterpret_cast<const __nv_bfloat16*>(y_relu), rows, C, saved_mean, saved_invstd, gamma,
                          (const float*)sums, dgamma, dbeta, reinterpret_cast<__nv_bfloat16*>(dx)));
  return 0;
}

int dk_gap_fwd(const void* x, int B, int P, int C, void* y, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(gap_fwd_kernel, ew_grid(static_cast<long>(B) * C), 256, 0, stream,
                          reinterpret_cast<const __nv_bfloat16*>(x), B, P, C, reinterpret_cast<__nv_bfloat16*>(y)));
  return 0;
}

int dk_gap_bwd(const void* dy, int B, int P, int C, void* dx, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(gap_bwd_kernel, ew_grid(static_cast<long>(B) * P * C), 256, 0, stream,
                          reinterpret_cast<const __nv_bfloat16*>(dy), B, P, C, reinterpret_cast<__nv_bfloat16*>(dx)));
  return 0;
}

int dk_relu_mask_bf16(void* dy, const void* act, long n, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(relu_mask_kernel, ew_grid(n), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<__nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(act), n));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

int dk_add_bf16(void* dst, const void* a, const void* b, long n, int relu, void* stream) {
  DK_HOST_CHECK(DK_LAUNCH(add_bf16_kernel, ew_grid(n), 256, 0, (cudaStream_t)stream, 
      reinterpret_cast<__nv_bfloat16*>(dst), reinterpret_cast<const __nv_bfloat16*>(a),
      reinterpret_cast<const __nv_bfloat16*>(b), n, relu));
  DK_HOST_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
