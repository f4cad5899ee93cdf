// Standalone correctness + timing harness for the tcgen05 GEMM (no Python needed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. gemm_test.cu ../gemm_tcgen05.cu -o gemm_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../gemm.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);         \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

__global__ void ref_gemm(const float* A, const float* B, float* D, int M, int N, int K,
                         const float* bias, int bias_m, int act, const __nv_bfloat16* mask) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += A[(size_t)m * K + k] * B[(size_t)n * K + k];
  if (bias) acc += bias_m ? bias[m] : bias[n];
  if (act) acc = fmaxf(acc, 0.f);
  if (mask && !(__bfloat162float(mask[(size_t)m * N + n]) > 0.f)) acc = 0.f;
  D[(size_t)m * N + n] = acc;
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int run_case(int M, int N, int K, int flags, int mode, int bn, bool timing, int splits = 1) {
  const int tf32 = flags & DK_GEMM_TF32, amn = (flags & DK_GEMM_A_MN) != 0, bmn = (flags & DK_GEMM_B_MN) != 0;
  // mode 0: plain bf16 out; 1: bias_n+relu bf16 out + transposed; 2: mask + fp32 out; 3: bias_m fp32
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hbias(M > N ? M : N);
  std::vector<__nv_bfloat16> hAb(hA.size()), hBb(hB.size()), hmask((size_t)M * N);
  for (size_t i = 0; i < hA.size(); ++i) {
    hAb[i] = __float2bfloat16(frand());
    hA[i] = tf32 ? frand() : __bfloat162float(hAb[i]);
  }
  for (size_t i = 0; i < hB.size(); ++i) {
    hBb[i] = __float2bfloat16(frand());
    hB[i] = tf32 ? frand() : __bfloat162float(hBb[i]);
  }
  for (auto& b : hbias) b = frand();
  for (auto& m : hmask) m = __float2bfloat16(frand());
  std::vector<__nv_bfloat16> hAmn(hA.size()), hBmn(hB.size());
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hAmn[(size_t)k * M + m] = hAb[(size_t)m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hBmn[(size_t)k * N + n] = hBb[(size_t)n * K + k];
  float *dA, *dB, *dRef, *dbias, *dOutF;
  __nv_bfloat16 *dAb, *dBb, *dmask, *dOutB, *dOutT;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dAb, hA.size() * 2));
  CK(cudaMalloc(&dBb, hB.size() * 2));
  CK(cudaMalloc(&dRef, (size_t)M * N * 4));
  CK(cudaMalloc(&dOutF, (size_t)M * N * 4));
  CK(cudaMalloc(&dOutB, (size_t)M * N * 2));
  CK(cudaMalloc(&dOutT, (size_t)M * N * 2));
  CK(cudaMalloc(&dmask, (size_t)M * N * 2));
  CK(cudaMalloc(&dbias, hbias.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAb, amn ? hAmn.data() : hAb.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBb, bmn ? hBmn.data() : hBb.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dmask, hmask.data(), hmask.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOutF, 0, (size_t)M * N * 4));
  CK(cudaMemset(dOutB, 0, (size_t)M * N * 2));
  CK(cudaMemset(dOutT, 0, (size_t)M * N * 2));

  DkGemmEpilogue ep = {};
  ep.alpha = 1.f;
  bool out_f32 = false, has_t = false;
  if (mode == 0) {
    ep.d = dOutB; ep.ldd = N;
  } else if (mode == 1) {
    ep.bias = dbias; ep.act = 1; ep.d = dOutB; ep.ldd = N; ep.dt = dOutT; ep.lddt = M; has_t = true;
  } else if (mode == 2) {
    ep.mask = dmask; ep.ld_mask = N; ep.d = dOutF; ep.ldd = N; ep.d_fp32 = 1; out_f32 = true;
  } else if (mode == 4) {
    ep.d = dOutF; ep.ldd = N; ep.d_fp32 = 1; out_f32 = true;
  } else {
    ep.bias = dbias; ep.bias_along_m = 1; ep.d = dOutF; ep.ldd = N; ep.d_fp32 = 1; out_f32 = true;
  }
  dim3 rg((N + 127) / 128, M);
  ref_gemm<<<rg, 128>>>(dA, dB, dRef, M, N, K, ep.bias, ep.bias_along_m, ep.act, ep.mask);
  CK(cudaGetLastError());
  const void* Ap = tf32 ? (const void*)dA : (const void*)dAb;
  const void* Bp = tf32 ? (const void*)dB : (const void*)dBb;
  const long lda = amn ? M : K, ldb = bmn ? N : K;
  if (splits != 1) ep.accumulate = 0;
  int r = dk_gemm_tn_ex(Ap, lda, Bp, ldb, &ep, M, N, K, flags, bn, splits, 0);
  if (r != 0) {
    printf("  launch failed r=%d\n", r);
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel failed: %s\n", cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> ref((size_t)M * N), outf((size_t)M * N);
  std::vector<__nv_bfloat16> outb((size_t)M * N), outt((size_t)M * N);
  CK(cudaMemcpy(ref.data(), dRef, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(outf.data(), dOutF, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(outb.data(), dOutB, ref.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(outt.data(), dOutT, ref.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  size_t bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float rv = ref[(size_t)m * N + n];
      float ov = out_f32 ? outf[(size_t)m * N + n] : __bfloat162float(outb[(size_t)m * N + n]);
      double tol = (out_f32 ? (tf32 ? 2e-2 : 3e-4) : 1e-2) * (fabs(rv) + sqrt((double)K) * 0.05);
      double err = fabs(rv - ov);
      if (err > tol) ++bad;
      if (has_t) {
        float tv = __bfloat162float(outt[(size_t)n * M + m]);
        if (tv != __bfloat162float(outb[(size_t)m * N + n])) ++bad;
      }
      if (err > max_err) max_err = err;
      if (fabs(rv) > max_ref) max_ref = fabs(rv);
    }
  printf("M=%5d N=%5d K=%5d flags=%d mode=%d bn=%3d  max_err=%.4g max_ref=%.4g bad=%zu %s\n", M, N,
         K, flags, mode, bn, max_err, max_ref, bad, bad ? "FAIL" : "ok");
  if (timing && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 5; ++i) { if (splits != 1) cudaMemsetAsync(dOutF, 0, (size_t)M * N * 4, 0); dk_gemm_tn_ex(Ap, lda, Bp, ldb, &ep, M, N, K, flags, bn, splits, 0); }
    const int iters = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) { if (splits != 1) cudaMemsetAsync(dOutF, 0, (size_t)M * N * 4, 0); dk_gemm_tn_ex(Ap, lda, Bp, ldb, &ep, M, N, K, flags, bn, splits, 0); }
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double us = ms * 1e3 / iters;
    printf("    time %.2f us  %.1f TFLOP/s\n", us, 2.0 * M * N * K / us * 1e-6);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dAb); cudaFree(dBb); cudaFree(dRef); cudaFree(dOutF);
  cudaFree(dOutB); cudaFree(dOutT); cudaFree(dmask); cudaFree(dbias);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  int fails = 0;
  const bool quick = argc > 1;
  fails += run_case(128, 128, 64, 0, 0, 128, false);
  fails += run_case(256, 256, 512, 0, 0, 128, false);
  fails += run_case(128, 16, 128, 0, 0, 16, false);
  // MN-major operands: B only (dgrad form), both (wgrad form), A only
  fails += run_case(128, 128, 64, DK_GEMM_B_MN, 0, 128, false);
  fails += run_case(128, 64, 64, DK_GEMM_B_MN, 0, 64, false);
  fails += run_case(128, 128, 256, DK_GEMM_B_MN, 2, 128, false);
  fails += run_case(128, 128, 64, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, false);
  fails += run_case(256, 256, 512, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, false);
  fails += run_case(128, 128, 128, DK_GEMM_A_MN, 2, 128, false);
  fails += run_case(1024, 1000, 200, DK_GEMM_B_MN, 2, 128, true);     // dgrad layer 2
  fails += run_case(1000, 784, 1024, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, true);  // wgrad layer 1
  fails += run_case(200, 1000, 1024, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, true);  // wgrad layer 2
  fails += run_case(16, 200, 1024, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, true);    // wgrad layer 3 (padded)
  fails += run_case(1024, 200, 16, DK_GEMM_B_MN, 2, 128, true);       // dgrad layer 3 (K padded)
  fails += run_case(304, 72, 136, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 64, false);
  // split-K (TMA add-reduction into a zeroed fp32 output), mode 4 = plain fp32
  fails += run_case(1000, 784, 4096, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 128, true, 6);
  fails += run_case(200, 1000, 4096, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 128, true, 16);
  fails += run_case(16, 200, 4096, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 128, true, 16);
  fails += run_case(32, 72, 4096, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 64, true, 8);
  fails += run_case(256, 256, 1024, 0, 4, 128, false, 4);
  // ragged shapes (MNIST MLP dims), fused epilogues
  fails += run_case(1024, 1000, 784, 0, 1, 128, true);
  fails += run_case(1024, 200, 1000, 0, 1, 64, true);
  fails += run_case(1024, 10, 200, 0, 3, 16, true);
  fails += run_case(37, 1000, 784, 0, 1, 128, false);
  fails += run_case(300, 77, 136, 0, 2, 128, false);
  // tf32
  fails += run_case(128, 128, 64, DK_GEMM_TF32, 2, 128, false);
  fails += run_case(1000, 256, 784, DK_GEMM_TF32, 3, 128, true);
  // cta_group::2 CTA-pair kernel
  fails += run_case(256, 256, 64, DK_GEMM_PAIR, 0, 256, false);
  fails += run_case(256, 256, 512, DK_GEMM_PAIR, 0, 256, false);
  fails += run_case(512, 384, 320, DK_GEMM_PAIR, 0, 128, false);
  fails += run_case(1000, 1000, 784, DK_GEMM_PAIR, 0, 256, true);
  fails += run_case(300, 200, 136, DK_GEMM_PAIR, 2, 256, false);
  fails += run_case(256, 256, 128, DK_GEMM_PAIR | DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, false);
  fails += run_case(256, 256, 128, DK_GEMM_PAIR | DK_GEMM_B_MN, 4, 256, false);
  fails += run_case(1000, 784, 4096, DK_GEMM_PAIR | DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, true, 4);
  fails += run_case(1000, 784, 16384, DK_GEMM_PAIR | DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, true, 4);
  fails += run_case(1000, 784, 16384, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, true, 4);
  fails += run_case(200, 1000, 16384, DK_GEMM_PAIR | DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, true, 18);
  fails += run_case(200, 1000, 16384, DK_GEMM_A_MN | DK_GEMM_B_MN, 4, 256, true, 18);
  if (!quick) {
    fails += run_case(8192, 8192, 8192, DK_GEMM_PAIR, 0, 256, true);
    fails += run_case(16384, 1000, 784, DK_GEMM_PAIR, 0, 256, true);
    fails += run_case(8192, 8192, 8192, 0, 0, 256, true);
    fails += run_case(8192, 1000, 784, 0, 1, 128, true);
    fails += run_case(8192, 1000, 784, 0, 0, 128, true);
    fails += run_case(1000, 784, 8192, DK_GEMM_A_MN | DK_GEMM_B_MN, 2, 128, true);
  }
  printf("FAILS=%d\n", fails);
  return fails ? 1 : 0;
}
