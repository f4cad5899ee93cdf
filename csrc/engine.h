// C ABI of the native op-list executor (engine.cu).
#pragma once
#include <stdint.h>

#include "dense_fused.h"
#include "gemm.h"

#define DK_OP_MAX_I 20
#define DK_OP_MAX_F 8
#define DK_ENGINE_SIDE_STREAMS 3
#define DK_ENGINE_SLOTS 16

enum {
  DK_OP_INPUT = 0,
  DK_OP_GEMM = 1,
  DK_OP_XENT = 2,
  DK_OP_ROWSUM = 3,
  DK_OP_TRANSPOSE = 4,
  DK_OP_OPTIM = 5,
  DK_OP_IM2COL = 6,
  DK_OP_COL2IM = 7,
  DK_OP_MAXPOOL_FWD = 8,
  DK_OP_MAXPOOL_BWD = 9,
  DK_OP_RELU_MASK = 10,
  DK_OP_ADD = 11,
  DK_OP_MEMSET = 12,
  DK_OP_PS_COMMIT = 13,
  DK_OP_PS_PULL = 14,
  DK_OP_PS_EXCHANGE = 15,
  DK_OP_PS_ELASTIC = 16,
  DK_OP_PS_DAMPED = 17,
  DK_OP_PS_TICKET = 18,
  DK_OP_LOCK_ACQUIRE = 19,
  DK_OP_LOCK_RELEASE = 20,
  DK_OP_EAMSGD_PRE = 21,
  DK_OP_EAMSGD_POST = 22,
  DK_OP_CAST = 23,
  DK_OP_ELOSS = 24,
  DK_OP_MEMCPY = 25,
  DK_OP_LABEL_INDEX = 26,
  DK_OP_COLSUM = 27,
  DK_OP_MEMCPY2D = 28,
  DK_OP_FORK = 29,
  DK_OP_JOIN = 30,
  DK_OP_GEMM_PULL = 31,
  DK_OP_BN_FWD = 32,
  DK_OP_BN_INF = 33,
  DK_OP_BN_BWD = 34,
  DK_OP_GAP_FWD = 35,
  DK_OP_GAP_BWD = 36,
  DK_OP_HEAD = 37,
  DK_OP_CONV_GEMM = 38,
  DK_OP_WFLIP = 39,
  DK_OP_CONV_WGRAD = 40,  // experimental
  DK_OP_BWD_UPDATE = 41,  // fused wgrad + bias-grad + optimizer (+ PS exchange) of every dense layer
  DK_OP_CONV_WGRAD_TMA = 42  // convolution weight (+ bias) gradient, im2col operand produced by TMA
};

#ifdef __cplusplus
extern "C" {
#endif

void* dk_engine_create();
void dk_engine_destroy(void* h);
int dk_engine_new_list(void* h);
int dk_engine_set_build_stream(void* h, int id);
int dk_engine_clear_list(void* h, int list);
int dk_engine_set_slot(void* h, int slot, void* p);
// pointer arguments: a device address, or -(slot + 1) to read the pointer from a slot at run time
int dk_engine_add_op(void* h, int list, int kind, const int64_t* iargs, int ni, const double* fargs,
                     int nf);
int dk_engine_add_gemm(void* h, int list, const void* A, long lda, const void* B, long ldb, int M, int N,
                       int K, int flags, int bn, int splits, const DkGemmEpilogue* ep);
// same, with the A operand base read from engine slot `a_slot` every time the list runs (the tensor map is
// re-encoded at enqueue time; inside a captured graph that is capture time only)
int dk_engine_add_gemm_slot(void* h, int list, int a_slot, long lda, const void* B, long ldb, int M, int N, int K,
                            int flags, int bn, int splits, const DkGemmEpilogue* ep);
// fused dense backward-update (dense_fused.cu); layers with x_slot >= 0 take their input from that slot
int dk_engine_add_bwd_update(void* h, int list, const DkBwdUpdateDesc* desc);
int dk_engine_add_gemm_pull(void* h, int list, const void* X, long ldx, const void* center_w, long ldc, int M, int N,
                            int K, void* w_local, void* w1_local, void* wb_local, const DkGemmEpilogue* ep);
// implicit-GEMM convolution (forward or dgrad form; see conv_gemm_kernel in gemm_tcgen05.cu)
int dk_engine_add_conv_gemm(void* h, int list, const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW,
                            int mul, int off, int div, const void* Bmat, long ldb, int M, int N, int K,
                            const DkGemmEpilogue* ep);
// EXPERIMENTAL implicit wgrad (conv_wgrad_kernel): dW [Cout, KH KW C] fp32 (zeroed) += dZ^T [rows, Cout] * gather(src)
int dk_engine_add_conv_wgrad(void* h, int list, const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW,
                             int stride, int pad, const void* dz, long lddz, float* dw, long lddw, int Cout, int rows);
// TMA-im2col weight gradient: dW [Cout, KH KW C] fp32 (zeroed) += dZ^T im2col(src); bias_grad (zeroed, may be NULL) += colsum(dZ).
// Adds one op per block of (tap, channel-chunk) units that fits TMEM / shared memory.
int dk_engine_add_conv_wgrad_tma(void* h, int list, const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH,
                                 int KW, int stride, int pad, const void* dz, long lddz, float* dw, long lddw, int Cout,
                                 float* bias_grad);
int dk_engine_run(void* h, int list, void* stream);
int dk_engine_list_size(void* h, int list);
int dk_engine_list_kernels(void* h, int list);
long dk_engine_launches(void* h);

// from fabric.cu
int dk_memcpy_async(void* dst, const void* src, long bytes, int kind, void* stream);
int dk_memset_async(void* dst, int value, long bytes, void* stream);
int dk_memcpy2d_async(void* dst, long dpitch, const void* src, long spitch, long width, long height, void* stream);

#ifdef __cplusplus
}
#endif
