// C ABI of the tcgen05 GEMM (see gemm_tcgen05.cu).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_bf16.h>
typedef __nv_bfloat16 dk_bf16;
#else
typedef uint16_t dk_bf16;
#endif

enum { DK_BF16 = 0, DK_F32 = 1 };
enum { DK_GEMM_TF32 = 1, DK_GEMM_A_MN = 2, DK_GEMM_B_MN = 4, DK_GEMM_PERSISTENT = 8, DK_GEMM_PAIR = 16,
       DK_GEMM_SHORT_A = 32 /* K-major A of one short M tile: TMA box of ceil8(M) rows (plain kernel only) */,
       DK_GEMM_MCAST_A = 64 /* with SHORT_A: the CTAs of a thread-block cluster (along N) each load 1/cluster of the A
                               k-block and multicast it to the others (dk_gemm_mcast_cluster(M) CTAs, A tensor map encoded
                               with dk_gemm_mcast_box_rows(M) rows) */ };
// With DK_GEMM_SHORT_A, bits 8..15 of `flags` may carry the height of the M tile (a multiple of 8, < 128): the GEMM then
// runs ceil(M / rows) CTAs along M, each loading only its `rows` rows of A -- at the reference's batch sizes the forward
// GEMMs are bound by the bytes one SM can ingest, so splitting the 64-row batch over 2-4 CTAs shortens the K loop.
#define DK_GEMM_TILE_ROWS(r) (((r) & 0xFF) << 8)
#define DK_GEMM_TILE_ROWS_OF(flags) (((flags) >> 8) & 0xFF)
// Bits 16..19: k-blocks per TMA request (plain kernel, K-major bf16 operands, bn <= 32).  A cp.async.bulk.tensor costs
// ~190 cycles of TMA-unit time whatever its size (tools/microbench/tma_request.cu), so a short-M GEMM that issues one A and
// one B request per 64-wide k-block is request-bound; with kch > 1 a pipeline stage holds kch k-blocks, filled by ONE 3-D
// request per operand (tensor viewed as [64 elements, rows, k-chunks]; maps passed in the tmap_d / tmap_m arguments, encoded
// by dk_tmap_encode_kchunks); the K tail (a partial last chunk cannot be expressed in that view) falls back to 2-D requests.
#define DK_GEMM_KCH(n) (((n) & 0xF) << 16)
#define DK_GEMM_KCH_OF(flags) (((flags) >> 16) & 0xF)

// Fused epilogue description: out = mask( act( alpha * acc + bias ) )
typedef struct DkGemmEpilogue {
  const float* bias;     // nullptr, or [N] (bias_along_m == 0) / [M] (bias_along_m == 1)
  int bias_along_m;
  int act;               // 0 = identity, 1 = ReLU
  const dk_bf16* mask;   // nullptr, or [M, ld_mask]: out = 0 where mask <= 0 (dReLU)
  int ld_mask;
  void* d;               // row-major output [M, ldd] (bf16, or fp32 if d_fp32), may be nullptr
  int ldd;
  int d_fp32;
  int accumulate;        // fp32 output only: d += result
  dk_bf16* dt;           // optional transposed bf16 copy [N, lddt]
  int lddt;
  float alpha;
  float drop_p;          // > 0: inverted dropout applied after the activation (training forward)
  uint32_t drop_seed;
  const int* step;       // device step counter mixed into the dropout hash (graph-replay safe)
  int tma_store;         // set by the launcher: output goes through smem staging + TMA store
  int tma_mask;          // set by the launcher: mask tile is fetched with TMA
  unsigned long long* trace;  // diagnostics: 8 clock stamps of CTA (0, 0) (nullptr = off)
  // ---- classifier head fused into the GEMM that produces its input (bn = 16, one M tile; head_w != nullptr) ----
  // The GEMM computes H = act(A B^T + bias) [M, N]; the head is logits = H W3^T + b3 (C <= 16 classes), softmax
  // cross-entropy against integer labels, dZ = (p - y) / M and dH = alpha (dZ W3) * (H > 0).  Every CTA adds its
  // 16-column slice's contribution to the logits (fp32 red.add into head_acc [128, 16]), the CTAs of the grid meet on
  // head_sync[0] (all co-resident: the grid is <= 32 CTAs), then every CTA finishes its own slice of dH; CTA 0 also
  // writes dZ and the loss / accuracy record.  Nothing is reset on the way out: the arrival counter is monotonic and
  // the logits scratch is double-buffered by launch parity (CTA 0 clears the other half at entry).
  const dk_bf16* head_w;   // W3 (bf16 shadow) [C, N], leading dimension head_ldw
  int head_ldw;
  const float* head_bias;  // b3 [C] or nullptr
  int head_c;
  const int* head_labels;  // [M] class indices
  int head_label_slot;     // engine: >= 0 -> head_labels is read from this slot when the list runs
  float* head_acc;         // [2][128, 16] fp32 (double-buffered by launch parity), zero at creation
  unsigned* head_sync;     // [2] u32: monotonic arrival counter, launch counter; zero at creation
  dk_bf16* head_dz;        // [M, head_ldz] bf16 (columns >= C zero)
  int head_ldz;
  dk_bf16* head_dh;        // [M, head_lddh] bf16
  int head_lddh;
  float head_alpha;        // dropout keep-scale of the layer that produced H (1 if none)
  float* head_hist;        // loss / accuracy record [hist_slots, 2]
  const int* head_step;
  int head_hist_slots;
} DkGemmEpilogue;

#ifdef __cplusplus
namespace dk {
typedef DkGemmEpilogue GemmEpilogue;
}
extern "C" {
#endif

// Encodes a CUtensorMap (128 bytes, 64-byte aligned) for a row-major [rows, cols] matrix with
// leading dimension ld (elements), box = [box_rows, 128 bytes], SWIZZLE_128B, zero OOB fill.
int dk_tmap_encode_2d(void* out_tmap, const void* base, int dtype, long rows, long cols, long ld,
                      int box_rows);
int dk_tmap_encode_kchunks(void* out_tmap, const void* base, long rows, long K, long ld, int box_rows, int kch);
int dk_gemm_pick_bn(int N);
int dk_gemm_a_box_rows(int M);
int dk_gemm_mcast_cluster(int M);    // CTAs per cluster for DK_GEMM_MCAST_A (1 = no multicast possible)
int dk_gemm_mcast_box_rows(int M);   // rows of the A box each CTA loads
int dk_gemm_pick_bn2(int M, int N);
int dk_gemm_pick_bn_splitk(int M, int N, int K);
int dk_gemm_tn_launch(const void* tmap_a, const void* tmap_b, const DkGemmEpilogue* ep, int M, int N,
                      int K, int bn, int flags, void* stream);
int dk_gemm_tn_launch2(const void* tmap_a, const void* tmap_b, const void* tmap_d, const void* tmap_m,
                       const DkGemmEpilogue* ep, int M, int N, int K, int bn, int flags, int splits, void* stream);
int dk_gemm_encode_output(void* tmap_d, const void* D, long ldd, int M, int N, int d_fp32);
int dk_gemm_pick_splits(int M, int N, int K, int bn, int tf32);
int dk_gemm_pick_splits_pair(int M, int N, int K, int bn);
// implicit-GEMM convolution (A gathered from an NHWC activation; see conv_gemm_kernel)
int dk_conv_gemm_launch(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off,
                        int div, const void* tmap_b, const void* tmap_d, const void* tmap_m, const DkGemmEpilogue* ep,
                        int M, int N, int K, int bn, void* stream);
int dk_conv_pick_bn(int N);
// implicit-GEMM convolution fed by TMA im2col descriptors (conv_tma.cu): persistent, double-buffered TMEM
int dk_conv_tma_supported(int C, int div, int N, int ldd, int d_fp32);
int dk_conv_tma_encode_a(void* out_tmap, const void* src, int B, int SH, int SW, int C, int GH, int GW, int mul, int off,
                         int chan);
int dk_conv_tma_encode_b(void* out_tmap, const void* W, long ldw, int N, int K, int bn, int chan);
int dk_conv_tma_bn(int N);
int dk_conv_tma_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, const void* tmap_m,
                       const DkGemmEpilogue* ep, int C, int GH, int GW, int KH, int KW, int mul, int off, int M, int N,
                       void* stream);
// TMA-im2col weight gradient (+ bias gradient), see conv_wgrad_tma_kernel
int dk_conv_wgrad_tma_units(int C);
int dk_conv_wgrad_tma_supported(int C, int Cout, long lddz, long lddw);
int dk_conv_wgrad_tma_encode(void* tmap_a, void* tmap_b, void* tmap_d, const void* src, int B, int SH, int SW, int C, int GH,
                             int GW, int KH, int KW, int stride, int pad, const void* dz, long lddz, float* dw, long lddw,
                             int Cout);
int dk_conv_wgrad_tma_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, int B, int C, int GH, int GW, int KH,
                             int KW, int stride, int pad, int Cout, int unit0, int units, float* bias_grad, void* stream);
int dk_conv_wgrad_tma(const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                      const void* dz, long lddz, float* dw, long lddw, int Cout, float* bias_grad, void* stream);
int dk_conv_tma(const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off,
                const void* Wmat, long ldw, const DkGemmEpilogue* ep, int M, int N, void* stream);
int dk_conv_gather_mode(int mode);
int dk_conv_gemm(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int mul, int off, int div,
                 const void* Bmat, long ldb, const DkGemmEpilogue* ep, int M, int N, int K, void* stream);
// EXPERIMENTAL implicit wgrad (not yet validated on hardware)
int dk_conv_wgrad_launch(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                         const void* tmap_a, const void* tmap_d, const DkGemmEpilogue* ep, int Cout, int rows, int splits,
                         void* stream);
int dk_conv_wgrad(const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW, int stride, int pad,
                  const void* dz, long lddz, float* dw, long lddw, int Cout, int rows, int splits, void* stream);
int dk_conv_weight_flip(const void* w, int ldw, void* wd, int ldwd, int Cout, int Cin, int KH, int KW, void* stream);
int dk_gemm_tn_ex(const void* A, long lda, const void* B, long ldb, const DkGemmEpilogue* ep, int M, int N,
                  int K, int flags, int bn, int splits, void* stream);
int dk_gemm_pull_launch(const void* tmap_a, const void* tmap_b, const void* tmap_d, const DkGemmEpilogue* ep, int M,
                        int N, int K, float* w_local, float* w1_local, void* wb_local, int ldw, void* stream);
int dk_gemm_pull(const float* X, long ldx, const float* center_w, long ldc, const DkGemmEpilogue* ep, int M, int N, int K,
                 float* w_local, float* w1_local, void* wb_local, void* stream);
int dk_gemm_encode_operands(void* tmap_a, void* tmap_b, const void* A, long lda, const void* B, long ldb,
                            int M, int N, int K, int bn, int flags);
int dk_gemm_tn(const void* A, long lda, const void* B, long ldb, const DkGemmEpilogue* ep, int M,
               int N, int K, int flags, int bn, void* stream);

#ifdef __cplusplus
}
#endif
