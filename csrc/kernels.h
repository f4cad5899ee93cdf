// C ABI of the elementwise / loss / layout kernels.
#pragma once
#include <stdint.h>

enum {
  DK_OPT_SGD = 0,
  DK_OPT_MOMENTUM = 1,
  DK_OPT_ADAGRAD = 2,
  DK_OPT_RMSPROP = 3,
  DK_OPT_ADAM = 4,
  DK_OPT_ADADELTA = 5,
  DK_OPT_ADAMAX = 6,
  DK_OPT_NADAM = 7   // Adam with Nesterov momentum (Dozat 2016), constant beta_1 (no momentum-decay schedule)
};

enum { DK_IN_U8 = 0, DK_IN_F32 = 1, DK_IN_BF16 = 2 };
enum { DK_LOSS_XENT = 0, DK_LOSS_MSE = 1, DK_LOSS_BCE = 2 };

#ifdef __cplusplus
extern "C" {
#endif

// optimizers
int dk_optim_step(int kind, float* w, const float* g, float* s0, float* s1, void* wb, long n, float lr,
                  float p0, float p1, float eps, float decay, int nesterov, const int* step,
                  float grad_scale, void* stream);
int dk_eamsgd_pre(float* w, float* r, float* wcopy, void* wb, long n, float mu, void* stream);
int dk_eamsgd_post(float* w, float* r, const float* wcopy, void* wb, long n, float eta, void* stream);
int dk_cast_bf16(const float* src, void* dst, long n, void* stream);

// loss: softmax + categorical cross-entropy forward/backward + accuracy, one launch.
//   logits fp32 [B, C]; labels: int32 class index [B] (labels_dense == nullptr) or dense fp32 [B, C].
//   dz bf16 [B, ldz] and dzt bf16 [C, ldzt] (either may be nullptr) receive (p - y) / B.
//   hist[2 * slot] += mean loss, hist[2 * slot + 1] += accuracy, slot = *step - 1 (or 0).
int dk_softmax_xent(const float* logits, int ldl, const int* labels, const float* labels_dense, int B,
                    int C, void* dz, int ldz, void* dzt, int ldzt, float* probs, float* hist,
                    const int* step, int hist_slots, void* stream);
int dk_dense_softmax_head(const void* H, int ldh, const void* Wb, int ldw, const float* bias, const int* labels,
                          const float* labels_dense, int B, int C, int K, void* dz, int ldz, void* dH, int lddh,
                          float alpha, int use_mask, float* hist, const int* step, int hist_slots, void* stream);
// generic elementwise losses on fp32 outputs (mse / binary cross-entropy on sigmoid outputs)
int dk_elementwise_loss(int kind, const float* out, const float* target, int B, int C, void* dz, int ldz,
                        void* dzt, int ldzt, float* hist, const int* step, int hist_slots, void* stream);

// input stage: x [B, F] (u8 / f32 / bf16) -> xb bf16 [B, ldx] (and optional xt bf16 [F, ldxt]),
// y = x * scale + shift (fused MinMaxTransformer); also increments the device step counter.
int dk_input_stage(const void* x, int in_dtype, int B, int F, float scale, float shift, void* xb,
                   int ldx, void* xt, int ldxt, int* step_counter, void* xf, int ldxf, void* stream);

// batched bf16 transposes: dst[c, r] = src[r, c] for a table of matrices
int dk_transpose_bf16(const void* src, int rows, int cols, int lds, void* dst, int ldd, void* stream);
// row sums of a bf16 matrix [rows, cols] (bias gradient from dZ^T): out[r] = sum_c src[r, c]
int dk_rowsum_bf16(const void* src, int rows, int cols, int lds, float* out, float scale, void* stream);

// column sums of a bf16 matrix [rows, cols]: out[c] += scale * sum_r src[r, c] (out pre-zeroed)
int dk_colsum_bf16(const void* src, int rows, int cols, int lds, float* out, float scale, void* stream);

// conv helpers (NHWC activations, bf16): im2col / col2im for KHxKW, stride, padding
int dk_im2col(const void* x, int B, int H, int W, int C, int KH, int KW, int stride, int pad, int OH,
              int OW, void* col, int ldcol, void* stream);
int dk_col2im(const void* col, int ldcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
              int OH, int OW, void* dx, void* stream);
int dk_col2im_ex(const void* col, int ldcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
                 int OH, int OW, void* dx, const void* mask, void* stream);
int dk_maxpool_fwd(const void* x, int B, int H, int W, int C, int k, int stride, void* y, void* stream);
int dk_maxpool_bwd(const void* x, const void* y, const void* dy, int B, int H, int W, int C, int k,
                   int stride, void* dx, void* stream);
int dk_maxpool_bwd_ex(const void* x, const void* y, const void* dy, int B, int H, int W, int C, int k,
                      int stride, void* dx, int relu, void* stream);
// BatchNormalization over the last axis of [rows, C] bf16 (C % 8 == 0) and global average pooling
int dk_bn_forward(const void* x, long rows, int C, float* sums, float* saved_mean, float* saved_invstd,
                  float* moving_mean, float* moving_var, const float* gamma, const float* beta, float eps,
                  float momentum, int relu, void* y, void* stream);
int dk_bn_inference(const void* x, long rows, int C, const float* moving_mean, const float* moving_var,
                    const float* gamma, const float* beta, float eps, int relu, void* y, void* stream);
int dk_bn_backward(const void* dy, const void* x, const void* y_relu, long rows, int C, const float* saved_mean,
                   const float* saved_invstd, const float* gamma, float* sums, float* dgamma, float* dbeta, void* dx,
                   void* stream);
int dk_gap_fwd(const void* x, int B, int P, int C, void* y, void* stream);
int dk_gap_bwd(const void* dy, int B, int P, int C, void* dx, void* stream);
int dk_relu_mask_bf16(void* dy, const void* act, long n, void* stream);
int dk_add_bf16(void* dst, const void* a, const void* b, long n, int relu, void* stream);

// inference post-processing (reference K14): LabelIndexTransformer rule + accuracy count
int dk_label_index(const float* probs, int B, int C, float threshold, int default_index, int* out_index,
                   const int* labels, int* correct_count, void* stream);

#ifdef __cplusplus
}
#endif
