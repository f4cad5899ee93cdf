// Native step executor.  The Python planner lowers a model + optimizer + PS algorithm into flat
// lists of device ops once (pointers, shapes, pre-encoded TMA descriptors); `dk_engine_run`
// then enqueues a whole list on a stream with no Python in the loop, which is what gets
// captured into the per-window CUDA graph (forward, loss, backward, optimizer, commit / pull).
//
// This replaces the reference's per-batch Python hot loop (distkeras/workers.py:327-342:
// train_on_batch -> get_weights -> numpy -> pickle -> socket) with a replayable device program.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "engine.h"
#include "gemm.h"
#include "kernels.h"
#include "ps.h"

namespace {

struct Op {
  int kind;
  int stream_id;  // 0 = caller's stream, 1.. = engine side streams (parallel graph branches)
  int64_t i[DK_OP_MAX_I];
  double f[DK_OP_MAX_F];
  // GEMM only
  alignas(64) CUtensorMap ta;
  alignas(64) CUtensorMap tb;
  alignas(64) CUtensorMap td;
  alignas(64) CUtensorMap tm;
  DkGemmEpilogue ep;
  int has_td, has_tm;
  int dyn_a_slot;   // >= 0: A operand base comes from this slot (tensor map re-encoded when the list runs)
  long dyn_lda;
  void* rec;        // DK_OP_BWD_UPDATE: launch record (owned)
  int bwd_slot[DK_BWD_MAX_LAYERS];
};

struct Engine {
  std::vector<std::vector<Op>> lists;
  void* slots[DK_ENGINE_SLOTS];
  long launches;  // kernels enqueued so far (bench "gpu_launches" accounting)
  int build_stream;            // stream id given to ops added from now on
  cudaStream_t side[DK_ENGINE_SIDE_STREAMS];  // independent work (wgrad / bias-grad branches) overlaps the dgrad chain
  std::vector<cudaEvent_t> events;
  size_t next_event;
};

cudaEvent_t next_event(Engine* e) {
  if (e->next_event >= e->events.size()) {
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    e->events.push_back(ev);
  }
  return e->events[e->next_event++];
}

inline void* resolve(const Engine* e, int64_t v) {
  if (v < 0) return e->slots[-v - 1];
  return reinterpret_cast<void*>(static_cast<uintptr_t>(v));
}

template <typename T>
inline T* rp(const Engine* e, int64_t v) {
  return reinterpret_cast<T*>(resolve(e, v));
}

// kernel launches performed by one op (memset / memcpy / fork / join nodes launch none)
int kernels_in_op(int kind) {
  switch (kind) {
    case DK_OP_MEMSET: case DK_OP_MEMCPY: case DK_OP_MEMCPY2D: case DK_OP_FORK: case DK_OP_JOIN: return 0;
    case DK_OP_BN_FWD: return 3;
    case DK_OP_BN_BWD: return 2;
    default: return 1;
  }
}

int run_op(Engine* e, Op& op, void* main_stream) {
  const int64_t* a = op.i;
  const double* f = op.f;
  if (op.kind == DK_OP_FORK || op.kind == DK_OP_JOIN) {
    // FORK: side stream waits for everything enqueued so far on the main stream.
    // JOIN: main stream waits for everything enqueued so far on the side stream.
    // a[0] = side stream id (1-based; 0 is accepted as 1)
    const int sid = a[0] <= 0 ? 0 : (int)a[0] - 1;
    if (sid >= DK_ENGINE_SIDE_STREAMS) return -1;
    cudaStream_t from = op.kind == DK_OP_FORK ? (cudaStream_t)main_stream : e->side[sid];
    cudaStream_t to = op.kind == DK_OP_FORK ? e->side[sid] : (cudaStream_t)main_stream;
    cudaEvent_t ev = next_event(e);
    DK_HOST_CHECK(cudaEventRecord(ev, from));
    DK_HOST_CHECK(cudaStreamWaitEvent(to, ev, 0));
    return 0;
  }
  void* st = op.stream_id >= 1 ? (void*)e->side[op.stream_id - 1] : main_stream;
  switch (op.kind) {
    case DK_OP_INPUT:
      // x, in_dtype, B, F, xb, ldx, xt, ldxt, step_counter, xf, ldxf | scale, shift
      return dk_input_stage(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (float)f[0], (float)f[1],
                            resolve(e, a[4]), (int)a[5], resolve(e, a[6]), (int)a[7], rp<int>(e, a[8]),
                            resolve(e, a[9]), (int)a[10], st);
    case DK_OP_GEMM_PULL:
      // M, N, K, w_local, w1_local, wb_local, ldw (tensor maps + epilogue pre-encoded)
      return dk_gemm_pull_launch(&op.ta, &op.tb, op.has_td ? &op.td : nullptr, &op.ep, (int)a[0], (int)a[1], (int)a[2],
                                 rp<float>(e, a[3]), rp<float>(e, a[4]), resolve(e, a[5]), (int)a[6], st);
    case DK_OP_CONV_GEMM:
      // src, SH, SW, C, GH, GW, KH, KW, mul, off, div, M, N, K, bn (weight / output / mask maps + epilogue pre-encoded)
      if (a[15])
        return dk_conv_tma_launch(&op.ta, &op.tb, op.has_td ? &op.td : nullptr, op.has_tm ? &op.tm : nullptr, &op.ep, (int)a[3],
                                  (int)a[4], (int)a[5], (int)a[6], (int)a[7], (int)a[8], (int)a[9], (int)a[11], (int)a[12], st);
      return dk_conv_gemm_launch(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5], (int)a[6],
                                 (int)a[7], (int)a[8], (int)a[9], (int)a[10], &op.tb, op.has_td ? &op.td : nullptr,
                                 op.has_tm ? &op.tm : nullptr, &op.ep, (int)a[11], (int)a[12], (int)a[13], (int)a[14], st);
    case DK_OP_CONV_WGRAD:
      // src, SH, SW, C, GH, GW, KH, KW, stride, pad, Cout, rows, splits (dZ / dW maps pre-encoded)
      return dk_conv_wgrad_launch(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5], (int)a[6],
                                  (int)a[7], (int)a[8], (int)a[9], &op.ta, &op.td, &op.ep, (int)a[10], (int)a[11],
                                  (int)a[12], st);
    case DK_OP_WFLIP:
      // w, ldw, wd, ldwd, Cout, Cin, KH, KW
      return dk_conv_weight_flip(resolve(e, a[0]), (int)a[1], resolve(e, a[2]), (int)a[3], (int)a[4], (int)a[5], (int)a[6],
                                 (int)a[7], st);
    case DK_OP_CONV_WGRAD_TMA:
      // B, C, GH, GW, KH, KW, stride, pad, Cout, unit0, units, bias_grad (dZ / im2col / dW maps pre-encoded)
      return dk_conv_wgrad_tma_launch(&op.ta, &op.tb, &op.td, (int)a[0], (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5],
                                      (int)a[6], (int)a[7], (int)a[8], (int)a[9], (int)a[10], rp<float>(e, a[11]), st);
    case DK_OP_BWD_UPDATE: {
      for (int l = 0; l < DK_BWD_MAX_LAYERS; ++l)
        if (op.bwd_slot[l] >= 0) {
          int r = dk_bwd_update_set_input(op.rec, l, e->slots[op.bwd_slot[l]]);
          if (r != 0) return r;
        }
      return dk_bwd_update_launch(op.rec, st);
    }
    case DK_OP_GEMM:
      // M, N, K, bn, flags (tensor maps + epilogue pre-encoded)
      if (op.dyn_a_slot >= 0) {
        int r = dk_tmap_encode_2d(&op.ta, e->slots[op.dyn_a_slot], DK_BF16, a[0], a[2], op.dyn_lda,
                                  DK_GEMM_TILE_ROWS_OF(a[4]) ? (int)DK_GEMM_TILE_ROWS_OF(a[4])
                                  : (a[4] & DK_GEMM_MCAST_A) ? dk_gemm_mcast_box_rows((int)a[0])
                                  : (a[4] & DK_GEMM_SHORT_A) ? dk_gemm_a_box_rows((int)a[0]) : 128);
        if (r != 0) return r;
        if (DK_GEMM_KCH_OF(a[4]) > 1) {
          r = dk_tmap_encode_kchunks(&op.td, e->slots[op.dyn_a_slot], a[0], a[2], op.dyn_lda,
                                     DK_GEMM_TILE_ROWS_OF(a[4]) ? (int)DK_GEMM_TILE_ROWS_OF(a[4]) : dk_gemm_a_box_rows((int)a[0]),
                                     (int)DK_GEMM_KCH_OF(a[4]));
          if (r != 0) return r;
        }
      }
      if (op.ep.head_w != nullptr && op.ep.head_label_slot >= 0)
        op.ep.head_labels = reinterpret_cast<const int*>(e->slots[op.ep.head_label_slot]);
      return dk_gemm_tn_launch2(&op.ta, &op.tb, op.has_td ? &op.td : nullptr, op.has_tm ? &op.tm : nullptr, &op.ep,
                                (int)a[0], (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5], st);
    case DK_OP_XENT:
      // logits, ldl, labels, labels_dense, B, C, dz, ldz, dzt, ldzt, probs, hist, step, hist_slots
      return dk_softmax_xent(rp<const float>(e, a[0]), (int)a[1], rp<const int>(e, a[2]),
                             rp<const float>(e, a[3]), (int)a[4], (int)a[5], resolve(e, a[6]), (int)a[7],
                             resolve(e, a[8]), (int)a[9], rp<float>(e, a[10]), rp<float>(e, a[11]),
                             rp<const int>(e, a[12]), (int)a[13], st);
    case DK_OP_HEAD:
      // H, ldh, Wb, ldw, bias, labels, labels_dense, B, C, K, dz, ldz, dH, lddh, use_mask, hist, step, hist_slots | alpha
      return dk_dense_softmax_head(resolve(e, a[0]), (int)a[1], resolve(e, a[2]), (int)a[3], rp<const float>(e, a[4]),
                                   rp<const int>(e, a[5]), rp<const float>(e, a[6]), (int)a[7], (int)a[8], (int)a[9],
                                   resolve(e, a[10]), (int)a[11], resolve(e, a[12]), (int)a[13], (float)f[0],
                                   (int)a[14], rp<float>(e, a[15]), rp<const int>(e, a[16]), (int)a[17], st);
    case DK_OP_ELOSS:
      // kind, out, target, B, C, dz, ldz, dzt, ldzt, hist, step, hist_slots
      return dk_elementwise_loss((int)a[0], rp<const float>(e, a[1]), rp<const float>(e, a[2]), (int)a[3],
                                 (int)a[4], resolve(e, a[5]), (int)a[6], resolve(e, a[7]), (int)a[8],
                                 rp<float>(e, a[9]), rp<const int>(e, a[10]), (int)a[11], st);
    case DK_OP_ROWSUM:
      // src, rows, cols, lds, out | scale
      return dk_rowsum_bf16(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], rp<float>(e, a[4]),
                            (float)f[0], st);
    case DK_OP_COLSUM:
      // src, rows, cols, lds, out | scale
      return dk_colsum_bf16(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], rp<float>(e, a[4]),
                            (float)f[0], st);
    case DK_OP_MEMCPY2D:
      // dst, dpitch, src, spitch, width_bytes, height
      return dk_memcpy2d_async(resolve(e, a[0]), (long)a[1], resolve(e, a[2]), (long)a[3], (long)a[4],
                               (long)a[5], st);
    case DK_OP_BN_FWD:
      // x, rows, C, sums, saved_mean, saved_invstd, moving_mean, moving_var, gamma, beta, relu, y | eps, momentum
      return dk_bn_forward(resolve(e, a[0]), (long)a[1], (int)a[2], rp<float>(e, a[3]), rp<float>(e, a[4]),
                           rp<float>(e, a[5]), rp<float>(e, a[6]), rp<float>(e, a[7]), rp<const float>(e, a[8]),
                           rp<const float>(e, a[9]), (float)f[0], (float)f[1], (int)a[10], resolve(e, a[11]), st);
    case DK_OP_BN_INF:
      // x, rows, C, moving_mean, moving_var, gamma, beta, relu, y | eps
      return dk_bn_inference(resolve(e, a[0]), (long)a[1], (int)a[2], rp<const float>(e, a[3]),
                             rp<const float>(e, a[4]), rp<const float>(e, a[5]), rp<const float>(e, a[6]),
                             (float)f[0], (int)a[7], resolve(e, a[8]), st);
    case DK_OP_BN_BWD:
      // dy, x, y_relu, rows, C, saved_mean, saved_invstd, gamma, sums, dgamma, dbeta, dx
      return dk_bn_backward(resolve(e, a[0]), resolve(e, a[1]), resolve(e, a[2]), (long)a[3], (int)a[4],
                            rp<const float>(e, a[5]), rp<const float>(e, a[6]), rp<const float>(e, a[7]),
                            rp<float>(e, a[8]), rp<float>(e, a[9]), rp<float>(e, a[10]), resolve(e, a[11]), st);
    case DK_OP_GAP_FWD:
      return dk_gap_fwd(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], resolve(e, a[4]), st);
    case DK_OP_GAP_BWD:
      return dk_gap_bwd(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], resolve(e, a[4]), st);
    case DK_OP_TRANSPOSE:
      // src, rows, cols, lds, dst, ldd
      return dk_transpose_bf16(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], resolve(e, a[4]),
                               (int)a[5], st);
    case DK_OP_OPTIM:
      // kind, w, g, s0, s1, wb, n, nesterov, step | lr, p0, p1, eps, decay, grad_scale
      return dk_optim_step((int)a[0], rp<float>(e, a[1]), rp<const float>(e, a[2]), rp<float>(e, a[3]),
                           rp<float>(e, a[4]), resolve(e, a[5]), (long)a[6], (float)f[0], (float)f[1],
                           (float)f[2], (float)f[3], (float)f[4], (int)a[7], rp<const int>(e, a[8]),
                           (float)f[5], st);
    case DK_OP_IM2COL:
      // x, B, H, W, C, KH, KW, stride, pad, OH, OW, col, ldcol
      return dk_im2col(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5], (int)a[6],
                       (int)a[7], (int)a[8], (int)a[9], (int)a[10], resolve(e, a[11]), (int)a[12], st);
    case DK_OP_COL2IM:
      // col, ldcol, B, H, W, C, KH, KW, stride, pad, OH, OW, dx, mask (0 = none)
      return dk_col2im_ex(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5], (int)a[6],
                          (int)a[7], (int)a[8], (int)a[9], (int)a[10], (int)a[11], resolve(e, a[12]),
                          resolve(e, a[13]), st);
    case DK_OP_MAXPOOL_FWD:
      // x, B, H, W, C, k, stride, y
      return dk_maxpool_fwd(resolve(e, a[0]), (int)a[1], (int)a[2], (int)a[3], (int)a[4], (int)a[5],
                            (int)a[6], resolve(e, a[7]), st);
    case DK_OP_MAXPOOL_BWD:
      // x, y, dy, B, H, W, C, k, stride, dx, relu
      return dk_maxpool_bwd_ex(resolve(e, a[0]), resolve(e, a[1]), resolve(e, a[2]), (int)a[3], (int)a[4],
                               (int)a[5], (int)a[6], (int)a[7], (int)a[8], resolve(e, a[9]), (int)a[10], st);
    case DK_OP_RELU_MASK:
      return dk_relu_mask_bf16(resolve(e, a[0]), resolve(e, a[1]), (long)a[2], st);
    case DK_OP_ADD:
      return dk_add_bf16(resolve(e, a[0]), resolve(e, a[1]), resolve(e, a[2]), (long)a[3], (int)a[4], st);
    case DK_OP_MEMSET:
      return dk_memset_async(resolve(e, a[0]), (int)a[1], (long)a[2], st);
    case DK_OP_MEMCPY:
      // dst, src, bytes, kind
      return dk_memcpy_async(resolve(e, a[0]), resolve(e, a[1]), (long)a[2], (int)a[3], st);
    case DK_OP_CAST:
      return dk_cast_bf16(rp<const float>(e, a[0]), resolve(e, a[1]), (long)a[2], st);
    case DK_OP_PS_COMMIT:
      // center, w, w1, n, scale_dev, ctrl, worker, iteration | scale
      return dk_ps_commit(rp<float>(e, a[0]), rp<const float>(e, a[1]), rp<const float>(e, a[2]), (long)a[3],
                          (float)f[0], rp<const float>(e, a[4]), rp<unsigned>(e, a[5]), (int)a[6],
                          (unsigned)a[7], st);
    case DK_OP_PS_PULL:
      // center, w, w1, wb, n, ctrl, last_update
      return dk_ps_pull(rp<const float>(e, a[0]), rp<float>(e, a[1]), rp<float>(e, a[2]), resolve(e, a[3]),
                        (long)a[4], rp<const unsigned>(e, a[5]), rp<unsigned>(e, a[6]), st);
    case DK_OP_PS_EXCHANGE:
      // center, w, w1, wb, n, scale_dev, ctrl, worker, iteration, last_update | scale
      return dk_ps_exchange(rp<float>(e, a[0]), rp<float>(e, a[1]), rp<float>(e, a[2]), resolve(e, a[3]),
                            (long)a[4], (float)f[0], rp<const float>(e, a[5]), rp<unsigned>(e, a[6]),
                            (int)a[7], (unsigned)a[8], rp<unsigned>(e, a[9]), st);
    case DK_OP_PS_ELASTIC:
      // center, w, wb, n, ctrl, worker, iteration | alpha
      return dk_ps_elastic(rp<float>(e, a[0]), rp<float>(e, a[1]), resolve(e, a[2]), (long)a[3], (float)f[0],
                           rp<unsigned>(e, a[4]), (int)a[5], (unsigned)a[6], st);
    case DK_OP_PS_DAMPED:
      // center, w, w1, wb, n, ctrl, worker, iteration | scale, inv_lr
      return dk_ps_damped_exchange(rp<float>(e, a[0]), rp<float>(e, a[1]), rp<float>(e, a[2]),
                                   resolve(e, a[3]), (long)a[4], (float)f[0], (float)f[1],
                                   rp<unsigned>(e, a[5]), (int)a[6], (unsigned)a[7], st);
    case DK_OP_PS_TICKET:
      return dk_ps_ticket(rp<unsigned>(e, a[0]), rp<const unsigned>(e, a[1]), rp<float>(e, a[2]), st);
    case DK_OP_LOCK_ACQUIRE:
      return dk_ps_lock_acquire(rp<unsigned>(e, a[0]), rp<unsigned>(e, a[1]), st);
    case DK_OP_LOCK_RELEASE:
      return dk_ps_lock_release(rp<unsigned>(e, a[0]), rp<const unsigned>(e, a[1]), st);
    case DK_OP_EAMSGD_PRE:
      // w, r, wcopy, wb, n | mu
      return dk_eamsgd_pre(rp<float>(e, a[0]), rp<float>(e, a[1]), rp<float>(e, a[2]), resolve(e, a[3]),
                           (long)a[4], (float)f[0], st);
    case DK_OP_EAMSGD_POST:
      return dk_eamsgd_post(rp<float>(e, a[0]), rp<float>(e, a[1]), rp<const float>(e, a[2]),
                            resolve(e, a[3]), (long)a[4], (float)f[0], st);
    case DK_OP_LABEL_INDEX:
      // probs, B, C, default_index, out_index, labels, correct | threshold
      return dk_label_index(rp<const float>(e, a[0]), (int)a[1], (int)a[2], (float)f[0], (int)a[3],
                            rp<int>(e, a[4]), rp<const int>(e, a[5]), rp<int>(e, a[6]), st);
    default:
      return -1000;
  }
}

}  // namespace

extern "C" {

void* dk_engine_create() {
  Engine* e = new Engine();
  memset(e->slots, 0, sizeof(e->slots));
  e->launches = 0;
  e->build_stream = 0;
  e->next_event = 0;
  for (int k = 0; k < DK_ENGINE_SIDE_STREAMS; ++k) {
    e->side[k] = nullptr;
    cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking);
  }
  return e;
}

void dk_engine_destroy(void* h) {
  Engine* e = reinterpret_cast<Engine*>(h);
  for (auto& lst : e->lists)
    for (Op& op : lst)
      if (op.rec != nullptr) {
        dk_bwd_update_release(op.rec);
        free(op.rec);
      }
  for (cudaEvent_t ev : e->events) cudaEventDestroy(ev);
  for (int k = 0; k < DK_ENGINE_SIDE_STREAMS; ++k)
    if (e->side[k] != nullptr) cudaStreamDestroy(e->side[k]);
  delete e;
}

// ops added after this call run on stream `id` (0 = caller's stream, 1 = engine side stream)
int dk_engine_set_build_stream(void* h, int id) {
  if (id < 0 || id > DK_ENGINE_SIDE_STREAMS) return -1;
  reinterpret_cast<Engine*>(h)->build_stream = id;
  return 0;
}

int dk_engine_new_list(void* h) {
  Engine* e = reinterpret_cast<Engine*>(h);
  e->lists.emplace_back();
  return static_cast<int>(e->lists.size()) - 1;
}

int dk_engine_clear_list(void* h, int list) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  e->lists[list].clear();
  return 0;
}

int dk_engine_set_slot(void* h, int slot, void* p) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (slot < 0 || slot >= DK_ENGINE_SLOTS) return -1;
  e->slots[slot] = p;
  return 0;
}

int dk_engine_add_op(void* h, int list, int kind, const int64_t* iargs, int ni, const double* fargs,
                     int nf) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size() || ni > DK_OP_MAX_I || nf > DK_OP_MAX_F) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = kind;
  op.stream_id = e->build_stream;
  for (int k = 0; k < ni; ++k) op.i[k] = iargs[k];
  for (int k = 0; k < nf; ++k) op.f[k] = fargs[k];
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

// GEMM: D = epilogue(A[M,K] * B[N,K]^T); operands must be fixed device buffers.
int dk_engine_add_gemm(void* h, int list, const void* A, long lda, const void* B, long ldb, int M, int N,
                       int K, int flags, int bn, int splits, const DkGemmEpilogue* ep) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_GEMM;
  op.stream_id = e->build_stream;
  const bool splitk_ok = ep->d_fp32 && ep->bias == nullptr && ep->act == 0 && ep->mask == nullptr;
  // bf16-output GEMMs of the forward / dgrad chain run on the persistent kernel (double-buffered
  // TMEM accumulator, epilogue overlapped with the next tile); DK_PERSISTENT=0 disables it
  static int persistent_env = -1;
  if (persistent_env < 0) {
    const char* pe = getenv("DK_PERSISTENT");
    persistent_env = (pe != nullptr && pe[0] == '0') ? 0 : 1;
  }
  if (bn <= 0 && persistent_env && !ep->d_fp32 && ep->d != nullptr && ep->dt == nullptr && !ep->bias_along_m &&
      !(flags & (DK_GEMM_A_MN | DK_GEMM_TF32 | DK_GEMM_SHORT_A)) && N >= 16 && (ep->ldd % 8) == 0 &&
      (ep->mask == nullptr || (ep->ld_mask % 8) == 0)) {
    bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
    flags |= DK_GEMM_PERSISTENT;
  }
  if (bn <= 0) bn = (splits == 0 && splitk_ok) ? dk_gemm_pick_bn_splitk(M, N, K) : dk_gemm_pick_bn2(M, N);
  if ((flags & DK_GEMM_B_MN) && bn < 64) bn = 64;
  // large wgrad GEMMs (both operands MN-major, split-K fp32 accumulation) run on CTA pairs
  // (cta_group::2: each CTA stages half of the B tile, one 256-row MMA per pair); DK_PAIR=0 disables
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* pe = getenv("DK_PAIR");
    pair_env = (pe != nullptr && pe[0] == '0') ? 0 : 1;
  }
  const bool pair = pair_env && splits == 0 && splitk_ok && bn == 256 && M >= 512 &&
                    (flags & (DK_GEMM_A_MN | DK_GEMM_B_MN)) == (DK_GEMM_A_MN | DK_GEMM_B_MN) &&
                    !(flags & (DK_GEMM_TF32 | DK_GEMM_PERSISTENT));
  if (pair) {
    flags |= DK_GEMM_PAIR;
    splits = dk_gemm_pick_splits_pair(M, N, K, bn);
  }
  int r = dk_gemm_encode_operands(&op.ta, &op.tb, A, lda, B, ldb, M, N, K, bn, flags);
  if (r != 0) return r;
  op.ep = *ep;
  if (DK_GEMM_KCH_OF(flags) > 1) {
    // several k-blocks per TMA request: the two extra tensor-map slots carry the 3-D [64, rows, k-chunks] operand views
    const int kch = DK_GEMM_KCH_OF(flags);
    const int a_rows = DK_GEMM_TILE_ROWS_OF(flags) ? DK_GEMM_TILE_ROWS_OF(flags) : dk_gemm_a_box_rows(M);
    if ((flags & (DK_GEMM_A_MN | DK_GEMM_B_MN | DK_GEMM_TF32)) || !(flags & DK_GEMM_SHORT_A)) return -6;
    r = dk_tmap_encode_kchunks(&op.td, A, M, K, lda, a_rows, kch);
    if (r == 0) r = dk_tmap_encode_kchunks(&op.tm, B, N, K, ldb, bn, kch);
    if (r != 0) return r;
    op.has_td = op.has_tm = 1;
  } else {
    op.has_td = ep->d != nullptr && dk_gemm_encode_output(&op.td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0;
    op.has_tm = ep->mask != nullptr && dk_gemm_encode_output(&op.tm, ep->mask, ep->ld_mask, M, N, 0) == 0;
  }
  if (splits == 0)  // auto: split-K only for plain fp32 accumulations (wgrad)
    splits = (ep->d_fp32 && ep->bias == nullptr && ep->act == 0 && ep->mask == nullptr)
                 ? dk_gemm_pick_splits(M, N, K, bn, flags & DK_GEMM_TF32)
                 : 1;
  op.i[0] = M; op.i[1] = N; op.i[2] = K; op.i[3] = bn; op.i[4] = flags; op.i[5] = splits;
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

int dk_engine_add_gemm_slot(void* h, int list, int a_slot, long lda, const void* B, long ldb, int M, int N, int K,
                            int flags, int bn, int splits, const DkGemmEpilogue* ep) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (a_slot < 0 || a_slot >= DK_ENGINE_SLOTS || (flags & (DK_GEMM_A_MN | DK_GEMM_TF32 | DK_GEMM_PAIR))) return -1;
  if (bn <= 0 && !(flags & DK_GEMM_SHORT_A)) return -1;  // slot-fed operands run on the plain kernel: explicit tile width
  // encode against a placeholder base (B is a valid, aligned device pointer); the real one is bound per run
  int r = dk_engine_add_gemm(h, list, B, lda, B, ldb, M, N, K, flags, bn, splits, ep);
  if (r < 0) return r;
  Op& op = e->lists[list][r];
  op.dyn_a_slot = a_slot;
  op.dyn_lda = lda;
  return r;
}

int dk_engine_add_bwd_update(void* h, int list, const DkBwdUpdateDesc* desc) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_BWD_UPDATE;
  op.stream_id = e->build_stream;
  void* raw = nullptr;
  if (posix_memalign(&raw, 128, dk_bwd_update_record_bytes()) != 0) return -2;
  op.rec = raw;
  // slot-fed layers are prepared against a placeholder input (their own dZ buffer) and re-pointed per run
  DkBwdUpdateDesc d = *desc;
  for (int l = 0; l < DK_BWD_MAX_LAYERS; ++l) {
    op.bwd_slot[l] = (l < d.nlayers && d.layer[l].x_slot >= 0) ? d.layer[l].x_slot : -1;
    if (op.bwd_slot[l] >= 0) d.layer[l].x = nullptr;
  }
  int r = dk_bwd_update_prepare(op.rec, &d);
  if (r != 0) {
    free(raw);
    return r < 0 ? r : -r;
  }
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

int dk_engine_add_conv_gemm(void* h, int list, const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW,
                            int mul, int off, int div, const void* Bmat, long ldb, int M, int N, int K,
                            const DkGemmEpilogue* ep) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_CONV_GEMM;
  op.stream_id = e->build_stream;
  // TMA-im2col persistent kernel when the geometry allows it (C a multiple of 64 or exactly 32, undilated);
  // the thread-gather kernel otherwise.  DK_CONV_TMA=0 forces the gather kernel (A/B runs).
  static int tma_env = -1;
  if (tma_env < 0) {
    const char* te = getenv("DK_CONV_TMA");
    tma_env = (te != nullptr && te[0] == '0') ? 0 : 1;
  }
  const bool use_tma = tma_env && dk_conv_tma_supported(C, div, N, ep->ldd, ep->d_fp32) && ep->d != nullptr && ep->dt == nullptr &&
                       (ep->mask == nullptr || (ep->ld_mask % 8) == 0) && M % (GH * GW) == 0;
  const int bn = use_tma ? dk_conv_tma_bn(N) : dk_conv_pick_bn(N);
  int r;
  if (use_tma) {
    const int chan = C % 64 == 0 ? 64 : 32;
    r = dk_conv_tma_encode_a(&op.ta, src, M / (GH * GW), SH, SW, C, GH, GW, mul, off, chan);
    if (r != 0) return r;
    r = dk_conv_tma_encode_b(&op.tb, Bmat, ldb, N, K, bn, chan);
  } else {
    r = dk_tmap_encode_2d(&op.tb, Bmat, DK_BF16, N, K, ldb, bn);
  }
  if (r != 0) return r;
  op.i[15] = use_tma ? 1 : 0;
  op.ep = *ep;
  op.has_td = ep->d != nullptr && dk_gemm_encode_output(&op.td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0;
  op.has_tm = ep->mask != nullptr && dk_gemm_encode_output(&op.tm, ep->mask, ep->ld_mask, M, N, 0) == 0;
  op.i[0] = (int64_t)(uintptr_t)src;
  op.i[1] = SH; op.i[2] = SW; op.i[3] = C; op.i[4] = GH; op.i[5] = GW; op.i[6] = KH; op.i[7] = KW;
  op.i[8] = mul; op.i[9] = off; op.i[10] = div; op.i[11] = M; op.i[12] = N; op.i[13] = K; op.i[14] = bn;
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

int dk_engine_add_conv_wgrad_tma(void* h, int list, const void* src, int B, int SH, int SW, int C, int GH, int GW, int KH,
                                 int KW, int stride, int pad, const void* dz, long lddz, float* dw, long lddw, int Cout,
                                 float* bias_grad) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  if (!dk_conv_wgrad_tma_supported(C, Cout, lddz, lddw)) return -2;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_CONV_WGRAD_TMA;
  op.stream_id = e->build_stream;
  int r = dk_conv_wgrad_tma_encode(&op.ta, &op.tb, &op.td, src, B, SH, SW, C, GH, GW, KH, KW, stride, pad, dz, lddz, dw, lddw, Cout);
  if (r != 0) return r;
  const int chan = C % 64 == 0 ? 64 : 32;
  const int total = KH * KW * (C / chan), per = dk_conv_wgrad_tma_units(C);
  // balanced blocks of units (e.g. 9 taps of 64 channels -> 5 + 4 rather than 7 + 2)
  const int launches = (total + per - 1) / per;
  const int each = (total + launches - 1) / launches;
  int last = -1;
  for (int u0 = 0; u0 < total; u0 += each) {
    const int n = total - u0 < each ? total - u0 : each;
    op.i[0] = B; op.i[1] = C; op.i[2] = GH; op.i[3] = GW; op.i[4] = KH; op.i[5] = KW; op.i[6] = stride; op.i[7] = pad;
    op.i[8] = Cout; op.i[9] = u0; op.i[10] = n;
    op.i[11] = u0 == 0 ? (int64_t)(uintptr_t)bias_grad : 0;
    e->lists[list].push_back(op);
    last = static_cast<int>(e->lists[list].size()) - 1;
  }
  return last;
}

int dk_engine_add_conv_wgrad(void* h, int list, const void* src, int SH, int SW, int C, int GH, int GW, int KH, int KW,
                             int stride, int pad, const void* dz, long lddz, float* dw, long lddw, int Cout, int rows) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_CONV_WGRAD;
  op.stream_id = e->build_stream;
  int r = dk_tmap_encode_2d(&op.ta, dz, DK_BF16, rows, Cout, lddz, 64);
  if (r != 0) return r;
  r = dk_gemm_encode_output(&op.td, dw, lddw, Cout, KH * KW * C, 1);
  if (r != 0) return r;
  op.has_td = 1;
  op.ep.alpha = 1.f;
  op.ep.d = dw;
  op.ep.ldd = static_cast<int>(lddw);
  op.ep.d_fp32 = 1;
  const int tiles = (KH * KW * C + 127) / 128 * ((Cout + 127) / 128);
  int splits = 148 / tiles;  // one 128-wide CTA per SM, single wave
  if (splits < 1) splits = 1;
  op.i[0] = (int64_t)(uintptr_t)src;
  op.i[1] = SH; op.i[2] = SW; op.i[3] = C; op.i[4] = GH; op.i[5] = GW; op.i[6] = KH; op.i[7] = KW;
  op.i[8] = stride; op.i[9] = pad; op.i[10] = Cout; op.i[11] = rows; op.i[12] = splits;
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

// First-layer forward with the weight pull fused in: X fp32 [M, K] local, center_w fp32 [N, K] in the
// parameter server's (peer-mapped) HBM; w / w1 / wb are the local copies refreshed by the kernel.
int dk_engine_add_gemm_pull(void* h, int list, const void* X, long ldx, const void* center_w, long ldc, int M, int N,
                            int K, void* w_local, void* w1_local, void* wb_local, const DkGemmEpilogue* ep) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  Op op;
  memset(&op, 0, sizeof(op));
  op.dyn_a_slot = -1;
  op.kind = DK_OP_GEMM_PULL;
  op.stream_id = e->build_stream;
  int r = dk_tmap_encode_2d(&op.ta, X, DK_F32, M, K, ldx, 128);
  if (r != 0) return r;
  r = dk_tmap_encode_2d(&op.tb, center_w, DK_F32, N, K, ldc, 128);
  if (r != 0) return r;
  op.ep = *ep;
  op.has_td = ep->d != nullptr && dk_gemm_encode_output(&op.td, ep->d, ep->ldd, M, N, ep->d_fp32) == 0;
  op.i[0] = M; op.i[1] = N; op.i[2] = K;
  op.i[3] = (int64_t)(uintptr_t)w_local; op.i[4] = (int64_t)(uintptr_t)w1_local; op.i[5] = (int64_t)(uintptr_t)wb_local;
  op.i[6] = ldc;
  e->lists[list].push_back(op);
  return static_cast<int>(e->lists[list].size()) - 1;
}

int dk_engine_run(void* h, int list, void* stream) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  e->next_event = 0;
  for (Op& op : e->lists[list]) {
    int r = run_op(e, op, stream);
    if (r != 0) return r;
    e->launches += kernels_in_op(op.kind);
  }
  return 0;
}

int dk_engine_list_size(void* h, int list) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  return static_cast<int>(e->lists[list].size());
}

// number of kernel launches a run of `list` performs (memset / memcpy nodes excluded)
int dk_engine_list_kernels(void* h, int list) {
  Engine* e = reinterpret_cast<Engine*>(h);
  if (list < 0 || list >= (int)e->lists.size()) return -1;
  int n = 0;
  for (const Op& op : e->lists[list])
    n += kernels_in_op(op.kind);
  return n;
}

long dk_engine_launches(void* h) { return reinterpret_cast<Engine*>(h)->launches; }

}  // extern "C"
