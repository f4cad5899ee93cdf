// C ABI of the fused dense backward-update kernel (dense_fused.cu).
#pragma once
#include <stdint.h>

#define DK_BWD_MAX_LAYERS 6
#define DK_BWD_MAX_SHARDS 16

enum { DK_COMM_NONE = 0, DK_COMM_EXCHANGE = 1, DK_COMM_ELASTIC = 2 };

// One dense layer of the group: dW[n_out, k_in] = dZ^T X, db = colsum(dZ), then the optimizer rule.
typedef struct DkBwdLayerDesc {
  const void* dz;   // bf16 [batch, lddz]: gradient w.r.t. the layer's pre-activation
  long lddz;
  const void* x;    // bf16 [batch, ldx]: the layer's input (ignored when x_slot >= 0)
  long ldx;
  int x_slot;       // >= 0: the input pointer is read from this engine slot when the list runs
  int n_out, k_in;
  long w_off;       // element offset of the kernel [n_out, k_in] in the flat buffers
  long b_off;       // element offset of the bias [n_out], or -1
  void* wb_pad;     // optional padded bf16 shadow [n_out, ldwb_pad] refreshed together with the flat one
  long ldwb_pad;
} DkBwdLayerDesc;

typedef struct DkBwdUpdateDesc {
  int nlayers, batch;
  DkBwdLayerDesc layer[DK_BWD_MAX_LAYERS];
  float* w;          // flat fp32 master
  float* s0;         // optimizer state (may be NULL)
  float* s1;
  float* w1;         // last pulled center (needed by comm_mode 1)
  void* wb;          // flat bf16 shadow
  int opt_kind;
  float lr, p0, p1, eps, decay;
  int nesterov;
  int* step;               // device step counter (>= 1 while a step runs); incremented by the kernel when step_inc
  unsigned* done_counter;  // zero-initialised device word (last-CTA detection)
  int step_inc;
  // parameter-server exchange fused into the epilogue
  int comm_mode;           // DK_COMM_*
  float comm_scale;        // 1/tau (ADAG), 1 (DOWNPOUR / DynSGD)
  const float* scale_dev;  // DynSGD: 1/staleness written by the ticket kernel (NULL otherwise)
  float alpha;             // elastic coefficient (comm_mode 2)
  int nshards;             // center = shard_center[idx / shard_per] + idx % shard_per
  long shard_per;
  float* shard_center[DK_BWD_MAX_SHARDS];
  unsigned* ctrl;          // PS control block (peer-mapped)
  int worker;
  unsigned* last_update;
  unsigned long long* trace;  // diagnostics: clock stamps of CTA 0 (nullptr = off)
} DkBwdUpdateDesc;

#ifdef __cplusplus
extern "C" {
#endif

// Size of / in-place construction of the opaque launch record (tensor maps encoded once).
long dk_bwd_update_record_bytes();
long dk_bwd_update_desc_bytes();  // sizeof(DkBwdUpdateDesc): checked by the ctypes binding
int dk_bwd_update_prepare(void* record, const DkBwdUpdateDesc* desc);
void dk_bwd_update_release(void* record);  // frees the device-side tensor maps of a prepared record
// Re-point layer `layer`'s input operand (slot-fed first layer) before a launch.
int dk_bwd_update_set_input(void* record, int layer, const void* x);
int dk_bwd_update_launch(const void* record, void* stream);
// One-shot convenience (tests).
int dk_bwd_update(const DkBwdUpdateDesc* desc, void* stream);

#ifdef __cplusplus
}
#endif
