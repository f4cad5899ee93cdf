// C ABI of the parameter-server kernels (ps_kernels.cu) and the control-block layout.
#pragma once
#include <stdint.h>

// Control block: an array of 32-bit words living next to the center variable in PS HBM.
enum {
  DK_CTRL_NUM_UPDATES = 0,      // commits applied so far (reference: ParameterServer.num_updates)
  DK_CTRL_LOCK_NEXT = 1,        // ticket lock (strict mode)
  DK_CTRL_LOCK_SERVING = 2,
  DK_CTRL_STOP = 3,             // stop flag (fault handling)
  DK_CTRL_SHARD_NEXT = 4,       // dynamic shard queue: next unclaimed data partition
  DK_CTRL_WORKERS_DONE = 5,     // workers that finished their shards
  DK_CTRL_BARRIER = 6,          // arrivals at the device-side rendezvous of the synchronous trainers (monotonic)
  DK_CTRL_HEARTBEAT = 16,       // + worker id (< 40): number of commits by that worker (liveness counter)
  DK_CTRL_DONE_FLAGS = 56,      // + worker id (< 40): non-zero once that worker finished its shards
  DK_CTRL_STALENESS_HIST = 96,  // 32 buckets
  DK_CTRL_WORDS = 128
};

#define DK_MAX_PEERS 16

#ifdef __cplusplus
extern "C" {
#endif

int dk_ps_commit(float* center, const float* w, const float* w1, long n, float scale,
                 const float* scale_dev, unsigned* ctrl, int worker, unsigned iteration, void* stream);
int dk_ps_pull(const float* center, float* w, float* w1, void* wb, long n, const unsigned* ctrl,
               unsigned* last_update, void* stream);
int dk_ps_exchange(float* center, float* w, float* w1, void* wb, long n, float scale,
                   const float* scale_dev, unsigned* ctrl, int worker, unsigned iteration,
                   unsigned* last_update, void* stream);
int dk_ps_elastic(float* center, float* w, void* wb, long n, float alpha, unsigned* ctrl, int worker,
                  unsigned iteration, void* stream);
int dk_ps_damped_exchange(float* center, float* w, float* w1, void* wb, long n, float scale,
                          float inv_lr, unsigned* ctrl, int worker, unsigned iteration, void* stream);
int dk_ps_ticket(unsigned* ctrl, const unsigned* last_update, float* scale_out, void* stream);
int dk_ps_fetch_add(unsigned* word, unsigned inc, unsigned* out, void* stream);
// Synchronous EASGD on the fabric.  dk_ps_barrier: rendezvous number *round (then ++*round) of `workers` ranks on
// the control block -- arrive with a release add, spin on an acquire load; gives up after timeout_ms (or when the
// stop flag is raised) and sets *broken.  dk_ps_easgd_read: E = alpha (W - C), W -= E (center only read);
// dk_ps_easgd_add: C += E (red.add) and one commit in the update counter.
int dk_ps_barrier(unsigned* ctrl, int workers, unsigned* round, unsigned* broken, int timeout_ms, void* stream);
int dk_ps_easgd_read(const float* center, float* w, void* wb, float* e, long n, float alpha, void* stream);
int dk_ps_easgd_add(float* center, const float* e, long n, unsigned* ctrl, int worker, void* stream);
int dk_ps_lock_acquire(unsigned* ctrl, unsigned* my_ticket, void* stream);
int dk_ps_lock_release(unsigned* ctrl, const unsigned* my_ticket, void* stream);
int dk_ps_average(float* const* peer_ptrs, int num_peers, long lo, long hi, void* stream);
int dk_ps_copy(float* dst, const float* src, long n, void* stream);

#ifdef __cplusplus
}
#endif
