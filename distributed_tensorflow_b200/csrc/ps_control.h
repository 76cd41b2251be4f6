// Control blocks shared between the parameter-server GPU and the worker GPUs (peer-mapped memory).
// Plain C layout: the Python side reads fields through the offsets exported by dtf_offsetof_ctl().
#pragma once
#define DTF_MAX_WORKERS 16

// Written by ONE worker over NVLink (its own 128-byte line), read by the ps.
struct __align__(128) WorkerSlotState {
  unsigned long long arrivals;   // monotonic: += 1 per pushing CTA (release, system scope)
  unsigned long long stamp;      // local_step (sync) / pulled version (async) of the latest push
  unsigned long long pad[14];
};

// Lives in the ps GPU's memory.
struct __align__(128) PsControl {
  unsigned long long global_step;        // += 1 per aggregate (sync) / per apply (async)   
  unsigned long long param_version;      // sequence number of the last completed ps_apply launch
  float beta1_power, beta2_power;        // Adam accumulators
  unsigned long long dropped_stale;      // stale pushes discarded (sync)
  unsigned long long applied_total;      // gradients folded into updates
  unsigned long long staleness_hist[16]; // async: histogram of (global_step at apply - step at pull)
  unsigned long long staleness_sum;
  unsigned long long decision_seq;       // block 0 -> other blocks handshake
  unsigned int decision_mask, decision_count, decision_ok;
  unsigned int done_ctas;
  unsigned int err;
  unsigned int last_async_worker;
  unsigned long long consumed[DTF_MAX_WORKERS];   // arrivals already consumed per worker (ps private)
  WorkerSlotState w[DTF_MAX_WORKERS];
};

// Lives in each worker GPU's memory; the ps writes it over NVLink, the worker spins locally.
struct __align__(128) WorkerMailbox {
  unsigned long long token;      // sync: global step carried by the latest token; async: number of acks
  unsigned long long version;    // global step of the parameters currently published
  unsigned long long pad[14];
};
