// Device-resident parameter-server protocol for one NVSwitch box.
//
//   * PsControl lives in the ps GPU's memory and is mapped (CUDA IPC / peer access) by every worker.
//     Workers push gradient tiles into their slot in ps HBM from GEMM epilogues (gemm_tcgen05.cu) and
//     from the fused MLP head below, then release-increment `w[rank].arrivals` over NVLink.
//   * ps_apply_kernel is the ps "service loop" body, one launch per aggregate: block 0 waits for the
//     required arrivals (sync: >= replicas_to_aggregate FRESH pushes, stale stamps dropped; async: any
//     one push), broadcasts the decision, then all blocks run ONE fused pass:
//     N-way reduce -> mean -> SGD/Momentum/TF-Adam apply -> write fp32 master + bf16 shadow (and,
//     in publish mode, store the new bf16 parameters straight into every worker's replica over
//     NVLink); the last block bumps global_step / beta powers and releases the tokens (one mailbox
//     store per worker, system scope).
//   * Tokens: the worker-side wait is fused into the first GEMM of the next step (its TMA producer
//     acquires the mailbox before loading the parameters).
#include <cstdio>
#include <cstring>

// DTF_HOST_EMU: the ps-side kernels of this file (ps_apply, publish, push/pull, token wait, fabric collectives) also
// compile with g++ against tests/emu/host_emu.h -- threads + barriers stand in for a thread block, plain atomics for the
// scoped PTX accesses, a registry of member buffers for multimem -- so the protocol (fresh/stale decision, mean, apply,
// tokens, staleness) and the fused MLP head are exercised by the CPU test tier from the SAME source.
#ifdef DTF_HOST_EMU
#include "host_emu.h"
#else
#include "common.cuh"
#endif
#include "ps_control.h"

namespace dtf {

struct PsApplyParams {
  PsControl* ctl;
  float* master;                 // [n] fp32 parameters (padded layout)
  float* slot_m;                 // optimizer slots (may be null)
  float* slot_v;
  const float* grad[DTF_MAX_WORKERS];      // per-worker gradient slots in ps memory
  float* grad_rw[DTF_MAX_WORKERS];         // same pointers, writable (zero-after-read ranges)
  __nv_bfloat16* shadow;                   // ps-local bf16 copy (peer-pull mode reads this)
  __nv_bfloat16* replica[DTF_MAX_WORKERS]; // optional per-worker bf16 replicas (publish mode), peer pointers
  WorkerMailbox* mailbox[DTF_MAX_WORKERS]; // peer pointers
  long long n;
  int num_workers;               // total_num_replicas
  int replicas_to_aggregate;
  unsigned int ctas_per_push;    // arrivals one complete push adds
  int mode;                      // 0 sync, 1 async
  int kind;                      // 0 sgd, 1 momentum, 2 adam
  float lr, momentum, beta1, beta2, eps;
  int nesterov;
  int publish_replicas;          // 1: also store the shadow into every worker replica
  int system_scope;              // 0: ps and workers share one GPU (gpu-scope fences suffice)
  long long zero_begin[4], zero_end[4];    // ranges of the slots to clear after reading (atomically accumulated grads)
  int num_zero;
  unsigned long long timeout_ns;
  unsigned long long* trace;     // optional ring: {kind, t0, t1, step} per launch
  int trace_cap;
  long long* phase_trace;        // optional [16] clock64 stamps of block 0 / the last block
  int idle_ok;                   // 1: a timeout is not an error (host-driven service loop polls)
  // NVLS mode (symmetric buffers, csrc/fabric_vmm.cu): the gradient slots live in the WORKERS' memories
  // (grad[w] = unicast peer pointer to worker w's copy) and the parameter replica is one symmetric buffer.
  const float* grad_mc;          // multicast mapping of the gradient buffer: multimem.ld_reduce = in-switch N-way sum
  float* grad_mc_rw;             // same mapping, writable (zero-after-read of atomically accumulated ranges)
  __nv_bfloat16* shadow_mc;      // multicast mapping of the bf16 replica: ONE multimem.st updates every GPU's copy
  float* master_mc;              // optional multicast mapping of an fp32 replica (models that consume fp32 parameters)
  unsigned int full_mask;        // all workers: the ld_reduce path needs every copy to hold a fresh gradient
  unsigned long long* token_mc;  // optional multicast address of a counter in every worker's replica buffer: each CTA bumps it once
                                 // (multimem.red) right after its own release fence -- the workers need no last-block round
};

#ifndef DTF_HOST_EMU
DTF_DEVICE unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DTF_DEVICE unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
DTF_DEVICE void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DTF_DEVICE void red_relaxed_sys_add_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DTF_DEVICE void st_relaxed_sys_ull(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// {token, version} of a worker mailbox as ONE 16-byte store (the pair shares an aligned 16-byte slot of the 128-byte line)
DTF_DEVICE void st_relaxed_sys_v2_u64(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
#endif

__global__ void __launch_bounds__(256, 2) ps_apply_kernel(const PsApplyParams p) {
  __shared__ unsigned int s_mask, s_count, s_ok;
  __shared__ unsigned long long s_seq;
  __shared__ float s_lr;
  PsControl* ctl = p.ctl;
  long long* tr = p.phase_trace;
#define PSTAMP(slot) do { if (tr && threadIdx.x == 0 && blockIdx.x == 0) tr[slot] = clock64(); } while (0)
  // programmatic dependent launch: this grid may have been started under the tail of the worker's step kernel (same GPU, same
  // stream) -- wait for it to complete before anything it wrote is read; only THEN let the next step kernel start its prologue
  // (it must not run beside the step kernel before it: it reads the device step counter that one advances at exit)
  griddep_wait();
  griddep_launch_dependents();
  PSTAMP(0);
  const unsigned long long t_start = threadIdx.x == 0 ? globaltimer_ns() : 0ull;     // (the LAST CTA to finish writes the ring row)

  // Sync mode with replicas_to_aggregate == total_num_replicas (the reference's setting, distributed_mnist.py:120-122): the
  // only possible decision is "all W pushes, all fresh" (a worker cannot run ahead of an aggregate it is part of, so no
  // stamp can be stale) -- every CTA waits for it BY ITSELF, no block-0 decision + broadcast round trip.
  const bool all_fresh_only = p.mode == 0 && p.replicas_to_aggregate == p.num_workers;
  if (threadIdx.x == 0 && all_fresh_only) {
    const unsigned long long seq = ld_acquire_gpu_u64(&ctl->param_version) + 1ull;
    s_seq = seq;
    unsigned int ok = 1, spins = 0;
    const unsigned long long t0 = globaltimer_ns();
    while (true) {
      unsigned long long arr[DTF_MAX_WORKERS];
#pragma unroll
      for (int w = 0; w < DTF_MAX_WORKERS; ++w)                // one batch of independent loads per poll
        arr[w] = w < p.num_workers ? ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->w[w].arrivals)) : ~0ull;
      bool all = true;
#pragma unroll
      for (int w = 0; w < DTF_MAX_WORKERS; ++w)
        if (w < p.num_workers && arr[w] < ctl->consumed[w] + p.ctas_per_push) all = false;
      if (all) break;
      if ((++spins & 0xFF) == 0 && (globaltimer_ns() - t0) > p.timeout_ns) {
        ok = 0;
        if (!p.idle_ok && blockIdx.x == 0) atomicExch(&ctl->err, 2u);
        break;
      }
      if (spins > 4096) __nanosleep(64);
    }
    if (p.system_scope) fence_acq_rel_sys(); else __threadfence();     // the gradient slots of all workers are now visible
    s_mask = p.full_mask;
    s_count = (unsigned int)p.num_workers;
    s_ok = ok;
    float lr = p.lr;
    if (p.kind == 2) lr = p.lr * sqrtf(1.0f - ctl->beta2_power) / (1.0f - ctl->beta1_power);
    s_lr = lr;
  }
  if (threadIdx.x == 0 && !all_fresh_only) {
    const unsigned long long seq = ld_acquire_gpu_u64(&ctl->param_version) + 1ull;
    s_seq = seq;
    unsigned int ok = 1;
    if (blockIdx.x == 0) {
      // ---------------- decision: which pushes does this aggregate consume? ----------------
      const unsigned long long gs = ctl->global_step;
      const unsigned long long t0 = globaltimer_ns();
      unsigned int mask = 0, count = 0;
      unsigned int spins = 0;
      while (true) {
        mask = 0;
        count = 0;
        if (p.mode == 0) {
          for (int w = 0; w < p.num_workers; ++w) {
            const unsigned long long arr = ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->w[w].arrivals));
            if (arr >= ctl->consumed[w] + p.ctas_per_push) {
              const unsigned long long stamp = ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->w[w].stamp));
              if (stamp >= gs) {
                mask |= 1u << w;
                ++count;
              } else {
                // stale gradient (stamped before the current global step): drop it, SURVEY A12
                ctl->consumed[w] += p.ctas_per_push;
                ctl->dropped_stale += 1;
              }
            }
          }
          if ((int)count >= p.replicas_to_aggregate) break;
        } else {
          const int start = (int)((ctl->last_async_worker + 1) % (unsigned)p.num_workers);
          for (int i = 0; i < p.num_workers; ++i) {
            const int w = (start + i) % p.num_workers;
            const unsigned long long arr = ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->w[w].arrivals));
            if (arr >= ctl->consumed[w] + p.ctas_per_push) {
              mask = 1u << w;
              count = 1;
              ctl->last_async_worker = (unsigned)w;
              // staleness = updates applied between this worker's pull and its apply
              const unsigned long long stamp = ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(&ctl->w[w].stamp));
              unsigned long long st = gs > stamp ? gs - stamp : 0ull;
              ctl->staleness_sum += st;
              if (st > 15) st = 15;
              ctl->staleness_hist[st] += 1;
              break;
            }
          }
          if (count) break;
        }
        if ((++spins & 0xFF) == 0 && (globaltimer_ns() - t0) > p.timeout_ns) {
          ok = 0;
          if (!p.idle_ok) atomicExch(&ctl->err, 2u);
          break;
        }
        if (spins > 4096) __nanosleep(64);
      }
      // relaxed polling above, ONE acquire fence here: the gradient slots of the chosen workers are now visible
      if (p.system_scope) fence_acq_rel_sys(); else __threadfence();
      ctl->decision_mask = mask;
      ctl->decision_count = count ? count : 1u;
      ctl->decision_ok = ok;
      __threadfence();
      st_release_gpu_u64(&ctl->decision_seq, seq);
    } else {
      const unsigned long long t0 = globaltimer_ns();
      unsigned int spins = 0;
      while (ld_acquire_gpu_u64(&ctl->decision_seq) < seq) {
        if ((++spins & 0xFF) == 0 && (globaltimer_ns() - t0) > 2 * p.timeout_ns) break;
        if (spins > 4096) __nanosleep(64);
      }
    }
    s_mask = ctl->decision_mask;
    s_count = ctl->decision_count;
    s_ok = ctl->decision_ok;
    float lr = p.lr;
    if (p.kind == 2) lr = p.lr * sqrtf(1.0f - ctl->beta2_power) / (1.0f - ctl->beta1_power);
    s_lr = lr;
  }
  __syncthreads();
  PSTAMP(1);          // decision known
  const unsigned int mask = s_mask;
  const float inv = 1.0f / (float)s_count;
  const float lr = s_lr;
  const bool ok = s_ok != 0;

  if (ok) {
    // ---------------- fused reduce + mean + apply + publish ----------------
    // NVLS: when every worker contributed, ONE multimem.ld_reduce per 16 bytes makes the switch sum the workers'
    // copies (the ps ingests n floats instead of N*n); otherwise (backup workers, stale drops, async) the chosen
    // workers' copies are read one by one through their unicast peer mappings.
    const bool mc_reduce = p.grad_mc != nullptr && mask == p.full_mask;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    // Up to PU float4 positions per thread per pass: ALL their gradient loads (multimem.ld_reduce round trips through the
    // switch, or the chosen workers' peer loads) and master / slot loads are issued before the first one is consumed, so a
    // large shard (ResNet-18: 11 M parameters) keeps PU x 16 B per thread in flight instead of one; the
    // MNIST-sized shards still take one position per thread (grid = n / 1024).
    constexpr int PU = 4;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < p.n; base += PU * stride) {
      float g[PU][4], xm[PU][4], sm[PU][4], sv[PU][4];
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const long long i0 = base + u * stride;
        g[u][0] = g[u][1] = g[u][2] = g[u][3] = 0.f;
        if (i0 >= p.n) continue;
        const bool vec = (i0 + 4 <= p.n);
        if (mc_reduce && vec) {
          const float4 t = multimem_ld_reduce_add_f32x4(p.grad_mc + i0);
          g[u][0] = t.x; g[u][1] = t.y; g[u][2] = t.z; g[u][3] = t.w;
        } else {
          for (int w = 0; w < p.num_workers; ++w) {
            if (!(mask & (1u << w))) continue;
            if (vec) {
              const float4 t = *reinterpret_cast<const float4*>(p.grad[w] + i0);
              g[u][0] += t.x; g[u][1] += t.y; g[u][2] += t.z; g[u][3] += t.w;
            } else {
              for (int j = 0; j < 4 && i0 + j < p.n; ++j) g[u][j] += p.grad[w][i0 + j];
            }
          }
        }
        if (vec) {
          const float4 t = *reinterpret_cast<const float4*>(p.master + i0);
          xm[u][0] = t.x; xm[u][1] = t.y; xm[u][2] = t.z; xm[u][3] = t.w;
          if (p.kind >= 1) {
            const float4 m4 = *reinterpret_cast<const float4*>(p.slot_m + i0);
            sm[u][0] = m4.x; sm[u][1] = m4.y; sm[u][2] = m4.z; sm[u][3] = m4.w;
          }
          if (p.kind == 2) {
            const float4 v4 = *reinterpret_cast<const float4*>(p.slot_v + i0);
            sv[u][0] = v4.x; sv[u][1] = v4.y; sv[u][2] = v4.z; sv[u][3] = v4.w;
          }
        } else {
          for (int j = 0; j < 4 && i0 + j < p.n; ++j) {
            xm[u][j] = p.master[i0 + j];
            if (p.kind >= 1) sm[u][j] = p.slot_m[i0 + j];
            if (p.kind == 2) sv[u][j] = p.slot_v[i0 + j];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PU; ++u) {
        const long long i0 = base + u * stride;
        if (i0 >= p.n) continue;
        const bool vec = (i0 + 4 <= p.n);
        float wv[4];
        for (int j = 0; j < 4 && i0 + j < p.n; ++j) {
          const float gj = g[u][j] * inv;
          float x = xm[u][j];
          if (p.kind == 0) {
            x -= lr * gj;
          } else if (p.kind == 1) {
            const float acc = p.momentum * sm[u][j] + gj;
            sm[u][j] = acc;
            x -= p.nesterov ? (lr * gj + lr * p.momentum * acc) : (lr * acc);
          } else {
            const float m = p.beta1 * sm[u][j] + (1.0f - p.beta1) * gj;
            const float v = p.beta2 * sv[u][j] + (1.0f - p.beta2) * gj * gj;
            sm[u][j] = m;
            sv[u][j] = v;
            x -= lr * m / (sqrtf(v) + p.eps);
          }
          wv[j] = x;
        }
        if (vec) {
          *reinterpret_cast<float4*>(p.master + i0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
          if (p.kind >= 1) *reinterpret_cast<float4*>(p.slot_m + i0) = make_float4(sm[u][0], sm[u][1], sm[u][2], sm[u][3]);
          if (p.kind == 2) *reinterpret_cast<float4*>(p.slot_v + i0) = make_float4(sv[u][0], sv[u][1], sv[u][2], sv[u][3]);
          const uint2 packed = make_uint2(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]));
          if (p.master_mc) multimem_st_f32x4(p.master_mc + i0, make_float4(wv[0], wv[1], wv[2], wv[3]));
          if (p.shadow_mc) {
            multimem_st_b64(p.shadow_mc + i0, packed.x, packed.y);      // the switch writes every GPU's replica (ours too)
          } else {
            if (p.shadow) *reinterpret_cast<uint2*>(p.shadow + i0) = packed;
            if (p.publish_replicas)
              for (int w = 0; w < p.num_workers; ++w)
                if (p.replica[w]) *reinterpret_cast<uint2*>(p.replica[w] + i0) = packed;     // NVLink store
          }
        } else {
          for (int j = 0; j < 4 && i0 + j < p.n; ++j) {
            p.master[i0 + j] = wv[j];
            if (p.kind >= 1) p.slot_m[i0 + j] = sm[u][j];
            if (p.kind == 2) p.slot_v[i0 + j] = sv[u][j];
            const __nv_bfloat16 b = __float2bfloat16(wv[j]);
            if (p.shadow) p.shadow[i0 + j] = b;
            if (p.publish_replicas)
              for (int w = 0; w < p.num_workers; ++w)
                if (p.replica[w]) p.replica[w][i0 + j] = b;
          }
        }
        for (int z = 0; z < p.num_zero; ++z) {
          if (i0 + 4 > p.zero_begin[z] && i0 < p.zero_end[z]) {
            if (mc_reduce && vec && i0 >= p.zero_begin[z] && i0 + 4 <= p.zero_end[z]) {
              multimem_st_f32x4(p.grad_mc_rw + i0, make_float4(0.f, 0.f, 0.f, 0.f));     // clears every worker's copy
              continue;
            }
            for (int w = 0; w < p.num_workers; ++w) {
              if (!(mask & (1u << w))) continue;
              for (int j = 0; j < 4; ++j) {
                const long long i = i0 + j;
                if (i >= p.zero_begin[z] && i < p.zero_end[z] && i < p.n) p.grad_rw[w][i] = 0.f;
              }
            }
          }
        }
      }
    }
  }

  // ---------------- completion: last block publishes the new step and releases the tokens ----------------
  __syncthreads();
  PSTAMP(2);          // block 0 finished its slice
  if (threadIdx.x == 0) {
    if (p.system_scope) __threadfence_system(); else __threadfence();      // one fence per CTA, after the barrier
    // NVLS: tell EVERY worker "this slice is published" with one switch-replicated add; a worker's step starts when all
    // gridDim.x slices of this aggregate have said so (sync, all-fresh aggregates only: a token for every replica)
    if (p.token_mc != nullptr && all_fresh_only && ok) multimem_red_add_u64(p.token_mc, 1ull);
    const unsigned int prev = atomicAdd(&ctl->done_ctas, 1u);
    if (prev == gridDim.x - 1) {
      // the ticket observed every CTA's fence + increment: ONE fence here makes all their parameter stores (local, peer,
      // multicast) precede the token stores below; the bookkeeping is ps-private (next launch = kernel boundary)
      if (p.system_scope) __threadfence_system(); else __threadfence();
      ctl->done_ctas = 0;
      if (ok) {
        for (int w = 0; w < p.num_workers; ++w)
          if (mask & (1u << w)) ctl->consumed[w] += p.ctas_per_push;
        const unsigned long long ngs = ctl->global_step + 1ull;
        ctl->global_step = ngs;
        ctl->applied_total += s_count;
        if (p.kind == 2) {
          ctl->beta1_power *= p.beta1;
          ctl->beta2_power *= p.beta2;
        }
        if (p.mode == 0) {
          // tokens (sync): every replica gets one carrying the NEW global step.  The fence above (after the ticket that
          // observed every CTA's release) + RELAXED system-scope stores = release; {token, version} travel as ONE
          // 16-byte store per mailbox, so no second fence is needed to order the pair
          for (int w = 0; w < p.num_workers; ++w)
            if (p.mailbox[w] != nullptr) st_relaxed_sys_v2_u64(&p.mailbox[w]->token, ngs, ngs);
        } else {
          // async: the pusher only; its token is a COUNT (red.add), so version and token are two accesses
          for (int w = 0; w < p.num_workers; ++w) {
            if (p.mailbox[w] == nullptr || !(mask & (1u << w))) continue;
            st_relaxed_sys_u64(reinterpret_cast<uint64_t*>(&p.mailbox[w]->version), ngs);
          }
          if (p.system_scope) __threadfence_system(); else __threadfence();
          for (int w = 0; w < p.num_workers; ++w) {
            if (p.mailbox[w] == nullptr || !(mask & (1u << w))) continue;
            red_relaxed_sys_add_u64(&p.mailbox[w]->token, 1ull);
          }
        }
      }
      if (p.trace && p.trace_cap > 0) {
        const unsigned long long slot = (s_seq - 1ull) % (unsigned long long)p.trace_cap;
        p.trace[slot * 4 + 0] = 1ull;
        p.trace[slot * 4 + 1] = t_start ? t_start : globaltimer_ns();
        p.trace[slot * 4 + 2] = globaltimer_ns();
        p.trace[slot * 4 + 3] = ctl->global_step;
      }
      __threadfence();
      st_release_gpu_u64(&ctl->param_version, s_seq);      // next launch's sequence number
      if (tr) { tr[3] = clock64(); tr[4] = blockIdx.x; }    // last block done (its own SM clock)
    }
    PSTAMP(5);
  }
}

// ---------------------------------------------------------------------------------------------
// ps init: build the bf16 shadow (and worker replicas) from the master, publish version/tokens
// ---------------------------------------------------------------------------------------------
__global__ void ps_publish_kernel(const float* __restrict__ master, __nv_bfloat16* __restrict__ shadow, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    shadow[i] = __float2bfloat16(master[i]);
}

// ---------------------------------------------------------------------------------------------
// Fused MLP head (SURVEY K2+K3+K4 small parts), ONE CTA:
//   logits = h.W2 + b2 ; softmax ; clipped cross-entropy (batch SUM) ; dlogits ;
//   dW2 = h^T.dlogits and db2 -> pushed to the ps slot ; dh = dlogits.W2^T * (h>0) (bf16, local) ;
//   db1 = colsum(dh) -> pushed ; stamp + arrival signal.
// B <= 128 rows, H <= 256 hidden units, C <= 16 classes.  W2 / b2 are read from the ps (peer
// pointers: bf16 shadow for W2, fp32 master for b2) -- the pull of the small parameters is fused here.
// ---------------------------------------------------------------------------------------------
struct MlpHeadParams {
  const __nv_bfloat16* h;       // [B, ldh] post-ReLU activations (bf16)
  long long ldh;
  const __nv_bfloat16* w2;      // [H, ldw2] bf16 (ps shadow; may be a peer pointer)
  long long ldw2;
  const float* b2;              // [C] fp32 (ps master; may be a peer pointer)
  const float* labels;          // [B, ldl]
  long long ldl;
  int B, H, C;
  float clip_min;
  float* loss_out;              // [1] local: batch-sum loss of this step
  float* loss_hist;             // optional [hist_cap] ring of losses
  unsigned long long* step_counter;   // local device step counter (incremented here)
  int hist_cap;
  __nv_bfloat16* dh;            // [B, lddh] local bf16 (feeds the dW1 GEMM)
  long long lddh;
  float* gw2;                   // ps slot: [H, ldgw2]
  long long ldgw2;
  float* gb2;                   // ps slot: [C]
  float* gb1;                   // ps slot: [H]
  float* logits_out;            // optional local [B, C]
  const WorkerMailbox* mailbox; // local mailbox (token already acquired by the first GEMM of the step)
  PsControl* ctl;               // ps control block (peer)
  int rank;                     // worker index
  int stamp_from_version;       // async: stamp = mailbox->version, sync: stamp = mailbox->token
  long long* phase_trace;       // optional [16] clock64 stamps
  float* h_acc;                 // optional [B, ld_acc] fp32 split-K partial sums of x.W1 (then h/ldh are unused)
  long long ld_acc;
  const float* b1;              // hidden bias (used with h_acc)
  int sys_scope;                // 1: ps is another GPU (system-scope fence before the arrival)
  int rows_per_cta;             // batch rows per CTA (grid = ceil(B / rows_per_cta))
};

__global__ void __launch_bounds__(512, 1) mlp_head_kernel(const MlpHeadParams p) {
  DTF_DYN_SMEM(float, sm);
  // Row-parallel: CTA i owns batch rows [i*rows_per_cta, ...).  Row-owned work (logits, softmax, dh) is disjoint;
  // the batch reductions (dW2, db2, db1, loss) are combined with fp32 atomics -- into the ps slot over NVLink
  // for the gradients (the ps clears those ranges after reading them).
  const int row_lo = blockIdx.x * p.rows_per_cta;
  const int B = min(p.rows_per_cta, p.B - row_lo);
  const int H = p.H, C = p.C;
  const bool multi = gridDim.x > 1;
  if (B <= 0) return;
  constexpr int CP = 16;                    // classes padded to 16 (4 x float4)
  constexpr int DS = 20;                    // row stride of the [B][16] tiles: 80 B keeps float4 rows conflict-free
  const int HP = H + 1;                     // padded activation row (bank-conflict-free column walks)
  float* s_h = sm;                          // [B][H+1]   activations h (fp32)
  float* s_w2 = s_h + (size_t)B * HP;       // [H][CP]
  float* s_dl = s_w2 + (size_t)H * CP;      // [B][DS]    logits -> dlogits
  float* s_lab = s_dl + (size_t)B * DS;     // [B][DS]    labels
  float* s_b2 = s_lab + (size_t)B * DS;     // [CP]
  float* s_red = s_b2 + CP;                 // [32]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31;
  long long* tr = p.phase_trace;
#define HSTAMP(slot) do { if (tr && tid == 0) tr[slot] = clock64(); } while (0)
  HSTAMP(0);

  // ---- phase A: issue every global load up front (128-bit where possible), then fill shared memory ----------
  {
    const int wv = (int)(p.ldw2 / 8);                      // uint4 per row of w2 (ldw2 == 16 -> 2)
    const int nwvec = H * wv;
    uint4 wreg[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = tid + u * nt;
      wreg[u] = (i < nwvec) ? reinterpret_cast<const uint4*>(p.w2)[i] : make_uint4(0, 0, 0, 0);
    }
    float lreg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * nt;
      lreg[u] = (i < B * C) ? p.labels[(long long)(row_lo + i / C) * p.ldl + (i % C)] : 0.f;
    }
    const float b2v = (tid < C) ? p.b2[tid] : 0.f;
    if (p.h_acc == nullptr) {
      // h arrives as bf16 (bias + ReLU already applied by the GEMM epilogue)
      const int hv = (int)(p.ldh / 8);
      const int nvec = B * hv;
      uint4 hreg[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * nt;
        hreg[u] = (i < nvec) ? reinterpret_cast<const uint4*>(p.h + (long long)row_lo * p.ldh)[i] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * nt;
        if (i < nvec) {
          const int b = i / hv, k0 = (i - b * hv) * 8;
          const uint32_t w[4] = {hreg[u].x, hreg[u].y, hreg[u].z, hreg[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + 2 * j;
            if (k < H) s_h[b * HP + k] = __uint_as_float(w[j] << 16);
            if (k + 1 < H) s_h[b * HP + k + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
          }
        }
      }
    } else {
      // h arrives as split-K fp32 partial sums (x.W1 without bias): finish it here -- h = relu(acc + b1) --
      // and clear the accumulator for the next step (zero-after-read)
      const int hv4 = (int)(p.ld_acc / 4);
      const int nvec = B * hv4;
      float4 areg[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * nt;
        areg[u] = (i < nvec) ? reinterpret_cast<const float4*>(p.h_acc + (long long)row_lo * p.ld_acc)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float* s_b1 = s_h + (size_t)B * HP - 0;       // staged below through s_w2 region? no: use registers
      (void)s_b1;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = tid + u * nt;
        if (i < nvec) {
          reinterpret_cast<float4*>(p.h_acc + (long long)row_lo * p.ld_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int b = i / hv4, k0 = (i - b * hv4) * 4;
          const float v[4] = {areg[u].x, areg[u].y, areg[u].z, areg[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = k0 + j;
            if (k < H) s_h[b * HP + k] = fmaxf(v[j] + __ldg(p.b1 + k), 0.f);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = tid + u * nt;
      if (i < nwvec) {
        const int k = i / wv, c0 = (i - k * wv) * 8;
        const uint32_t w[4] = {wreg[u].x, wreg[u].y, wreg[u].z, wreg[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + 2 * j;
          if (c < CP) s_w2[k * CP + c] = (c < C) ? __uint_as_float(w[j] << 16) : 0.f;
          if (c + 1 < CP) s_w2[k * CP + c + 1] = (c + 1 < C) ? __uint_as_float(w[j] & 0xFFFF0000u) : 0.f;
        }
      }
    }
    for (int i = tid; i < B * DS; i += nt) s_lab[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * nt;
      if (i < B * C) s_lab[(i / C) * DS + (i % C)] = lreg[u];
    }
    if (tid < CP) s_b2[tid] = b2v;
  }
  __syncthreads();
  HSTAMP(1);          // inputs in shared memory

  // The three small products are shared-memory-bandwidth bound on CUDA cores (ncu: mio_throttle), so the
  // thread mappings are chosen to minimise shared-memory wavefronts per FMA, not instruction count.
  // ---- logits[b][:] = h[b][:] . W2 + b2 : a warp covers 8 rows x 4 class-quads; per k it issues one scalar LDS
  //      (8 distinct rows, broadcast to the 4 quads) and one LDS.128 (one 64-byte W2 row, broadcast to the 8 rows)
  {
    const int warp = tid >> 5;
    const int r = lane >> 2, j = lane & 3;
    for (int row0 = warp * 8; row0 < B; row0 += (nt >> 5) * 8) {
      const int b = row0 + r;
      const float* hr = s_h + (b < B ? b : 0) * HP;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float hv = hr[k];
        const float4 w = *reinterpret_cast<const float4*>(s_w2 + k * CP + 4 * j);
        a0 = fmaf(hv, w.x, a0);
        a1 = fmaf(hv, w.y, a1);
        a2 = fmaf(hv, w.z, a2);
        a3 = fmaf(hv, w.w, a3);
      }
      if (b < B)
        *reinterpret_cast<float4*>(s_dl + b * DS + 4 * j) =
            make_float4(a0 + s_b2[4 * j], a1 + s_b2[4 * j + 1], a2 + s_b2[4 * j + 2], a3 + s_b2[4 * j + 3]);
    }
  }
  __syncthreads();
  HSTAMP(2);          // logits done
  if (p.logits_out) {
    for (int i = tid; i < B * C; i += nt) p.logits_out[(long long)row_lo * C + i] = s_dl[(i / C) * DS + (i % C)];
    __syncthreads();    // the softmax below rewrites s_dl in place (uniform branch: logits_out is a kernel argument);
                        // found by the host-emulation tier -- without it the optional copy raced with the row owners
  }

  // ---- softmax + clipped xent + dlogits, one thread per row (rows read/written as float4) ---------------------------
  float my_loss = 0.f;
  if (tid < B) {
    float z[CP], lab[CP];
    float4* zr = reinterpret_cast<float4*>(s_dl + tid * DS);
    const float4* lr = reinterpret_cast<const float4*>(s_lab + tid * DS);
#pragma unroll
    for (int j = 0; j < CP / 4; ++j) {
      const float4 a = zr[j], l4 = lr[j];
      z[4 * j] = a.x; z[4 * j + 1] = a.y; z[4 * j + 2] = a.z; z[4 * j + 3] = a.w;
      lab[4 * j] = l4.x; lab[4 * j + 1] = l4.y; lab[4 * j + 2] = l4.z; lab[4 * j + 3] = l4.w;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CP; ++c) if (c < C) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    float y[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      y[c] = (c < C) ? __expf(z[c] - mx) : 0.f;
      se += y[c];
    }
    const float inv = 1.f / se;
    const float lse = mx + __logf(se);
    float tsum = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      if (c < C) {
        y[c] *= inv;
        const float l = lab[c];
        if (p.clip_min > 0.f) {
          // log(clamp(y, clip, 1)) == max(log y, log clip) for y <= 1: one exact log-softmax, no extra MUFU per class
          my_loss -= l * fmaxf(z[c] - lse, __logf(p.clip_min));
          tsum += (y[c] >= p.clip_min) ? l : 0.f;
        } else {
          my_loss -= l * (z[c] - lse);
          tsum += l;
        }
      }
    }
    float g[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      g[c] = 0.f;
      if (c < C) {
        const float t = (p.clip_min > 0.f && y[c] < p.clip_min) ? 0.f : lab[c];
        g[c] = y[c] * tsum - t;
      }
    }
#pragma unroll
    for (int j = 0; j < CP / 4; ++j) zr[j] = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
  }
  for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  if (lane == 0) s_red[tid >> 5] = my_loss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < (nt >> 5); ++i) t += s_red[i];
    p.loss_out[blockIdx.x] = t;            // per-CTA partial of the batch-sum loss (the reader adds them up)
    if (blockIdx.x == 0) {
      unsigned long long step = 0;
      if (p.step_counter) { step = *p.step_counter; *p.step_counter = step + 1; }
      if (!multi && p.loss_hist && p.hist_cap > 0) p.loss_hist[step % (unsigned long long)p.hist_cap] = t;
    }
  }
  HSTAMP(3);          // softmax/xent/dlogits done

  // ---- dW2[k][:] = sum_b h[b][k] * dl[b][:] : a warp covers 8 hidden units x 4 class-quads; per row b one scalar
  //      LDS (8 consecutive k of row b) and one LDS.128 (the 64-byte dl row, broadcast) ---------------------------------
  {
    const int warp = tid >> 5;
    const int r = lane >> 2, j = lane & 3;
    for (int k0 = warp * 8; k0 < H; k0 += (nt >> 5) * 8) {
      const int k = k0 + r;
      const int kk = k < H ? k : 0;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
      for (int b = 0; b < B; ++b) {
        const float hv = s_h[b * HP + kk];
        const float4 d = *reinterpret_cast<const float4*>(s_dl + b * DS + 4 * j);
        a0 = fmaf(hv, d.x, a0);
        a1 = fmaf(hv, d.y, a1);
        a2 = fmaf(hv, d.z, a2);
        a3 = fmaf(hv, d.w, a3);
      }
      if (k < H) {
        float* gw = p.gw2 + (long long)k * p.ldgw2 + 4 * j;      // NVLink stores straight into the ps slot
        if (multi) {
          if (4 * j < C) atomicAdd(gw, a0);
          if (4 * j + 1 < C) atomicAdd(gw + 1, a1);
          if (4 * j + 2 < C) atomicAdd(gw + 2, a2);
          if (4 * j + 3 < C) atomicAdd(gw + 3, a3);
        } else {
          if (4 * j < C) gw[0] = a0;
          if (4 * j + 1 < C) gw[1] = a1;
          if (4 * j + 2 < C) gw[2] = a2;
          if (4 * j + 3 < C) gw[3] = a3;
        }
      }
    }
  }
  // ---- dh[b][k] = (h[b][k] > 0) * dl[b][:] . W2[k][:] ; db1[k] = sum_b dh[b][k] : a thread owns hidden unit k (its W2
  //      row lives in registers) for a quarter of the rows; per row one scalar LDS (32 consecutive k: conflict-free)
  //      and four broadcast LDS.128 (the dl row).  dh rows are written 64 contiguous bytes per warp. -------------------
  {
    const int QR = (nt / ((H + 31) / 32 * 32)) > 0 ? (nt / ((H + 31) / 32 * 32)) : 1;    // row groups
    const int HPAD = (H + 31) / 32 * 32;
    const int k = tid % HPAD, g = tid / HPAD;
    float db1 = 0.f;
    if (g < QR && k < H) {
      float w2r[CP];
      const float4* wr = reinterpret_cast<const float4*>(s_w2 + k * CP);
#pragma unroll
      for (int jj = 0; jj < CP / 4; ++jj) {
        const float4 w = wr[jj];
        w2r[4 * jj] = w.x; w2r[4 * jj + 1] = w.y; w2r[4 * jj + 2] = w.z; w2r[4 * jj + 3] = w.w;
      }
      for (int b = g; b < B; b += QR) {
        const float hv = s_h[b * HP + k];
        const float4* dr = reinterpret_cast<const float4*>(s_dl + b * DS);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int jj = 0; jj < CP / 4; ++jj) {
          const float4 dl = dr[jj];
          d0 = fmaf(dl.x, w2r[4 * jj], d0);
          d1 = fmaf(dl.y, w2r[4 * jj + 1], d1);
          d0 = fmaf(dl.z, w2r[4 * jj + 2], d0);
          d1 = fmaf(dl.w, w2r[4 * jj + 3], d1);
        }
        const float d = hv > 0.f ? (d0 + d1) : 0.f;
        p.dh[(long long)(row_lo + b) * p.lddh + k] = __float2bfloat16(d);
        db1 += d;
      }
    }
    // combine the row groups through shared memory (the logits tile rows >= ... are free: use s_lab as scratch)
    __syncthreads();
    float* s_part = s_red + 32;              // [QR][HPAD] <= 512 floats, reserved by the launcher
    if (g < QR && k < HPAD) s_part[g * HPAD + k] = db1;
    __syncthreads();
    if (tid < H) {
      float a = 0.f;
      for (int gg = 0; gg < QR; ++gg) a += s_part[gg * HPAD + tid];
      if (multi) atomicAdd(p.gb1 + tid, a); else p.gb1[tid] = a;
    }
  }
  if (tid >= nt - 32 && tid - (nt - 32) < C) {
    const int c = tid - (nt - 32);
    float a0 = 0.f, a1 = 0.f;
    int b = 0;
    for (; b + 2 <= B; b += 2) { a0 += s_dl[b * DS + c]; a1 += s_dl[(b + 1) * DS + c]; }
    if (b < B) a0 += s_dl[b * DS + c];
    if (multi) atomicAdd(p.gb2 + c, a0 + a1); else p.gb2[c] = a0 + a1;
  }
  HSTAMP(4);          // dW2/dh/db1 done and stored
  // ---- stamp + arrival (optional: a later kernel of the same step may signal for the whole push instead) --------
  if (p.ctl != nullptr) {
    __syncthreads();
    if (tid == 0) {
      // ONE fence by one thread after the CTA barrier: release is cumulative, and a system-scope membar per
      // thread serialises for tens of microseconds
      if (p.sys_scope) __threadfence_system(); else __threadfence();
      const unsigned long long stamp = p.stamp_from_version ? p.mailbox->version : p.mailbox->token;
      st_relaxed_sys_ull(&p.ctl->w[p.rank].stamp, stamp);
      red_release_sys_add_u64(reinterpret_cast<uint64_t*>(&p.ctl->w[p.rank].arrivals), 1ull);
    }
  }
  HSTAMP(6);          // fenced + signalled
}

// Input-pipeline stage for the device-resident dataset: batch index = (step * stride + offset) % nbatches,
// where step is the worker's DEVICE step counter (so the launch is CUDA-graph replayable).  Converts the
// fp32 images of that batch to the bf16 staging tile consumed by both GEMMs and copies the labels.
__global__ void stage_from_dataset_kernel(const float* __restrict__ images, const float* __restrict__ labels,
                                          long long nbatches, int B, int D, int C, long long stride, long long offset,
                                          const unsigned long long* __restrict__ step_counter,
                                          __nv_bfloat16* __restrict__ x16, float* __restrict__ lab_out) {
  const unsigned long long step = step_counter ? *step_counter : 0ull;
  const long long bi = (long long)((step * (unsigned long long)stride + (unsigned long long)offset) %
                                   (unsigned long long)nbatches);
  const float4* src = reinterpret_cast<const float4*>(images + bi * (long long)B * D);
  uint2* dst = reinterpret_cast<uint2*>(x16);
  const long long n4 = (long long)B * D / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  const float* ls = labels + bi * (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * C;
       i += (long long)gridDim.x * blockDim.x)
    lab_out[i] = ls[i];
}

// wait (on the worker) until the mailbox token reaches `target`: used when the step's first kernel
// is not a GEMM with a fused wait (generic models), and by tests.
__global__ void wait_token_kernel(const WorkerMailbox* mb, unsigned long long target,
                                  const unsigned long long* target_ptr, unsigned long long timeout_ns,
                                  unsigned int* err) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long t = target + (target_ptr ? *target_ptr : 0ull);
    if (!wait_flag_ge_u64(reinterpret_cast<const uint64_t*>(&mb->token), t, timeout_ns) && err) atomicExch(err, 3u);
  }
}

// push an already-computed local gradient buffer into the ps slot (generic models / graph mode):
// vectorised NVLink stores + stamp + arrival.  One "push" = gridDim.x arrivals.
__global__ void push_grad_kernel(const float* __restrict__ src, float* __restrict__ dst_peer, long long n, PsControl* ctl,
                                 const WorkerMailbox* mb, int rank, int stamp_from_version, int write_stamp) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst_peer)[i] = reinterpret_cast<const float4*>(src)[i];
  if (blockIdx.x == 0)
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) dst_peer[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (write_stamp && blockIdx.x == 0) {
      const unsigned long long stamp = stamp_from_version ? mb->version : mb->token;
      st_relaxed_sys_ull(&ctl->w[rank].stamp, stamp);
      __threadfence_system();
    }
    red_release_sys_add_u64(reinterpret_cast<uint64_t*>(&ctl->w[rank].arrivals), 1ull);
  }
}

// pull the bf16 shadow from the ps into a local replica (generic models): peer loads, 128-bit
__global__ void pull_shadow_kernel(const uint4* __restrict__ src_peer, uint4* __restrict__ dst, long long n16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src_peer[i];
}

// ---------------------------------------------------------------------------------------------
// Fabric collectives as plain kernels (generic models, fabric tests, push/pull bandwidth measurements):
//   broadcast: one source buffer -> every GPU's copy.   NVLS: one multimem.st per 16 bytes (the switch fans out);
//              unicast: one peer store per destination (npeers x the egress bytes).
//   reduce:    element-wise sum of every GPU's copy.     NVLS: one multimem.ld_reduce per 16 bytes (in-switch add);
//              unicast: one peer load per source (npeers x the ingress bytes).
// ---------------------------------------------------------------------------------------------
struct PeerList {
  void* p[DTF_MAX_WORKERS];
};

// Each thread keeps U 16-byte accesses in flight (independent loads first, then the stores / adds): a single multimem
// round trip through the switch is microseconds, so bandwidth comes from memory-level parallelism -- U x 16 B per thread
// x 256 threads x the resident CTAs must cover (link rate x round-trip time) ~ 900 GB/s x 3 us ~ 2.7 MB.
template <int U>
__global__ void __launch_bounds__(256) fabric_bcast_kernel(const uint4* __restrict__ src, uint4* mc_dst, PeerList peers, int npeers,
                                                           long long n16) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n16) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long j = i + u * stride;
      if (j >= n16) break;
      if (mc_dst != nullptr) {
        multimem_st_b128(mc_dst + j, v[u]);
      } else {
        for (int k = 0; k < npeers; ++k) reinterpret_cast<uint4*>(peers.p[k])[j] = v[u];
      }
    }
  }
}

template <int U>
__global__ void __launch_bounds__(256) fabric_reduce_kernel(const float* mc_src, PeerList peers, int npeers, float* __restrict__ dst,
                                                            long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += U * stride) {
    float4 acc[U];
    if (mc_src != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (i + u * stride < n4) acc[u] = multimem_ld_reduce_add_f32x4(mc_src + 4 * (i + u * stride));
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < npeers; ++k) {
        float4 t[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + u * stride < n4) t[u] = reinterpret_cast<const float4*>(peers.p[k])[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (i + u * stride < n4) { acc[u].x += t[u].x; acc[u].y += t[u].y; acc[u].z += t[u].z; acc[u].w += t[u].w; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n4) reinterpret_cast<float4*>(dst)[i + u * stride] = acc[u];
  }
}

}  // namespace dtf

extern "C" {
using namespace dtf;

// grid <= 0 / unroll <= 0: defaults tuned on 8 x B200 (tools/nvls_check.py sweep): 8 accesses in flight per thread, 4 CTAs / SM
int dtf_fabric_bcast_ex(const void* src, void* mc_dst, void* const* peer_dst, int npeers, long long nbytes, int grid, int unroll,
                        cudaStream_t s) {
  if (nbytes % 16 || npeers > DTF_MAX_WORKERS) return -2;
  PeerList pl;
  memset(&pl, 0, sizeof(pl));
  for (int k = 0; k < npeers; ++k) pl.p[k] = peer_dst[k];
  const int g = grid > 0 ? grid : 592;
  const uint4* sp = reinterpret_cast<const uint4*>(src);
  uint4* mp = reinterpret_cast<uint4*>(mc_dst);
  if (unroll >= 16) { DTF_LAUNCH(fabric_bcast_kernel<16>, g, 256, s, sp, mp, pl, npeers, nbytes / 16); }
  else if (unroll == 4) { DTF_LAUNCH(fabric_bcast_kernel<4>, g, 256, s, sp, mp, pl, npeers, nbytes / 16); }
  else { DTF_LAUNCH(fabric_bcast_kernel<8>, g, 256, s, sp, mp, pl, npeers, nbytes / 16); }
  return (int)cudaGetLastError();
}

int dtf_fabric_reduce_ex(const void* mc_src, void* const* peer_src, int npeers, float* dst, long long nfloats, int grid, int unroll,
                         cudaStream_t s) {
  if (nfloats % 4 || npeers > DTF_MAX_WORKERS) return -2;
  PeerList pl;
  memset(&pl, 0, sizeof(pl));
  for (int k = 0; k < npeers; ++k) pl.p[k] = peer_src[k];
  const int g = grid > 0 ? grid : 592;
  const float* mp = reinterpret_cast<const float*>(mc_src);
  if (unroll >= 16) { DTF_LAUNCH(fabric_reduce_kernel<16>, g, 256, s, mp, pl, npeers, dst, nfloats / 4); }
  else if (unroll == 4) { DTF_LAUNCH(fabric_reduce_kernel<4>, g, 256, s, mp, pl, npeers, dst, nfloats / 4); }
  else { DTF_LAUNCH(fabric_reduce_kernel<8>, g, 256, s, mp, pl, npeers, dst, nfloats / 4); }
  return (int)cudaGetLastError();
}

int dtf_fabric_bcast(const void* src, void* mc_dst, void* const* peer_dst, int npeers, long long nbytes, int grid, cudaStream_t s) {
  return dtf_fabric_bcast_ex(src, mc_dst, peer_dst, npeers, nbytes, grid, 0, s);
}

int dtf_fabric_reduce(const void* mc_src, void* const* peer_src, int npeers, float* dst, long long nfloats, int grid, cudaStream_t s) {
  return dtf_fabric_reduce_ex(mc_src, peer_src, npeers, dst, nfloats, grid, 0, s);
}

int dtf_sizeof_ps_control() { return (int)sizeof(PsControl); }
int dtf_sizeof_mailbox() { return (int)sizeof(WorkerMailbox); }
int dtf_offsetof_ctl(int which) {
  switch (which) {
    case 0: return (int)offsetof(PsControl, global_step);
    case 1: return (int)offsetof(PsControl, param_version);
    case 2: return (int)offsetof(PsControl, beta1_power);
    case 3: return (int)offsetof(PsControl, beta2_power);
    case 4: return (int)offsetof(PsControl, dropped_stale);
    case 5: return (int)offsetof(PsControl, applied_total);
    case 6: return (int)offsetof(PsControl, staleness_hist);
    case 7: return (int)offsetof(PsControl, staleness_sum);
    case 8: return (int)offsetof(PsControl, err);
    case 9: return (int)offsetof(PsControl, w);
    case 10: return (int)sizeof(WorkerSlotState);
    case 11: return (int)offsetof(PsControl, consumed);
  }
  return -1;
}

struct DtfPsApplyArgs {
  void* ctl;
  float* master;
  float* slot_m;
  float* slot_v;
  float* grad[DTF_MAX_WORKERS];
  void* shadow;
  void* replica[DTF_MAX_WORKERS];
  void* mailbox[DTF_MAX_WORKERS];
  long long n;
  int num_workers, replicas_to_aggregate;
  unsigned int ctas_per_push;
  int mode, kind;
  float lr, momentum, beta1, beta2, eps;
  int nesterov, publish_replicas;
  long long zero_begin[4], zero_end[4];
  int num_zero;
  unsigned long long timeout_ns;
  unsigned long long* trace;
  int trace_cap;
  int grid;
  int system_scope;
  long long* phase_trace;
  int idle_ok;
  const float* grad_mc;
  void* shadow_mc;
  float* master_mc;
  unsigned long long* token_mc;
};

// default grid: one float4 per thread (a single load round trip) for small shards, up to 4 per thread per pass for large
// ones; all CTAs must be co-resident (blocks other than the decider spin): 2 per SM by launch bounds -> at most 296
int dtf_ps_apply_grid(long long n) {
  long long want = (n / 4 + 255) / 256;
  return (int)(want < 1 ? 1 : (want > 296 ? 296 : want));
}

int dtf_ps_apply(const DtfPsApplyArgs* a, cudaStream_t s) {
  PsApplyParams p;
  memset(&p, 0, sizeof(p));
  p.ctl = reinterpret_cast<PsControl*>(a->ctl);
  p.master = a->master; p.slot_m = a->slot_m; p.slot_v = a->slot_v;
  for (int w = 0; w < DTF_MAX_WORKERS; ++w) {
    p.grad[w] = a->grad[w]; p.grad_rw[w] = a->grad[w];
    p.replica[w] = reinterpret_cast<__nv_bfloat16*>(a->replica[w]);
    p.mailbox[w] = reinterpret_cast<WorkerMailbox*>(a->mailbox[w]);
  }
  p.shadow = reinterpret_cast<__nv_bfloat16*>(a->shadow);
  p.n = a->n; p.num_workers = a->num_workers; p.replicas_to_aggregate = a->replicas_to_aggregate;
  p.ctas_per_push = a->ctas_per_push; p.mode = a->mode; p.kind = a->kind;
  p.lr = a->lr; p.momentum = a->momentum; p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps;
  p.nesterov = a->nesterov; p.publish_replicas = a->publish_replicas; p.system_scope = a->system_scope;
  for (int z = 0; z < 4; ++z) { p.zero_begin[z] = a->zero_begin[z]; p.zero_end[z] = a->zero_end[z]; }
  p.num_zero = a->num_zero;
  p.timeout_ns = a->timeout_ns ? a->timeout_ns : 2000000000ull;
  p.trace = a->trace; p.trace_cap = a->trace_cap; p.phase_trace = a->phase_trace; p.idle_ok = a->idle_ok;
  p.grad_mc = a->grad_mc; p.grad_mc_rw = const_cast<float*>(a->grad_mc);
  p.shadow_mc = reinterpret_cast<__nv_bfloat16*>(a->shadow_mc);
  p.master_mc = a->master_mc;
  p.token_mc = a->token_mc;
  p.full_mask = a->num_workers >= 32 ? 0xFFFFFFFFu : ((1u << a->num_workers) - 1u);
  int grid = a->grid;
  if (grid <= 0) grid = dtf_ps_apply_grid(a->n);
#ifndef DTF_HOST_EMU
  if (pdl_enabled() && !a->system_scope) {      // colocated ps (same GPU and stream as the worker): see csrc/mlp_step.cu
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, ps_apply_kernel, p);
  }
#endif
  DTF_LAUNCH(ps_apply_kernel, grid, 256, s, p);
  return (int)cudaGetLastError();
}

int dtf_ps_publish(const float* master, void* shadow, long long n, cudaStream_t s) {
  long long g = (n + 255) / 256;
  if (g > 1184) g = 1184;
  DTF_LAUNCH(ps_publish_kernel, (int)g, 256, s, master, reinterpret_cast<__nv_bfloat16*>(shadow), n);
  return (int)cudaGetLastError();
}

struct DtfMlpHeadArgs {
  const void* h; long long ldh;
  const void* w2; long long ldw2;
  const float* b2;
  const float* labels; long long ldl;
  int B, H, C;
  float clip_min;
  float* loss_out; float* loss_hist; unsigned long long* step_counter; int hist_cap;
  void* dh; long long lddh;
  float* gw2; long long ldgw2;
  float* gb2; float* gb1;
  float* logits_out;
  const void* mailbox; void* ctl;
  int rank; int stamp_from_version;
  long long* phase_trace;
  float* h_acc; long long ld_acc; const float* b1; int sys_scope;
  int ctas;
};

int dtf_mlp_head(const DtfMlpHeadArgs* a, cudaStream_t s) {
  if (a->B > 128 || a->H > 256 || a->C > 16) return -2;
  MlpHeadParams p;
  p.h = reinterpret_cast<const __nv_bfloat16*>(a->h); p.ldh = a->ldh;
  p.w2 = reinterpret_cast<const __nv_bfloat16*>(a->w2); p.ldw2 = a->ldw2;
  p.b2 = a->b2; p.labels = a->labels; p.ldl = a->ldl; p.B = a->B; p.H = a->H; p.C = a->C; p.clip_min = a->clip_min;
  p.loss_out = a->loss_out; p.loss_hist = a->loss_hist; p.step_counter = a->step_counter; p.hist_cap = a->hist_cap;
  p.dh = reinterpret_cast<__nv_bfloat16*>(a->dh); p.lddh = a->lddh;
  p.gw2 = a->gw2; p.ldgw2 = a->ldgw2; p.gb2 = a->gb2; p.gb1 = a->gb1; p.logits_out = a->logits_out;
  p.mailbox = reinterpret_cast<const WorkerMailbox*>(a->mailbox); p.ctl = reinterpret_cast<PsControl*>(a->ctl);
  p.rank = a->rank; p.stamp_from_version = a->stamp_from_version; p.phase_trace = a->phase_trace;
  p.h_acc = a->h_acc; p.ld_acc = a->ld_acc; p.b1 = a->b1; p.sys_scope = a->sys_scope;
  int ctas = a->ctas > 1 ? a->ctas : 1;
  p.rows_per_cta = (a->B + ctas - 1) / ctas;
  p.rows_per_cta = (p.rows_per_cta + 7) / 8 * 8;            // whole 8-row warp groups
  ctas = (a->B + p.rows_per_cta - 1) / p.rows_per_cta;
  const size_t smem = sizeof(float) * ((size_t)a->B * (a->H + 1) + (size_t)a->H * 16 + (size_t)a->B * 40 + 16 + 32 + 512);
  if (a->B > 512 || a->H > 512) return -2;
  if ((a->ldh % 8) || (a->ldw2 % 8) || a->ldh > 8 * 8 * 512 / a->B) return -3;
#ifndef DTF_HOST_EMU
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(mlp_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    configured[dev] = true;
  }
#endif
  DTF_LAUNCH_SMEM(mlp_head_kernel, ctas, 512, smem, s, p);
  return (int)cudaGetLastError();
}

int dtf_stage_from_dataset(const float* images, const float* labels, long long nbatches, int B, int D, int C,
                           long long stride, long long offset, const unsigned long long* step_counter, void* x16,
                           float* lab_out, cudaStream_t s) {
  if (((long long)B * D) % 4) return -2;
  DTF_LAUNCH(stage_from_dataset_kernel, 40, 256, s, images, labels, nbatches, B, D, C, stride, offset, step_counter,
             reinterpret_cast<__nv_bfloat16*>(x16), lab_out);
  return (int)cudaGetLastError();
}

int dtf_wait_token(const void* mailbox, unsigned long long target, const unsigned long long* target_ptr,
                   unsigned long long timeout_ns, unsigned int* err, cudaStream_t s) {
  DTF_LAUNCH(wait_token_kernel, 1, 32, s, reinterpret_cast<const WorkerMailbox*>(mailbox), target, target_ptr,
             timeout_ns ? timeout_ns : 2000000000ull, err);
  return (int)cudaGetLastError();
}

int dtf_push_grad(const float* src, float* dst_peer, long long n, void* ctl, const void* mailbox, int rank,
                  int stamp_from_version, int write_stamp, int grid, cudaStream_t s) {
  DTF_LAUNCH(push_grad_kernel, grid, 256, s, src, dst_peer, n, reinterpret_cast<PsControl*>(ctl),
             reinterpret_cast<const WorkerMailbox*>(mailbox), rank, stamp_from_version, write_stamp);
  return (int)cudaGetLastError();
}

int dtf_pull_shadow(const void* src_peer, void* dst, long long nbytes, int grid, cudaStream_t s) {
  DTF_LAUNCH(pull_shadow_kernel, grid, 256, s, reinterpret_cast<const uint4*>(src_peer), reinterpret_cast<uint4*>(dst), nbytes / 16);
  return (int)cudaGetLastError();
}

#ifndef DTF_HOST_EMU
// ---------------------------------------------------------------------------------------------
// fabric: IPC-exportable device allocations + peer access
// ---------------------------------------------------------------------------------------------
int dtf_fabric_alloc(void** out, long long bytes) {
  cudaError_t e = cudaMalloc(out, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*out, 0, (size_t)bytes);
}
int dtf_fabric_free(void* p) { return (int)cudaFree(p); }
int dtf_fabric_export(void* p, void* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int dtf_fabric_import(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}
int dtf_fabric_close(void* p) { return (int)cudaIpcCloseMemHandle(p); }
int dtf_enable_peer(int peer) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return (int)e;
}
int dtf_can_access_peer(int dev, int peer) {
  int ok = 0;
  cudaDeviceCanAccessPeer(&ok, dev, peer);
  return ok;
}
int dtf_ipc_handle_size() { return (int)sizeof(cudaIpcMemHandle_t); }
int dtf_memcpy_d2h(void* dst, const void* src, long long n) { return (int)cudaMemcpy(dst, src, (size_t)n, cudaMemcpyDeviceToHost); }
int dtf_memcpy_h2d(void* dst, const void* src, long long n) { return (int)cudaMemcpy(dst, src, (size_t)n, cudaMemcpyHostToDevice); }
int dtf_memset(void* p, int v, long long n, cudaStream_t s) { return (int)cudaMemsetAsync(p, v, (size_t)n, s); }

#endif  // !DTF_HOST_EMU

}  // extern "C"
