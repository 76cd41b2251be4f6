// Native step executor: one C call enqueues a whole training step (host->device copies of the batch, every kernel
// of the step, the device->host read of the loss) and -- optionally -- waits for it.
//
// The Python engines build an op list ONCE per rank (pointers into their ctypes launch descriptors); per step only
// the host source pointers of the H2D ops change.  This is the host half of the "launch-bound inner loop" story:
// the device half is CUDA-graph replay (bench kernel-only number), this is the end-to-end path a user's
// `engine.step(x_pinned, y_pinned) -> loss` call takes (SURVEY A8 feed_dict + fetch of the loss, C7).
//
// Role in the reference: the per-step `mon_sess.run([train_step, global_step, loss], feed_dict=...)` round trip
// (distributed_mnist.py:152) -- feed copied to the device, the step executed, the fetched loss copied back.
#include <cstdio>
#include <cstring>

// DTF_HOST_EMU: compiled by g++ against tests/emu/step_exec_stubs.h, where the CUDA runtime calls and the kernel launchers
// record a trace instead of running -- the op dispatch below is unit-tested without a GPU (tests/test_step_exec_host.py).
#ifdef DTF_HOST_EMU
#include "step_exec_stubs.h"
#else
#include <cuda_runtime.h>
#endif

extern "C" {

#ifndef DTF_HOST_EMU
// entry points of the other translation units (same shared object)
struct DtfGemmArgs;
struct DtfMlpHeadArgs;
struct DtfPsApplyArgs;
struct DtfMlpStepArgs;
int dtf_gemm_bf16(const DtfGemmArgs* g, cudaStream_t stream);
int dtf_mlp_head(const DtfMlpHeadArgs* a, cudaStream_t s);
int dtf_ps_apply(const DtfPsApplyArgs* a, cudaStream_t s);
int dtf_mlp_step(const DtfMlpStepArgs* a, cudaStream_t s);
int dtf_convert_f32_bf16(const float* in, long long ld_in, void* out, long long ld_out, long long rows, long long cols,
                         long long cols_pad, cudaStream_t s);
int dtf_wait_token(const void* mailbox, unsigned long long target, const unsigned long long* target_ptr,
                   unsigned long long timeout_ns, unsigned int* err, cudaStream_t s);
int dtf_push_grad(const float* src, float* dst_peer, long long n, void* ctl, const void* mailbox, int rank,
                  int stamp_from_version, int write_stamp, int grid, cudaStream_t s);
int dtf_stage_from_dataset(const float* images, const float* labels, long long nbatches, int B, int D, int C,
                           long long stride, long long offset, const unsigned long long* step_counter, void* x16,
                           float* lab_out, cudaStream_t s);
#endif

enum DtfOpKind {
  DTF_OP_H2D = 1,         // p0 = dst (device), p1 = src (pinned host), i0 = bytes
  DTF_OP_D2H = 2,         // p0 = dst (pinned host), p1 = src (device), i0 = bytes
  DTF_OP_CONVERT = 3,     // p0 = in f32, p1 = out bf16, i0 = ld_in, i1 = ld_out, i2 = rows, i3 = cols, i4 = cols_pad
  DTF_OP_GEMM = 4,        // p0 = DtfGemmArgs*
  DTF_OP_HEAD = 5,        // p0 = DtfMlpHeadArgs*
  DTF_OP_PS_APPLY = 6,    // p0 = DtfPsApplyArgs*
  DTF_OP_WAIT_TOKEN = 7,  // p0 = mailbox, p1 = target_ptr, p2 = err, i0 = target, u0 = timeout_ns
  DTF_OP_SIGNAL = 8,      // p0 = ctl, p1 = mailbox, i0 = worker index, i1 = stamp_from_version
  DTF_OP_STAGE = 9,       // p0 = images, p1 = labels, p2 = step counter, p3 = x16, p4 = labels out, i0 = nbatches, i1 = B,
                          // i2 = D, i3 = C, i4 = stride, i5 = offset
  DTF_OP_SYNC = 10,       // cudaStreamSynchronize
  DTF_OP_EVENT_RECORD = 11,   // p0 = cudaEvent_t
  DTF_OP_EVENT_WAIT = 12,     // p0 = cudaEvent_t (cross-stream dependency)
  DTF_OP_GRAPH = 13,          // p0 = cudaGraphExec_t captured from a kernel-only op range (dtf_capture_ops)
  DTF_OP_MLP_STEP = 14        // p0 = DtfMlpStepArgs*: the whole worker step of the MLP as one kernel (csrc/mlp_step.cu)
};

struct DtfStepOp {
  int kind;
  int is_kernel;          // counted by the launch counter (filled by the executor's caller for bookkeeping only)
  void* p0;
  void* p1;
  void* p2;
  void* p3;
  void* p4;
  long long i0, i1, i2, i3, i4, i5;
  unsigned long long u0;
};

// Runs ops[0..n) on `stream` of `device` (>= 0: made current for the call and restored afterwards).
// Returns 0, or (index + 1) * 100000 + |code| of the first failing op.  *kernels_out += kernels launched.
int dtf_run_ops(const DtfStepOp* ops, int n, int device, cudaStream_t stream, int* kernels_out) {
  int prev = -1;
  if (device >= 0) {
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device); else prev = -1;
  }
  int rc = 0, kernels = 0, i = 0;
  for (; i < n && rc == 0; ++i) {
    const DtfStepOp& o = ops[i];
    switch (o.kind) {
      case DTF_OP_H2D:
        rc = (int)cudaMemcpyAsync(o.p0, o.p1, (size_t)o.i0, cudaMemcpyHostToDevice, stream);
        break;
      case DTF_OP_D2H:
        rc = (int)cudaMemcpyAsync(o.p0, o.p1, (size_t)o.i0, cudaMemcpyDeviceToHost, stream);
        break;
      case DTF_OP_CONVERT:
        rc = dtf_convert_f32_bf16(reinterpret_cast<const float*>(o.p0), o.i0, o.p1, o.i1, o.i2, o.i3, o.i4, stream);
        ++kernels;
        break;
      case DTF_OP_GEMM:
        rc = dtf_gemm_bf16(reinterpret_cast<const DtfGemmArgs*>(o.p0), stream);
        ++kernels;
        break;
      case DTF_OP_HEAD:
        rc = dtf_mlp_head(reinterpret_cast<const DtfMlpHeadArgs*>(o.p0), stream);
        ++kernels;
        break;
      case DTF_OP_PS_APPLY:
        rc = dtf_ps_apply(reinterpret_cast<const DtfPsApplyArgs*>(o.p0), stream);
        ++kernels;
        break;
      case DTF_OP_MLP_STEP:
        rc = dtf_mlp_step(reinterpret_cast<const DtfMlpStepArgs*>(o.p0), stream);
        ++kernels;
        break;
      case DTF_OP_WAIT_TOKEN:
        rc = dtf_wait_token(o.p0, (unsigned long long)o.i0, reinterpret_cast<const unsigned long long*>(o.p1), o.u0,
                            reinterpret_cast<unsigned int*>(o.p2), stream);
        ++kernels;
        break;
      case DTF_OP_SIGNAL:
        rc = dtf_push_grad(nullptr, nullptr, 0, o.p0, o.p1, (int)o.i0, (int)o.i1, 1, 1, stream);
        ++kernels;
        break;
      case DTF_OP_STAGE:
        rc = dtf_stage_from_dataset(reinterpret_cast<const float*>(o.p0), reinterpret_cast<const float*>(o.p1), o.i0, (int)o.i1,
                                    (int)o.i2, (int)o.i3, o.i4, o.i5, reinterpret_cast<const unsigned long long*>(o.p2), o.p3,
                                    reinterpret_cast<float*>(o.p4), stream);
        ++kernels;
        break;
      case DTF_OP_SYNC:
        rc = (int)cudaStreamSynchronize(stream);
        break;
      case DTF_OP_EVENT_RECORD:
        rc = (int)cudaEventRecord(reinterpret_cast<cudaEvent_t>(o.p0), stream);
        break;
      case DTF_OP_EVENT_WAIT:
        rc = (int)cudaStreamWaitEvent(stream, reinterpret_cast<cudaEvent_t>(o.p0), 0);
        break;
      case DTF_OP_GRAPH:
        rc = (int)cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(o.p0), stream);
        kernels += (int)o.i0;       // kernels inside the graph (recorded at capture time)
        break;
      default:
        rc = -99;
    }
  }
  if (kernels_out) *kernels_out += kernels;
  if (prev >= 0) cudaSetDevice(prev);
  if (rc != 0) return i * 100000 + (rc < 0 ? -rc : rc);
  return 0;
}

// Capture ops[0..n) (kernel launches whose step-dependent inputs all come from device memory: wait targets and
// batch indices are read from device counters, never from launch arguments) into an executable CUDA graph.
// Every kernel must have been launched eagerly once before (first-launch attribute setup is not capturable).
int dtf_capture_ops(const DtfStepOp* ops, int n, int device, cudaStream_t stream, void** exec_out, int* kernels_out) {
  int prev = -1;
  if (device >= 0) {
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device); else prev = -1;
  }
  int rc = (int)cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
  int kernels = 0;
  if (rc == 0) {
    rc = dtf_run_ops(ops, n, -1, stream, &kernels);
    cudaGraph_t graph = nullptr;
    const int rc2 = (int)cudaStreamEndCapture(stream, &graph);
    if (rc == 0) rc = rc2;
    if (rc == 0) {
      cudaGraphExec_t exec = nullptr;
      rc = (int)cudaGraphInstantiate(&exec, graph, 0);
      if (rc == 0) *exec_out = exec;
    }
    if (graph) cudaGraphDestroy(graph);
  }
  if (kernels_out) *kernels_out = kernels;
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

// ---- native multi-step loop: the end-to-end inner loop without the interpreter between steps -----------------------------
// One call runs `steps` training steps of ONE local worker: per step the H2D copy of THAT step's batch from pinned host memory
// (copy stream, double-buffered: issued one step ahead so it overlaps the previous step's kernels), the step's kernels (the
// CUDA-graphed compute plan of the buffer parity), the ps shard's apply when it runs on its own stream, and the D2H copy of
// the step's loss partials into row i of a pinned host array.  The host keeps at most `depth` steps in flight (it waits for
// the loss of step i - depth to have LANDED before enqueuing step i).  Exactly what `PSTrainEngine.step(x, y,
// sync_loss="deferred", prefetch=next)` does per step, minus the Python between the calls.
// Reference role: the `while` loop around `mon_sess.run([train_step, global_step, loss], feed_dict=...)`
// (distributed_mnist.py:148-152), i.e. next_batch -> feed -> step -> fetched loss, K times.
#define DTF_LOOP_MAX_DEVICES 64

struct DtfLoopArgs {
  int device;               // made current for the call (>= 0)
  int steps;
  int depth;                // >= 1
  int parity;               // buffer set of the first step; on return: of the next step
  int prefetched;           // in: copy[parity] was already issued for the first batch.  out: prefetch_next
  int prefetch_next;        // also issue the copy of the batch FOLLOWING the last step (a continuous loop across calls)
  int x_op, y_op;           // index of the H2D ops (x, labels) inside both copy plans
  int n_copy[2], n_compute[2], n_ps;
  DtfStepOp* copy_ops[2];   // per parity: wait done[p] -> H2D x -> H2D y -> (convert) -> record ready[p]
  DtfStepOp* compute_ops[2];// per parity: wait ready[p] -> kernels / graph -> record done[p]
  DtfStepOp* ps_ops;        // optional: the ps shard's applies on its own stream
  cudaStream_t copy_stream, stream, ps_stream;
  const char* x_base;       // pinned host batches: batch b at x_base + b * x_stride (bytes)
  const char* y_base;
  long long x_stride, y_stride;
  long long nbatches, first, batch_step;   // step i trains on batch (first + i * batch_step) % nbatches
  const void* loss_src;     // device: the step's loss partials
  long long loss_bytes;
  char* loss_host;          // pinned host: row i (loss_row_bytes apart) receives step i's partials
  long long loss_row_bytes;
  long long kernels;        // out: kernels launched
  long long waited;         // out: event waits the host issued (run-ahead bound + the final one)
};

static cudaEvent_t loop_event_pool[DTF_LOOP_MAX_DEVICES][64];
static int loop_events_made[DTF_LOOP_MAX_DEVICES];

// Destroys the cached events of every device (tests; a process that wants its handles back).  Not while a loop is running.
int dtf_loop_release_events() {
  int n = 0, prev = -1;
  cudaGetDevice(&prev);
  for (int d = 0; d < DTF_LOOP_MAX_DEVICES; ++d) {
    if (loop_events_made[d] == 0) continue;
    cudaSetDevice(d);
    for (int e = 0; e < loop_events_made[d]; ++e, ++n) cudaEventDestroy(loop_event_pool[d][e]);
    loop_events_made[d] = 0;
  }
  if (prev >= 0) cudaSetDevice(prev);
  return n;
}

int dtf_run_loop(DtfLoopArgs* a) {
  if (!a || a->steps < 0 || a->depth < 1 || a->depth > 64 || a->nbatches < 1) return -1;
  if (a->x_op < 0 || a->y_op < 0 || a->x_op >= a->n_copy[0] || a->y_op >= a->n_copy[0] || a->x_op >= a->n_copy[1] ||
      a->y_op >= a->n_copy[1]) return -2;
  int prev = -1;
  if (a->device >= 0) {
    cudaGetDevice(&prev);
    if (prev != a->device) cudaSetDevice(a->device); else prev = -1;
  }
  // "loss row i has landed" events: created once per device and reused by later calls (a K = 20 step call should not pay
  // for creating and destroying them every time); a loop of one process is driven by one thread at a time
  int cur = 0;
  if (a->device < 0) cudaGetDevice(&cur); else cur = a->device;
  if (cur < 0 || cur >= DTF_LOOP_MAX_DEVICES) {
    if (prev >= 0) cudaSetDevice(prev);
    return -3;
  }
  cudaEvent_t* landed = loop_event_pool[cur];
  int rc = 0, kernels = 0;
  const int need = a->depth < a->steps ? a->depth : a->steps;
  while (loop_events_made[cur] < need && rc == 0) {
    rc = (int)cudaEventCreateWithFlags(&landed[loop_events_made[cur]], cudaEventDisableTiming);
    if (rc == 0) ++loop_events_made[cur];
  }
  auto issue_copy = [&](int par, long long step) -> int {
    long long b = (a->first + step * a->batch_step) % a->nbatches;
    if (b < 0) b += a->nbatches;
    a->copy_ops[par][a->x_op].p1 = const_cast<char*>(a->x_base) + b * a->x_stride;
    a->copy_ops[par][a->y_op].p1 = const_cast<char*>(a->y_base) + b * a->y_stride;
    return dtf_run_ops(a->copy_ops[par], a->n_copy[par], -1, a->copy_stream, &kernels);
  };
  int par = a->parity & 1;
  long long waits = 0;
  if (rc == 0 && a->steps > 0 && !a->prefetched) rc = issue_copy(par, 0);
  for (int i = 0; i < a->steps && rc == 0; ++i) {
    const int slot = i % a->depth;
    if (i >= a->depth) {                       // the loss of step i - depth has landed: its event slot is free again
      rc = (int)cudaEventSynchronize(landed[slot]);
      ++waits;
      if (rc != 0) break;
    }
    rc = dtf_run_ops(a->compute_ops[par], a->n_compute[par], -1, a->stream, &kernels);
    if (rc == 0 && a->n_ps > 0) rc = dtf_run_ops(a->ps_ops, a->n_ps, -1, a->ps_stream, &kernels);
    if (rc == 0 && (i + 1 < a->steps || a->prefetch_next))
      rc = issue_copy(par ^ 1, i + 1);                                           // next batch travels under this step's kernels
    if (rc == 0 && a->loss_bytes > 0)
      rc = (int)cudaMemcpyAsync(a->loss_host + (long long)i * a->loss_row_bytes, a->loss_src, (size_t)a->loss_bytes,
                                cudaMemcpyDeviceToHost, a->stream);
    if (rc == 0) rc = (int)cudaEventRecord(landed[slot], a->stream);
    par ^= 1;
  }
  if (rc == 0 && a->steps > 0) {               // every loss row is on the host when the call returns
    rc = (int)cudaEventSynchronize(landed[(a->steps - 1) % a->depth]);
    ++waits;
  }
  a->parity = par;
  a->prefetched = (rc == 0 && a->steps > 0 && a->prefetch_next) ? 1 : 0;
  a->kernels = kernels;
  a->waited = waits;
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

int dtf_sizeof_loop_args() { return (int)sizeof(DtfLoopArgs); }

int dtf_graph_destroy(void* exec) { return (int)cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(exec)); }

int dtf_sizeof_step_op() { return (int)sizeof(DtfStepOp); }

}  // extern "C"
