// Recording stand-ins for the CUDA runtime calls and the kernel launchers csrc/step_exec.cu uses, so that the native step
// executor's op dispatch (argument mapping of every op kind, kernel counting, error index, graph capture sequence) is tested
// by g++ builds without a GPU.  Every call appends one line to a trace that the test reads back.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };

namespace step_emu {
inline std::string trace;
inline int current_device = 0;
inline int fail_kind = 0, fail_code = 0;      // make the launcher / call named by fail_kind return fail_code
template <class... A>
static inline void log(const char* fmt, A... a) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), fmt, a...);
  trace += buf;
  trace += "\n";
}
}  // namespace step_emu

static inline cudaError_t cudaGetDevice(int* d) { *d = step_emu::current_device; return 0; }
static inline cudaError_t cudaSetDevice(int d) { step_emu::log("setdevice %d", d); step_emu::current_device = d; return 0; }
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind k, cudaStream_t s) {
  step_emu::log("memcpy %s dst=%p src=%p n=%zu stream=%p", k == cudaMemcpyHostToDevice ? "h2d" : "d2h", dst, src, n, s);
  return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { step_emu::log("sync stream=%p", s); return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { step_emu::log("event_record ev=%p stream=%p", e, s); return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) { step_emu::log("event_wait ev=%p stream=%p", e, s); return 0; }
enum { cudaEventDisableTiming = 2 };
namespace step_emu { inline long long next_event = 0xe000; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags) {
  if (step_emu::fail_kind == 100) return step_emu::fail_code;
  *e = (void*)(step_emu::next_event++);
  step_emu::log("event_create ev=%p flags=%u", *e, flags);
  return 0;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) { step_emu::log("event_sync ev=%p", e); return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { step_emu::log("event_destroy ev=%p", e); return 0; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t s) { step_emu::log("graph_launch exec=%p stream=%p", g, s); return 0; }
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) { step_emu::log("begin_capture stream=%p", s); return 0; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g) { step_emu::log("end_capture stream=%p", s); *g = (void*)0x6a; return 0; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { step_emu::log("instantiate graph=%p", g); *e = (void*)0xe1; return 0; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { step_emu::log("graph_destroy %p", g); return 0; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { step_emu::log("exec_destroy %p", e); return 0; }

extern "C" {
struct DtfGemmArgs;
struct DtfMlpHeadArgs;
struct DtfPsApplyArgs;
struct DtfMlpStepArgs;
__attribute__((weak)) int dtf_gemm_bf16(const DtfGemmArgs* g, cudaStream_t s) { step_emu::log("gemm args=%p stream=%p", (const void*)g, s); return step_emu::fail_kind == 4 ? step_emu::fail_code : 0; }
__attribute__((weak)) int dtf_mlp_head(const DtfMlpHeadArgs* a, cudaStream_t s) { step_emu::log("head args=%p stream=%p", (const void*)a, s); return 0; }
__attribute__((weak)) int dtf_ps_apply(const DtfPsApplyArgs* a, cudaStream_t s) { step_emu::log("ps_apply args=%p stream=%p", (const void*)a, s); return 0; }
__attribute__((weak)) int dtf_mlp_step(const DtfMlpStepArgs* a, cudaStream_t s) { step_emu::log("mlp_step args=%p stream=%p", (const void*)a, s); return 0; }
__attribute__((weak)) int dtf_convert_f32_bf16(const float* in, long long ld_in, void* out, long long ld_out, long long rows, long long cols,
                                               long long cols_pad, cudaStream_t s) {
  step_emu::log("convert in=%p ld_in=%lld out=%p ld_out=%lld rows=%lld cols=%lld pad=%lld", (const void*)in, ld_in, out, ld_out, rows, cols, cols_pad);
  return 0;
}
__attribute__((weak)) int dtf_wait_token(const void* mailbox, unsigned long long target, const unsigned long long* target_ptr,
                                         unsigned long long timeout_ns, unsigned int* err, cudaStream_t s) {
  step_emu::log("wait_token mb=%p target=%llu ptr=%p timeout=%llu err=%p", mailbox, target, (const void*)target_ptr, timeout_ns, (void*)err);
  return 0;
}
__attribute__((weak)) int dtf_push_grad(const float* src, float* dst_peer, long long n, void* ctl, const void* mailbox, int rank,
                                        int stamp_from_version, int write_stamp, int grid, cudaStream_t s) {
  step_emu::log("push_grad src=%p dst=%p n=%lld ctl=%p mb=%p rank=%d from_version=%d write_stamp=%d grid=%d", (const void*)src, (void*)dst_peer, n,
                ctl, mailbox, rank, stamp_from_version, write_stamp, grid);
  return 0;
}
__attribute__((weak)) int dtf_stage_from_dataset(const float* images, const float* labels, long long nbatches, int B, int D, int C,
                                                 long long stride, long long offset, const unsigned long long* step_counter, void* x16,
                                                 float* lab_out, cudaStream_t s) {
  step_emu::log("stage images=%p labels=%p nb=%lld B=%d D=%d C=%d stride=%lld offset=%lld ctr=%p x16=%p lab=%p", (const void*)images,
                (const void*)labels, nbatches, B, D, C, stride, offset, (const void*)step_counter, x16, (void*)lab_out);
  return 0;
}
__attribute__((weak)) const char* step_emu_trace() { return step_emu::trace.c_str(); }
__attribute__((weak)) void step_emu_reset(int fail_kind, int fail_code) {
  step_emu::trace.clear();
  step_emu::fail_kind = fail_kind;
  step_emu::fail_code = fail_code;
  step_emu::current_device = 0;
  step_emu::next_event = 0xe000;
}
}
