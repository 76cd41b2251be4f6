// Host emulation of the CUDA execution model, just enough to run bandwidth-style kernels (no tensor cores, no TMA) on a
// machine without a GPU: one std::thread per CUDA thread of a block, a pthread barrier for __syncthreads(), blocks executed
// one after another (so "last block done" ticket patterns see the same order a serialised grid would give), `__shared__`
// = function-local static (one block alive at a time).  What it checks: indexing, reduction trees, ticket logic, bounds.
// What it cannot check: memory-model races between concurrently running blocks, alignment faults, performance.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

// fp32 -> bf16 round-to-nearest-even (what cvt.rn.bf16x2.f32 does), NaN kept quiet
static inline unsigned dtf_emu_bf16(float f) {
  unsigned u;
  __builtin_memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
namespace dtf {
static inline unsigned pack_bf16x2(float lo, float hi) { return dtf_emu_bf16(lo) | (dtf_emu_bf16(hi) << 16); }
}  // namespace dtf

struct alignas(4) uchar4 {
  unsigned char x, y, z, w;
};
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return uchar4{a, b, c, d}; }

struct alignas(8) uint2 {
  unsigned x, y;
};
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

struct __nv_bfloat16 {
  uint16_t v;
};
static inline __nv_bfloat16 __float2bfloat16(float f) { return __nv_bfloat16{(uint16_t)dtf_emu_bf16(f)}; }

#define __align__(n) alignas(n)
#define DTF_DEVICE static inline
#define DTF_LAUNCH(kernel, grid, block, stream, ...) dtf_emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define DTF_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
  dtf_emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); }, (size_t)(smem))
#define DTF_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(dtf_emu::ctx->dyn_smem)

typedef void* cudaStream_t;
static inline int cudaGetLastError() { return 0; }
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }

namespace dtf_emu {
// Everything a running block needs lives in a per-launch context, reached through a thread_local pointer: several host
// threads can launch kernels at the same time (the multi-"GPU" protocol simulations do: one host thread per emulated
// stream).  `__shared__` variables are function statics, so two launches of the SAME kernel must not overlap unless that
// kernel declares no static shared memory.
struct LaunchCtx {
  dim3 b_dim, g_dim;
  pthread_barrier_t* barrier = nullptr;          // __syncthreads()
  pthread_barrier_t* warp_bars = nullptr;        // one barrier + one exchange row per warp (warp shuffles)
  uint64_t (*warp_slots)[32] = nullptr;
  void* dyn_smem = nullptr;                      // dynamic shared memory of the block being executed
};
inline thread_local LaunchCtx* ctx = nullptr;
inline thread_local dim3 t_idx, t_bidx;          // threadIdx / blockIdx of the calling emulated thread
inline thread_local unsigned lin_tid;            // x + y * dim.x + z * dim.x * dim.y: warps are 32 consecutive ids

static inline void sync() { pthread_barrier_wait(ctx->barrier); }

// One host thread per CUDA thread of a block, created ONCE per launch; every thread walks the grid block by block with a
// barrier between blocks, so blocks still run strictly one after another (block 0 first) without paying a thread
// creation per block.
static inline void launch(dim3 grid, dim3 block, const std::function<void()>& body, size_t smem_bytes = 0) {
  const unsigned nthreads = block.x * block.y * block.z;
  // DTF_EMU_BLOCK_ORDER=reverse walks the grid from the last block to the first: a kernel whose result depends on the order
  // in which blocks run (other than by design, like ps_apply's "block 0 decides") has a bug the hardware will find too
  static const bool reverse = [] { const char* e = std::getenv("DTF_EMU_BLOCK_ORDER"); return e && std::strcmp(e, "reverse") == 0; }();
  LaunchCtx c;
  std::vector<float4> smem_buf((smem_bytes + 15) / 16 + 1);        // 16-byte aligned like the hardware's
  c.dyn_smem = smem_buf.data();
  pthread_barrier_t bar, block_bar;
  pthread_barrier_init(&bar, nullptr, nthreads);
  pthread_barrier_init(&block_bar, nullptr, nthreads);
  c.barrier = &bar;
  const unsigned nwarps = (nthreads + 31) / 32;
  std::vector<pthread_barrier_t> wb(nwarps);
  for (unsigned wi = 0; wi < nwarps; ++wi) pthread_barrier_init(&wb[wi], nullptr, std::min(32u, nthreads - 32 * wi));
  std::vector<uint64_t> slots((size_t)nwarps * 32);
  c.warp_bars = wb.data();
  c.warp_slots = reinterpret_cast<uint64_t(*)[32]>(slots.data());
  c.g_dim = grid;
  c.b_dim = block;
  std::vector<std::thread> ts;
  ts.reserve(nthreads);
  for (unsigned tz = 0; tz < block.z; ++tz)
    for (unsigned ty = 0; ty < block.y; ++ty)
      for (unsigned tx = 0; tx < block.x; ++tx)
        ts.emplace_back([&, tx, ty, tz]() {
          ctx = &c;
          t_idx = dim3(tx, ty, tz);
          lin_tid = tx + ty * block.x + tz * block.x * block.y;
          const unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
          for (unsigned long long i = 0; i < total; ++i) {
            const unsigned long long idx = reverse ? total - 1 - i : i;      // x fastest, like the hardware's linear block id
            t_bidx = dim3((unsigned)(idx % grid.x), (unsigned)((idx / grid.x) % grid.y), (unsigned)(idx / ((unsigned long long)grid.x * grid.y)));
            body();
            pthread_barrier_wait(&block_bar);        // the whole block is done before the next one starts
          }
        });
  for (auto& t : ts) t.join();
  pthread_barrier_destroy(&bar);
  pthread_barrier_destroy(&block_bar);
  for (auto& b : wb) pthread_barrier_destroy(&b);
}
}  // namespace dtf_emu

#define threadIdx (dtf_emu::t_idx)
#define blockIdx (dtf_emu::t_bidx)
#define blockDim (dtf_emu::ctx->b_dim)
#define gridDim (dtf_emu::ctx->g_dim)
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __syncthreads() dtf_emu::sync()
#define __threadfence() std::atomic_thread_fence(std::memory_order_seq_cst)

template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) {       // cache-global load: a plain read here (aggregates have no volatile copy)
  T v;
  std::memcpy(&v, p, sizeof(T));
  return v;
}
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned int atomicExch(unsigned int* p, unsigned int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline long long clock64() { return (long long)std::chrono::steady_clock::now().time_since_epoch().count(); }
static inline void __nanosleep(unsigned ns) { std::this_thread::sleep_for(std::chrono::nanoseconds(ns)); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---- multimem emulation: a "multicast address range" is a registered placeholder range plus its member buffers -------------
namespace dtf_emu {
struct McRange {
  char* base;
  size_t bytes;
  std::vector<char*> members;
};
inline std::vector<McRange> mc_ranges;
inline std::mutex mc_mu;
static inline const McRange& mc_find(const void* p) {
  const char* c = reinterpret_cast<const char*>(p);
  for (const auto& r : mc_ranges)
    if (c >= r.base && c < r.base + r.bytes) return r;
  std::terminate();            // a multimem access outside every registered range is a test bug
}
}  // namespace dtf_emu

extern "C" __attribute__((weak)) void dtf_emu_mc_register(void* base, long long bytes, void* const* members, int n) {
  std::lock_guard<std::mutex> g(dtf_emu::mc_mu);
  dtf_emu::McRange r{reinterpret_cast<char*>(base), (size_t)bytes, {}};
  for (int i = 0; i < n; ++i) r.members.push_back(reinterpret_cast<char*>(members[i]));
  dtf_emu::mc_ranges.push_back(r);
}
extern "C" __attribute__((weak)) void dtf_emu_mc_clear() {
  std::lock_guard<std::mutex> g(dtf_emu::mc_mu);
  dtf_emu::mc_ranges.clear();
}

// ---- the scoped / ordered accesses of common.cuh as sequentially consistent host atomics -----------------------------------
namespace dtf {
static inline void griddep_launch_dependents() {}
static inline void griddep_wait() {}
static inline uint64_t globaltimer_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline uint64_t ld_relaxed_sys_u64(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline uint64_t ld_acquire_sys_u64(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void st_relaxed_sys_u64(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline void red_release_sys_add_u64(uint64_t* p, uint64_t v) { __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void fence_acq_rel_sys() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned int ld_acquire_gpu_u32(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void st_release_gpu_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline void red_relaxed_sys_add_u64(unsigned long long* p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void st_relaxed_sys_ull(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline void st_relaxed_sys_v2_u64(unsigned long long* p, unsigned long long a, unsigned long long b) {
  __atomic_store_n(p + 1, b, __ATOMIC_SEQ_CST);      // version first, token last: a reader that sees the token sees the version
  __atomic_store_n(p, a, __ATOMIC_SEQ_CST);
}
static inline bool wait_flag_ge_u64(const uint64_t* flag, uint64_t target, uint64_t timeout_ns) {
  const uint64_t t0 = globaltimer_ns();
  while (__atomic_load_n(flag, __ATOMIC_SEQ_CST) < target) {
    if (globaltimer_ns() - t0 > timeout_ns) return false;
    std::this_thread::yield();
  }
  return true;
}
static inline float4 multimem_ld_reduce_add_f32x4(const float* mc) {
  const auto& r = dtf_emu::mc_find(mc);
  const size_t off = reinterpret_cast<const char*>(mc) - r.base;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (char* m : r.members) {
    const float4 v = *reinterpret_cast<const float4*>(m + off);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  return acc;
}
template <class T>
static inline void dtf_emu_mc_store(void* mc, const T& v) {
  const auto& r = dtf_emu::mc_find(mc);
  const size_t off = reinterpret_cast<char*>(mc) - r.base;
  for (char* m : r.members) *reinterpret_cast<T*>(m + off) = v;
}
static inline void multimem_red_add_u64(unsigned long long* mc, unsigned long long v) {
  const auto& r = dtf_emu::mc_find(mc);
  const size_t off = reinterpret_cast<char*>(mc) - r.base;
  for (char* m : r.members) __atomic_fetch_add(reinterpret_cast<unsigned long long*>(m + off), v, __ATOMIC_SEQ_CST);
}
static inline void multimem_st_f32x4(float* mc, float4 v) { dtf_emu_mc_store(mc, v); }
static inline void multimem_st_b64(void* mc, uint32_t lo, uint32_t hi) { dtf_emu_mc_store(mc, uint2{lo, hi}); }
static inline void multimem_st_b128(void* mc, uint4 v) { dtf_emu_mc_store(mc, v); }
}  // namespace dtf

// ---- warp shuffles (full-mask, all lanes of the warp call it -- which is how the kernels use them) --------------------------
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  const unsigned warp = dtf_emu::lin_tid / 32, lane = dtf_emu::lin_tid % 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  dtf_emu::ctx->warp_slots[warp][lane] = bits;
  pthread_barrier_wait(&dtf_emu::ctx->warp_bars[warp]);
  const uint64_t other = dtf_emu::ctx->warp_slots[warp][lane ^ (unsigned)lane_mask];
  pthread_barrier_wait(&dtf_emu::ctx->warp_bars[warp]);
  T r;
  std::memcpy(&r, &other, sizeof(T));
  return r;
}
#define __expf(x) std::exp((float)(x))       // glibc declares __expf / __logf itself: map the CUDA fast-math names by macro
#define __logf(x) std::log((float)(x))
static inline float __bfloat162float(__nv_bfloat16 b) {
  const unsigned u = (unsigned)b.v << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline float atomicAdd(float* p, float v) {
  static std::mutex mu;                  // blocks run one after another; threads of a block contend rarely
  std::lock_guard<std::mutex> g(mu);
  const float old = *p;
  *p = old + v;
  return old;
}

static inline float __uint_as_float(unsigned u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
using std::max;
using std::min;
