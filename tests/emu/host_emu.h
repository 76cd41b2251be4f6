// Host emulation of the CUDA execution model, just enough to run bandwidth-style kernels (no tensor cores, no TMA) on a
// machine without a GPU: one std::thread per CUDA thread of a block, a pthread barrier for __syncthreads(), blocks executed
// one after another (so "last block done" ticket patterns see the same order a serialised grid would give), `__shared__`
// = function-local static (one block alive at a time).  What it checks: indexing, reduction trees, ticket logic, bounds.
// What it cannot check: memory-model races between concurrently running blocks, alignment faults, performance.
#pragma once
#include <pthread.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <functional>
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

typedef void* cudaStream_t;
static inline int cudaGetLastError() { return 0; }

namespace dtf_emu {
inline thread_local dim3 t_idx;
inline dim3 b_idx, b_dim, g_dim;
inline pthread_barrier_t* barrier = nullptr;

static inline void sync() { pthread_barrier_wait(barrier); }

static inline void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, nthreads);
  barrier = &bar;
  g_dim = grid;
  b_dim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        b_idx = dim3(bx, by, bz);
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx)
              ts.emplace_back([&, tx, ty, tz]() {
                t_idx = dim3(tx, ty, tz);
                body();
              });
        for (auto& t : ts) t.join();
      }
  pthread_barrier_destroy(&bar);
  barrier = nullptr;
}
}  // namespace dtf_emu

#define threadIdx (dtf_emu::t_idx)
#define blockIdx (dtf_emu::b_idx)
#define blockDim (dtf_emu::b_dim)
#define gridDim (dtf_emu::g_dim)
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __syncthreads() dtf_emu::sync()
#define __threadfence() std::atomic_thread_fence(std::memory_order_seq_cst)

template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
