// Host (CPU-tier) optimizer apply kernels: one fused pass per variable, the role of TF's C++ ApplyGradientDescent /
// ApplyMomentum / ApplyAdam kernels on a /cpu:0 parameter server (the apply op runs on the variable's
// device, i.e. inside the ps task; reference distributed_mnist.py:115,126 and example_between_graph.py:61,73).
// The GPU tier's equivalents are optimizer_apply_kernel / ps_apply_kernel (csrc/elementwise.cu, csrc/ps_engine.cu).
//
// Semantics are TF-1.x's: momentum  accum = momentum*accum + g ; var -= lr*accum  (nesterov: var -= lr*g + lr*momentum*accum)
//                         adam      m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; var -= lr_t * m / (sqrt(v) + eps)
// with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller (epsilon is NOT bias-corrected: unlike torch.optim.Adam).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../philox.h"

// The loops are compiled twice (AVX2+FMA and baseline x86-64) and dispatched at load time (GCC function
// multi-versioning): the library is built on one machine and shipped to others, so -march=native is not an option.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define DTF_CPU_CLONES __attribute__((target_clones("avx2,fma", "default")))
#else
#define DTF_CPU_CLONES
#endif

// Denormals: a weight whose gradient is exactly zero step after step (MNIST's border pixels) has Adam / momentum slots that decay
// geometrically INTO the denormal range (0.9^t * m0 < 1.2e-38 after ~800 steps) and stay there for a hundred steps and more; x86
// handles every denormal operand with a micro-code assist -- measured here: 40 us -> 1.2 ms per apply of the 784x100 matrix.
// TensorFlow's kernels run with flush-to-zero + denormals-are-zero set in its worker threads (ScopedFlushDenormal); so do
// these loops, for their own duration (the caller's MXCSR is restored).
#if defined(__x86_64__)
#include <xmmintrin.h>
struct ScopedFlushDenormals {
  unsigned int saved;
  ScopedFlushDenormals() : saved(_mm_getcsr()) { _mm_setcsr(saved | 0x8040u); }     // FTZ (bit 15) | DAZ (bit 6)
  ~ScopedFlushDenormals() { _mm_setcsr(saved); }
};
#else
struct ScopedFlushDenormals {};
#endif

extern "C" {

// kind: 0 sgd, 1 momentum, 2 adam.  All buffers fp32, contiguous, n elements, NOT overlapping.  Returns 0, or -1 on bad
// arguments.
DTF_CPU_CLONES
int dtf_cpu_optimizer_apply(int kind, float* __restrict__ var, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ g, long long n, float lr, float momentum, int nesterov, float beta1,
                            float beta2, float eps) {
  if (n < 0 || var == nullptr || g == nullptr) return -1;
  ScopedFlushDenormals no_denormals;
  if (kind == 0) {
    for (long long i = 0; i < n; ++i) var[i] -= lr * g[i];
    return 0;
  }
  if (kind == 1) {
    if (m == nullptr) return -1;
    if (nesterov) {
      const float lm = lr * momentum;
      for (long long i = 0; i < n; ++i) {
        const float acc = momentum * m[i] + g[i];
        m[i] = acc;
        var[i] -= lr * g[i] + lm * acc;
      }
    } else {
      for (long long i = 0; i < n; ++i) {
        const float acc = momentum * m[i] + g[i];
        m[i] = acc;
        var[i] -= lr * acc;
      }
    }
    return 0;
  }
  if (kind == 2) {
    if (m == nullptr || v == nullptr) return -1;
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    for (long long i = 0; i < n; ++i) {
      const float gi = g[i];
      const float mi = beta1 * m[i] + omb1 * gi;
      const float vi = beta2 * v[i] + omb2 * gi * gi;
      m[i] = mi;
      v[i] = vi;
      var[i] -= lr * mi / (std::sqrt(vi) + eps);
    }
    return 0;
  }
  return -1;
}

// Random fill of the CPU tier (initialisers of variables living on a /cpu:0 parameter server): the stream defined in
// csrc/philox.h, i.e. the values the GPU kernel philox_fill_kernel (csrc/elementwise.cu) draws for the same (key, offset).
// kind: 0 uniform [p0, p1), 1 normal(mean p0, stddev p1), 2 truncated normal (|z| <= 2).  Returns 0, or -1 on bad arguments.
int dtf_cpu_philox_fill(float* out, long long n, unsigned long long key, unsigned long long offset, unsigned long long stream_id,
                        int kind, float p0, float p1) {
  if (n < 0 || kind < 0 || kind > 2 || (n > 0 && out == nullptr)) return -1;
  const long long nblk = (n + 3) / 4;
  for (long long b = 0; b < nblk; ++b) {
    const dtf_rng::Block blk = dtf_rng::philox4x32_10(offset + (unsigned long long)b, stream_id, key);
    float v[4];
    dtf_rng::block_values(blk, kind, p0, p1, v);
    for (int j = 0; j < 4 && b * 4 + j < n; ++j) out[b * 4 + j] = v[j];
  }
  return 0;
}

// The raw 32-bit words of `nblk` blocks (known-answer tests of the generator itself).
int dtf_cpu_philox_words(unsigned int* out, long long nblk, unsigned long long key, unsigned long long ctr_lo, unsigned long long ctr_hi) {
  if (nblk < 0 || (nblk > 0 && out == nullptr)) return -1;
  for (long long b = 0; b < nblk; ++b) {
    const dtf_rng::Block blk = dtf_rng::philox4x32_10(ctr_lo + (unsigned long long)b, ctr_hi, key);
    for (int j = 0; j < 4; ++j) out[b * 4 + j] = blk.w[j];
  }
  return 0;
}

// Input pipeline: dst row i = src row idx[i] (rows of row_bytes bytes), split over up to `threads` host threads -- the gather
// behind a shuffled epoch of batches laid out [batch, row] in (pinned) host memory, i.e. the role of `mnist.train.next_batch`
// (distributed_mnist.py:149) done once per epoch, off the training thread and without the GIL.
// Returns 0, -1 on bad arguments, -2 when an index is outside [0, n_src).
int dtf_gather_rows(const void* src, long long n_src, const long long* idx, long long n, long long row_bytes, void* dst, int threads) {
  if (n < 0 || n_src < 0 || row_bytes < 0 || (n > 0 && (src == nullptr || dst == nullptr || idx == nullptr))) return -1;
  for (long long i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= n_src) return -2;
  const char* s = static_cast<const char*>(src);
  char* d = static_cast<char*>(dst);
  auto work = [&](long long lo, long long hi) {
    for (long long i = lo; i < hi; ++i) std::memcpy(d + i * row_bytes, s + idx[i] * row_bytes, (size_t)row_bytes);
  };
  long long t = threads < 1 ? 1 : threads;
  if (n * row_bytes < (1ll << 20)) t = 1;                 // small gathers are not worth a thread start
  if (t > n) t = n > 0 ? n : 1;
  if (t == 1) {
    work(0, n);
    return 0;
  }
  std::vector<std::thread> pool;
  const long long per = (n + t - 1) / t;
  for (long long k = 0; k < t; ++k) {
    const long long lo = k * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return 0;
}

}  // extern "C"
