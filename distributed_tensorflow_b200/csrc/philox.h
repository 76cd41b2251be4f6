// Counter-based random numbers shared by the device kernel (csrc/elementwise.cu: philox_fill_kernel), the host-emulation
// build of that kernel and the CPU tier (csrc/runtime/cpu_kernels.cpp: dtf_cpu_philox_fill): ONE definition of the stream,
// so a seeded initialiser draws the same values on a /cpu:0 parameter server and on a GPU (what TF guarantees by using
// Philox on both; reference: tf.truncated_normal / tf.random_normal / tf.zeros initialisers of distributed_mnist.py:98-105,
// example_between_graph.py:50-51 -- SURVEY A7 / K13).
//
//   * generator: Philox4x32-10 (Salmon et al., SC'11); key = the op's 64-bit stream seed, counter = (block index, stream id)
//   * element i of a fill uses word i % 4 of block offset + i / 4  -> any element is computable independently (one thread
//     per block of four), and a stateful op simply advances its offset by ceil(n / 4) per execution
//   * uniform [lo, hi): 24 mantissa bits; normal: Box-Muller on the word pairs (0,1) and (2,3); truncated normal (|z| <= 2,
//     TF's definition): inverse CDF on ONE word per element -- no rejection loop, so consumption is data-independent
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__) && !defined(DTF_HOST_EMU)
#define DTF_RNG_FN __host__ __device__ __forceinline__
#else
#define DTF_RNG_FN inline
#endif

namespace dtf_rng {

enum Kind { UNIFORM = 0, NORMAL = 1, TRUNCATED_NORMAL = 2 };

struct Block {
  uint32_t w[4];
};

DTF_RNG_FN Block philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Block b;
  b.w[0] = c0; b.w[1] = c1; b.w[2] = c2; b.w[3] = c3;
  return b;
}

// [0, 1): k * 2^-24;  (0, 1]: (k + 1) * 2^-24;  (0, 1): (k + 0.5) * 2^-24   with k = the word's top 24 bits
DTF_RNG_FN float u01_closed_open(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
DTF_RNG_FN float u01_open_closed(uint32_t x) { return (float)((x >> 8) + 1u) * 5.9604644775390625e-8f; }
DTF_RNG_FN float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f; }

// erfinv for |x| < 0.9966 (M. Giles, "Approximating the erfinv function", central branch, single precision)
DTF_RNG_FN float erfinv_central(float x) {
  float w = -logf((1.0f - x) * (1.0f + x)) - 2.5f;
  float p = 2.81022636e-08f;
  p = 3.43273939e-07f + p * w;
  p = -3.5233877e-06f + p * w;
  p = -4.39150654e-06f + p * w;
  p = 0.00021858087f + p * w;
  p = -0.00125372503f + p * w;
  p = -0.00417768164f + p * w;
  p = 0.246640727f + p * w;
  p = 1.50140941f + p * w;
  return p * x;
}

// The four values of one block.  p0/p1: (lo, hi) for UNIFORM, (mean, stddev) otherwise.
DTF_RNG_FN void block_values(const Block& b, int kind, float p0, float p1, float out[4]) {
  if (kind == UNIFORM) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = p0 + (p1 - p0) * u01_closed_open(b.w[j]);
  } else if (kind == NORMAL) {
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
      const float r = sqrtf(-2.0f * logf(u01_open_closed(b.w[j])));
      const float t = 6.283185307179586f * u01_closed_open(b.w[j + 1]);
      out[j] = p0 + p1 * (r * cosf(t));
      out[j + 1] = p0 + p1 * (r * sinf(t));
    }
  } else {
    // z = sqrt(2) * erfinv((2u - 1) * erf(sqrt(2))),  u in (0, 1)  ->  z in (-2, 2) with the normal density
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = (2.0f * u01_open(b.w[j]) - 1.0f) * 0.9544997361036416f;
      out[j] = p0 + p1 * (1.4142135623730951f * erfinv_central(s));
    }
  }
}

}  // namespace dtf_rng
