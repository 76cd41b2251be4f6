// Non-GEMM sm_100a kernels: conversion/staging, fused softmax-cross-entropy fwd+bwd, ReLU backward,
// column sums, optimizer applies (SGD / Momentum / TF-Adam), im2col / col2im for NHWC convolution,
// and a plain CUDA-core reference GEMM used by tests and for shapes TMA cannot address.
//
// These are bandwidth- or latency-bound; they use 128-bit vector accesses where alignment allows and
// warp-shuffle reductions.  SURVEY kernel sites: K2/K3 (xent), K5/K6 (apply), K8 (argmax),
// K9 (element-wise), K12 (tower mean), K14 (conv lowering).
#include <cstdio>
#include <cstring>

// DTF_HOST_EMU: this file also compiles with g++ against tests/emu/host_emu.h, so the CPU test tier runs these kernels
// from the same source (docs/TESTING.md, "Host emulation of the CUDA block model").
#ifdef DTF_HOST_EMU
#include "host_emu.h"
#else
#include "common.cuh"
#endif
#include "philox.h"

namespace dtf {

DTF_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DTF_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// fp32 [rows, cols] (ld_in) -> bf16 [rows, cols] (ld_out), optionally zero-filling pad columns
// ---------------------------------------------------------------------------------------------
__global__ void convert_f32_bf16_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                                        long long ld_out, long long rows, long long cols, long long cols_pad) {
  const long long total = rows * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad, c = i - r * cols_pad;
    out[r * ld_out + c] = __float2bfloat16(c < cols ? in[r * ld_in + c] : 0.0f);
  }
}

__global__ void convert_f32_bf16_vec_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

__global__ void convert_u8_bf16_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n,
                                       float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16((float)in[i] * scale);
}

// ---------------------------------------------------------------------------------------------
// K13: random initialisers.  One thread per Philox block of four values (csrc/philox.h defines the stream; the CPU tier
// draws the same values from the same definition).  out[i], i in [0, n): word i % 4 of block `offset + i / 4`.
// ---------------------------------------------------------------------------------------------
__global__ void philox_fill_kernel(float* __restrict__ out, long long n, unsigned long long key, unsigned long long offset,
                                   unsigned long long stream_id, int kind, float p0, float p1) {
  const long long nblk = (n + 3) / 4;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (long long)gridDim.x * blockDim.x) {
    const dtf_rng::Block blk = dtf_rng::philox4x32_10(offset + (unsigned long long)b, stream_id, key);
    float v[4];
    dtf_rng::block_values(blk, kind, p0, p1, v);
    const long long i = b * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i + j < n) out[i + j] = v[j];
  }
}

// ---------------------------------------------------------------------------------------------
// K9: element-wise graph ops (the linear-regression programs: example_between_graph.py:55-60 `y = weight * x + biase`,
// `tf.square(y_ - y)`, tf.reduce_mean; the in-graph example's matmul results added up).  fp32, contiguous.
//   binary: out[i] = f(a[ia], b[ib]);  an operand is indexed by i (full), 0 (scalar) or i % inner (a trailing-dims vector
//   broadcast over the leading dims) -- the broadcasts these programs use.
// ---------------------------------------------------------------------------------------------
enum { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_DIV = 3, EW_MAX = 4, EW_MIN = 5, EW_SQDIFF = 6 };
enum { EW_NEG = 0, EW_SQUARE = 1, EW_SQRT = 2, EW_RSQRT = 3, EW_EXP = 4, EW_LOG = 5, EW_ABS = 6, EW_SIGMOID = 7, EW_TANH = 8,
       EW_RELU = 9 };
enum { EW_FULL = 0, EW_SCALAR = 1, EW_INNER = 2 };

template <int OP>
DTF_DEVICE float ew_binary(float x, float y) {
  if (OP == EW_ADD) return x + y;
  if (OP == EW_SUB) return x - y;
  if (OP == EW_MUL) return x * y;
  if (OP == EW_DIV) return x / y;
  if (OP == EW_MAX) return fmaxf(x, y);
  if (OP == EW_MIN) return fminf(x, y);
  const float d = x - y;
  return d * d;
}

template <int OP>
DTF_DEVICE float ew_unary(float x) {
  if (OP == EW_NEG) return -x;
  if (OP == EW_SQUARE) return x * x;
  if (OP == EW_SQRT) return sqrtf(x);
  if (OP == EW_RSQRT) return 1.0f / sqrtf(x);
  if (OP == EW_EXP) return expf(x);
  if (OP == EW_LOG) return logf(x);
  if (OP == EW_ABS) return fabsf(x);
  if (OP == EW_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (OP == EW_TANH) return tanhf(x);
  return x > 0.0f ? x : 0.0f;
}

// The op is a template parameter (one instantiation per op: no switch in the loop); indices are 32-bit (the launcher splits
// larger tensors), so the trailing-vector broadcast costs one 32-bit remainder per element.
template <int OP>
__global__ void ew_binary_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, unsigned int n,
                                 int mode_a, int mode_b, unsigned int inner) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int r = (mode_a == EW_INNER || mode_b == EW_INNER) ? i % inner : 0u;
    const float x = a[mode_a == EW_FULL ? i : (mode_a == EW_SCALAR ? 0u : r)];
    const float y = b[mode_b == EW_FULL ? i : (mode_b == EW_SCALAR ? 0u : r)];
    out[i] = ew_binary<OP>(x, y);
  }
}

template <int OP>
__global__ void ew_unary_kernel(const float* __restrict__ x, float* __restrict__ out, unsigned int n) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = ew_unary<OP>(x[i]);
}

// out[i] = alpha * x[i or 0] + beta: scalings, negation, and the broadcast of a scalar gradient (the backward of a full reduction)
__global__ void ew_affine_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int mode, float alpha, float beta) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = alpha * x[mode == EW_FULL ? i : 0] + beta;
}

// out[0] += scale * sum(x) (out zeroed by the launcher): warp shuffles -> one shared-memory row per block -> one atomic per
// block.  `square`: sum of squares (the mean-squared-error reduction in one pass).
__global__ void ew_reduce_sum_kernel(const float* __restrict__ x, long long n, float scale, int square, float* __restrict__ out) {
  __shared__ float part[8];
  float acc = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc += square ? v * v : v;
  }
  acc = warp_sum(acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(out, v * scale);
  }
}

// element-wise binary / unary kernels above walk fp32; everything else stays with the op layer

// ---------------------------------------------------------------------------------------------
// K10: split / concat as ONE gather-scatter kernel over up to 16 parts.  Part p is `outer` rows of `row_len[p]` contiguous
// floats; source and destination rows are `*_stride[p]` apart and start at `*_off[p]`.  concat along an axis = every part
// reads its own tensor and writes a column band of the result; the scatter of a global batch = every part reads a row band
// and writes its own destination -- which may live on a PEER GPU (in-graph replication: example_in_graph.py:38 splits the
// batch across workers, :58 concatenates their results; with peer access enabled the stores / loads travel over NVLink).
// ---------------------------------------------------------------------------------------------
struct CopyParts {
  const float* src[16];
  float* dst[16];
  long long src_stride[16], dst_stride[16], src_off[16], dst_off[16], row_len[16];
  long long cum[17];            // cum[p] = elements of parts 0..p-1 (outer * row_len each)
  long long outer;
  int nparts;
};

__global__ void copy_parts_kernel(const CopyParts t) {
  const long long total = t.cum[t.nparts];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int p = 0;
    while (p + 1 < t.nparts && i >= t.cum[p + 1]) ++p;
    const long long j = i - t.cum[p];
    const long long o = j / t.row_len[p], c = j - o * t.row_len[p];
    t.dst[p][o * t.dst_stride[p] + t.dst_off[p] + c] = t.src[p][o * t.src_stride[p] + t.src_off[p] + c];
  }
}

// ---------------------------------------------------------------------------------------------
// fused softmax + cross-entropy, forward and backward in one pass (one warp per row)
//   y = softmax(logits);  t_j = labels_j * [y_j >= clip_min]
//   loss_row = -sum_j labels_j * log(clamp(y_j, clip_min, 1))          (clip_min = 0: -sum labels*log_softmax)
//   dlogits_i = -t_i + y_i * sum_j t_j                                 (== y - labels when nothing clips)
// loss is either accumulated into one scalar (batch SUM, reference distributed_mnist.py:113) or per row.
// ---------------------------------------------------------------------------------------------
__global__ void softmax_xent_kernel(const float* __restrict__ logits, long long ld_logits,
                                    const float* __restrict__ labels, long long ld_labels, int rows, int cols,
                                    float clip_min, float* __restrict__ loss_sum, float* __restrict__ loss_rows,
                                    float* __restrict__ dlogits, long long ld_d, __nv_bfloat16* __restrict__ dlogits_bf16,
                                    long long ld_db, int cols_pad_bf16, float* __restrict__ probs, long long ld_p,
                                    float grad_scale) {
  const int warps_per_block = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * warps_per_block + warp;
  if (row >= rows) return;
  const float* z = logits + (long long)row * ld_logits;
  const float* lab = labels + (long long)row * ld_labels;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < cols; c += 32) se += __expf(z[c] - mx);
  se = warp_sum(se);
  const float inv = 1.0f / se;
  const float lse = mx + __logf(se);
  float loss = 0.f, tsum = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float y = __expf(z[c] - mx) * inv;
    const float l = lab[c];
    if (clip_min > 0.f) {
      loss -= l * __logf(fminf(fmaxf(y, clip_min), 1.0f));
      tsum += (y >= clip_min) ? l : 0.f;
    } else {
      loss -= l * (z[c] - lse);
      tsum += l;
    }
  }
  loss = warp_sum(loss);
  tsum = warp_sum(tsum);
  for (int c = lane; c < cols; c += 32) {
    const float y = __expf(z[c] - mx) * inv;
    const float l = lab[c];
    const float t = (clip_min > 0.f && y < clip_min) ? 0.f : l;
    const float g = (y * tsum - t) * grad_scale;
    if (dlogits) dlogits[(long long)row * ld_d + c] = g;
    if (dlogits_bf16) dlogits_bf16[(long long)row * ld_db + c] = __float2bfloat16(g);
    if (probs) probs[(long long)row * ld_p + c] = y;
  }
  if (dlogits_bf16)
    for (int c = cols + lane; c < cols_pad_bf16; c += 32) dlogits_bf16[(long long)row * ld_db + c] = __float2bfloat16(0.f);
  if (lane == 0) {
    if (loss_rows) loss_rows[row] = loss;
    if (loss_sum) atomicAdd(loss_sum, loss);
  }
}

// ---------------------------------------------------------------------------------------------
// relu backward, column sums, argmax, tower mean
// ---------------------------------------------------------------------------------------------
__global__ void relu_grad_kernel(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ out,
                                 long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = y[i] > 0.f ? g[i] : 0.f;
}

// out[c] = sum_r in[r, c]; one block per 32 columns, 8 warps stride the rows
__global__ void colsum_kernel(const float* __restrict__ in, long long ld, int rows, int cols, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = w; r < rows; r += 8) s += in[(long long)r * ld + c];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    out[c] = t;
  }
}

__global__ void argmax_rows_kernel(const float* __restrict__ in, long long ld, int rows, int cols, long long* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < cols; c += 32) {
    const float v = in[(long long)row * ld + c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

// out = mean over `n_in` same-shaped inputs (tower gradient averaging)
__global__ void mean_of_n_kernel(const float* const* __restrict__ ins, int n_in, float* __restrict__ out, long long n) {
  const float inv = 1.0f / (float)n_in;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < n_in; ++k) s += ins[k][i];
    out[i] = s * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// optimizer applies on flat fp32 buffers (TF formulations), optional bf16 shadow copy of the params
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
  float* var;
  float* m;        // momentum accum / Adam m
  float* v;        // Adam v
  const float* g;
  __nv_bfloat16* shadow;   // optional bf16 copy of var
  long long n;
  int kind;        // 0 sgd, 1 momentum, 2 adam
  float lr;        // sgd/momentum: lr ; adam: lr_t (bias-corrected)
  float momentum;
  int nesterov;
  float beta1, beta2, eps;
  float grad_scale;
};

DTF_DEVICE float apply_one(const ApplyArgs& a, long long i, float g) {
  float w = a.var[i];
  if (a.kind == 0) {
    w -= a.lr * g;
  } else if (a.kind == 1) {
    const float acc = a.momentum * a.m[i] + g;
    a.m[i] = acc;
    w -= a.nesterov ? (a.lr * g + a.lr * a.momentum * acc) : (a.lr * acc);
  } else {
    const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
    a.m[i] = m;
    a.v[i] = v;
    w -= a.lr * m / (sqrtf(v) + a.eps);          // epsilon OUTSIDE the bias correction (TF Adam)
  }
  a.var[i] = w;
  return w;
}

__global__ void optimizer_apply_kernel(const ApplyArgs a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const float w = apply_one(a, i, a.g[i] * a.grad_scale);
    if (a.shadow) a.shadow[i] = __float2bfloat16(w);
  }
}

// ---------------------------------------------------------------------------------------------
// NHWC im2col (bf16 out, K padded to a multiple of 8) and col2im (fp32 accumulate) for conv lowering
// ---------------------------------------------------------------------------------------------
__global__ void im2col_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols, int n, int h, int w,
                                   int c, int kh, int kw, int sh, int sw, int pt, int pl, int ho, int wo, long long ldc) {
  const long long kdim = (long long)kh * kw * c;
  const long long total = (long long)n * ho * wo * ldc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / ldc;
    const long long kk = i - row * ldc;
    float v = 0.f;
    if (kk < kdim) {
      const int ci = (int)(kk % c);
      const int kx = (int)((kk / c) % kw);
      const int ky = (int)(kk / ((long long)c * kw));
      const int ox = (int)(row % wo);
      const int oy = (int)((row / wo) % ho);
      const int b = (int)(row / ((long long)wo * ho));
      const int iy = oy * sh - pt + ky, ix = ox * sw - pl + kx;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = x[(((long long)b * h + iy) * w + ix) * c + ci];
    }
    cols[i] = __float2bfloat16(v);
  }
}

__global__ void col2im_nhwc_kernel(const float* __restrict__ gcols, long long ldg, float* __restrict__ gx, int n, int h,
                                   int w, int c, int kh, int kw, int sh, int sw, int pt, int pl, int ho, int wo) {
  // gather form: one thread per input element sums the contributions of every patch covering it
  const long long total = (long long)n * h * w * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % c);
    const int ix = (int)((i / c) % w);
    const int iy = (int)((i / ((long long)c * w)) % h);
    const int b = (int)(i / ((long long)c * w * h));
    float s = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int ty = iy + pt - ky;
      if (ty < 0 || ty % sh) continue;
      const int oy = ty / sh;
      if (oy >= ho) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int tx = ix + pl - kx;
        if (tx < 0 || tx % sw) continue;
        const int ox = tx / sw;
        if (ox >= wo) continue;
        const long long row = ((long long)b * ho + oy) * wo + ox;
        s += gcols[row * ldg + ((long long)ky * kw + kx) * c + ci];
      }
    }
    gx[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// reference GEMM on CUDA cores (fp32 accumulate over bf16-rounded inputs); tests + odd shapes
// ---------------------------------------------------------------------------------------------
__global__ void gemm_ref_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                float* __restrict__ c, int M, int N, int K, long long lda, long long ldb, long long ldc,
                                int a_mn, int b_mn, const float* __restrict__ bias, int relu, float alpha) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float av = __bfloat162float(a_mn ? a[(long long)k * lda + m] : a[(long long)m * lda + k]);
    const float bv = __bfloat162float(b_mn ? b[(long long)k * ldb + n] : b[(long long)n * ldb + k]);
    acc += av * bv;
  }
  acc *= alpha;
  if (bias) acc += bias[n];
  if (relu) acc = fmaxf(acc, 0.f);
  c[(long long)m * ldc + n] = acc;
}

static inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

// Chunks of at most 2^30 elements per launch keep the kernels' index arithmetic in 32 bits; a chunk boundary is a multiple of
// `inner`, so the trailing-vector broadcast stays aligned.
template <int OP>
static int ew_binary_launch(const float* a, const float* b, float* out, long long n, int mode_a, int mode_b, long long inner,
                            cudaStream_t s) {
  long long chunk = 1ll << 30;
  if (mode_a == EW_INNER || mode_b == EW_INNER) {
    if (inner >= (1ll << 31)) return -1;
    chunk = inner > chunk ? inner : chunk / inner * inner;
  }
  for (long long off = 0; off < n; off += chunk) {
    const long long m = n - off < chunk ? n - off : chunk;
    DTF_LAUNCH(ew_binary_kernel<OP>, grid_for(m), 256, s, a + (mode_a == EW_FULL ? off : 0), b + (mode_b == EW_FULL ? off : 0), out + off,
               (unsigned int)m, mode_a, mode_b, (unsigned int)(inner > 0 ? inner : 1));
  }
  return (int)cudaGetLastError();
}

template <int OP>
static int ew_unary_launch(const float* x, float* out, long long n, cudaStream_t s) {
  const long long chunk = 1ll << 30;
  for (long long off = 0; off < n; off += chunk) {
    const long long m = n - off < chunk ? n - off : chunk;
    DTF_LAUNCH(ew_unary_kernel<OP>, grid_for(m), 256, s, x + off, out + off, (unsigned int)m);
  }
  return (int)cudaGetLastError();
}

}  // namespace dtf

extern "C" {
using namespace dtf;

int dtf_convert_f32_bf16(const float* in, long long ld_in, void* out, long long ld_out, long long rows, long long cols,
                         long long cols_pad, cudaStream_t s) {
  if (ld_in == cols && ld_out == cols && cols_pad == cols && (rows * cols) % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
    const long long n4 = rows * cols / 4;
    DTF_LAUNCH(convert_f32_bf16_vec_kernel, grid_for(n4), 256, s, reinterpret_cast<const float4*>(in),
                                                             reinterpret_cast<uint2*>(out), n4);
  } else {
    DTF_LAUNCH(convert_f32_bf16_kernel, grid_for(rows * cols_pad), 256, s, in, ld_in, reinterpret_cast<__nv_bfloat16*>(out),
                                                                      ld_out, rows, cols, cols_pad);
  }
  return (int)cudaGetLastError();
}

int dtf_convert_u8_bf16(const void* in, void* out, long long n, float scale, cudaStream_t s) {
  DTF_LAUNCH(convert_u8_bf16_kernel, grid_for(n), 256, s, reinterpret_cast<const uint8_t*>(in),
                                                     reinterpret_cast<__nv_bfloat16*>(out), n, scale);
  return (int)cudaGetLastError();
}

int dtf_softmax_xent(const float* logits, long long ld_logits, const float* labels, long long ld_labels, int rows, int cols,
                     float clip_min, float* loss_sum, float* loss_rows, float* dlogits, long long ld_d,
                     void* dlogits_bf16, long long ld_db, int cols_pad_bf16, float* probs, long long ld_p, float grad_scale,
                     cudaStream_t s) {
  const int wpb = 4;
  DTF_LAUNCH(softmax_xent_kernel, (rows + wpb - 1) / wpb, wpb * 32, s,
      logits, ld_logits, labels, ld_labels, rows, cols, clip_min, loss_sum, loss_rows, dlogits, ld_d,
      reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ld_db, cols_pad_bf16, probs, ld_p, grad_scale);
  return (int)cudaGetLastError();
}

int dtf_philox_fill(float* out, long long n, unsigned long long key, unsigned long long offset, unsigned long long stream_id,
                    int kind, float p0, float p1, cudaStream_t s) {
  if (n < 0 || kind < 0 || kind > 2 || (n > 0 && out == nullptr)) return -1;
  if (n == 0) return 0;
  DTF_LAUNCH(philox_fill_kernel, grid_for((n + 3) / 4), 256, s, out, n, key, offset, stream_id, kind, p0, p1);
  return (int)cudaGetLastError();
}

int dtf_ew_binary(const float* a, const float* b, float* out, long long n, int op, int mode_a, int mode_b, long long inner,
                  cudaStream_t s) {
  if (n < 0 || op < 0 || op > 6 || mode_a < 0 || mode_a > 2 || mode_b < 0 || mode_b > 2) return -1;
  if ((mode_a == 2 || mode_b == 2) && inner < 1) return -1;
  if (n == 0) return 0;
  switch (op) {
    case EW_ADD: return ew_binary_launch<EW_ADD>(a, b, out, n, mode_a, mode_b, inner, s);
    case EW_SUB: return ew_binary_launch<EW_SUB>(a, b, out, n, mode_a, mode_b, inner, s);
    case EW_MUL: return ew_binary_launch<EW_MUL>(a, b, out, n, mode_a, mode_b, inner, s);
    case EW_DIV: return ew_binary_launch<EW_DIV>(a, b, out, n, mode_a, mode_b, inner, s);
    case EW_MAX: return ew_binary_launch<EW_MAX>(a, b, out, n, mode_a, mode_b, inner, s);
    case EW_MIN: return ew_binary_launch<EW_MIN>(a, b, out, n, mode_a, mode_b, inner, s);
    default: return ew_binary_launch<EW_SQDIFF>(a, b, out, n, mode_a, mode_b, inner, s);
  }
}

int dtf_ew_unary(const float* x, float* out, long long n, int op, cudaStream_t s) {
  if (n < 0 || op < 0 || op > 9) return -1;
  if (n == 0) return 0;
  switch (op) {
    case EW_NEG: return ew_unary_launch<EW_NEG>(x, out, n, s);
    case EW_SQUARE: return ew_unary_launch<EW_SQUARE>(x, out, n, s);
    case EW_SQRT: return ew_unary_launch<EW_SQRT>(x, out, n, s);
    case EW_RSQRT: return ew_unary_launch<EW_RSQRT>(x, out, n, s);
    case EW_EXP: return ew_unary_launch<EW_EXP>(x, out, n, s);
    case EW_LOG: return ew_unary_launch<EW_LOG>(x, out, n, s);
    case EW_ABS: return ew_unary_launch<EW_ABS>(x, out, n, s);
    case EW_SIGMOID: return ew_unary_launch<EW_SIGMOID>(x, out, n, s);
    case EW_TANH: return ew_unary_launch<EW_TANH>(x, out, n, s);
    default: return ew_unary_launch<EW_RELU>(x, out, n, s);
  }
}

int dtf_ew_affine(const float* x, float* out, long long n, int mode, float alpha, float beta, cudaStream_t s) {
  if (n < 0 || mode < 0 || mode > 1) return -1;
  if (n == 0) return 0;
  DTF_LAUNCH(ew_affine_kernel, grid_for(n), 256, s, x, out, n, mode, alpha, beta);
  return (int)cudaGetLastError();
}

// out[0] = scale * sum(x) or scale * sum(x^2).  `out` must not alias x.
int dtf_ew_reduce_sum(const float* x, long long n, float scale, int square, float* out, cudaStream_t s) {
  if (n < 0 || out == nullptr) return -1;
  int rc = (int)cudaMemsetAsync(out, 0, sizeof(float), s);
  if (rc != 0 || n == 0) return rc;
  long long blocks = (n + 256 * 8 - 1) / (256 * 8);           // >= 8 elements per thread before another block is worth an atomic
  if (blocks > 592) blocks = 592;
  DTF_LAUNCH(ew_reduce_sum_kernel, (int)blocks, 256, s, x, n, scale, square, out);
  return (int)cudaGetLastError();
}

// parts: nparts <= 16; arrays of nparts entries (host memory).  See copy_parts_kernel.
int dtf_copy_parts(int nparts, long long outer, const void* const* src, void* const* dst, const long long* src_stride,
                   const long long* dst_stride, const long long* src_off, const long long* dst_off, const long long* row_len,
                   cudaStream_t s) {
  if (nparts < 1 || nparts > 16 || outer < 0) return -1;
  CopyParts t;
  memset(&t, 0, sizeof(t));
  t.nparts = nparts;
  t.outer = outer;
  t.cum[0] = 0;
  for (int p = 0; p < nparts; ++p) {
    if (row_len[p] < 0 || src_stride[p] < 0 || dst_stride[p] < 0 || src_off[p] < 0 || dst_off[p] < 0) return -1;
    t.src[p] = reinterpret_cast<const float*>(src[p]);
    t.dst[p] = reinterpret_cast<float*>(dst[p]);
    t.src_stride[p] = src_stride[p];
    t.dst_stride[p] = dst_stride[p];
    t.src_off[p] = src_off[p];
    t.dst_off[p] = dst_off[p];
    t.row_len[p] = row_len[p] > 0 ? row_len[p] : 1;          // an empty part contributes no element; keep the divisor sane
    t.cum[p + 1] = t.cum[p] + outer * row_len[p];
  }
  if (t.cum[nparts] == 0) return 0;
  DTF_LAUNCH(copy_parts_kernel, grid_for(t.cum[nparts]), 256, s, t);
  return (int)cudaGetLastError();
}

int dtf_relu_grad(const float* g, const float* y, float* out, long long n, cudaStream_t s) {
  DTF_LAUNCH(relu_grad_kernel, grid_for(n), 256, s, g, y, out, n);
  return (int)cudaGetLastError();
}

int dtf_colsum(const float* in, long long ld, int rows, int cols, float* out, cudaStream_t s) {
  DTF_LAUNCH(colsum_kernel, (cols + 31) / 32, 256, s, in, ld, rows, cols, out);
  return (int)cudaGetLastError();
}

int dtf_argmax_rows(const float* in, long long ld, int rows, int cols, long long* out, cudaStream_t s) {
  DTF_LAUNCH(argmax_rows_kernel, (rows + 3) / 4, 128, s, in, ld, rows, cols, out);
  return (int)cudaGetLastError();
}

int dtf_mean_of_n(const float* const* ins_dev, int n_in, float* out, long long n, cudaStream_t s) {
  DTF_LAUNCH(mean_of_n_kernel, grid_for(n), 256, s, ins_dev, n_in, out, n);
  return (int)cudaGetLastError();
}

int dtf_optimizer_apply(float* var, float* m, float* v, const float* g, void* shadow_bf16, long long n, int kind, float lr,
                        float momentum, int nesterov, float beta1, float beta2, float eps, float grad_scale,
                        cudaStream_t s) {
  ApplyArgs a;
  a.var = var; a.m = m; a.v = v; a.g = g; a.shadow = reinterpret_cast<__nv_bfloat16*>(shadow_bf16); a.n = n;
  a.kind = kind; a.lr = lr; a.momentum = momentum; a.nesterov = nesterov; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.grad_scale = grad_scale;
  DTF_LAUNCH(optimizer_apply_kernel, grid_for(n), 256, s, a);
  return (int)cudaGetLastError();
}

int dtf_im2col_nhwc(const float* x, void* cols, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int pt, int pl,
                    int ho, int wo, long long ldc, cudaStream_t s) {
  DTF_LAUNCH(im2col_nhwc_kernel, grid_for((long long)n * ho * wo * ldc), 256, s,
      x, reinterpret_cast<__nv_bfloat16*>(cols), n, h, w, c, kh, kw, sh, sw, pt, pl, ho, wo, ldc);
  return (int)cudaGetLastError();
}

int dtf_col2im_nhwc(const float* gcols, long long ldg, float* gx, int n, int h, int w, int c, int kh, int kw, int sh, int sw,
                    int pt, int pl, int ho, int wo, cudaStream_t s) {
  DTF_LAUNCH(col2im_nhwc_kernel, grid_for((long long)n * h * w * c), 256, s, gcols, ldg, gx, n, h, w, c, kh, kw, sh, sw, pt,
                                                                       pl, ho, wo);
  return (int)cudaGetLastError();
}

int dtf_gemm_ref(const void* a, const void* b, float* c, int M, int N, int K, long long lda, long long ldb, long long ldc,
                 int a_mn, int b_mn, const float* bias, int relu, float alpha, cudaStream_t s) {
  dim3 block(32, 8), grid((N + 31) / 32, (M + 7) / 8);
  DTF_LAUNCH(gemm_ref_kernel, grid, block, s, reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
                                         c, M, N, K, lda, ldb, ldc, a_mn, b_mn, bias, relu, alpha);
  return (int)cudaGetLastError();
}

}  // extern "C"
