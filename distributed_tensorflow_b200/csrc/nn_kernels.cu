// Fused training-mode batch normalisation for NHWC activations (the BN / ReLU / residual glue between the
// convolutions of ResNet-18).  The activation tensor is viewed as fp32 [rows, C] with rows = N*H*W, C % 4 == 0.
//
//   forward   bn_reduce<0>  : per-channel sum(x), sum(x^2) -> mean, rstd          (one pass over x)
//             bn_apply      : y = relu?( (x - mean) * rstd * scale + offset (+ residual) )   (one pass, float4)
//   backward  bn_reduce<1>  : g = dy * [y > 0]?;  doffset = sum(g), dscale = sum(g * xhat)   (one pass over dy, x, y)
//             bn_bwd_apply  : dx = scale * rstd * (g - doffset/rows - xhat * dscale/rows);  dresidual = g
//
// i.e. 2 launches forward + 2 backward per BN layer instead of ~25 element-wise / reduction launches of the eager
// formulation, and x is read twice + written once instead of ~10 round trips through HBM.
//
// Reduction layout: block (32, 8) -- x-threads own one float4 channel group each (a warp reads 512 contiguous bytes of a
// row; for C < 128 the spare lanes take further rows so the warp still covers 512 contiguous bytes), y-threads stride over
// rows with 4 (statistics) / 2 (backward) independent rows in flight per thread; grid (ceil(C/128), G).  Every block writes its partial sums to a workspace
// [G][2][C]; the LAST block to finish a channel column (ticket counter, threadfence pattern) folds the G partials in
// double precision and writes the per-channel results, then resets the ticket so the kernel is CUDA-graph replayable.
// Deterministic: no floating-point atomics.
//
// Also here: channel-vectorised convolution lowering for C % 8 == 0 (every ResNet layer but the 3-channel stem):
//   im2col_nhwc_vec8 : one thread = one (output pixel, filter tap, 8-channel group): 2 x 16-byte loads of fp32, one
//                      16-byte store of 8 bf16 -- 8x fewer index computations than the scalar kernel in elementwise.cu
//   col2im_nhwc_vec4 : one thread = one (input pixel, 4-channel group), gathers its <= kh*kw contributions as float4
//
// And the pooling of the model family (K14 "pool"), NHWC, 4 channels per thread:
//   maxpool_nhwc_fwd / _bwd : window max with the argmax (position inside the window) kept as one byte per element; the
//                            backward GATHERS -- every input element sums dy of the windows whose argmax it is -- so it is
//                            deterministic and needs no atomics.  Padding behaves like -inf (TF SAME semantics).
//   global_avgpool_nhwc_fwd / _bwd : mean over H*W per (image, channel) and its broadcast gradient.
//
// STATUS: validated on hardware in round 2 (tests/test_gpu_nn_fused.py, memcheck / racecheck / synccheck clean, ncu summaries in
// profiles/prof_nn.ncu-summary.txt); formulas also checked on CPU against autograd (tests/test_nn_fused_reference.py) and the
// source runs under the host emulation (tests/test_nn_kernels_host_emulation.py).  Default on (DTF_FUSED_NN=0 / DTF_FUSED_BN=0
// switch back to the element-wise PyTorch formulation in ops/native.py).
#include <cstdint>

// DTF_HOST_EMU: the same source compiled by g++ against tests/emu/host_emu.h (threads + barriers stand in for a thread
// block, blocks run one after another) so that indexing / reduction / ticket logic is checked on machines without a GPU.
#ifdef DTF_HOST_EMU
#include "host_emu.h"
#else
#include "common.cuh"
#endif

namespace dtf {

constexpr int kBnTX = 32;          // float4 channel groups per block (128 channels)
constexpr int kBnTY = 8;           // row lanes per block
constexpr int kBnCh = kBnTX * 4;   // channels per block column

// MODE 0: a = x,              s1 += a,  s2 += a * a
// MODE 1: g = dy * mask(y),   s1 += g,  s2 += g * (x - mean) * rstd
template <int MODE>
__global__ void __launch_bounds__(kBnTX* kBnTY)
bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y_mask,
                 const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, int C,
                 float* __restrict__ partial, unsigned int* __restrict__ tickets, float* __restrict__ r1,
                 float* __restrict__ r2, float eps) {
  __shared__ float sm[2][kBnTY][kBnCh];
  __shared__ double smd[2][kBnCh];
  __shared__ int is_last;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * kBnTX + tx;
  const int cbase = blockIdx.x * kBnCh;
  // Narrow tensors (C < 128, one column block): the x-lanes beyond C/4 would idle, so a warp covers SUB consecutive rows
  // instead -- lane = sub * (C/4) + channel group.  Wide tensors: SUB = 1, lane = channel group.
  const int c4n = C / 4;
  const int SUB = (gridDim.x == 1 && c4n < kBnTX && kBnTX % c4n == 0) ? kBnTX / c4n : 1;
  const int sub = SUB > 1 ? tx / c4n : 0;
  const int c = SUB > 1 ? (tx % c4n) * 4 : cbase + tx * 4;
  const bool live = c < C;
  // independent rows in flight per thread: the reduction is latency-bound otherwise (one 16-byte load per thread per trip)
  constexpr int UNR = MODE == 0 ? 4 : 2;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), rs = m;
    if (MODE == 1) {
      m = *reinterpret_cast<const float4*>(mean + c);
      rs = *reinterpret_cast<const float4*>(rstd + c);
    }
    const long long rstride = (long long)gridDim.y * kBnTY * SUB;
    for (long long r0 = ((long long)blockIdx.y * kBnTY + ty) * SUB + sub; r0 < rows; r0 += rstride * UNR) {
      float4 v[UNR], g[UNR], o[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const long long r = r0 + u * rstride;
        if (r < rows) {
          const long long off = r * C + c;
          v[u] = *reinterpret_cast<const float4*>(x + off);
          if (MODE == 1) {
            g[u] = *reinterpret_cast<const float4*>(dy + off);
            if (y_mask != nullptr) o[u] = *reinterpret_cast<const float4*>(y_mask + off);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (r0 + u * rstride >= rows) break;
        if (MODE == 0) {
          s1[0] += v[u].x; s1[1] += v[u].y; s1[2] += v[u].z; s1[3] += v[u].w;
          s2[0] += v[u].x * v[u].x; s2[1] += v[u].y * v[u].y; s2[2] += v[u].z * v[u].z; s2[3] += v[u].w * v[u].w;
        } else {
          float4 gg = g[u];
          if (y_mask != nullptr) {
            gg.x = o[u].x > 0.f ? gg.x : 0.f; gg.y = o[u].y > 0.f ? gg.y : 0.f;
            gg.z = o[u].z > 0.f ? gg.z : 0.f; gg.w = o[u].w > 0.f ? gg.w : 0.f;
          }
          s1[0] += gg.x; s1[1] += gg.y; s1[2] += gg.z; s1[3] += gg.w;
          s2[0] += gg.x * (v[u].x - m.x) * rs.x; s2[1] += gg.y * (v[u].y - m.y) * rs.y;
          s2[2] += gg.z * (v[u].z - m.z) * rs.z; s2[3] += gg.w * (v[u].w - m.w) * rs.w;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sm[0][ty][tx * 4 + j] = s1[j];
    sm[1][ty][tx * 4 + j] = s2[j];
  }
  __syncthreads();
  // 256 threads = 2 quantities x 128 lane-slots: fold the 8 row lanes ...
  const int q = t / kBnCh, ch = t % kBnCh;
  float acc = 0.f;
#pragma unroll
  for (int yy = 0; yy < kBnTY; ++yy) acc += sm[q][yy][ch];
  if (SUB > 1) {
    // ... and, for narrow tensors, the SUB row groups that share a channel: slot = sub * C + channel
    __syncthreads();
    sm[q][0][ch] = acc;
    __syncthreads();
    acc = 0.f;
    if (ch < C)
      for (int sidx = 0; sidx < SUB; ++sidx) acc += sm[q][0][sidx * C + ch];
  }
  if (cbase + ch < C) partial[((long long)blockIdx.y * 2 + q) * C + cbase + ch] = acc;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int ticket = atomicAdd(&tickets[blockIdx.x], 1u);
    is_last = (ticket == gridDim.y - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  {
    double dacc = 0.0;
    if (cbase + ch < C)
      for (unsigned int g = 0; g < gridDim.y; ++g) dacc += (double)__ldcg(partial + ((long long)g * 2 + q) * C + cbase + ch);
    smd[q][ch] = dacc;
  }
  __syncthreads();
  if (q == 0 && cbase + ch < C) {
    const double S1 = smd[0][ch], S2 = smd[1][ch];
    if (MODE == 0) {
      const double mu = S1 / (double)rows;
      double var = S2 / (double)rows - mu * mu;
      if (var < 0.0) var = 0.0;
      r1[cbase + ch] = (float)mu;
      r2[cbase + ch] = (float)(1.0 / sqrt(var + (double)eps));
    } else {
      r1[cbase + ch] = (float)S1;      // d offset
      r2[cbase + ch] = (float)S2;      // d scale
    }
  }
  if (t == 0) tickets[blockIdx.x] = 0u;     // replayable: the next launch starts from a clean ticket
}

// The grid stride (gridDim.x * blockDim.x float4 elements) is a multiple of C/4 for every power-of-two channel count, so a
// thread keeps ONE channel group for its whole walk and loads the per-channel coefficients once; otherwise they are
// re-read per element (they sit in L1).
__global__ void bn_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ res, float4* __restrict__ y,
                                const float4* __restrict__ mean, const float4* __restrict__ rstd,
                                const float4* __restrict__ scale, const float4* __restrict__ offset, long long n4, int c4n,
                                int relu) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fixed = stride % c4n == 0;
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f), a = m, of = m;
  if (fixed && i0 < n4) {
    const int c4 = (int)(i0 % c4n);
    const float4 rs = __ldg(rstd + c4), sc = __ldg(scale + c4);
    m = __ldg(mean + c4);
    of = __ldg(offset + c4);
    a = make_float4(rs.x * sc.x, rs.y * sc.y, rs.z * sc.z, rs.w * sc.w);
  }
  for (long long i = i0; i < n4; i += stride) {
    if (!fixed) {
      const int c4 = (int)(i % c4n);
      const float4 rs = __ldg(rstd + c4), sc = __ldg(scale + c4);
      m = __ldg(mean + c4);
      of = __ldg(offset + c4);
      a = make_float4(rs.x * sc.x, rs.y * sc.y, rs.z * sc.z, rs.w * sc.w);
    }
    const float4 v = x[i];
    float4 o;
    o.x = (v.x - m.x) * a.x + of.x;
    o.y = (v.y - m.y) * a.y + of.y;
    o.z = (v.z - m.z) * a.z + of.z;
    o.w = (v.w - m.w) * a.w + of.w;
    if (res != nullptr) {
      const float4 r = res[i];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    y[i] = o;
  }
}

__global__ void bn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ y_mask,
                                    const float4* __restrict__ x, const float4* __restrict__ mean,
                                    const float4* __restrict__ rstd, const float4* __restrict__ scale,
                                    const float4* __restrict__ doffset, const float4* __restrict__ dscale,
                                    float4* __restrict__ dx, float4* __restrict__ dres, long long n4, int c4n,
                                    float inv_rows) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fixed = stride % c4n == 0;
  // dx = k1 * (g - k2 - (x - m) * k3) with k1 = scale * rstd, k2 = doffset / rows, k3 = rstd * dscale / rows
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f), rs = m, k1 = m, k2 = m, k3 = m;
  auto coeffs = [&](int c4) {
    const float4 sc = __ldg(scale + c4), s1 = __ldg(doffset + c4), s2 = __ldg(dscale + c4);
    m = __ldg(mean + c4);
    rs = __ldg(rstd + c4);
    k1 = make_float4(sc.x * rs.x, sc.y * rs.y, sc.z * rs.z, sc.w * rs.w);
    k2 = make_float4(s1.x * inv_rows, s1.y * inv_rows, s1.z * inv_rows, s1.w * inv_rows);
    k3 = make_float4(rs.x * s2.x * inv_rows, rs.y * s2.y * inv_rows, rs.z * s2.z * inv_rows, rs.w * s2.w * inv_rows);
  };
  if (fixed && i0 < n4) coeffs((int)(i0 % c4n));
  for (long long i = i0; i < n4; i += stride) {
    if (!fixed) coeffs((int)(i % c4n));
    float4 g = dy[i];
    if (y_mask != nullptr) {
      const float4 o = y_mask[i];
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
      g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    const float4 v = x[i];
    float4 d;
    d.x = k1.x * (g.x - k2.x - (v.x - m.x) * k3.x);
    d.y = k1.y * (g.y - k2.y - (v.y - m.y) * k3.y);
    d.z = k1.z * (g.z - k2.z - (v.z - m.z) * k3.z);
    d.w = k1.w * (g.w - k2.w - (v.w - m.w) * k3.w);
    dx[i] = d;
    if (dres != nullptr) dres[i] = g;
  }
}

// cols[row, (ky*kw + kx)*c + ci] = bf16(x[b, oy*sh - pt + ky, ox*sw - pl + kx, ci]) (0 outside), row = (b*ho + oy)*wo + ox
__global__ void im2col_nhwc_vec8_kernel(const float* __restrict__ x, uint4* __restrict__ cols, int n, int h, int w, int c,
                                        int kh, int kw, int sh, int sw, int pt, int pl, int ho, int wo, long long ldc8) {
  const int c8n = c / 8, taps = kh * kw;
  const long long per_row = (long long)taps * c8n;
  const long long total = (long long)n * ho * wo * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int tap = rem / c8n, c8 = rem - tap * c8n;
    const int ky = tap / kw, kx = tap - ky * kw;
    const int ox = (int)(row % wo);
    const int oy = (int)((row / wo) % ho);
    const int b = (int)(row / ((long long)wo * ho));
    const int iy = oy * sh - pt + ky, ix = ox * sw - pl + kx;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
      const float4* src = reinterpret_cast<const float4*>(x + (((long long)b * h + iy) * w + ix) * c + c8 * 8);
      const float4 a = src[0], bq = src[1];
      out.x = pack_bf16x2(a.x, a.y); out.y = pack_bf16x2(a.z, a.w);
      out.z = pack_bf16x2(bq.x, bq.y); out.w = pack_bf16x2(bq.z, bq.w);
    }
    cols[row * ldc8 + (long long)tap * c8n + c8] = out;
  }
}

// gx[b, iy, ix, ci] = sum over taps (ky, kx) with (iy + pt - ky) % sh == 0, (ix + pl - kx) % sw == 0 and the output
// pixel in range of gcols[row(b, oy, ox), (ky*kw + kx)*c + ci]
__global__ void col2im_nhwc_vec4_kernel(const float* __restrict__ gcols, long long ldg, float4* __restrict__ gx, int n, int h,
                                        int w, int c, int kh, int kw, int sh, int sw, int pt, int pl, int ho, int wo) {
  const int c4n = c / 4;
  const long long total = (long long)n * h * w * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long pix = i / c4n;
    const int ix = (int)(pix % w);
    const int iy = (int)((pix / w) % h);
    const int b = (int)(pix / ((long long)w * h));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
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
        const float4 v = *reinterpret_cast<const float4*>(gcols + row * ldg + ((long long)ky * kw + kx) * c + c4 * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    gx[i] = s;
  }
}

// y[b, oy, ox, c] = max over the window; arg = ky * kw + kx of the (first) maximum
__global__ void maxpool_nhwc_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, uchar4* __restrict__ arg, int n,
                                        int h, int w, int c4n, int kh, int kw, int sh, int sw, int pt, int pl, int ho, int wo) {
  const long long total = (long long)n * ho * wo * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long pix = i / c4n;
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int b = (int)(pix / ((long long)wo * ho));
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 bi = make_uchar4(0, 0, 0, 0);
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = oy * sh - pt + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = ox * sw - pl + kx;
        if (ix < 0 || ix >= w) continue;
        const float4 v = x[(((long long)b * h + iy) * w + ix) * c4n + c4];
        const unsigned char t = (unsigned char)(ky * kw + kx);
        if (v.x > best.x) { best.x = v.x; bi.x = t; }
        if (v.y > best.y) { best.y = v.y; bi.y = t; }
        if (v.z > best.z) { best.z = v.z; bi.z = t; }
        if (v.w > best.w) { best.w = v.w; bi.w = t; }
      }
    }
    y[i] = best;
    arg[i] = bi;
  }
}

// dx[b, iy, ix, c] = sum of dy over the windows that cover (iy, ix) AND whose argmax is this position
__global__ void maxpool_nhwc_bwd_kernel(const float4* __restrict__ dy, const uchar4* __restrict__ arg, float4* __restrict__ dx,
                                        int n, int h, int w, int c4n, int kh, int kw, int sh, int sw, int pt, int pl, int ho,
                                        int wo) {
  const long long total = (long long)n * h * w * c4n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long pix = i / c4n;
    const int ix = (int)(pix % w);
    const int iy = (int)((pix / w) % h);
    const int b = (int)(pix / ((long long)w * h));
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
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
        const long long o = (((long long)b * ho + oy) * wo + ox) * c4n + c4;
        const uchar4 a = arg[o];
        const float4 g = dy[o];
        const unsigned char t = (unsigned char)(ky * kw + kx);
        if (a.x == t) s.x += g.x;
        if (a.y == t) s.y += g.y;
        if (a.z == t) s.z += g.z;
        if (a.w == t) s.w += g.w;
      }
    }
    dx[i] = s;
  }
}

// out[b, c] = mean over h*w of x[b, :, :, c]; one thread per (image, 4 channels): consecutive threads read consecutive
// channel groups, so every row of the walk is one coalesced segment
__global__ void global_avgpool_nhwc_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ out, int n, int hw, int c4n) {
  const long long total = (long long)n * c4n;
  const float inv = 1.0f / (float)hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long b = i / c4n;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < hw; ++p) {
      const float4 v = x[(b * hw + p) * c4n + c4];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  }
}

__global__ void global_avgpool_nhwc_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int n, int hw, int c4n) {
  const long long total = (long long)n * hw * c4n;
  const float inv = 1.0f / (float)hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const long long b = i / ((long long)hw * c4n);
    const float4 g = dy[b * c4n + c4];
    dx[i] = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static inline int bn_grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace dtf

extern "C" {
using namespace dtf;

// Row-split factor G the launchers use for [rows, C]: enough blocks to fill 148 SMs a few times over, at least 64 rows
// per block.  The workspace must hold dtf_bn_workspace_floats(rows, C) floats, tickets ceil(C/128) zeroed uint32.
int dtf_bn_row_splits(long long rows, int C) {
  const int gx = (C + kBnCh - 1) / kBnCh;
  long long g = (148 * 4 + gx - 1) / gx;
  const long long by_rows = (rows + 63) / 64;
  if (g > by_rows) g = by_rows;
  if (g < 1) g = 1;
  return (int)g;
}

long long dtf_bn_workspace_floats(long long rows, int C) { return (long long)dtf_bn_row_splits(rows, C) * 2 * C; }

// mode 0: (x) -> r1 = mean, r2 = rstd.   mode 1: (x, dy, y_mask?, mean, rstd) -> r1 = doffset, r2 = dscale.
int dtf_bn_reduce(int mode, const float* x, const float* dy, const float* y_mask, const float* mean, const float* rstd,
                  long long rows, int C, float* workspace, unsigned int* tickets, float* r1, float* r2, float eps,
                  cudaStream_t s) {
  if (C % 4 != 0 || rows <= 0 || !aligned16(x) || (mode == 1 && (!aligned16(dy) || !aligned16(mean) || !aligned16(rstd))) ||
      (y_mask != nullptr && !aligned16(y_mask)))
    return -1;
  dim3 block(kBnTX, kBnTY), grid((C + kBnCh - 1) / kBnCh, dtf_bn_row_splits(rows, C));
  const float* none = nullptr;
  if (mode == 0)
    DTF_LAUNCH(bn_reduce_kernel<0>, grid, block, s, x, none, none, none, none, rows, C, workspace, tickets, r1, r2, eps);
  else
    DTF_LAUNCH(bn_reduce_kernel<1>, grid, block, s, x, dy, y_mask, mean, rstd, rows, C, workspace, tickets, r1, r2, eps);
  return (int)cudaGetLastError();
}

int dtf_bn_apply(const float* x, const float* residual, float* y, const float* mean, const float* rstd, const float* scale,
                 const float* offset, long long rows, int C, int relu, cudaStream_t s) {
  if (C % 4 != 0 || !aligned16(x) || !aligned16(y) || !aligned16(mean) || !aligned16(rstd) || !aligned16(scale) ||
      !aligned16(offset) || (residual != nullptr && !aligned16(residual)))
    return -1;
  const long long n4 = rows * C / 4;
  DTF_LAUNCH(bn_apply_kernel, bn_grid_for(n4), 256, s, reinterpret_cast<const float4*>(x),
             reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(mean),
             reinterpret_cast<const float4*>(rstd), reinterpret_cast<const float4*>(scale),
             reinterpret_cast<const float4*>(offset), n4, C / 4, relu);
  return (int)cudaGetLastError();
}

int dtf_bn_bwd_apply(const float* dy, const float* y_mask, const float* x, const float* mean, const float* rstd,
                     const float* scale, const float* doffset, const float* dscale, float* dx, float* dres, long long rows,
                     int C, cudaStream_t s) {
  if (C % 4 != 0 || !aligned16(dy) || !aligned16(x) || !aligned16(dx) || !aligned16(mean) || !aligned16(rstd) ||
      !aligned16(scale) || !aligned16(doffset) || !aligned16(dscale) || (y_mask != nullptr && !aligned16(y_mask)) ||
      (dres != nullptr && !aligned16(dres)))
    return -1;
  const long long n4 = rows * C / 4;
  DTF_LAUNCH(bn_bwd_apply_kernel, bn_grid_for(n4), 256, s, reinterpret_cast<const float4*>(dy),
             reinterpret_cast<const float4*>(y_mask), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(mean),
             reinterpret_cast<const float4*>(rstd), reinterpret_cast<const float4*>(scale),
             reinterpret_cast<const float4*>(doffset), reinterpret_cast<const float4*>(dscale), reinterpret_cast<float4*>(dx),
             reinterpret_cast<float4*>(dres), n4, C / 4, 1.0f / (float)rows);
  return (int)cudaGetLastError();
}

// -1: shape / alignment not eligible (the caller falls back to the scalar kernels of elementwise.cu)
int dtf_im2col_nhwc_vec8(const float* x, void* cols, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int pt, int pl,
                         int ho, int wo, long long ldc, cudaStream_t s) {
  if (c % 8 != 0 || ldc % 8 != 0 || ldc != (long long)kh * kw * c || !aligned16(x) || !aligned16(cols)) return -1;
  const long long total = (long long)n * ho * wo * kh * kw * (c / 8);
  DTF_LAUNCH(im2col_nhwc_vec8_kernel, bn_grid_for(total), 256, s, x, reinterpret_cast<uint4*>(cols), n, h, w, c, kh, kw, sh, sw,
             pt, pl, ho, wo, ldc / 8);
  return (int)cudaGetLastError();
}

int dtf_col2im_nhwc_vec4(const float* gcols, long long ldg, float* gx, int n, int h, int w, int c, int kh, int kw, int sh, int sw,
                         int pt, int pl, int ho, int wo, cudaStream_t s) {
  if (c % 4 != 0 || ldg % 4 != 0 || !aligned16(gcols) || !aligned16(gx)) return -1;
  const long long total = (long long)n * h * w * (c / 4);
  DTF_LAUNCH(col2im_nhwc_vec4_kernel, bn_grid_for(total), 256, s, gcols, ldg, reinterpret_cast<float4*>(gx), n, h, w, c, kh, kw,
             sh, sw, pt, pl, ho, wo);
  return (int)cudaGetLastError();
}

// pooling: -1 = not eligible (C % 4, alignment, window of more than 255 positions)
int dtf_maxpool_nhwc_fwd(const float* x, float* y, void* arg, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int pt,
                         int pl, int ho, int wo, cudaStream_t s) {
  if (c % 4 != 0 || kh * kw > 255 || !aligned16(x) || !aligned16(y) || (reinterpret_cast<uintptr_t>(arg) & 3)) return -1;
  const long long total = (long long)n * ho * wo * (c / 4);
  DTF_LAUNCH(maxpool_nhwc_fwd_kernel, bn_grid_for(total), 256, s, reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y),
             reinterpret_cast<uchar4*>(arg), n, h, w, c / 4, kh, kw, sh, sw, pt, pl, ho, wo);
  return (int)cudaGetLastError();
}

int dtf_maxpool_nhwc_bwd(const float* dy, const void* arg, float* dx, int n, int h, int w, int c, int kh, int kw, int sh, int sw,
                         int pt, int pl, int ho, int wo, cudaStream_t s) {
  if (c % 4 != 0 || kh * kw > 255 || !aligned16(dy) || !aligned16(dx) || (reinterpret_cast<uintptr_t>(arg) & 3)) return -1;
  const long long total = (long long)n * h * w * (c / 4);
  DTF_LAUNCH(maxpool_nhwc_bwd_kernel, bn_grid_for(total), 256, s, reinterpret_cast<const float4*>(dy),
             reinterpret_cast<const uchar4*>(arg), reinterpret_cast<float4*>(dx), n, h, w, c / 4, kh, kw, sh, sw, pt, pl, ho, wo);
  return (int)cudaGetLastError();
}

int dtf_global_avgpool_nhwc(const float* in, float* out, int n, int hw, int c, int backward, cudaStream_t s) {
  if (c % 4 != 0 || hw <= 0 || !aligned16(in) || !aligned16(out)) return -1;
  if (!backward)
    DTF_LAUNCH(global_avgpool_nhwc_fwd_kernel, bn_grid_for((long long)n * (c / 4)), 256, s, reinterpret_cast<const float4*>(in),
               reinterpret_cast<float4*>(out), n, hw, c / 4);
  else
    DTF_LAUNCH(global_avgpool_nhwc_bwd_kernel, bn_grid_for((long long)n * hw * (c / 4)), 256, s,
               reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, hw, c / 4);
  return (int)cudaGetLastError();
}

}  // extern "C"
