// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues and fused NVLink push/pull.
//
//   C[M,N] = alpha * op(A)[M,K] . op(B)[K,N]  (+ bias[N]) (ReLU) (* relu-mask) ; fp32 accumulate in tensor memory.
//   Operand precision: bf16 (tcgen05.mma.kind::f16) or fp32 storage consumed as TF32 (tcgen05.mma.kind::tf32: the
//   reference model is fp32 end to end, /root/reference/distributed_mnist.py:98-113 -- fp32 tensors stay fp32 in HBM,
//   TMA moves them untouched and the tensor core reads the top 19 bits).  One shared-memory stage is 128 bytes of K
//   per row for both: 64 bf16 or 32 fp32 elements, four MMAs (K = 16 / K = 8) per stage.
//
// * Operands are loaded by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage
//   shared-memory ring; one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) with the
//   accumulator in tensor memory; four epilogue warps read it back with tcgen05.ld.
// * Both operands may be K-major or MN-major (UMMA descriptor + instruction-descriptor major bits), so the
//   three GEMMs of a dense layer -- y = x.W, dW = x^T.dy, dx = dy.W^T with TF-layout W[in,out] --
//   run without any transpose pass.
// * Fused pull (SURVEY C1+K1): the B tensor map may point at a parameter-server GPU's published
//   parameter buffer (peer memory over NVLink); the TMA producer first acquires the ps's version flag.
// * Fused push (SURVEY K4+C2): C may be a gradient slot in the ps GPU's memory; the epilogue stores
//   tiles straight from TMEM to the peer and then release-increments the ps's arrival counter.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..9 = epilogue (warp w owns TMEM lanes [32*(w%4), +32); two warps per quarter split the columns).
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

// DTF_HOST_EMU: only the HOST half of this file (argument checks, tile / stage / kernel selection, tensor-map boxes, grid
// and shared-memory sizes) is compiled by g++, against tests/emu/gemm_host_stubs.h which records what would have been
// encoded and launched -- the dispatch logic is unit-tested on machines without a GPU.  The kernels are hardware-only.
#ifdef DTF_HOST_EMU
#include "gemm_host_stubs.h"
#else
#include "common.cuh"
#endif

namespace dtf {

static constexpr int kBlockM = 128;
static constexpr int kABytes = kBlockM * 128;   // 16 KB: 128 rows x one 128-byte swizzle row (64 bf16 / 32 fp32 of K)
static constexpr int kEpilogueWarps = 8;
static constexpr int kThreads = 64 + 32 * kEpilogueWarps;   // TMA warp + MMA warp + epilogue warps

struct GemmParams {
  int M, N, K;
  int block_n;           // multiple of 16 (multiple of 64 when B is MN-major), <= 256
  int a_mn, b_mn;        // 1 = operand stored MN-major ([K, M] / [K, N] row-major)
  int num_kb;            // ceil(K / 64)
  int kb_per_split;      // K blocks handled by one blockIdx.z
  int stages;
  void* c;
  long long ldc;
  int c_bf16;
  const float* bias;
  int relu;
  const __nv_bfloat16* mask;   // optional [M, ldmask]: output *= (mask > 0)   (ReLU backward)
  long long ldmask;
  float alpha;
  int atomic;            // split-K / accumulate: red.add into fp32 C
  float* colsum;         // optional [N]: += column sums of the (post-mask, post-alpha) tile (bias gradient)
  const unsigned long long* wait_flag;   // optional: acquire until *wait_flag >= wait_target before loading
  unsigned long long wait_target;
  const unsigned long long* wait_target_ptr;   // optional: target = wait_target + *wait_target_ptr (device step counter)
  unsigned long long* signal;            // optional: release-increment by 1 per CTA after the tile is stored
  unsigned int* err;                     // optional: set to 1 when the wait timed out
  unsigned long long timeout_ns;
  long long* phase_trace;                // optional [gridsize][16] clock64 stamps (profiling builds of the step)
  int signal_gpu_scope;                  // 1: consumer is on the same GPU (gpu-scope fence suffices)
  const unsigned long long* stamp_src;   // optional: *stamp_dst = *stamp_src before the arrival (block 0,0,0 only)
  unsigned long long* stamp_dst;
  // operand precision (host-filled): bf16 -> {0, 64, 8192, 128};  fp32-as-tf32 -> {1, 32, 4096, 64}
  int tf32;              // 1: fp32 operands, tcgen05.mma.kind::tf32
  int kbk;               // K elements per shared-memory stage (= one 128-byte swizzle row)
  int mn_lbo;            // MN-major tiles: bytes between adjacent 128-byte-wide MN chunks (= kbk rows x 128 B)
  int mn_kstep16;        // MN-major tiles: descriptor advance (>> 4) per MMA = K-per-MMA rows x 128 B
  int mn_sbo;            // MN-major tiles: bytes between swizzle atoms along K (8 rows x 128 B; fp32: 4 rows x 128 B)
  int mn_type;           // MN-major tiles: descriptor layout type (2 = SWIZZLE_128B; fp32: 1 = SWIZZLE_128B_BASE32B)
  // implicit-GEMM convolution (bf16, stride 1): map_a is a 4-D map {C, W, H, images} over the NHWC activation and the
  // A operand is its patch matrix, never materialised -- one shifted box per (tap, 64-channel chunk); the zero fill of the
  // out-of-bounds part of a box IS the padding.
  //   conv == 1: A = patches [pixels, taps*C] (K-major: fprop, and dgrad on dY with the flipped filter); box = 128 pixels
  //   conv == 2: A = patches^T [taps*C, pixels] (MN-major: wgrad, K = pixels); box = 64 pixels, one per 64-channel M chunk
  int conv;
  int cv_w, cv_hw;       // W and H*W of the activation (pixels are ordered image, row, column)
  int cv_cpt;            // 64-channel chunks per tap (C / 64)
  int cv_kw, cv_taps;    // filter width, kh*kw
  int cv_pt, cv_pl;      // top / left padding
  int cv_n;              // images (coordinate of an all-out-of-bounds box: the M tail of wgrad)
};


#ifndef DTF_HOST_EMU   // ---- device code (tcgen05 / TMEM / TMA): hardware only ----
// One MMA of the stage: bf16 (K = 16) or tf32 (K = 8); CTAS == 2 -> cta_group::2 (leader thread of the pair).
template <int CTAS>
DTF_DEVICE void umma_any(const GemmParams& p, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if (p.tf32) {
    if (CTAS == 2) umma_tf32_2cta(d_tmem, a_desc, b_desc, idesc, acc); else umma_tf32(d_tmem, a_desc, b_desc, idesc, acc);
  } else {
    if (CTAS == 2) umma_bf16_2cta(d_tmem, a_desc, b_desc, idesc, acc); else umma_bf16(d_tmem, a_desc, b_desc, idesc, acc);
  }
}
// One 16-column slice of the epilogue: alpha, bias, ReLU, ReLU-backward mask, bias-gradient column sums, store
// (fp32 / bf16 / atomic accumulate).  `r` holds the fp32 accumulators of this thread's row.
DTF_DEVICE void epilogue_chunk16(const GemmParams& p, const uint32_t* r, int c0, int n0, long long grow, bool row_ok,
                                 bool add_bias, const float* s_bias, int lane, float* stage_row = nullptr) {
  float* cf = reinterpret_cast<float*>(p.c);
  __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.c);
  const int gc0 = n0 + c0;
  if (gc0 >= p.N) return;                     // warp-uniform
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = __uint_as_float(r[j]) * p.alpha;
    if (add_bias) x += s_bias[c0 + j];
    if (p.relu) x = fmaxf(x, 0.0f);
    v[j] = x;
  }
  if (p.mask != nullptr) {
    const __nv_bfloat16* mrow = p.mask + grow * p.ldmask + gc0;
    if (row_ok && gc0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(mrow) & 15) == 0)) {
      const uint4 m0 = reinterpret_cast<const uint4*>(mrow)[0];
      const uint4 m1 = reinterpret_cast<const uint4*>(mrow)[1];
      const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // bf16 > 0  <=>  sign bit clear and magnitude non-zero
        const uint32_t lo = mw[j] & 0xFFFFu, hi = mw[j] >> 16;
        v[2 * j] = (lo != 0u && lo < 0x8000u) ? v[2 * j] : 0.0f;
        v[2 * j + 1] = (hi != 0u && hi < 0x8000u) ? v[2 * j + 1] : 0.0f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int gc = gc0 + j;
        float mk = 0.0f;
        if (row_ok && gc < p.N) mk = __bfloat162float(p.mask[grow * p.ldmask + gc]);
        v[j] = mk > 0.0f ? v[j] : 0.0f;
      }
    }
  }
  if (p.colsum != nullptr) {
    float mine = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float s = row_ok ? v[j] : 0.0f;
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (lane == j) mine = s;
    }
    if (lane < 16 && gc0 + lane < p.N) atomicAdd(p.colsum + gc0 + lane, mine);
  }
  if (stage_row != nullptr) {
    // coalescing path: park this thread's 16 values in its row of the warp's staging tile
#pragma unroll
    for (int j = 0; j < 16; ++j) stage_row[j] = v[j];
    return;
  }
  if (row_ok) {
    const bool full = gc0 + 16 <= p.N;
    if (p.atomic) {
      float* dst = cf + grow * p.ldc + gc0;
      if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        // split-K partials: one 16-byte reduction per four columns (REDG.E.ADD.F32x4) -- the scalar form made the L2 atomic
        // units the bottleneck of every split GEMM (conv wgrad / small-M fprop: ~16 M scalar atomics = ~60 us per layer)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                       "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (gc0 + j < p.N) atomicAdd(dst + j, v[j]);
      }
    } else if (p.c_bf16) {
      __nv_bfloat16* dst = cb + grow * p.ldc + gc0;
      if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        uint4 w0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                              pack_bf16x2(v[6], v[7]));
        uint4 w1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                              pack_bf16x2(v[14], v[15]));
        reinterpret_cast<uint4*>(dst)[0] = w0;
        reinterpret_cast<uint4*>(dst)[1] = w1;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (gc0 + j < p.N) dst[j] = __float2bfloat16(v[j]);
      }
    } else {
      float* dst = cf + grow * p.ldc + gc0;
      if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (gc0 + j < p.N) dst[j] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_holder;
  __shared__ float s_bias[256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* tr = p.phase_trace ? p.phase_trace + 16 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
#define DTF_STAMP(slot) do { if (tr) tr[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) DTF_STAMP(0);
  const int m0 = blockIdx.x * kBlockM;
  const int n0 = blockIdx.y * p.block_n;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int b_bytes = p.block_n * 128;                 // 128 bytes of K per row: 64 bf16 / 32 fp32
  const int stage_bytes = kABytes + b_bytes;
  const int mnc = p.kbk;                               // elements per 128-byte MN chunk (== K elements per stage)
  // 128B-swizzled tiles must start on 1024-byte boundaries
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.block_n) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_holder, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  if (threadIdx.x == 0) DTF_STAMP(1);          // setup done (barriers, TMEM alloc)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (p.wait_flag != nullptr) {
        // fused pull: the parameters behind map_b (or map_a) are published by another GPU
        const unsigned long long target = p.wait_target + (p.wait_target_ptr ? *p.wait_target_ptr : 0ull);
        if (!wait_flag_ge_u64(reinterpret_cast<const uint64_t*>(p.wait_flag), target, p.timeout_ns)) {
          if (p.err) atomicExch(p.err, 1u);
        }
        fence_proxy_async();
      }
      DTF_STAMP(2);                              // token acquired
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* a_dst = tiles + s * stage_bytes;
        uint8_t* b_dst = a_dst + kABytes;
        const int k0 = kb * p.kbk;
        if (p.conv == 1) {
          const int tap = kb / p.cv_cpt, c0 = (kb - tap * p.cv_cpt) * 64;
          const int ky = tap / p.cv_kw, kx = tap - ky * p.cv_kw;
          const int img = m0 / p.cv_hw, h0 = (m0 - img * p.cv_hw) / p.cv_w;
          tma_load_4d(a_dst, &map_a, &full_bar[s], c0, kx - p.cv_pl, h0 + ky - p.cv_pt, img);   // box {64 ch, W, rows, images} = 128 pixels
        } else if (p.conv == 2) {
          const int img = k0 / p.cv_hw, h0 = (k0 - img * p.cv_hw) / p.cv_w;                     // k0 = first of 64 pixels
          for (int j = 0; j < 2; ++j) {
            const int mc = m0 / 64 + j, tap = mc / p.cv_cpt, c0 = (mc - tap * p.cv_cpt) * 64;
            const int ky = tap / p.cv_kw, kx = tap - ky * p.cv_kw;
            tma_load_4d(a_dst + j * p.mn_lbo, &map_a, &full_bar[s], c0, kx - p.cv_pl, h0 + ky - p.cv_pt,
                        tap < p.cv_taps ? img : p.cv_n);
          }
        } else if (!p.a_mn) {
          tma_load_2d(a_dst, &map_a, &full_bar[s], k0, m0);                 // box {128 B of k, 128 m}
        } else {
          for (int j = 0; j < kBlockM / mnc; ++j)                           // box {128 B of m, kbk k} x (2 | 4)
            tma_load_2d(a_dst + j * p.mn_lbo, &map_a, &full_bar[s], m0 + mnc * j, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(b_dst, &map_b, &full_bar[s], k0, n0);                 // box {128 B of k, block_n}
        } else {
          for (int j = 0; j < p.block_n / mnc; ++j)
            tma_load_2d(b_dst + j * p.mn_lbo, &map_b, &full_bar[s], n0 + mnc * j, k0);   // box {128 B of n, kbk k}
        }
      }
      DTF_STAMP(3);                              // last TMA issued
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc = make_idesc(kBlockM, p.block_n, p.a_mn, p.b_mn, p.tf32);
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        if (it == 0) DTF_STAMP(4);               // first stage landed
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * stage_bytes);
        const uint32_t b_addr = a_addr + kABytes;
        const uint64_t a_desc = p.a_mn ? make_smem_desc(a_addr, p.mn_lbo, p.mn_sbo, p.mn_type) : make_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t b_desc = p.b_mn ? make_smem_desc(b_addr, p.mn_lbo, p.mn_sbo, p.mn_type) : make_smem_desc_sw128(b_addr, 16, 1024);
        const uint32_t a_step = p.a_mn ? (uint32_t)p.mn_kstep16 : (32u >> 4);   // advance one MMA of K (32 bytes K-major)
        const uint32_t b_step = p.b_mn ? (uint32_t)p.mn_kstep16 : (32u >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                           // 4 MMAs per 128-byte stage
          umma_any<1>(p, tmem_base, a_desc + (uint64_t)(a_step * k), b_desc + (uint64_t)(b_step * k), idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);      // frees the smem slot once these MMAs have consumed it
      }
      umma_commit(&tmem_full_bar);       // accumulator complete
      DTF_STAMP(5);                              // last MMA issued
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const long long grow = (long long)m0 + row;
    const bool row_ok = grow < p.M;
    const bool have_k = kb_end > kb_begin;
    const bool add_bias = p.bias != nullptr && (p.atomic == 0 || blockIdx.z == 0);
    if (add_bias) {
      // stage the bias slice in shared memory while the mainloop runs (no global-load latency in the epilogue)
      for (int i = threadIdx.x - 64; i < p.block_n; i += 32 * kEpilogueWarps) s_bias[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
      asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
    }
    if (have_k) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
    if (threadIdx.x == 64) DTF_STAMP(6);         // accumulator ready
    // Eight epilogue warps: two per TMEM lane quarter, each taking half of the columns.  With one warp per
    // scheduler every dependent latency (tcgen05.ld, bias LDS, address math, store issue) is exposed, so the
    // epilogue is latency- not bandwidth-bound: measured 575 cycles per 16-column slice with four warps; wider
    // TMEM loads (x32, double-buffered) and a shared-memory transpose for coalesced stores were both SLOWER.
    const int half = (warp - 2) >> 2;
    const bool split_cols = (p.block_n % 32) == 0;
    const int c_begin = split_cols ? half * (p.block_n / 2) : 0;
    const int c_end = split_cols ? c_begin + p.block_n / 2 : (half == 0 ? p.block_n : 0);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      uint32_t r[16];
      if (have_k) {
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0;
      }
      epilogue_chunk16(p, r, c0, n0, grow, row_ok, add_bias, s_bias, lane);
    }
    if (threadIdx.x == 64) DTF_STAMP(7);         // tile stored
    if (p.signal != nullptr) {
      // fused push: make this CTA's tile visible system-wide, then bump the consumer's arrival counter
      // (one fence by one thread after the CTA-level barrier: release is cumulative, and a system-scope
      //  membar per thread serialises for tens of microseconds)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      if (warp == 2 && lane == 0) {
        if (p.stamp_dst != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
          asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.stamp_dst), "l"(*p.stamp_src) : "memory");
        if (p.signal_gpu_scope) __threadfence(); else fence_acq_rel_sys();
        red_release_sys_add_u64(reinterpret_cast<uint64_t*>(p.signal), 1ull);
        DTF_STAMP(8);                            // fence + arrival signalled
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  if (threadIdx.x == 32) DTF_STAMP(9);           // end
}

// =================================================================================================
// Persistent variant for compute-sized problems (plain GEMMs: conv-as-GEMM, dense layers at scale).
//
// One CTA (or CTA PAIR) per SM loops over output tiles (static round-robin, M fastest so concurrently running
// CTAs share B panels in L2).  The accumulator is DOUBLE-BUFFERED in tensor memory (2 x BLOCK_N columns <= 512):
// while the epilogue warps drain tile i out of TMEM buffer i%2, the MMA thread is already accumulating tile i+1
// into the other buffer and the TMA producer is prefetching its operands -- the epilogue, the kernel prologue
// (barrier init, TMEM alloc, descriptor prefetch) and the pipeline fill are paid once per SM instead of once per
// tile.  Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), the tile loop.
//
// CTAS == 2 (tcgen05 cta_group::2): the two SMs of a TPC form a 2-CTA cluster working on ONE 256 x BLOCK_N tile.
// Each CTA loads its own 128 rows of A and only HALF of the B tile (BLOCK_N/2 rows); the leader CTA's single
// MMA thread issues M=256 instructions that read both CTAs' shared memory and write both CTAs' tensor memory.
// Per output element this moves 1.5x fewer operand bytes from L2 into the SMs than the 1-CTA 128 x 256 tile --
// measured: the 1-CTA kernel saturates the L2 -> SM fabric (~12 TB/s) at ~1.15 PFLOP/s.
// Barrier topology (pair): full[s] lives in the leader (one arrive.expect_tx for BOTH CTAs' bytes; each CTA's TMA
// counts its bytes there), empty[s] / tmem_full[a] live in both CTAs (multicast tcgen05.commit), tmem_empty[a]
// lives in the leader (epilogue warps of both CTAs arrive, the peer's remotely).
// =================================================================================================
template <int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                    const GemmParams p, const int tiles_m, const int tiles_n) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_holder;
  __shared__ float s_bias[2][256];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int bn_cta = p.block_n / CTAS;                         // rows of the B tile THIS CTA loads
  const int b_bytes = bn_cta * 128;
  const int stage_bytes = kABytes + b_bytes;                    // per CTA
  const int mnc = p.kbk;
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t acc_cols = 32;                       // columns of ONE accumulator buffer (power of two >= BLOCK_N)
  while (acc_cols < (uint32_t)p.block_n) acc_cols <<= 1;
  const uint32_t tmem_cols = 2 * acc_cols;
  const int num_tiles = tiles_m * tiles_n;                      // tiles of (CTAS * 128) x BLOCK_N
  const int worker = blockIdx.x / CTAS, num_workers = gridDim.x / CTAS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], CTAS * kEpilogueWarps);      // one arrival per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_2cta(&tmem_holder, tmem_cols); tmem_relinquish_2cta(); }
    else { tmem_alloc(&tmem_holder, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  // (pair: barrier.cluster orders the allocator's write of the holder in BOTH CTAs; the CTA barrier behind it is free here
  //  -- once per kernel -- and is the ordering compute-sanitizer's racecheck models for shared memory)
  if (CTAS == 2) { cluster_sync_all(); __syncthreads(); } else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    if (elect_one()) {
      int it = 0;
      for (int t = worker; t < num_tiles; t += num_workers) {
        const int m0 = (t % tiles_m) * (kBlockM * CTAS) + (int)cta_rank * kBlockM;
        const int n0 = (t / tiles_m) * p.block_n + (int)cta_rank * bn_cta;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[s], stage_bytes * CTAS);
          uint8_t* a_dst = tiles + s * stage_bytes;
          uint8_t* b_dst = a_dst + kABytes;
          const int k0 = kb * p.kbk;
          if (CTAS == 2) {
            if (!p.a_mn) {
              tma_load_2d_2cta(a_dst, &map_a, &full_bar[s], k0, m0);
            } else {
              for (int j = 0; j < kBlockM / mnc; ++j)
                tma_load_2d_2cta(a_dst + j * p.mn_lbo, &map_a, &full_bar[s], m0 + mnc * j, k0);
            }
            if (!p.b_mn) {
              tma_load_2d_2cta(b_dst, &map_b, &full_bar[s], k0, n0);
            } else {
              for (int j = 0; j < bn_cta / mnc; ++j)
                tma_load_2d_2cta(b_dst + j * p.mn_lbo, &map_b, &full_bar[s], n0 + mnc * j, k0);
            }
          } else {
            if (p.conv == 1) {
              // implicit-GEMM fprop / dgrad (see gemm_bf16_tcgen05_kernel): one shifted 4-D box of the activation per K block
              const int tap = kb / p.cv_cpt, c0 = (kb - tap * p.cv_cpt) * 64;
              const int ky = tap / p.cv_kw, kx = tap - ky * p.cv_kw;
              const int img = m0 / p.cv_hw, h0 = (m0 - img * p.cv_hw) / p.cv_w;
              tma_load_4d(a_dst, &map_a, &full_bar[s], c0, kx - p.cv_pl, h0 + ky - p.cv_pt, img);
            } else if (!p.a_mn) {
              tma_load_2d(a_dst, &map_a, &full_bar[s], k0, m0);
            } else {
              for (int j = 0; j < kBlockM / mnc; ++j)
                tma_load_2d(a_dst + j * p.mn_lbo, &map_a, &full_bar[s], m0 + mnc * j, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(b_dst, &map_b, &full_bar[s], k0, n0);
            } else {
              for (int j = 0; j < bn_cta / mnc; ++j)
                tma_load_2d(b_dst + j * p.mn_lbo, &map_b, &full_bar[s], n0 + mnc * j, k0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      const uint32_t idesc = make_idesc(kBlockM * CTAS, p.block_n, p.a_mn, p.b_mn, p.tf32);
      const uint32_t a_step = p.a_mn ? (uint32_t)p.mn_kstep16 : (32u >> 4);
      const uint32_t b_step = p.b_mn ? (uint32_t)p.mn_kstep16 : (32u >> 4);
      int it = 0, ti = 0;
      for (int t = worker; t < num_tiles; t += num_workers, ++ti) {
        const int acc = ti & 1;
        const uint32_t acc_ph = (ti >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);        // the epilogues have drained this accumulator buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + s * stage_bytes);
          const uint32_t b_addr = a_addr + kABytes;
          const uint64_t a_desc = p.a_mn ? make_smem_desc(a_addr, p.mn_lbo, p.mn_sbo, p.mn_type) : make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = p.b_mn ? make_smem_desc(b_addr, p.mn_lbo, p.mn_sbo, p.mn_type) : make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_any<CTAS>(p, d_tmem, a_desc + (uint64_t)(a_step * k), b_desc + (uint64_t)(b_step * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (CTAS == 2) umma_commit_2cta_mc(&empty_bar[s]); else umma_commit(&empty_bar[s]);
        }
        if (CTAS == 2) umma_commit_2cta_mc(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs: each drains its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool split_cols = (p.block_n % 32) == 0;
    const int c_begin = split_cols ? half * (p.block_n / 2) : 0;
    const int c_end = split_cols ? c_begin + p.block_n / 2 : (half == 0 ? p.block_n : 0);
    const bool add_bias = p.bias != nullptr;
    int ti = 0;
    for (int t = worker; t < num_tiles; t += num_workers, ++ti) {
      const int acc = ti & 1;
      const uint32_t acc_ph = (ti >> 1) & 1;
      const int m0 = (t % tiles_m) * (kBlockM * CTAS) + (int)cta_rank * kBlockM;
      const int n0 = (t / tiles_m) * p.block_n;
      const long long grow = (long long)m0 + q * 32 + lane;
      const bool row_ok = grow < p.M;
      if (add_bias) {
        // bias slice of this tile (double-buffered with the accumulator: the previous tile may still be read)
        for (int i = threadIdx.x - 64; i < p.block_n; i += 32 * kEpilogueWarps) s_bias[acc][i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(q * 32) << 16);
      bool released = false;
      for (int c0 = c_begin; c0 < c_end; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= c_end) {
          // all of this warp's TMEM reads of the buffer have completed: hand it back to the MMA thread BEFORE the
          // last slice's global stores, so the next-but-one tile's accumulation can start as early as possible
          tc_fence_before();
          if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
          released = true;
        }
        epilogue_chunk16(p, r, c0, n0, grow, row_ok, add_bias, s_bias[acc], lane);
      }
      if (!released) {               // (BLOCK_N % 32 != 0: the upper half-warps own no columns but still arrive)
        tc_fence_before();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
      }
    }
  }
  __syncwarp();                    // role branches leave warps 0/1 diverged; the cluster barrier is warp-aligned
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { if (CTAS == 2) tmem_dealloc_2cta(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols); }
}

#endif  // !DTF_HOST_EMU

// =================================================================================================
// host side
// =================================================================================================
#ifndef DTF_HOST_EMU
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

// 2-D bf16 / fp32 tensor map over a row-major [rows, cols] matrix with leading dimension ld (elements),
// box = {box_cols (x esize = 128 bytes), box_rows}, 128-byte swizzle, zero fill out of bounds.
// sw32 = 1: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks XOR row % 4) -- the only layout tcgen05 accepts for MN-major
// 32-bit operands (descriptor layout type SWIZZLE_128B_BASE32B).
static int make_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                    int box_rows, int esize, int sw32) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
// 4-D bf16 map over a dense NHWC activation: dims {C, W, H, images}, box {64 channels (128 B), bw, bh, bn}, 128-byte swizzle,
// zero fill outside the tensor.  A box lands in shared memory as [bn*bh*bw pixel rows] x 128 B -- the same image as the 2-D
// box {64, rows} of a materialised patch matrix, so the UMMA descriptors do not change.
static int make_map4(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2ull, (cuuint64_t)w * c * 2ull, (cuuint64_t)h * w * c * 2ull};
  cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
#endif  // !DTF_HOST_EMU (the emulation stubs provide make_map / make_map4: they record the requested tensor map)

struct MapKey {
  const void* ptr;
  long long rows, cols, ld;
  int bc, br, es, sw32;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, rows, cols, ld, bc, br, es, sw32) < std::tie(o.ptr, o.rows, o.cols, o.ld, o.bc, o.br, o.es, o.sw32);
  }
};
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;
static int cached_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int bc, int br, int es, int sw32);
static int cached_map4(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn);
// fp32 maps for the other translation units (csrc/mlp_step.cu)
int cached_map_f32(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                   int sw32) {
  return cached_map(out, ptr, rows, cols, ld, box_cols, box_rows, 4, sw32);
}

static int cached_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int bc, int br,
                      int es, int sw32) {
  MapKey k{ptr, rows, cols, ld, bc, br, es, sw32};
  std::lock_guard<std::mutex> g(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  int rc = make_map(out, ptr, rows, cols, ld, bc, br, es, sw32);
  if (rc == 0) {
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[k] = *out;
  }
  return rc;
}

// 4-D activation maps share the cache: the key packs the geometry into the 2-D fields (es = 0 marks a 4-D entry)
static int cached_map4(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn) {
  MapKey k{ptr, ((long long)n << 32) | (unsigned)h, ((long long)w << 32) | (unsigned)c, (long long)bn, bw, bh, 0, 0};
  std::lock_guard<std::mutex> g(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  int rc = make_map4(out, ptr, n, h, w, c, bw, bh, bn);
  if (rc == 0) {
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[k] = *out;
  }
  return rc;
}

}  // namespace dtf

extern "C" {

struct DtfGemmArgs {
  const void* a;         // bf16
  const void* b;         // bf16
  void* c;               // fp32 or bf16
  long long M, N, K;
  long long lda, ldb, ldc;
  int a_mn, b_mn;        // storage: a_mn=0 -> A[M,K] row-major, 1 -> A^T stored as [K,M]; same for B ([N,K] / [K,N])
  int c_bf16;
  const float* bias;
  int relu;
  const void* mask;      // bf16 [M, ldmask]
  long long ldmask;
  float alpha;
  int splits;            // >1: split-K with atomic accumulation into fp32 C (C must be zeroed, no relu)
  int accumulate;        // 1: atomically add into existing fp32 C
  float* colsum;
  const unsigned long long* wait_flag;
  unsigned long long wait_target;
  unsigned long long* signal;
  unsigned int* err;
  unsigned long long timeout_ns;
  int block_n_override;
  const unsigned long long* wait_target_ptr;
  long long* phase_trace;
  int signal_gpu_scope;
  const unsigned long long* stamp_src;
  unsigned long long* stamp_dst;
  int persistent;        // 0: auto (persistent kernel when tiles > SMs), 1: force persistent 1-CTA, 2: force CTA pairs, -1: never
  int cta_pair;          // -1: never use cta_group::2 in auto mode
  int tf32;              // 1: A and B are fp32 in memory, multiplied as TF32 (tcgen05.mma.kind::tf32); lda/ldb % 4 == 0
  // implicit-GEMM convolution: a = bf16 NHWC activation [cv_n, cv_h, cv_w, cv_c] (dense), M/K describe the patch matrix
  int conv;              // 0 plain, 1 A = patches (M = pixels, K = taps*C), 2 A = patches^T (M = taps*C, K = pixels)
  int cv_n, cv_h, cv_w, cv_c, cv_kh, cv_kw, cv_pt, cv_pl;
};

// Returns 0 on success, <0 for argument errors, >0 for CUDA errors.
int dtf_gemm_bf16(const DtfGemmArgs* g, cudaStream_t stream) {
  using namespace dtf;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return -2;
  const int es = g->tf32 ? 4 : 2;                                               // operand element size
  const int kbk = 128 / es;                                                     // K elements per stage (128-byte rows)
  if ((g->lda % (16 / es)) || (g->ldb % (16 / es))) return -3;                  // TMA: 16-byte row pitch
  if ((reinterpret_cast<uintptr_t>(g->a) & 15) || (reinterpret_cast<uintptr_t>(g->b) & 15)) return -4;
  if (g->splits > 1 && (g->relu || g->c_bf16)) return -5;
  int cv_bw = 0, cv_bh = 0, cv_bn = 0;
  if (g->conv) {
    // implicit-GEMM convolution: bf16, stride 1, whole 64-channel chunks, and a tile's pixels = one box of whole rows /
    // whole images (W | tile, tile | H*W or H*W | tile)
    const int tile = g->conv == 1 ? kBlockM : 64;
    const long long hw = (long long)g->cv_h * g->cv_w, pixels = (long long)g->cv_n * hw, taps = (long long)g->cv_kh * g->cv_kw;
    if (g->tf32 || (g->conv != 1 && g->conv != 2) || g->cv_c % 64 || g->cv_w > tile || tile % g->cv_w) return -8;
    if (!((hw % tile) == 0 || (tile % hw) == 0) || pixels % tile) return -8;
    if (g->conv == 1 && (g->a_mn || g->M != pixels || g->K != taps * g->cv_c)) return -8;
    if (g->conv == 2 && (!g->a_mn || g->K != pixels || g->M != taps * g->cv_c)) return -8;
    cv_bw = g->cv_w;
    cv_bh = (int)(hw >= tile ? tile / g->cv_w : g->cv_h);
    cv_bn = (int)(hw >= tile ? 1 : tile / hw);
    if (cv_bh > 256 || cv_bn > 256) return -8;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)g->M; p.N = (int)g->N; p.K = (int)g->K;
  p.a_mn = g->a_mn; p.b_mn = g->b_mn;
  int bn;
  if (g->block_n_override > 0) bn = g->block_n_override;
  else if (g->b_mn) bn = g->N <= 64 ? 64 : (g->N <= 128 ? 128 : (g->N <= 192 ? 192 : 256));
  else { bn = (int)((g->N + 15) / 16 * 16); if (bn > 256) bn = (g->N % 256 == 0 || g->N > 1024) ? 256 : 128; }
  if (g->b_mn && (bn % kbk)) return -6;                                         // whole 128-byte MN chunks
  if (bn % 16 || bn > 256 || bn < 16) return -6;
  p.block_n = bn;
  p.tf32 = g->tf32 ? 1 : 0; p.kbk = kbk; p.mn_lbo = kbk * 128; p.mn_kstep16 = ((g->tf32 ? 8 : 16) * 128) >> 4;
  p.mn_sbo = g->tf32 ? 512 : 1024; p.mn_type = g->tf32 ? 1 : 2;
  const int sw32 = g->tf32 ? 1 : 0;                                             // MN-major fp32 tiles: 32-byte-chunk swizzle
  p.num_kb = (int)((g->K + kbk - 1) / kbk);
  int splits = g->splits > 1 ? g->splits : 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  const int stage_bytes = kABytes + bn * 128;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (g->conv) {
    // short K loops (taps x chunks) under many tiles: two CTAs per SM, so one tile's epilogue runs under the other's mainloop
    p.conv = g->conv; p.cv_w = g->cv_w; p.cv_hw = g->cv_h * g->cv_w; p.cv_cpt = g->cv_c / 64;
    p.cv_kw = g->cv_kw; p.cv_taps = g->cv_kh * g->cv_kw; p.cv_pt = g->cv_pt; p.cv_pl = g->cv_pl; p.cv_n = g->cv_n;
    // ... when there ARE two CTAs' worth of tiles per SM; a grid of at most one CTA per SM keeps the deep pipeline (it is
    // TMA-latency bound: 3 stages in flight measured 30 us for 18 K blocks)
    const long long ctas_total = ((g->M + kBlockM - 1) / kBlockM) * ((g->N + bn - 1) / bn) * splits;
    const int two = (100 * 1024) / stage_bytes;
    if (ctas_total > 148 && two >= 3 && stages > two) stages = two;
  }
  if (stages > p.kb_per_split) stages = p.kb_per_split < 2 ? 2 : p.kb_per_split;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.c = g->c; p.ldc = g->ldc; p.c_bf16 = g->c_bf16;
  p.bias = g->bias; p.relu = g->relu;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(g->mask); p.ldmask = g->ldmask;
  p.alpha = g->alpha;
  p.atomic = (splits > 1 || g->accumulate) ? 1 : 0;
  p.colsum = g->colsum;
  p.wait_flag = g->wait_flag; p.wait_target = g->wait_target; p.wait_target_ptr = g->wait_target_ptr;
  p.signal = g->signal; p.err = g->err;
  p.timeout_ns = g->timeout_ns ? g->timeout_ns : 2000000000ull;
  p.phase_trace = g->phase_trace;
  p.signal_gpu_scope = g->signal_gpu_scope; p.stamp_src = g->stamp_src; p.stamp_dst = g->stamp_dst;

  CUtensorMap ma, mb;
  int rc;
  if (g->conv)       rc = cached_map4(&ma, g->a, g->cv_n, g->cv_h, g->cv_w, g->cv_c, cv_bw, cv_bh, cv_bn);
  else if (!g->a_mn) rc = cached_map(&ma, g->a, g->M, g->K, g->lda, kbk, kBlockM, es, 0);  // [M rows, K cols]
  else               rc = cached_map(&ma, g->a, g->K, g->M, g->lda, kbk, kbk, es, sw32);   // [K rows, M cols]
  if (rc) return rc < 0 ? -7 : 1000 + rc;
  if (!g->b_mn) rc = cached_map(&mb, g->b, g->N, g->K, g->ldb, kbk, bn, es, 0);       // [N rows, K cols]
  else          rc = cached_map(&mb, g->b, g->K, g->N, g->ldb, kbk, kbk, es, sw32);   // [K rows, N cols]
  if (rc) return rc < 0 ? -7 : 1000 + rc;

  const size_t smem = (size_t)stages * stage_bytes + 1024;
  // the opt-in shared-memory limit is a per-device function attribute (one process may drive all 8 GPUs)
  static bool configured[64] = {false};
  static int sm_count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
#ifndef DTF_HOST_EMU
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(224 * 1024));   // 227 KB minus static barriers + bias stage
    if (e != cudaSuccess) return 2000 + (int)e;
    e = cudaFuncSetAttribute(gemm_bf16_tcgen05_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(220 * 1024));               // more static shared memory (double-buffered bias)
    if (e != cudaSuccess) return 2000 + (int)e;
    e = cudaFuncSetAttribute(gemm_bf16_tcgen05_persistent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(220 * 1024));
    if (e != cudaSuccess) return 2000 + (int)e;
#endif
    cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    configured[dev] = true;
  }
  const unsigned tiles_m = (unsigned)((g->M + kBlockM - 1) / kBlockM), tiles_n = (unsigned)((g->N + bn - 1) / bn);
  const int sms = (dev >= 0 && dev < 64 && sm_count[dev] > 0) ? sm_count[dev] : 148;
  // Persistent path: plain GEMMs with more tiles than SMs (no split-K, no fused wait/signal, no phase stamps).
  // (implicit-GEMM fprop / dgrad with more tiles than SMs takes the persistent 1-CTA kernel too: per-tile prologue / epilogue of
  //  a 9-K-block tile is most of its time -- measured 26 us for 512 tiles of 128 x 64 x 576 in the tile kernel)
  const bool plain = splits == 1 && !p.atomic && p.wait_flag == nullptr && p.signal == nullptr && p.phase_trace == nullptr &&
                     g->conv != 2;
  if (plain && g->persistent >= 0 && (g->persistent > 0 || (long long)tiles_m * tiles_n > sms)) {
    // CTA pairs (cta_group::2, 256 x BLOCK_N tiles) when the tile shape allows it: BLOCK_N a multiple of 32 (each
    // CTA loads BLOCK_N/2 rows of B; 128 when B is MN-major) and at least two 128-row blocks of M.
    int ctas = 1;
    const bool pair_ok = (bn % 32 == 0) && (!g->b_mn || (bn / 2) % kbk == 0) && g->M > kBlockM && !g->conv;
    // measured (tools/gemm_perf.py): pairs win only with the widest tile (BLOCK_N = 256: 1292 vs 1144 TFLOP/s at 4096^3);
    // with narrower tiles the B half-tile is too small to matter and the 1-CTA kernel's finer tile granularity wins
    if (g->persistent == 2 || (g->persistent != 1 && pair_ok && g->cta_pair >= 0 && bn >= 192)) ctas = pair_ok ? 2 : 1;
    if (ctas == 2 && !g->b_mn) {
      // K-major B: the TMA box of one CTA covers only ITS half of the tile's N rows
      rc = cached_map(&mb, g->b, g->N, g->K, g->ldb, kbk, bn / 2, es, 0);
      if (rc) return rc < 0 ? -7 : 1000 + rc;
    }
    const int st_bytes = kABytes + (bn / ctas) * 128;
    int pst = (int)((216 * 1024 - 1024) / st_bytes);
    if (pst > 8) pst = 8;
    if (pst > p.num_kb) pst = p.num_kb < 2 ? 2 : p.num_kb;
    p.stages = pst;
    const size_t psmem = (size_t)pst * st_bytes + 1024;
    const unsigned tm = (unsigned)((g->M + kBlockM * ctas - 1) / (kBlockM * ctas));
    long long nt = (long long)tm * tiles_n;
    if (ctas == 2) {
      const unsigned pairs = (unsigned)(nt < sms / 2 ? nt : sms / 2);
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(2 * pairs, 1, 1);
      cfg.blockDim = dim3(kThreads, 1, 1);
      cfg.dynamicSmemBytes = psmem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
#ifdef DTF_HOST_EMU
      return dtf_emu_record_gemm_launch(2, cfg.gridDim, cfg.dynamicSmemBytes, (int)attr[0].val.clusterDim.x, ma, mb, p, (int)tm, (int)tiles_n);
#else
      cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_persistent_kernel<2>, ma, mb, p, (int)tm, (int)tiles_n);
      return (int)e;
#endif
    }
    const unsigned pgrid = (unsigned)(nt < sms ? nt : sms);
#ifdef DTF_HOST_EMU
    return dtf_emu_record_gemm_launch(1, dim3(pgrid, 1, 1), psmem, 1, ma, mb, p, (int)tm, (int)tiles_n);
#else
    gemm_bf16_tcgen05_persistent_kernel<1><<<pgrid, kThreads, psmem, stream>>>(ma, mb, p, (int)tm, (int)tiles_n);
    return (int)cudaGetLastError();
#endif
  }
  dim3 grid(tiles_m, tiles_n, (unsigned)splits);
#ifdef DTF_HOST_EMU
  return dtf_emu_record_gemm_launch(0, grid, smem, 1, ma, mb, p, (int)tiles_m, (int)tiles_n);
#else
  gemm_bf16_tcgen05_kernel<<<grid, kThreads, smem, stream>>>(ma, mb, p);
  return (int)cudaGetLastError();
#endif
}

void dtf_gemm_clear_map_cache() {
  std::lock_guard<std::mutex> g(dtf::g_maps_mu);
  dtf::g_maps.clear();
}

}  // extern "C"
