// Stand-ins for the CUDA runtime / driver pieces the HOST half of csrc/gemm_tcgen05.cu touches, so that dtf_gemm_bf16's
// dispatch logic (argument checks, BLOCK_N / stage / kernel selection, tensor-map boxes, grid, shared memory, cluster
// size) compiles with g++ and can be unit-tested: tensor maps and launches are RECORDED instead of encoded / issued.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __nv_bfloat16 {
  uint16_t v;
};
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttribute {
  cudaLaunchAttributeID id;
  struct {
    struct {
      unsigned x, y, z;
    } clusterDim;
  } val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes;
  cudaStream_t stream;
  cudaLaunchAttribute* attrs;
  unsigned numAttrs;
};

// what a tensor map was asked to describe (cuTensorMapEncodeTiled is not called)
struct CUtensorMap {
  const void* ptr;
  long long rows, cols, ld;
  int box_cols, box_rows;
};

struct DtfEmuGemmRecord {
  int kind;                  // 0 tile kernel, 1 persistent 1-CTA, 2 persistent CTA pairs
  unsigned gx, gy, gz;
  long long smem;
  int cluster;
  int block_n, stages, num_kb, kb_per_split, atomic;
  int tiles_m, tiles_n;      // tile kernel: grid tiles; persistent: tile counts handed to the kernel (M tiles of 128 x CTAS rows)
  long long a_rows, a_cols, a_ld;
  int a_box_cols, a_box_rows;
  long long b_rows, b_cols, b_ld;
  int b_box_cols, b_box_rows;
};

namespace dtf_emu_gemm {
inline DtfEmuGemmRecord last;
inline int sm_count = 148;
}  // namespace dtf_emu_gemm

static inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = dtf_emu_gemm::sm_count;
  return cudaSuccess;
}

namespace dtf {
static int make_map(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows, int /*esize*/ = 2, int /*sw32*/ = 0) {
  *out = CUtensorMap{ptr, rows, cols, ld, box_cols, box_rows};
  return 0;
}
static int make_map4(CUtensorMap* out, const void* ptr, int n, int h, int w, int c, int bw, int bh, int bn) {
  *out = CUtensorMap{ptr, (long long)n * h * w, (long long)c, (long long)c, 64, bw * bh * bn};     // pixels x channels, box = pixels of one tile
  return 0;
}
}  // namespace dtf

template <class P>
static int dtf_emu_record_gemm_launch(int kind, dim3 grid, size_t smem, int cluster, const CUtensorMap& ma, const CUtensorMap& mb,
                                      const P& p, int tm, int tn) {
  DtfEmuGemmRecord& r = dtf_emu_gemm::last;
  r = DtfEmuGemmRecord{kind, grid.x, grid.y, grid.z, (long long)smem, cluster, p.block_n, p.stages, p.num_kb, p.kb_per_split,
                       p.atomic, tm, tn, ma.rows, ma.cols, ma.ld, ma.box_cols, ma.box_rows, mb.rows, mb.cols, mb.ld, mb.box_cols,
                       mb.box_rows};
  return 0;
}

extern "C" __attribute__((weak)) void dtf_emu_gemm_last(DtfEmuGemmRecord* out) { *out = dtf_emu_gemm::last; }
extern "C" __attribute__((weak)) void dtf_emu_gemm_set_sm_count(int n) { dtf_emu_gemm::sm_count = n; }
