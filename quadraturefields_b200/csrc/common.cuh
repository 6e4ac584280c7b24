// Shared host/device helpers for libquadfield.so (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "quadfield.h"

namespace qf {

void set_error(const char* fmt, ...);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace qf

#define QF_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      qf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));        \
      return QF_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

#define QF_REQUIRE(cond, ...)                                                                    \
  do {                                                                                           \
    if (!(cond)) {                                                                               \
      qf::set_error(__VA_ARGS__);                                                                \
      return QF_ERR_INVALID;                                                                     \
    }                                                                                            \
  } while (0)

#define QF_LAUNCH_CHECK() QF_CUDA_CHECK(cudaGetLastError())
// Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute is per device, so it is set once per
// (call site, device); a bit mask per call site remembers which devices are done (thread-safe, devices 0..63).
#define QF_ENSURE_DYNAMIC_SMEM(kernel, bytes)                                                                      \
  do {                                                                                                             \
    static std::atomic<unsigned long long> qf_smem_done_{0};                                                       \
    int qf_dev_ = 0;                                                                                               \
    QF_CUDA_CHECK(cudaGetDevice(&qf_dev_));                                                                        \
    const unsigned long long qf_bit_ = 1ull << (qf_dev_ & 63);                                                     \
    if (!(qf_smem_done_.load(std::memory_order_acquire) & qf_bit_)) {                                              \
      QF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      qf_smem_done_.fetch_or(qf_bit_, std::memory_order_release);                                                  \
    }                                                                                                              \
  } while (0)

// ---- opaque handle layouts (shared between translation units) --------------------------------

struct qf_mesh {
  int64_t n_vertices = 0, n_faces = 0, n_nodes = 0;
  float* d_vertices = nullptr;    // (V,3) copy
  int32_t* d_faces = nullptr;     // (F,3) copy
  float4* d_tris = nullptr;       // Morton-sorted: 3 x float4 per triangle: (v0,id) (v1,0) (v2,0)
  float4* d_planes = nullptr;     // per ORIGINAL triangle id: (n.xyz from fp64, d=-(n.v0))
  float4* d_nodes = nullptr;      // 4 x float4 per node: both children's boxes + refs
  float4* d_wnodes = nullptr;     // wide BVH: 32 children x 2 float4 per node (traverse.cuh), collapsed from the binary tree
  int32_t* d_wqueue = nullptr;    // binary node behind every wide node (build scratch)
  int32_t* d_wstate = nullptr;    // [0] level begin, [1] level end, [2] wide nodes allocated, [3] ok flag, [4] level
  int64_t wide_capacity = 0;      // wide nodes that fit in d_wnodes
  float* d_scene = nullptr;       // [0..2] lo, [3..5] hi, [6] pad
  // build scratch (bvh.cu alloc_build_scratch / free_build_scratch)
  uint64_t *d_keys = nullptr, *d_keys_sorted = nullptr;
  uint32_t *d_idx = nullptr, *d_idx_sorted = nullptr;
  int32_t *d_left = nullptr, *d_right = nullptr, *d_parent = nullptr, *d_leaf_parent = nullptr;
  int32_t *d_first = nullptr, *d_last = nullptr, *d_flags = nullptr;
  float4* d_ibox = nullptr;       // 2 x float4 per internal node (lo, hi)
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  int32_t* d_call_slots = nullptr;  // ring of per-call scratch: [coherent chunks, total chunks, ray counter, pad]
  mutable std::atomic<unsigned> call_id{0};   // concurrent qf_trace_firstk calls on one mesh take distinct slots
  size_t bytes = 0;
  size_t scratch_bytes = 0;       // part of `bytes` that is build scratch (released after create, back on the first vertex update)
  bool want_wide = true;
  float h_pad = 0.f;
  float restart_eps = 0.f;        // > 0: keep hits like the reference's Embree restart loop (qf_mesh_set_restart_eps)
  // baked path: per-triangle barycentric setup (128 B: v0, e0, e1, d00, d01, d11, 1/den as doubles + the three uv pairs),
  // built lazily for the uv array of the last baked render and rebuilt when that pointer or the vertices change
  mutable double* d_bary = nullptr;
  mutable const float* bary_uv = nullptr;
  mutable unsigned bary_version = 0;
  unsigned geometry_version = 1;
};
constexpr int kCallSlots = 64;

struct qf_ngp {
  qf_grid_desc desc;
  int64_t n_entries = 0;
  __half2* d_table = nullptr;     // fp16 working copy, 2 features per entry
  __half* d_weights = nullptr;    // fp16 working copy of all five matrices, smem-ready layout (mma.sync kernels)
  unsigned char* d_weights_tc = nullptr;  // the same matrices as UMMA K-major canonical images (tcgen05 kernel)
};

struct qf_texture {
  int size = 0, num_lobes = 0, colour_logit = 0, record_bytes = 0;
  float lambda_thres = 7.5f;
  uint8_t* d_records = nullptr;   // interleaved texel records (see baked.cu)
  float* d_tables = nullptr;      // dequantiser lookup tables, 7 x 256 floats (see baked.cu)
};
