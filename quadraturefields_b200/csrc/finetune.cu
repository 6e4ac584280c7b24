// Mesh finetuning accumulators: the per-triangle / per-vertex scatters the reference does with
// torch_scatter atomics, kept on the device so the finetune loop needs no host round trip before the BVH refit.
//
//   triangle_accumulate_kernel : MeshFinetune.update_d   (mesh_utils.py:126-133)   cache_d[tri] += d*w, cache_w[tri] += w
//   vertex_scatter/apply       : MeshFinetune.update_faces (mesh_utils.py:135-144) clip(cache_d/cache_w, ±scaling) per
//                                triangle, mean over the face corners incident to each vertex, vertices += mean
//   triangle_weight_max_kernel : prune pass (prune_mesh_after_finetuning.py:354-356) tri_w[tri] = max(tri_w[tri], w)
#include "common.cuh"

namespace qf {

__global__ void triangle_accumulate_kernel(const float* __restrict__ d, const float* __restrict__ w,
                                           const int64_t* __restrict__ index_tri, int64_t M, int64_t F,
                                           float* __restrict__ cache_d, float* __restrict__ cache_w) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t f = index_tri[i];
  if (f < 0 || f >= F) return;
  const float wi = w[i];
  atomicAdd(cache_d + 3 * f, d[3 * i] * wi);
  atomicAdd(cache_d + 3 * f + 1, d[3 * i + 1] * wi);
  atomicAdd(cache_d + 3 * f + 2, d[3 * i + 2] * wi);
  atomicAdd(cache_w + f, wi);
}

// scratch (V,4): xyz = sum of the clipped deformation of incident face corners, w = corner count
__global__ void vertex_scatter_kernel(const float* __restrict__ cache_d, const float* __restrict__ cache_w,
                                      const int32_t* __restrict__ faces, int64_t F, int64_t V, float scaling,
                                      float* __restrict__ scratch) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float w = cache_w[f];
  float dx = fminf(fmaxf(__fdiv_rn(cache_d[3 * f], w), -scaling), scaling);
  float dy = fminf(fmaxf(__fdiv_rn(cache_d[3 * f + 1], w), -scaling), scaling);
  float dz = fminf(fmaxf(__fdiv_rn(cache_d[3 * f + 2], w), -scaling), scaling);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t v = faces[3 * f + k];
    if (v < 0 || v >= V) continue;
    atomicAdd(scratch + 4 * v, dx);
    atomicAdd(scratch + 4 * v + 1, dy);
    atomicAdd(scratch + 4 * v + 2, dz);
    atomicAdd(scratch + 4 * v + 3, 1.0f);
  }
}

__global__ void vertex_apply_kernel(const float* __restrict__ scratch, int64_t V, float* __restrict__ vertices) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float4 s = reinterpret_cast<const float4*>(scratch)[v];
  const float cnt = fmaxf(s.w, 1.0f);   // scatter_mean clamps the count to >= 1
  vertices[3 * v] += __fdiv_rn(s.x, cnt);
  vertices[3 * v + 1] += __fdiv_rn(s.y, cnt);
  vertices[3 * v + 2] += __fdiv_rn(s.z, cnt);
}

__global__ void triangle_weight_max_kernel(const float* __restrict__ w, int64_t stride, const int64_t* __restrict__ index_tri,
                                           int64_t M, int64_t F, float* __restrict__ tri_w) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t f = index_tri[i];
  if (f < 0 || f >= F) return;
  const float wi = w[i * stride];
  // the running maximum starts at 0 (zeros + scatter_max + torch.maximum), so only positive weights can raise it;
  // positive floats order like their bit patterns
  if (wi > 0.f) atomicMax(reinterpret_cast<int*>(tri_w) + f, __float_as_int(wi));
}

}  // namespace qf

using namespace qf;

extern "C" int qf_triangle_accumulate(const float* d_disp, const float* d_w, const int64_t* d_index_tri, int64_t M,
                                      int64_t n_faces, float* d_cache_d, float* d_cache_w, void* stream) {
  QF_REQUIRE(M >= 0 && n_faces >= 0, "qf_triangle_accumulate: negative size");
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_disp && d_w && d_index_tri && d_cache_d && d_cache_w, "qf_triangle_accumulate: NULL argument");
  triangle_accumulate_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d_disp, d_w, d_index_tri, M, n_faces,
                                                                                           d_cache_d, d_cache_w);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" size_t qf_vertex_displace_workspace_bytes(int64_t n_vertices) { return (size_t)(n_vertices > 0 ? n_vertices : 1) * 16; }

extern "C" int qf_vertex_displace(const float* d_cache_d, const float* d_cache_w, const int32_t* d_faces, int64_t n_faces,
                                  int64_t n_vertices, float scaling, float* d_vertices, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
  QF_REQUIRE(n_faces >= 0 && n_vertices >= 0, "qf_vertex_displace: negative size");
  if (n_faces == 0 || n_vertices == 0) return QF_OK;
  QF_REQUIRE(d_cache_d && d_cache_w && d_faces && d_vertices && d_workspace, "qf_vertex_displace: NULL argument");
  QF_REQUIRE(workspace_bytes >= qf_vertex_displace_workspace_bytes(n_vertices), "qf_vertex_displace: workspace %zu < %zu bytes",
             workspace_bytes, qf_vertex_displace_workspace_bytes(n_vertices));
  QF_REQUIRE(((uintptr_t)d_workspace & 15) == 0, "qf_vertex_displace: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* scratch = (float*)d_workspace;
  QF_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)n_vertices * 16, st));
  vertex_scatter_kernel<<<(unsigned)ceil_div(n_faces, 256), 256, 0, st>>>(d_cache_d, d_cache_w, d_faces, n_faces, n_vertices, scaling, scratch);
  QF_LAUNCH_CHECK();
  vertex_apply_kernel<<<(unsigned)ceil_div(n_vertices, 256), 256, 0, st>>>(scratch, n_vertices, d_vertices);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_triangle_weight_max(const float* d_weights, int64_t stride, const int64_t* d_index_tri, int64_t M,
                                      int64_t n_faces, float* d_tri_w, void* stream) {
  QF_REQUIRE(M >= 0 && n_faces >= 0 && stride >= 1, "qf_triangle_weight_max: bad size");
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_weights && d_index_tri && d_tri_w, "qf_triangle_weight_max: NULL argument");
  triangle_weight_max_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d_weights, stride, d_index_tri, M, n_faces, d_tri_w);
  QF_LAUNCH_CHECK();
  return QF_OK;
}
