// (5) Compositing: segmented transmittance scans.
//
//  * derive_properties_kernel — the mesh path (utils.py:863-898): one thread per ray walks its <= K
//    depth-sorted hit samples once and produces colour, opacity, depth and the per-sample weights;
//    replaces 2 kaolin pack scans + 3 pack reductions + 3 scatters.
//  * render_weights_kernel — the nerfacc surface behind field_rendering.py (exclusive_prod/exclusive_sum,
//    :203,:261): one warp per ray, 32 samples per step, warp-shuffle inclusive scan with a running carry,
//    so arbitrarily long ray segments (volumetric samples) are handled; its reverse twin does the backward.
//  * accumulate kernels — field_rendering.py:483-573, deterministic segment order when packed_info is given.
// All are HBM-streaming kernels: every sample is read once and written once.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace qf {

__global__ void derive_properties_kernel(const float* __restrict__ color, const float* __restrict__ density,
                                         const float* __restrict__ depths, float delta,
                                         const int64_t* __restrict__ offsets, int64_t N, int bg_mode,
                                         const float* __restrict__ bkgd, float* __restrict__ rgb,
                                         float* __restrict__ alpha_out, float* __restrict__ depth_out,
                                         float* __restrict__ weights) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t s = offsets[i], e = offsets[i + 1];
  float fill = bg_mode == QF_BG_BLACK ? 0.f : 1.f;  // quirk Q2
  float r = fill, g = fill, b = fill, A = 0.f, D = 0.f;
  if (e > s) {
    float cum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    for (int64_t j = s; j < e; ++j) {
      float tau = density[j] * delta;
      float w = expf(-cum) * (1.0f - expf(-tau));
      cum += tau;
      cr += w * color[3 * j]; cg += w * color[3 * j + 1]; cb += w * color[3 * j + 2];
      D += w * depths[j];
      A += w;
      if (weights) weights[j] = w;
    }
    if (bg_mode == QF_BG_WHITE) { r = (1.f - A) + A * cr; g = (1.f - A) + A * cg; b = (1.f - A) + A * cb; }        // quirk Q1
    else if (bg_mode == QF_BG_BLACK) { r = A * cr; g = A * cg; b = A * cb; }
    else { r = A * cr + (1.f - A) * bkgd[0]; g = A * cg + (1.f - A) * bkgd[1]; b = A * cb + (1.f - A) * bkgd[2]; }
  }
  rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
  alpha_out[i] = A;
  depth_out[i] = D;
}

// backward of derive_properties w.r.t. color (M,3) and density (M): thread per ray, two sweeps over its <= K samples.
//   out_c = bgmix(A, C_c),  A = sum w,  C_c = sum w c_c,  Depth = sum w t,  w_i = exp(-sum_{j<i} tau_j)(1-exp(-tau_i))
//   dL/dw_i = sum_c g_c (dOut_c/dA + dOut_c/dC_c c_ic) + g_A + g_D t_i ;  dL/dtau_k = gw_k T_k e^{-tau_k} - sum_{i>k} gw_i w_i
__global__ void derive_properties_bwd_kernel(const float* __restrict__ color, const float* __restrict__ density,
                                             const float* __restrict__ depths, float delta,
                                             const int64_t* __restrict__ offsets, int64_t N, int bg_mode,
                                             const float* __restrict__ bkgd, const float* __restrict__ g_rgb,
                                             const float* __restrict__ g_alpha, const float* __restrict__ g_depth,
                                             float* __restrict__ g_color, float* __restrict__ g_density) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t s = offsets[i], e = offsets[i + 1];
  if (e <= s) return;
  float cum = 0.f, C[3] = {0.f, 0.f, 0.f}, A = 0.f;
  for (int64_t j = s; j < e; ++j) {
    float tau = density[j] * delta;
    float w = expf(-cum) * (1.0f - expf(-tau));
    cum += tau;
    C[0] += w * color[3 * j]; C[1] += w * color[3 * j + 1]; C[2] += w * color[3 * j + 2];
    A += w;
  }
  const float g[3] = {g_rgb ? g_rgb[3 * i] : 0.f, g_rgb ? g_rgb[3 * i + 1] : 0.f, g_rgb ? g_rgb[3 * i + 2] : 0.f};
  const float gA = g_alpha ? g_alpha[i] : 0.f, gD = g_depth ? g_depth[i] : 0.f;
  // out_c = (1-A) + A C_c (white) | A C_c (black) | A C_c + (1-A) b_c (random)
  float dA = gA, dC[3];
  for (int c = 0; c < 3; ++c) {
    float b = bg_mode == QF_BG_WHITE ? 1.f : (bg_mode == QF_BG_BLACK ? 0.f : bkgd[c]);
    dA += g[c] * (C[c] - b);
    dC[c] = g[c] * A;
  }
  // reverse sweep: suffix = sum_{i>k} gw_i w_i
  float suffix = 0.f;
  for (int64_t j = e - 1; j >= s; --j) {
    float tau = density[j] * delta;
    cum -= tau;                                  // exclusive prefix of sample j
    float T = expf(-cum), ea = expf(-tau);
    float w = T * (1.0f - ea);
    float gw = dA + gD * depths[j] + dC[0] * color[3 * j] + dC[1] * color[3 * j + 1] + dC[2] * color[3 * j + 2];
    if (g_color) { g_color[3 * j] = dC[0] * w; g_color[3 * j + 1] = dC[1] * w; g_color[3 * j + 2] = dC[2] * w; }
    if (g_density) g_density[j] = delta * (gw * T * ea - suffix);
    suffix += gw * w;
  }
}

// ---------------------------------------------------------------- nerfacc-style scans
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}
__device__ __forceinline__ float warp_suffix_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

template <int MODE>  // 0: alphas ; 1: sigmas * (t_ends - t_starts)
__global__ void render_weights_kernel(const float* __restrict__ in, const float* __restrict__ ts, const float* __restrict__ te,
                                      const int64_t* __restrict__ packed, int64_t n_rays, const float* __restrict__ prefix,
                                      float* __restrict__ w_out, float* __restrict__ T_out, float* __restrict__ a_out) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const int64_t start = packed[2 * ray], cnt = packed[2 * ray + 1];
  float carry = MODE == 0 ? 1.f : 0.f;
  for (int64_t b = 0; b < cnt; b += 32) {
    const int64_t j = start + b + lane;
    const bool ok = b + lane < cnt;
    float x = 0.f, alpha = 0.f;
    if (ok) {
      x = in[j];
      if (MODE == 1) { x = x * (te[j] - ts[j]); alpha = 1.0f - expf(-x); }
      else alpha = x;
    }
    float T;
    if (MODE == 0) {
      float f = ok ? 1.0f - x : 1.f;
      float inc = warp_incl_prod(f, lane);
      float exc = __shfl_up_sync(0xffffffffu, inc, 1);
      T = carry * (lane == 0 ? 1.f : exc);
      carry *= __shfl_sync(0xffffffffu, inc, 31);
    } else {
      float inc = warp_incl_sum(x, lane);
      float exc = carry + (inc - x);
      T = expf(-exc);
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (ok) {
      if (prefix) T *= prefix[j];
      if (T_out) T_out[j] = T;
      if (w_out) w_out[j] = T * alpha;
      if (a_out) a_out[j] = alpha;
    }
  }
}

// d in / d(weights, trans):  g_k = gw_k * dW_k/dx_k|direct - (1/f_k) * sum_{i>k} (gw_i w_i + gT_i T_i)
template <int MODE>
__global__ void render_weights_bwd_kernel(const float* __restrict__ in, const float* __restrict__ ts,
                                          const float* __restrict__ te, const int64_t* __restrict__ packed,
                                          int64_t n_rays, const float* __restrict__ prefix, const float* __restrict__ gw,
                                          const float* __restrict__ gT, float* __restrict__ gin) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const int64_t start = packed[2 * ray], cnt = packed[2 * ray + 1];
  if (cnt == 0) return;
  // pass 1 (forward): per-chunk carries are recomputed on the fly; pass 2 (reverse) needs T_j, so we
  // first compute the total and then walk chunks backwards dividing the carry out again would be
  // ill-conditioned — instead recompute the prefix of each chunk start by a forward sweep.
  const int64_t n_chunks = (cnt + 31) / 32;
  float suffix = 0.f;  // sum over samples after the current chunk of (gw_i w_i + gT_i T_i)
  for (int64_t c = n_chunks - 1; c >= 0; --c) {
    // forward sweep to the start of chunk c (cnt is small on the mesh path; volumetric rays have a few hundred samples)
    float carry = MODE == 0 ? 1.f : 0.f;
    for (int64_t b = 0; b < c * 32; b += 32) {
      const int64_t j = start + b + lane;
      float x = in[j];
      if (MODE == 1) x = x * (te[j] - ts[j]);
      if (MODE == 0) { float inc = warp_incl_prod(1.0f - x, lane); carry *= __shfl_sync(0xffffffffu, inc, 31); }
      else { float inc = warp_incl_sum(x, lane); carry += __shfl_sync(0xffffffffu, inc, 31); }
    }
    const int64_t b = c * 32;
    const int64_t j = start + b + lane;
    const bool ok = b + lane < cnt;
    float x = 0.f, alpha = 0.f, dt = 0.f;
    if (ok) {
      x = in[j];
      if (MODE == 1) { dt = te[j] - ts[j]; x = x * dt; alpha = 1.0f - expf(-x); }
      else alpha = x;
    }
    float T;
    if (MODE == 0) {
      float f = ok ? 1.0f - x : 1.f;
      float inc = warp_incl_prod(f, lane);
      float exc = __shfl_up_sync(0xffffffffu, inc, 1);
      T = carry * (lane == 0 ? 1.f : exc);
    } else {
      float inc = warp_incl_sum(x, lane);
      T = expf(-(carry + (inc - x)));
    }
    if (ok && prefix) T *= prefix[j];
    float gwj = (ok && gw) ? gw[j] : 0.f, gTj = (ok && gT) ? gT[j] : 0.f;
    float term = ok ? gwj * T * alpha + gTj * T : 0.f;
    float suf_incl = warp_suffix_sum(term, lane);
    float after = suffix + (suf_incl - term);  // sum over samples strictly after j
    if (ok) {
      if (MODE == 0) gin[j] = gwj * T - after / fmaxf(1.0f - alpha, 1e-10f);
      else gin[j] = dt * (gwj * T * (1.0f - alpha) - after);
    }
    suffix += __shfl_sync(0xffffffffu, suf_incl, 0);
  }
}

__global__ void accumulate_packed_kernel(const float* __restrict__ w, const float* __restrict__ v, int D,
                                         const int64_t* __restrict__ packed, int64_t n_rays, float* __restrict__ out,
                                         int accumulate_into) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const int64_t start = packed[2 * ray], cnt = packed[2 * ray + 1];
  for (int d = 0; d < D; ++d) {
    float s = 0.f;
    for (int64_t b = lane; b < cnt; b += 32) {
      const int64_t j = start + b;
      s += v ? w[j] * v[j * D + d] : w[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[ray * D + d] = accumulate_into ? out[ray * D + d] + s : s;
  }
}

__global__ void accumulate_indexed_kernel(const float* __restrict__ w, const float* __restrict__ v, int D,
                                          const int64_t* __restrict__ ray_indices, int64_t M, float* __restrict__ out) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int64_t r = ray_indices[j];
  const float wj = w[j];
  for (int d = 0; d < D; ++d) atomicAdd(out + r * D + d, v ? wj * v[j * D + d] : wj);
}

// grad of out[ray] = sum w v:  gw_j = sum_d gout[ray_j,d] v_jd ; gv_jd = w_j gout[ray_j,d]
__global__ void accumulate_bwd_kernel(const float* __restrict__ w, const float* __restrict__ v, int D,
                                      const int64_t* __restrict__ ray_indices, int64_t M, const float* __restrict__ gout,
                                      float* __restrict__ gw, float* __restrict__ gv) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int64_t r = ray_indices[j];
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    float go = gout[r * D + d];
    s += v ? go * v[j * D + d] : go;
    if (gv) gv[j * D + d] = w[j] * go;
  }
  if (gw) gw[j] = s;
}

__global__ void count_kernel(const int64_t* __restrict__ ray_indices, int64_t M, unsigned long long* __restrict__ cnt) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < M) atomicAdd(cnt + ray_indices[j], 1ull);
}
__global__ void interleave_kernel(const int64_t* __restrict__ starts, const int64_t* __restrict__ cnt, int64_t n,
                                  int64_t* __restrict__ packed) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { packed[2 * i] = starts[i]; packed[2 * i + 1] = cnt[i]; }
}

}  // namespace qf

using namespace qf;

extern "C" int qf_derive_properties(const float* d_color, const float* d_density, const float* d_depths, float delta,
                                    const int64_t* d_offsets, int64_t n_rays, int bg_mode, const float* d_bkgd,
                                    float* d_rgb, float* d_alpha, float* d_depth_out, float* d_weights, void* stream) {
  QF_REQUIRE(d_offsets && d_rgb && d_alpha && d_depth_out, "qf_derive_properties: NULL argument");
  QF_REQUIRE(bg_mode >= 0 && bg_mode <= 2, "qf_derive_properties: bg_mode=%d", bg_mode);
  QF_REQUIRE(bg_mode != QF_BG_RANDOM || d_bkgd, "qf_derive_properties: bg 'random' needs render_bkgd");
  if (n_rays == 0) return QF_OK;
  derive_properties_kernel<<<(int)ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
      d_color, d_density, d_depths, delta, d_offsets, n_rays, bg_mode, d_bkgd, d_rgb, d_alpha, d_depth_out, d_weights);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_derive_properties_backward(const float* d_color, const float* d_density, const float* d_depths, float delta,
                                             const int64_t* d_offsets, int64_t n_rays, int64_t n_hits, int bg_mode,
                                             const float* d_bkgd, const float* d_grad_rgb, const float* d_grad_alpha,
                                             const float* d_grad_depth, float* d_grad_color, float* d_grad_density,
                                             void* stream) {
  QF_REQUIRE(bg_mode >= 0 && bg_mode <= 2, "qf_derive_properties_backward: bg_mode=%d", bg_mode);
  if (n_rays == 0 || n_hits == 0) return QF_OK;
  QF_REQUIRE(d_color && d_density && d_depths && d_offsets, "qf_derive_properties_backward: NULL argument");
  QF_REQUIRE(bg_mode != QF_BG_RANDOM || d_bkgd, "qf_derive_properties_backward: bg 'random' needs render_bkgd");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_grad_color) QF_CUDA_CHECK(cudaMemsetAsync(d_grad_color, 0, sizeof(float) * 3 * n_hits, st));
  if (d_grad_density) QF_CUDA_CHECK(cudaMemsetAsync(d_grad_density, 0, sizeof(float) * n_hits, st));
  derive_properties_bwd_kernel<<<(int)ceil_div(n_rays, 256), 256, 0, st>>>(d_color, d_density, d_depths, delta, d_offsets, n_rays,
                                                                             bg_mode, d_bkgd, d_grad_rgb, d_grad_alpha, d_grad_depth,
                                                                             d_grad_color, d_grad_density);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_render_weights(int mode, const float* d_in, const float* d_t_starts, const float* d_t_ends,
                                 const int64_t* d_packed_info, int64_t n_rays, int64_t n_samples,
                                 const float* d_prefix_trans, float* d_weights, float* d_trans, float* d_alphas_out,
                                 void* stream) {
  QF_REQUIRE(mode == 0 || mode == 1, "qf_render_weights: mode=%d", mode);
  if (n_rays == 0 || n_samples == 0) return QF_OK;
  QF_REQUIRE(d_in && d_packed_info, "qf_render_weights: NULL argument");
  QF_REQUIRE(mode == 0 || (d_t_starts && d_t_ends), "qf_render_weights: density mode needs t_starts/t_ends");
  int blocks = (int)ceil_div(n_rays * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) render_weights_kernel<0><<<blocks, 256, 0, st>>>(d_in, d_t_starts, d_t_ends, d_packed_info, n_rays, d_prefix_trans, d_weights, d_trans, d_alphas_out);
  else render_weights_kernel<1><<<blocks, 256, 0, st>>>(d_in, d_t_starts, d_t_ends, d_packed_info, n_rays, d_prefix_trans, d_weights, d_trans, d_alphas_out);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_render_weights_backward(int mode, const float* d_in, const float* d_t_starts, const float* d_t_ends,
                                          const int64_t* d_packed_info, int64_t n_rays, int64_t n_samples,
                                          const float* d_prefix_trans, const float* d_grad_weights,
                                          const float* d_grad_trans, float* d_grad_in, void* stream) {
  QF_REQUIRE(mode == 0 || mode == 1, "qf_render_weights_backward: mode=%d", mode);
  if (n_rays == 0 || n_samples == 0) return QF_OK;
  QF_REQUIRE(d_in && d_packed_info && d_grad_in, "qf_render_weights_backward: NULL argument");
  int blocks = (int)ceil_div(n_rays * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  QF_CUDA_CHECK(cudaMemsetAsync(d_grad_in, 0, sizeof(float) * n_samples, st));
  if (mode == 0) render_weights_bwd_kernel<0><<<blocks, 256, 0, st>>>(d_in, d_t_starts, d_t_ends, d_packed_info, n_rays, d_prefix_trans, d_grad_weights, d_grad_trans, d_grad_in);
  else render_weights_bwd_kernel<1><<<blocks, 256, 0, st>>>(d_in, d_t_starts, d_t_ends, d_packed_info, n_rays, d_prefix_trans, d_grad_weights, d_grad_trans, d_grad_in);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_accumulate_along_rays(const float* d_weights, const float* d_values, int D,
                                        const int64_t* d_packed_info, int64_t n_rays, float* d_out, int accumulate_into,
                                        void* stream) {
  QF_REQUIRE(d_weights && d_packed_info && d_out && D >= 1, "qf_accumulate_along_rays: bad argument");
  if (n_rays == 0) return QF_OK;
  accumulate_packed_kernel<<<(int)ceil_div(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      d_weights, d_values, D, d_packed_info, n_rays, d_out, accumulate_into);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_accumulate_along_rays_indexed(const float* d_weights, const float* d_values, int D,
                                                const int64_t* d_ray_indices, int64_t n_samples, float* d_out,
                                                void* stream) {
  if (n_samples == 0) return QF_OK;
  QF_REQUIRE(d_weights && d_ray_indices && d_out && D >= 1, "qf_accumulate_along_rays_indexed: bad argument");
  accumulate_indexed_kernel<<<(int)ceil_div(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(d_weights, d_values, D, d_ray_indices, n_samples, d_out);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_accumulate_along_rays_backward(const float* d_weights, const float* d_values, int D,
                                                 const int64_t* d_ray_indices, int64_t n_samples, const float* d_grad_out,
                                                 float* d_grad_weights, float* d_grad_values, void* stream) {
  if (n_samples == 0) return QF_OK;
  QF_REQUIRE(d_weights && d_ray_indices && d_grad_out && D >= 1, "qf_accumulate_along_rays_backward: bad argument");
  accumulate_bwd_kernel<<<(int)ceil_div(n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
      d_weights, d_values, D, d_ray_indices, n_samples, d_grad_out, d_grad_weights, d_grad_values);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_pack_info(const int64_t* d_ray_indices, int64_t n_samples, int64_t n_rays, int64_t* d_packed_info,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
  QF_REQUIRE(d_packed_info && d_workspace, "qf_pack_info: NULL argument");
  if (n_rays == 0) return QF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  size_t arr = (sizeof(int64_t) * (size_t)n_rays + 255) / 256 * 256;
  QF_REQUIRE(workspace_bytes >= 2 * arr + 1024, "qf_pack_info: workspace too small");
  int64_t* cnt = (int64_t*)d_workspace;
  int64_t* starts = (int64_t*)((char*)d_workspace + arr);
  void* tmp = (char*)d_workspace + 2 * arr;
  size_t tmp_bytes = workspace_bytes - 2 * arr;
  QF_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int64_t) * n_rays, st));
  if (n_samples > 0) count_kernel<<<(int)ceil_div(n_samples, 256), 256, 0, st>>>(d_ray_indices, n_samples, (unsigned long long*)cnt);
  QF_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, starts, n_rays, st));
  interleave_kernel<<<(int)ceil_div(n_rays, 256), 256, 0, st>>>(starts, cnt, n_rays, d_packed_info);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" size_t qf_pack_info_workspace_bytes(int64_t n_rays) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int64_t*)nullptr, (int64_t*)nullptr, n_rays);
  size_t arr = (sizeof(int64_t) * (size_t)n_rays + 255) / 256 * 256;
  return 2 * arr + tmp + 1024;
}
