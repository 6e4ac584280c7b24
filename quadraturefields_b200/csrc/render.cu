// Fused frame render: rays -> BVH first-K hits -> compact hit records -> field / baked shading ->
// per-ray composite, all resident in HBM with no host synchronisation and no Python chunk loops
// (the reference runs utils.py:465-607 / :998-1095 with a CPU intersector, a GPU->CPU lexsort and
// 160 000-sample batches).
//
//   trace_compact_kernel : one thread per ray; traversal (traverse.cuh), plane-hit point + depth per hit
//                          (mesh_utils.py:33-40,371), stable depth order (:375), then the CTA reserves a
//                          contiguous run of the compact hit array with ONE atomicAdd.  Inside a warp's run the
//                          records are slot-major: all first hits of the 32 rays, then all second hits, ... so
//                          the 32 samples of a shading warp are same-depth-order hits of neighbouring pixels and
//                          share hash-grid cells (fused field kernel -8 % at c2, -22 % at c4 against ray-major).
//   shading              : ngp_forward_kernel (field.cu) or baked_shade_kernel (baked.cu) over the M live
//                          samples only (M is read from device memory; persistent grid of 148 x k CTAs).
//   composite_rays_kernel: derive_properties (utils.py:863-898) per ray.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "traverse.cuh"

namespace qf {

struct FieldArgs;  // field.cu
int launch_ngp_forward_hits(const qf_ngp* f, const float4* hit_pd, const int2* hit_rt, const float* d_viewdirs,
                            const int32_t* d_M, float4* out4, cudaStream_t st);
int launch_baked_shade(const qf_texture* tex, const qf_mesh* mesh, const float* d_uv, const float4* hit_pd,
                       const int2* hit_rt, const float* d_viewdirs, const int32_t* d_M, float4* out4, cudaStream_t st);

struct Workspace {
  int32_t* cursor;     // [0] live hit samples of the current chunk
  int32_t* ray_start;  // (CH) first record of the ray's WARP run (slot-major inside the run, see warp_slot_layout)
  int32_t* ray_count;  // (CH)
  float4* hit_pd;      // (CH*K) psi.xyz, depth
  int2* hit_rt;        // (CH*K) ray id (global), triangle id
  float4* hit_out;     // (CH*K) rgb, sigma
};

constexpr int64_t kChunkRays = 1ll << 21;

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static size_t workspace_layout(int64_t n_rays, int K, Workspace* w, char* base) {
  int64_t ch = n_rays < kChunkRays ? n_rays : kChunkRays;
  if (ch < 1) ch = 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  size_t o_cursor = take(256), o_start = take(sizeof(int32_t) * ch), o_count = take(sizeof(int32_t) * ch);
  size_t o_pd = take(sizeof(float4) * ch * K), o_rt = take(sizeof(int2) * ch * K), o_out = take(sizeof(float4) * ch * K);
  if (w) {
    w->cursor = (int32_t*)(base + o_cursor); w->ray_start = (int32_t*)(base + o_start); w->ray_count = (int32_t*)(base + o_count);
    w->hit_pd = (float4*)(base + o_pd); w->hit_rt = (int2*)(base + o_rt); w->hit_out = (float4*)(base + o_out);
  }
  return off;
}

// Ray handled by a thread: linear, or an 8x4 pixel tile per warp for image-ordered rays (tighter packets, and the
// compacted hit samples of a warp stay close in space for the gather-bound shading kernel).  trace and composite
// must agree on it: the slot-major record layout is defined per warp.
// With n_bands > 0 (the chunk is n_bands whole 4-row bands) the bands are taken CENTRE-OUT: warp order mid, mid-1, mid+1, ...
// The deepest packets of an object-centred frame (NeRF-synthetic, Shelly: every camera looks at the origin) then start
// first and the cheap background rows fill the tail of the launch — the longest packet no longer ends the kernel.
#ifndef QF_CENTER_OUT
#define QF_CENTER_OUT 1
#endif
__device__ __forceinline__ int64_t ray_of_thread(int64_t linear, int lane, int img_w, int n_bands) {
  if (img_w <= 0) return linear;
  const int64_t gw = linear >> 5;
  const int tiles = img_w >> 3;
  int64_t band = gw / tiles;
#if QF_CENTER_OUT
  if (n_bands > 0 && band < n_bands) {
    const int64_t mid = n_bands >> 1;
    band = (band & 1) ? mid - ((band + 1) >> 1) : mid + (band >> 1);
  }
#endif
  return (band * 4 + (lane >> 3)) * img_w + (gw % tiles) * 8 + (lane & 7);
}

// K <= 8 hit buffer of the fused frame: 8 shared-memory slots per ray (8 KB per CTA) — O(1) append, sorted once on output.
// r2i A/B on the c2 frame: trace 0.1139 ms with HitBufReg<8> (sorted registers, carry insertion), 0.1172 ms with an unsorted
// register buffer + sorting network, 0.1043 ms with this one.
#ifndef QF_TRACE_K8_SMEM
#define QF_TRACE_K8_SMEM 1
#endif
#if QF_TRACE_K8_SMEM
using HitBufK8 = HitBufSmemT<8>;
#else
using HitBufK8 = HitBufReg<8>;
#endif
#ifndef QF_TRACE_MIN_CTAS
#define QF_TRACE_MIN_CTAS 8   // 64 registers: c2 trace 0.124 -> 0.114 ms (r2 A/B); the K=32 variant is shared-memory limited anyway
#endif
template <class HB>
__global__ void __launch_bounds__(128, QF_TRACE_MIN_CTAS) trace_compact_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                            const float4* __restrict__ planes, const float* __restrict__ scene,
                                                            const float* __restrict__ origins, const float* __restrict__ dirs,
                                                            int64_t ray0, int64_t n, int K, int img_w, int mode,
                                                            int32_t* __restrict__ cursor,
                                                            int32_t* __restrict__ ray_start, int32_t* __restrict__ ray_count,
                                                            float4* __restrict__ hit_pd, int2* __restrict__ hit_rt,
                                                            const float4* __restrict__ wnodes, const int32_t* __restrict__ wstate,
                                                            int k_trav, float restart_eps, int n_bands) {
  __shared__ int s_stack[4][kWideStack];
  if (wnodes && !__ldg(wstate + 3)) wnodes = nullptr;    // the collapse gave up on this mesh: binary tree only
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t li = ray_of_thread(blockIdx.x * (int64_t)blockDim.x + tid, lane, img_w, n_bands);  // ray inside the chunk
  __shared__ float s_ht[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  __shared__ int s_hi[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  const bool valid = li < n;
  HB hb(s_ht, s_hi, tid);
  int total = 0;
  Ray r = make_ray(origins, dirs, ray0 + (valid ? li : n - 1));
  trace_ray<HB>(r, valid, nodes, wnodes, tris, k_trav, hb, total, s_stack[warp], mode);
  if (restart_eps > 0.f && valid) hb.restart_filter(restart_eps, K);   // k_trav = QF_MAX_HITS raw hits -> K kept (Embree restart loop)
  // the warp reserves a contiguous run of the compact hit array with ONE atomicAdd (no CTA barrier: packets differ a lot in
  // cost, and a __syncthreads here made every warp wait for the slowest of its CTA — 15 % of the stall samples in r2f)
  // the traversal is over: its stack now holds the slot tables of the warp's run (records before slot j | lanes with a j-th hit)
  static_assert(kWideStack >= 2 * QF_MAX_HITS, "slot tables alias the packet stack");
  __syncwarp();
  int* const slot_off = s_stack[warp];
  unsigned* const slot_mask = reinterpret_cast<unsigned*>(s_stack[warp] + QF_MAX_HITS);
  int c = valid ? hb.count(k_trav) : 0;
  const int wsum = __reduce_add_sync(0xffffffffu, c);
  int wbase = 0;
  if (lane == 0 && wsum) wbase = atomicAdd(cursor, wsum);
  wbase = __shfl_sync(0xffffffffu, wbase, 0);
  // slot-major layout of the warp's run: record (lane, j) sits at warp_start + off[j] + rank of lane among mask[j]
  const int warp_start = wbase;
  const int cmax = __reduce_max_sync(0xffffffffu, c);
  for (int j = 0, off = 0; j < cmax; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, c > j);
    if (lane == 0) { slot_off[j] = off; slot_mask[j] = m; }
    off += __popc(m);
  }
  __syncwarp();
  if (!valid) return;
  const unsigned lt = (1u << lane) - 1u;
  auto pos = [&](int j) { return warp_start + slot_off[j] + __popc(slot_mask[j] & lt); };
  ray_start[li] = warp_start;
  ray_count[li] = c;
  float prev = -1.f;
  bool unsorted = false;
  hb.for_each(k_trav, [&](int j, float, int id) {
    float px, py, pz;
    plane_hit(r, __ldg(planes + id), px, py, pz);
    float d = norm3(__fsub_rn(px, r.ox), __fsub_rn(py, r.oy), __fsub_rn(pz, r.oz));
    unsorted |= d < prev;
    prev = d;
    const int p = pos(j);
    hit_pd[p] = make_float4(px, py, pz, d);
    hit_rt[p] = make_int2((int)(ray0 + li), id);
  });
  if (unsorted) {  // rare: plane-hit depth order differs from Möller–Trumbore t order; stable insertion sort
    for (int s = 1; s < c; ++s) {
      float4 pd = hit_pd[pos(s)];
      int2 rt = hit_rt[pos(s)];
      int q = s;
      while (q > 0) {
        float4 o = hit_pd[pos(q - 1)];
        if (!(o.w > pd.w)) break;
        hit_pd[pos(q)] = o;
        hit_rt[pos(q)] = hit_rt[pos(q - 1)];
        --q;
      }
      hit_pd[pos(q)] = pd;
      hit_rt[pos(q)] = rt;
    }
  }
}

// Output placement of composite_rays_kernel.  band_rows == 0: pixel of ray i at index i (compact).  band_rows > 0: the rays
// are the band-cyclic share (qf_generate_rays_banded) of a frame `width` pixels wide and the outputs address the WHOLE
// frame — possibly in the memory of a peer GPU (qf_peer_open): the final image gather of a ray-sharded frame then is the
// composite kernel's own stores over NVLink, no collective and no reassembly pass.
struct FrameMap { int band_rows, band_stride, band_offset, width; };

__global__ void composite_rays_kernel(const int32_t* __restrict__ ray_start, const int32_t* __restrict__ ray_count,
                                      const float4* __restrict__ hit_pd, const float4* __restrict__ hit_out, float delta,
                                      int64_t ray0, int64_t n, int bg_mode, const float* __restrict__ bkgd,
                                      float* __restrict__ rgb, float* __restrict__ alpha_out, float* __restrict__ depth_out,
                                      const int32_t* __restrict__ cursor, int32_t* __restrict__ hits_total, int img_w,
                                      FrameMap fm, int n_bands) {
  const int64_t linear = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (linear == 0 && hits_total) atomicAdd(hits_total, *cursor);
  const int64_t li = ray_of_thread(linear, lane, img_w, n_bands);  // same warp <-> rays mapping as trace_compact_kernel
  const bool valid = li < n;
  const int s = valid ? ray_start[li] : 0, c = valid ? ray_count[li] : 0;
  float fill = bg_mode == QF_BG_BLACK ? 0.f : 1.f;
  float r = fill, g = fill, b = fill, A = 0.f, D = 0.f;
  float cum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
  // walk the warp's slot-major run: the j-th hits of the 32 rays are adjacent records
  const int cmax = __reduce_max_sync(0xffffffffu, c);
  const unsigned lt = (1u << lane) - 1u;
  for (int j = 0, off = 0; j < cmax; ++j) {
    const unsigned m = __ballot_sync(0xffffffffu, c > j);
    if (c > j) {
      const int p = s + off + __popc(m & lt);
      float4 o = hit_out[p];
      float tau = o.w * delta;
      float w = expf(-cum) * (1.0f - expf(-tau));
      cum += tau;
      cr += w * o.x; cg += w * o.y; cb += w * o.z;
      D += w * hit_pd[p].w;
      A += w;
    }
    off += __popc(m);
  }
  if (!valid) return;
  if (c > 0) {
    if (bg_mode == QF_BG_WHITE) { r = (1.f - A) + A * cr; g = (1.f - A) + A * cg; b = (1.f - A) + A * cb; }
    else if (bg_mode == QF_BG_BLACK) { r = A * cr; g = A * cg; b = A * cb; }
    else { r = A * cr + (1.f - A) * bkgd[0]; g = A * cg + (1.f - A) * bkgd[1]; b = A * cb + (1.f - A) * bkgd[2]; }
  }
  int64_t i = ray0 + li;
  if (fm.band_rows > 0) {   // the rays are a band-cyclic share of a frame: write the pixel where it sits in the WHOLE frame
    const int64_t lrow = i / fm.width, col = i - lrow * fm.width;
    const int64_t grow = ((lrow / fm.band_rows) * fm.band_stride + fm.band_offset) * fm.band_rows + lrow % fm.band_rows;
    i = grow * fm.width + col;
  }
  rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
  alpha_out[i] = A;
  depth_out[i] = D;
}

// band_rows > 0: only the rows y with (y / band_rows) % band_stride == band_offset are generated, compactly and in
// order — the band-cyclic share of one rank of a ray-sharded frame (n_local_rows of them); 0: the whole image.
__global__ void generate_rays_kernel(float r00, float r01, float r02, float r10, float r11, float r12, float r20, float r21,
                                     float r22, float tx, float ty, float tz, int W, int H, float focal, float cx, float cy,
                                     float sgn, float* __restrict__ origins, float* __restrict__ viewdirs, int band_rows,
                                     int band_stride, int band_offset, int n_local_rows) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)W * (band_rows > 0 ? n_local_rows : H)) return;
  int x = (int)(i % W), y = (int)(i / W);  // meshgrid(indexing="xy") flattened row-major (nerf_synthetic.py:311-317)
  if (band_rows > 0) y = ((y / band_rows) * band_stride + band_offset) * band_rows + y % band_rows;
  float cxd = __fdiv_rn(__fadd_rn(__fsub_rn((float)x, cx), 0.5f), focal);
  float cyd = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn((float)y, cy), 0.5f), focal), sgn);
  float czd = sgn;
  float dx = __fadd_rn(__fadd_rn(__fmul_rn(cxd, r00), __fmul_rn(cyd, r01)), __fmul_rn(czd, r02));
  float dy = __fadd_rn(__fadd_rn(__fmul_rn(cxd, r10), __fmul_rn(cyd, r11)), __fmul_rn(czd, r12));
  float dz = __fadd_rn(__fadd_rn(__fmul_rn(cxd, r20), __fmul_rn(cyd, r21)), __fmul_rn(czd, r22));
  float n = norm3(dx, dy, dz);
  viewdirs[3 * i] = __fdiv_rn(dx, n); viewdirs[3 * i + 1] = __fdiv_rn(dy, n); viewdirs[3 * i + 2] = __fdiv_rn(dz, n);
  origins[3 * i] = tx; origins[3 * i + 1] = ty; origins[3 * i + 2] = tz;
}

// Training-mode rays (nerf_synthetic.py:293-309, 341-370): ray i looks through pixel (x[i], y[i]) of camera image_id[i];
// x / y arrive as floats (integer pixel indices, or index + U[0,1) with add_ray_direction_noise).  Same operation order
// as generate_rays_kernel, per-ray camera matrix.
__global__ void generate_rays_indexed_kernel(const float* __restrict__ c2w, int64_t n_views, const int64_t* __restrict__ image_id,
                                             const float* __restrict__ xs, const float* __restrict__ ys, int64_t n, float focal,
                                             float cx, float cy, float sgn, float* __restrict__ origins,
                                             float* __restrict__ viewdirs) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = image_id ? image_id[i] : 0;
  if (v < 0 || v >= n_views) {   // an id outside the camera array: poison the ray instead of reading out of bounds
    const float q = __int_as_float(0x7fc00000);
    for (int k = 0; k < 3; ++k) { origins[3 * i + k] = q; viewdirs[3 * i + k] = q; }
    return;
  }
  const float* m = c2w + 12 * v;
  float cxd = __fdiv_rn(__fadd_rn(__fsub_rn(xs[i], cx), 0.5f), focal);
  float cyd = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(ys[i], cy), 0.5f), focal), sgn);
  float czd = sgn;
  float dx = __fadd_rn(__fadd_rn(__fmul_rn(cxd, m[0]), __fmul_rn(cyd, m[1])), __fmul_rn(czd, m[2]));
  float dy = __fadd_rn(__fadd_rn(__fmul_rn(cxd, m[4]), __fmul_rn(cyd, m[5])), __fmul_rn(czd, m[6]));
  float dz = __fadd_rn(__fadd_rn(__fmul_rn(cxd, m[8]), __fmul_rn(cyd, m[9])), __fmul_rn(czd, m[10]));
  float nrm = norm3(dx, dy, dz);
  viewdirs[3 * i] = __fdiv_rn(dx, nrm); viewdirs[3 * i + 1] = __fdiv_rn(dy, nrm); viewdirs[3 * i + 2] = __fdiv_rn(dz, nrm);
  origins[3 * i] = m[3]; origins[3 * i + 1] = m[7]; origins[3 * i + 2] = m[11];
}

// The images the reference's eval loops write to disk (train_finetune.py:639-646, test_baking_texture_images.py:401-407):
//   rgb8 = uint8(clamp(rgb, 0, 1) * 255),  depth8 = uint8(depth / depth.max() * 255)   (fp32, truncation as numpy's astype)
// made on the device so that 4 bytes per pixel instead of 16 cross PCIe when the caller only wants the PNG-ready frame.
__global__ void depth_max_kernel(const float* __restrict__ depth, int64_t n, int* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, depth[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_int(m));   // depths are >= 0: int order = float order
}
__device__ __forceinline__ unsigned char to_u8(float v) { return (v >= 0.f && v < 256.f) ? (unsigned char)(int)v : (unsigned char)0; }
__global__ void frame_to_u8_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, int64_t n,
                                   const int* __restrict__ dmax_bits, unsigned char* __restrict__ rgb8,
                                   unsigned char* __restrict__ depth8) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) rgb8[3 * i + c] = to_u8(__fmul_rn(fminf(fmaxf(rgb[3 * i + c], 0.f), 1.f), 255.f));
  if (depth8) depth8[i] = to_u8(__fmul_rn(__fdiv_rn(depth[i], __int_as_float(*dmax_bits)), 255.f));
}

enum class Shade { NGP, BAKED };

// Optional per-stage CUDA-event timing of the fused render (bench.py's roofline uses it): events are recorded on
// the launching stream around trace / shade / composite; reading them back synchronises.
struct StageProfile {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;   // 4 events per recorded chunk
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    cudaEvent_t e;
    if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
  }
};
static StageProfile g_prof;

static int render_common(Shade mode, const qf_mesh* mesh, const qf_ngp* field, const qf_texture* tex, const float* d_uv,
                         const float* d_origins, const float* d_viewdirs, int64_t n_rays, int image_width, int K, float delta, int bg_mode,
                         const float* d_bkgd, float* d_rgb, float* d_alpha, float* d_depth, int32_t* d_hits_total,
                         void* d_workspace, size_t workspace_bytes, cudaStream_t st, FrameMap fm = FrameMap{0, 0, 0, 0}) {
  QF_REQUIRE(mesh, "qf_render: NULL mesh");
  if (fm.band_rows > 0) {
    QF_REQUIRE(fm.width > 0 && fm.band_stride > 0 && fm.band_offset >= 0 && fm.band_offset < fm.band_stride,
               "qf_render_*_to_frame: band_rows=%d band_stride=%d band_offset=%d frame_width=%d", fm.band_rows, fm.band_stride,
               fm.band_offset, fm.width);
    QF_REQUIRE(n_rays % ((int64_t)fm.band_rows * fm.width) == 0, "qf_render_*_to_frame: n_rays=%lld is not whole bands of %d x %d",
               (long long)n_rays, fm.band_rows, fm.width);
  }
  QF_REQUIRE(n_rays >= 0, "qf_render: n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) {   // empty tensors carry NULL data pointers
    if (d_hits_total) QF_CUDA_CHECK(cudaMemsetAsync(d_hits_total, 0, sizeof(int32_t), st));
    return QF_OK;
  }
  QF_REQUIRE(d_origins && d_viewdirs && d_rgb && d_alpha && d_depth && d_workspace, "qf_render: NULL argument");
  QF_REQUIRE(K >= 1 && K <= QF_MAX_HITS, "qf_render: K=%d outside [1,%d]", K, QF_MAX_HITS);
  QF_REQUIRE(bg_mode >= 0 && bg_mode <= 2, "qf_render: bg_mode=%d", bg_mode);
  QF_REQUIRE(bg_mode != QF_BG_RANDOM || d_bkgd, "qf_render: bg 'random' needs render_bkgd");
  QF_REQUIRE(n_rays < (1ll << 31), "qf_render: at most 2^31-1 rays per call");
  QF_REQUIRE(workspace_bytes >= qf_render_workspace_bytes(n_rays, K), "qf_render: workspace %zu < %zu bytes", workspace_bytes,
             qf_render_workspace_bytes(n_rays, K));
  if (d_hits_total) QF_CUDA_CHECK(cudaMemsetAsync(d_hits_total, 0, sizeof(int32_t), st));
  if (n_rays == 0) return QF_OK;
  Workspace w;
  workspace_layout(n_rays, K, &w, (char*)d_workspace);
  static const int trace_mode = getenv("QF_TRACE_MODE") ? atoi(getenv("QF_TRACE_MODE")) : 0;
  // the 8x4 tile mapping needs whole tiles: width % 8 == 0 and every chunk made of whole 4-row bands
  const int64_t band = (int64_t)image_width * 4;
  const bool tiled = image_width > 0 && image_width % 8 == 0 && n_rays % band == 0 && band <= kChunkRays;
  const int img_w = tiled ? image_width : 0;
  const int64_t chunk = tiled ? (kChunkRays / band) * band : kChunkRays;
  for (int64_t ray0 = 0; ray0 < n_rays; ray0 += chunk) {
    const int64_t n = (n_rays - ray0) < chunk ? (n_rays - ray0) : chunk;
    QF_CUDA_CHECK(cudaMemsetAsync(w.cursor, 0, sizeof(int32_t), st));
    const int blocks = (int)ceil_div(n, 128);
    // centre-out band order for the K <= 8 frames (r2i, c2 trace alone 0.104 -> 0.092 ms; the pipelined frame rate is
    // unchanged) and for K = 32 launches of at most 600 K rays — a rank's share of a ray-sharded 1080p frame at N >= 4, where
    // the longest packets end the launch (share of 1/4: trace 0.525 -> 0.483 ms, of 1/8: 0.346 -> 0.332 ms,
    // tools/diag_share.py).  Larger K = 32 launches lose more BVH locality by alternating between the two halves of the image
    // than the shorter tail returns (c4 1.647 -> 1.664 ms, c5 4.65 -> 4.75 ms at N=1).
    static const int64_t co_max = getenv("QF_CENTER_OUT_K32_MAX_RAYS") ? atoll(getenv("QF_CENTER_OUT_K32_MAX_RAYS")) : 600000;
    const int n_bands = (tiled && ((mesh->restart_eps > 0.f ? QF_MAX_HITS : K) <= 8 || n <= co_max)) ? (int)(n / band) : 0;
    cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
    if (g_prof.enabled) { for (auto& e : pe) e = g_prof.get(); cudaEventRecord(pe[0], st); }
    const float eps = mesh->restart_eps;
    const int k_trav = eps > 0.f ? QF_MAX_HITS : K;
    if (k_trav <= 8)
      trace_compact_kernel<HitBufK8><<<blocks, 128, 0, st>>>(mesh->d_nodes, mesh->d_tris, mesh->d_planes, mesh->d_scene, d_origins,
                                                      d_viewdirs, ray0, n, K, img_w, trace_mode, w.cursor, w.ray_start, w.ray_count, w.hit_pd, w.hit_rt, mesh->d_wnodes, mesh->d_wstate, k_trav, eps, n_bands);
    else
      trace_compact_kernel<HitBufSmem><<<blocks, 128, 0, st>>>(mesh->d_nodes, mesh->d_tris, mesh->d_planes, mesh->d_scene, d_origins,
                                                       d_viewdirs, ray0, n, K, img_w, trace_mode, w.cursor, w.ray_start, w.ray_count, w.hit_pd, w.hit_rt, mesh->d_wnodes, mesh->d_wstate, k_trav, eps, n_bands);
    QF_LAUNCH_CHECK();
    if (g_prof.enabled) cudaEventRecord(pe[1], st);
    int rc = mode == Shade::NGP ? launch_ngp_forward_hits(field, w.hit_pd, w.hit_rt, d_viewdirs, w.cursor, w.hit_out, st)
                                : launch_baked_shade(tex, mesh, d_uv, w.hit_pd, w.hit_rt, d_viewdirs, w.cursor, w.hit_out, st);
    if (rc != QF_OK) return rc;
    if (g_prof.enabled) cudaEventRecord(pe[2], st);
    composite_rays_kernel<<<(int)ceil_div(n, 256), 256, 0, st>>>(w.ray_start, w.ray_count, w.hit_pd, w.hit_out, delta, ray0, n,
                                                                 bg_mode, d_bkgd, d_rgb, d_alpha, d_depth, w.cursor, d_hits_total, img_w, fm, n_bands);
    QF_LAUNCH_CHECK();
    if (g_prof.enabled) { cudaEventRecord(pe[3], st); for (auto e : pe) g_prof.ev.push_back(e); }
  }
  return QF_OK;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_render_mesh_ngp_to_frame(const qf_mesh* mesh, const qf_ngp* field, const float* d_origins, const float* d_viewdirs,
                                           int64_t n_rays, int K, float delta, int bg_mode, const float* d_bkgd, int band_rows,
                                           int band_stride, int band_offset, int frame_width, float* d_frame_rgb,
                                           float* d_frame_alpha, float* d_frame_depth, int32_t* d_hits_total, void* d_workspace,
                                           size_t workspace_bytes, void* stream) {
  QF_REQUIRE(field, "qf_render_mesh_ngp_to_frame: NULL field");
  QF_REQUIRE(band_rows > 0, "qf_render_mesh_ngp_to_frame: band_rows=%d", band_rows);
  return render_common(Shade::NGP, mesh, field, nullptr, nullptr, d_origins, d_viewdirs, n_rays, frame_width, K, delta, bg_mode, d_bkgd,
                       d_frame_rgb, d_frame_alpha, d_frame_depth, d_hits_total, d_workspace, workspace_bytes, (cudaStream_t)stream,
                       FrameMap{band_rows, band_stride, band_offset, frame_width});
}

extern "C" int qf_render_mesh_baked_to_frame(const qf_mesh* mesh, const qf_texture* tex, const float* d_uv_scaled,
                                             const float* d_origins, const float* d_viewdirs, int64_t n_rays, int K, float delta,
                                             int bg_mode, const float* d_bkgd, int band_rows, int band_stride, int band_offset,
                                             int frame_width, float* d_frame_rgb, float* d_frame_alpha, float* d_frame_depth,
                                             int32_t* d_hits_total, void* d_workspace, size_t workspace_bytes, void* stream) {
  QF_REQUIRE(tex && d_uv_scaled, "qf_render_mesh_baked_to_frame: NULL texture / uv");
  QF_REQUIRE(band_rows > 0, "qf_render_mesh_baked_to_frame: band_rows=%d", band_rows);
  return render_common(Shade::BAKED, mesh, nullptr, tex, d_uv_scaled, d_origins, d_viewdirs, n_rays, frame_width, K, delta, bg_mode,
                       d_bkgd, d_frame_rgb, d_frame_alpha, d_frame_depth, d_hits_total, d_workspace, workspace_bytes,
                       (cudaStream_t)stream, FrameMap{band_rows, band_stride, band_offset, frame_width});
}

extern "C" int qf_frame_to_u8(const float* d_rgb, const float* d_depth, int64_t n_pixels, unsigned char* d_rgb8,
                              unsigned char* d_depth8, int32_t* d_scratch, void* stream) {
  if (n_pixels == 0) return QF_OK;
  QF_REQUIRE(d_rgb && d_rgb8 && n_pixels > 0, "qf_frame_to_u8: NULL argument");
  QF_REQUIRE(!d_depth8 || (d_depth && d_scratch), "qf_frame_to_u8: the depth image needs d_depth and a 4-byte scratch");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_depth8) {
    QF_CUDA_CHECK(cudaMemsetAsync(d_scratch, 0, sizeof(int32_t), st));
    depth_max_kernel<<<kNumSMs * 2, 256, 0, st>>>(d_depth, n_pixels, d_scratch);
  }
  frame_to_u8_kernel<<<(unsigned)ceil_div(n_pixels, 256), 256, 0, st>>>(d_rgb, d_depth, n_pixels, d_scratch, d_rgb8, d_depth8);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

// ---- peer memory (one process per GPU): a frame buffer allocated on one rank and mapped into the others, so that their
// composite kernels store their pixels straight into it over NVLink.  Plain cudaMalloc + CUDA IPC handles (64 opaque bytes
// a rank passes to its peers through any host channel, e.g. torch.distributed.broadcast_object_list).
extern "C" int qf_peer_alloc(size_t bytes, void** d_ptr) {
  QF_REQUIRE(d_ptr && bytes > 0, "qf_peer_alloc: NULL / empty");
  QF_CUDA_CHECK(cudaMalloc(d_ptr, bytes));
  return QF_OK;
}
extern "C" int qf_peer_free(void* d_ptr) {
  if (d_ptr) QF_CUDA_CHECK(cudaFree(d_ptr));
  return QF_OK;
}
extern "C" int qf_peer_export(const void* d_ptr, unsigned char* handle64) {
  QF_REQUIRE(d_ptr && handle64, "qf_peer_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  QF_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
  memcpy(handle64, &h, 64);
  return QF_OK;
}
extern "C" int qf_peer_open(const unsigned char* handle64, void** d_ptr) {
  QF_REQUIRE(d_ptr && handle64, "qf_peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  QF_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QF_OK;
}
extern "C" int qf_peer_close(void* d_ptr) {
  if (d_ptr) QF_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
  return QF_OK;
}

// diagnostics (not part of the ABI): work counters of the wide packet traversal, non-zero only in a -DQF_TRACE_STATS build;
// reading resets them
extern "C" int qf_debug_trace_stats(unsigned long long* out8) {
#ifdef QF_TRACE_STATS
  QF_CUDA_CHECK(cudaDeviceSynchronize());
  QF_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_trace_stats, sizeof(unsigned long long) * 8));
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  QF_CUDA_CHECK(cudaMemcpyToSymbol(g_trace_stats, z, sizeof(z)));
#else
  for (int i = 0; i < 8; ++i) out8[i] = 0;
#endif
  return QF_OK;
}

extern "C" size_t qf_render_workspace_bytes(int64_t n_rays, int K) { return workspace_layout(n_rays, K, nullptr, nullptr); }

extern "C" int qf_render_mesh_ngp(const qf_mesh* mesh, const qf_ngp* field, const float* d_origins, const float* d_viewdirs,
                                  int64_t n_rays, int image_width, int K, float delta, int bg_mode, const float* d_bkgd, float* d_rgb,
                                  float* d_alpha, float* d_depth, int32_t* d_hits_total, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
  QF_REQUIRE(field, "qf_render_mesh_ngp: NULL field");
  return render_common(Shade::NGP, mesh, field, nullptr, nullptr, d_origins, d_viewdirs, n_rays, image_width, K, delta, bg_mode, d_bkgd, d_rgb,
                       d_alpha, d_depth, d_hits_total, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int qf_render_mesh_baked(const qf_mesh* mesh, const qf_texture* tex, const float* d_uv_scaled, const float* d_origins,
                                    const float* d_viewdirs, int64_t n_rays, int image_width, int K, float delta, int bg_mode,
                                    const float* d_bkgd, float* d_rgb, float* d_alpha, float* d_depth, int32_t* d_hits_total,
                                    void* d_workspace, size_t workspace_bytes, void* stream) {
  QF_REQUIRE(tex && d_uv_scaled, "qf_render_mesh_baked: NULL texture / uv");
  return render_common(Shade::BAKED, mesh, nullptr, tex, d_uv_scaled, d_origins, d_viewdirs, n_rays, image_width, K, delta, bg_mode, d_bkgd,
                       d_rgb, d_alpha, d_depth, d_hits_total, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int qf_generate_rays(const float* c, int W, int H, float focal, float cx, float cy, int opengl, float* d_origins,
                                float* d_viewdirs, void* stream) {
  QF_REQUIRE(c && d_origins && d_viewdirs && W > 0 && H > 0, "qf_generate_rays: bad argument");
  const int64_t n = (int64_t)W * H;
  generate_rays_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(c[0], c[1], c[2], c[4], c[5], c[6], c[8], c[9], c[10],
                                                                                c[3], c[7], c[11], W, H, focal, cx, cy,
                                                                                opengl ? -1.0f : 1.0f, d_origins, d_viewdirs, 0, 1, 0, H);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int64_t qf_band_rows(int H, int band_rows, int band_stride, int band_offset) {
  if (H <= 0 || band_rows <= 0 || band_stride <= 0 || band_offset < 0 || band_offset >= band_stride) return -1;
  int64_t rows = 0;
  for (int b = band_offset; b * band_rows < H; b += band_stride) rows += (H - b * band_rows) < band_rows ? (H - b * band_rows) : band_rows;
  return rows;
}

extern "C" int qf_generate_rays_banded(const float* c, int W, int H, float focal, float cx, float cy, int opengl, int band_rows,
                                       int band_stride, int band_offset, float* d_origins, float* d_viewdirs, void* stream) {
  QF_REQUIRE(c && d_origins && d_viewdirs && W > 0 && H > 0, "qf_generate_rays_banded: bad argument");
  const int64_t rows = qf_band_rows(H, band_rows, band_stride, band_offset);
  QF_REQUIRE(rows >= 0, "qf_generate_rays_banded: band_rows=%d band_stride=%d band_offset=%d", band_rows, band_stride, band_offset);
  QF_REQUIRE(H % band_rows == 0, "qf_generate_rays_banded: H=%d is not a multiple of band_rows=%d", H, band_rows);
  const int64_t n = (int64_t)W * rows;
  if (n == 0) return QF_OK;
  generate_rays_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(c[0], c[1], c[2], c[4], c[5], c[6], c[8], c[9], c[10],
                                                                                c[3], c[7], c[11], W, H, focal, cx, cy,
                                                                                opengl ? -1.0f : 1.0f, d_origins, d_viewdirs, band_rows,
                                                                                band_stride, band_offset, (int)rows);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_generate_rays_indexed(const float* d_c2w, int64_t n_views, const int64_t* d_image_id, const float* d_x,
                                        const float* d_y, int64_t n, float focal, float cx, float cy, int opengl,
                                        float* d_origins, float* d_viewdirs, void* stream) {
  QF_REQUIRE(n >= 0 && n_views >= 1 && focal > 0.f, "qf_generate_rays_indexed: n=%lld views=%lld focal=%f", (long long)n,
             (long long)n_views, focal);
  if (n == 0) return QF_OK;
  QF_REQUIRE(d_c2w && d_x && d_y && d_origins && d_viewdirs, "qf_generate_rays_indexed: NULL argument");
  generate_rays_indexed_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      d_c2w, n_views, d_image_id, d_x, d_y, n, focal, cx, cy, opengl ? -1.0f : 1.0f, d_origins, d_viewdirs);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_profile_enable(int on) {
  g_prof.enabled = on != 0;
  if (on) {  // create the events up front so that recording inside a timed region does no driver allocation
    while (g_prof.pool.size() < 4096) {
      cudaEvent_t e;
      QF_CUDA_CHECK(cudaEventCreate(&e));
      g_prof.pool.push_back(e);
    }
  }
  return QF_OK;
}

// Sums the recorded stage times (ms) since the last read: ms3 = {trace, shade, composite}; n_chunks = launches of each.
extern "C" int qf_profile_read(double* ms3, int64_t* n_chunks) {
  QF_REQUIRE(ms3 && n_chunks, "qf_profile_read: NULL argument");
  ms3[0] = ms3[1] = ms3[2] = 0.0;
  *n_chunks = 0;
  for (size_t i = 0; i + 3 < g_prof.ev.size(); i += 4) {
    QF_CUDA_CHECK(cudaEventSynchronize(g_prof.ev[i + 3]));
    for (int k = 0; k < 3; ++k) {
      float ms = 0.f;
      QF_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof.ev[i + k], g_prof.ev[i + k + 1]));
      ms3[k] += ms;
    }
    ++*n_chunks;
  }
  for (auto e : g_prof.ev) g_prof.pool.push_back(e);
  g_prof.ev.clear();
  return QF_OK;
}
