// (1) Quadrature-mesh BVH: GPU LBVH build (63-bit Morton, Karras 2012) and first-K all-hits traversal.
//
// Replaces the OptiX `intersector.Intersector` pybind object (mesh_utils.py:77-96) and the Embree
// default (mesh_utils.py:223,350-354).  B200 has no RT cores, so traversal is a software
// loop on the SMs (traverse.cuh): 64-byte two-child nodes fetched as 4 x LDG.128 through the read-only
// path, one triangle per leaf so that the child box in the parent IS the triangle's padded box,
// Morton-ordered 48-byte triangle records, a per-ray sorted K-buffer held in registers, and the
// K-th hit's t used as the culling distance once the buffer is full.  Coherent warps traverse as a packet.
//
// Exactness: a node is entered iff the fp32 slab test of its box passes; the triangle predicate
// contains the same slab test on the triangle's own padded box, and node boxes are exact unions
// of those, so by monotonicity of rounding no triangle that the brute-force oracle accepts can be
// culled.
//
// This file is compiled with -fmad=false; all predicate arithmetic additionally uses explicit
// round-to-nearest intrinsics.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <float.h>
#include <stdlib.h>

#include "traverse.cuh"

namespace qf {

// ---------------------------------------------------------------- scene bounds + pad
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void bounds_init_kernel(int* __restrict__ ob) {
  if (threadIdx.x < 3) ob[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) ob[threadIdx.x] = INT_MIN;
}

__global__ void bounds_kernel(const float* __restrict__ v, int64_t n, int* __restrict__ ob) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = __ldg(v + 3 * i + c);
      lo[c] = fminf(lo[c], x);
      hi[c] = fmaxf(hi[c], x);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(ob + c, float_to_ordered(lo[c]));
      atomicMax(ob + 3 + c, float_to_ordered(hi[c]));
    }
  }
}

// scene[0..5] = bounds, scene[6] = pad = (largest extent) * 1e-4f  (oracle: mesh_box_pad)
__global__ void bounds_finish_kernel(const int* __restrict__ ob, float* __restrict__ scene) {
  float lo[3], hi[3];
  for (int c = 0; c < 3; ++c) {
    lo[c] = ordered_to_float(ob[c]);
    hi[c] = ordered_to_float(ob[3 + c]);
    scene[c] = lo[c];
    scene[3 + c] = hi[c];
  }
  float ext = fmaxf(fmaxf(__fsub_rn(hi[0], lo[0]), __fsub_rn(hi[1], lo[1])), __fsub_rn(hi[2], lo[2]));
  scene[6] = __fmul_rn(ext, 1e-4f);
}

// ---------------------------------------------------------------- per-triangle setup
__device__ __forceinline__ uint64_t expand21(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void tri_setup_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int64_t F,
                                 const float* __restrict__ scene, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ idx, float4* __restrict__ planes) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= F) return;
  int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
  float v0[3], v1[3], v2[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { v0[k] = verts[3 * (int64_t)a + k]; v1[k] = verts[3 * (int64_t)b + k]; v2[k] = verts[3 * (int64_t)c + k]; }
  // Morton key of the box centre, 21 bits per axis
  uint64_t q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float lo = scene[k], hi = scene[3 + k];
    float cen = 0.5f * (min3(v0[k], v1[k], v2[k]) + max3(v0[k], v1[k], v2[k]));
    float ext = hi - lo;
    float u = ext > 0.f ? (cen - lo) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    q[k] = (uint64_t)fminf(u * 2097152.0f, 2097151.0f);
  }
  keys[f] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
  idx[f] = (uint32_t)f;
  // trimesh face_normals: fp64 cross(v1-v0, v2-v1) / |.|, zero if degenerate; cast to fp32 (mesh_utils.py:104)
  double ax = (double)v1[0] - (double)v0[0], ay = (double)v1[1] - (double)v0[1], az = (double)v1[2] - (double)v0[2];
  double bx = (double)v2[0] - (double)v1[0], by = (double)v2[1] - (double)v1[1], bz = (double)v2[2] - (double)v1[2];
  double nx = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by));
  double ny = __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz));
  double nz = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
  double ln = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
  float fx = 0.f, fy = 0.f, fz = 0.f;
  if (ln > 0.0) { fx = (float)__ddiv_rn(nx, ln); fy = (float)__ddiv_rn(ny, ln); fz = (float)__ddiv_rn(nz, ln); }
  float d = -__fadd_rn(__fadd_rn(__fmul_rn(fx, v0[0]), __fmul_rn(fy, v0[1])), __fmul_rn(fz, v0[2]));
  planes[f] = make_float4(fx, fy, fz, d);
}

__global__ void tri_gather_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int64_t F,
                                  const uint32_t* __restrict__ idx_sorted, float4* __restrict__ tris) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= F) return;
  uint32_t f = idx_sorted[k];
  int a = faces[3 * (int64_t)f], b = faces[3 * (int64_t)f + 1], c = faces[3 * (int64_t)f + 2];
  tris[3 * k + 0] = make_float4(verts[3 * (int64_t)a], verts[3 * (int64_t)a + 1], verts[3 * (int64_t)a + 2], __int_as_float((int)f));
  tris[3 * k + 1] = make_float4(verts[3 * (int64_t)b], verts[3 * (int64_t)b + 1], verts[3 * (int64_t)b + 2], 0.f);
  tris[3 * k + 2] = make_float4(verts[3 * (int64_t)c], verts[3 * (int64_t)c + 1], verts[3 * (int64_t)c + 2], 0.f);
}

// ---------------------------------------------------------------- Karras hierarchy
__device__ __forceinline__ int delta_fn(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

__global__ void karras_kernel(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                              int* __restrict__ parent, int* __restrict__ leaf_parent, int* __restrict__ first,
                              int* __restrict__ last) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta_fn(keys, n, i, i - d);
  int lmax = 2;
  while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta_fn(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  // child refs: >=0 internal, ~k leaf k
  int L = (lo == gamma) ? ~gamma : gamma;
  int R = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  left[i] = L; right[i] = R; first[i] = lo; last[i] = hi;
  if (L >= 0) parent[L] = i; else leaf_parent[gamma] = i;
  if (R >= 0) parent[R] = i; else leaf_parent[gamma + 1] = i;
  if (i == 0) parent[0] = -1;
}

__device__ __forceinline__ void tri_box(const float4* __restrict__ tris, int k, float pad, float* lo, float* hi) {
  float4 a = tris[3 * (int64_t)k], b = tris[3 * (int64_t)k + 1], c = tris[3 * (int64_t)k + 2];
  lo[0] = __fsub_rn(min3(a.x, b.x, c.x), pad); lo[1] = __fsub_rn(min3(a.y, b.y, c.y), pad); lo[2] = __fsub_rn(min3(a.z, b.z, c.z), pad);
  hi[0] = __fadd_rn(max3(a.x, b.x, c.x), pad); hi[1] = __fadd_rn(max3(a.y, b.y, c.y), pad); hi[2] = __fadd_rn(max3(a.z, b.z, c.z), pad);
}

// bottom-up box fit: the second thread to reach a node unions its two children
__global__ void fit_kernel(const float4* __restrict__ tris, int n, const float* __restrict__ scene,
                           const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent,
                           const int* __restrict__ leaf_parent, int* __restrict__ flags, float4* __restrict__ ibox) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float pad = scene[6];
  int cur = leaf_parent[k];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(flags + cur, 1) == 0) return;
    __threadfence();
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int refs[2] = {left[cur], right[cur]};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float cl[3], ch[3];
      if (refs[c] < 0) tri_box(tris, ~refs[c], pad, cl, ch);
      else {
        float4 a = __ldcg(ibox + 2 * (int64_t)refs[c]), b = __ldcg(ibox + 2 * (int64_t)refs[c] + 1);
        cl[0] = a.x; cl[1] = a.y; cl[2] = a.z; ch[0] = b.x; ch[1] = b.y; ch[2] = b.z;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) { lo[q] = fminf(lo[q], cl[q]); hi[q] = fmaxf(hi[q], ch[q]); }
    }
    ibox[2 * (int64_t)cur] = make_float4(lo[0], lo[1], lo[2], 0.f);
    ibox[2 * (int64_t)cur + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    cur = parent[cur];
  }
}

// final 64-byte traversal nodes: both children's boxes + references (leaf = one triangle)
__global__ void emit_nodes_kernel(const float4* __restrict__ tris, int n, const float* __restrict__ scene,
                                  const int* __restrict__ left, const int* __restrict__ right,
                                  const float4* __restrict__ ibox, float4* __restrict__ nodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  float pad = scene[6];
  int refs[2] = {left[i], right[i]};
  float lo[2][3], hi[2][3];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    int r = refs[c];
    if (r < 0) tri_box(tris, ~r, pad, lo[c], hi[c]);
    else {
      float4 a = ibox[2 * (int64_t)r], b = ibox[2 * (int64_t)r + 1];
      lo[c][0] = a.x; lo[c][1] = a.y; lo[c][2] = a.z; hi[c][0] = b.x; hi[c][1] = b.y; hi[c][2] = b.z;
    }
  }
  nodes[4 * (int64_t)i + 0] = make_float4(lo[0][0], lo[0][1], lo[0][2], hi[0][0]);
  nodes[4 * (int64_t)i + 1] = make_float4(hi[0][1], hi[0][2], lo[1][0], lo[1][1]);
  nodes[4 * (int64_t)i + 2] = make_float4(lo[1][2], hi[1][0], hi[1][1], hi[1][2]);
  nodes[4 * (int64_t)i + 3] = make_float4(__int_as_float(refs[0]), __int_as_float(refs[1]), 0.f, 0.f);
}

// single-triangle mesh: a root whose first child is that triangle
__global__ void emit_tiny_root_kernel(const float4* __restrict__ tris, const float* __restrict__ scene,
                                      float4* __restrict__ nodes) {
  float lo[3], hi[3];
  tri_box(tris, 0, scene[6], lo, hi);
  nodes[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
  nodes[1] = make_float4(hi[1], hi[2], 0.f, 0.f);
  nodes[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  nodes[3] = make_float4(__int_as_float(~0), __int_as_float(kEmptyRef), 0.f, 0.f);
}

// ---------------------------------------------------------------- wide BVH: collapse of the binary tree
// One warp per wide node, lane = child slot.  The frontier starts as the two children of the node's binary root and is
// expanded one binary step at a time:
//   1. while a frontier entry covers more than 32 triangles, the one with the largest surface area is replaced by its two
//      children (slots permitting) — the upper levels tile space by area, like a SAH collapse;
//   2. entries of <= 32 triangles stay whole subtrees (each becomes one well-filled node of the next level), except that
//      the smallest ones are dissolved into single triangles while the free slots allow, so no node of 2-3 triangles is
//      ever emitted next to free slots.
// Children that are still internal get a wide node of the next level; triangle children carry the triangle's own padded
// box.  Levels are processed by successive launches (kWideLevels of them: the packet stack, traverse.cuh kWideStack, holds
// 31 * kWideLevels + 1 pushes); a
// tree that needs more, or more nodes than were allocated, clears the ok flag and the traversal keeps to the binary tree.
constexpr int kWideLevels = 6;
static_assert(31 * kWideLevels + 1 <= kWideStack, "packet stack too small for the collapse depth");

__global__ void wide_init_kernel(int32_t* __restrict__ wstate, int32_t* __restrict__ wqueue) {
  wstate[0] = 0; wstate[1] = 1; wstate[2] = 1; wstate[3] = 1; wstate[4] = 0;
  wqueue[0] = 0;
}

__global__ void wide_advance_kernel(int32_t* __restrict__ wstate, int capacity) {
  wstate[0] = wstate[1];
  wstate[1] = wstate[2] < capacity ? wstate[2] : capacity;
  if (wstate[2] > capacity) wstate[3] = 0;
  wstate[4] += 1;
  if (wstate[4] >= kWideLevels && wstate[0] < wstate[1]) wstate[3] = 0;    // deeper than the packet stack allows
}

__global__ void __launch_bounds__(128) wide_collapse_kernel(const float4* __restrict__ tris, int n_tris, const float* __restrict__ scene,
                                                            const int* __restrict__ left, const int* __restrict__ right,
                                                            const int* __restrict__ first, const int* __restrict__ last,
                                                            const float4* __restrict__ ibox, int32_t* __restrict__ wqueue,
                                                            int32_t* __restrict__ wstate, int capacity, float4* __restrict__ wnodes) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int begin = wstate[0], end = wstate[1];
  const float pad = scene[6];
  for (int w = begin + gw; w < end; w += nw) {
    int ref = kEmptyRef;          // binary reference held by this slot
    int cnt = 0;                  // triangles below it
    float area = -1.f;
    auto fill = [&](int r) {
      ref = r;
      if (r < 0) { cnt = 1; area = -1.f; }
      else {
        cnt = last[r] - first[r] + 1;
        const float4 a = ibox[2 * (int64_t)r], b = ibox[2 * (int64_t)r + 1];
        const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
        area = dx * dy + dy * dz + dz * dx;
      }
    };
    int n;
    if (n_tris == 1) { if (lane == 0) fill(~0); n = 1; }
    else {
      const int b = wqueue[w];
      if (lane == 0) fill(left[b]);
      if (lane == 1) fill(right[b]);
      n = 2;
    }
    // ---- phase 1: split the largest-area entry that covers more than 32 triangles
    while (n < 32) {
      float key = (ref >= 0 && cnt > 32) ? area : -1.f;
      float best = key;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (best < 0.f) break;
      const int who = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
      const int r = __shfl_sync(0xffffffffu, ref, who);
      if (lane == who) fill(left[r]);
      if (lane == n) fill(right[r]);
      ++n;
    }
    // ---- phase 2: dissolve the smallest remaining subtrees into triangles while the slots allow
    while (n < 32) {
      int key = (ref >= 0 && cnt <= 32) ? cnt : 0x7fffffff;
      int best = key;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (best == 0x7fffffff || n + best - 1 > 32) break;
      const int who = __ffs(__ballot_sync(0xffffffffu, key == best)) - 1;
      const int r = __shfl_sync(0xffffffffu, ref, who);
      if (lane == who) fill(left[r]);
      if (lane == n) fill(right[r]);
      ++n;
    }
    // ---- emit: internal entries get wide nodes of the next level
    const bool internal = lane < n && ref >= 0;
    const unsigned mi = __ballot_sync(0xffffffffu, internal);
    int base = 0;
    if (lane == 0 && mi) base = atomicAdd(wstate + 2, __popc(mi));
    base = __shfl_sync(0xffffffffu, base, 0);
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
    int out_ref = kEmptyRef;
    if (lane < n) {
      if (ref < 0) { tri_box(tris, ~ref, pad, lo, hi); out_ref = ref; }
      else {
        const float4 a = ibox[2 * (int64_t)ref], b = ibox[2 * (int64_t)ref + 1];
        lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
        const int idx = base + __popc(mi & ((1u << lane) - 1u));
        if (idx < capacity) { wqueue[idx] = ref; out_ref = idx; }
        else out_ref = kEmptyRef;      // overflow: the ok flag goes down in wide_advance_kernel, the node is never used
      }
    }
    if (wnodes) {   // NULL: dry run that only counts the nodes (qf_mesh_create sizes the array with it)
      float4* rec = wnodes + ((int64_t)w * 32 + lane) * 2;
      rec[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
      rec[1] = make_float4(hi[1], hi[2], __int_as_float(out_ref), 0.f);
    }
  }
}

// Pre-pass over a ray list.  slot[0] += warps (32 consecutive rays) that would traverse as a packet, slot[1] += warps,
// slot[3] = number of rays whose slab test against the whole scene box passes; their ids are compacted into `list`.
// The outputs are pre-filled with "no hit" (memsets), so the refill kernel only writes rays that have hits.
__global__ void classify_rays_kernel(const float* __restrict__ scene, const float* __restrict__ origins,
                                     const float* __restrict__ dirs, int64_t N, int32_t* __restrict__ slot,
                                     int32_t* __restrict__ list) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = i < N;
  Ray r = make_ray(origins, dirs, valid ? i : N - 1);
  const bool coh = warp_is_coherent(r, valid);
  // all padded triangle boxes lie inside [scene_lo - pad, scene_hi + pad]: by monotonicity a ray that fails this slab
  // test fails every triangle's own box test
  const float pad = __ldg(scene + 6);
  float tn, tf;
  const bool in_scene = valid && slab(r, __fsub_rn(__ldg(scene), pad), __fsub_rn(__ldg(scene + 1), pad), __fsub_rn(__ldg(scene + 2), pad),
                                      __fadd_rn(__ldg(scene + 3), pad), __fadd_rn(__ldg(scene + 4), pad), __fadd_rn(__ldg(scene + 5), pad), tn, tf);
  const unsigned m = __ballot_sync(0xffffffffu, in_scene);
  int base = 0;
  if (lane == 0 && valid) {
    atomicAdd(slot + 1, 1);
    if (coh) atomicAdd(slot, 1);
    if (m && list) base = atomicAdd(slot + 3, __popc(m));
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  if (in_scene && list) list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
}

__global__ void fill_inf_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = __int_as_float(0x7f800000);
}
__device__ __forceinline__ bool mostly_coherent(const int32_t* slot) { return 2 * __ldg(slot) >= __ldg(slot + 1); }

template <class HB>
__device__ __forceinline__ void write_hits(HB& hb, int total, int64_t i, int K, int32_t* __restrict__ out_tri,
                                           float* __restrict__ out_t, int32_t* __restrict__ out_count,
                                           int32_t* __restrict__ out_total) {
  const int c = hb.count(K);
  hb.for_each(K, [&](int j, float t, int id) {
    out_tri[i * K + j] = id;
    if (out_t) out_t[i * K + j] = t;
  });
  for (int j = c; j < K; ++j) {
    out_tri[i * K + j] = -1;
    if (out_t) out_t[i * K + j] = __int_as_float(0x7f800000);
  }
  out_count[i] = c;
  if (out_total) out_total[i] = total;
}

// Incoherent ray lists (training batches of random pixels): persistent warps whose lanes fetch a new ray as soon as
// enough of them have finished, so one long ray no longer idles the other 31 lanes (thread efficiency was 3.4 / 32
// with one ray per thread, profiles/r1c).  Runs only when classify_rays_kernel found the list mostly incoherent.
template <class HB, bool COUNT_ALL>
__global__ void __launch_bounds__(128) trace_refill_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                           const float* __restrict__ origins, const float* __restrict__ dirs,
                                                           int64_t N, int K, int32_t* __restrict__ out_tri,
                                                           float* __restrict__ out_t, int32_t* __restrict__ out_count,
                                                           int32_t* __restrict__ out_total, int32_t* __restrict__ slot,
                                                           const int32_t* __restrict__ list, int k_trav, float restart_eps) {
  if (mostly_coherent(slot)) return;
  N = __ldg(slot + 3);   // rays that reach the scene box; the rest keep the pre-filled "no hit"
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  __shared__ float s_ht[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  __shared__ int s_hi[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  int sref[kStackDepth];
  float stn[kStackDepth];
  HB hb(s_ht, s_hi, threadIdx.x);
  Ray r;
  int64_t ri = -1;
  int sp = 0, cur = kDoneRef, total = 0;
  float cur_tn = 0.f;
  bool exhausted = false;
  hb.init(k_trav);
  while (true) {
    const unsigned idle = __ballot_sync(0xffffffffu, ri < 0);
    if (!exhausted && __popc(idle) >= 12) {
      const int n = __popc(idle);
      int base = 0;
      if (lane == 0) base = atomicAdd(slot + 2, n);
      base = __shfl_sync(0xffffffffu, base, 0);
      const int64_t mine = (int64_t)base + __popc(idle & lt);
      if (ri < 0 && mine < N) {
        ri = __ldg(list + mine);
        r = make_ray(origins, dirs, ri);
        hb.init(k_trav);
        total = 0; sp = 0; cur = 0; cur_tn = 0.f;
      }
      exhausted = (int64_t)base + n >= N;
    }
    if (__all_sync(0xffffffffu, ri < 0)) { if (exhausted) break; else continue; }
    int steps = 0;
    while (ri >= 0 && cur >= 0 && cur != kDoneRef && steps < 24) {
      single_node_step<HB, !COUNT_ALL>(r, nodes, hb, sref, stn, sp, cur, cur_tn);
      ++steps;
    }
    while (ri >= 0 && cur < 0) {
      const float tcull = COUNT_ALL ? __int_as_float(0x7f800000) : hb.cull_distance();
      if (cur_tn <= tcull) leaf_intersect<HB>(r, tris, cur, cur_tn, hb, total);
      if (sp) { cur = sref[--sp]; cur_tn = stn[sp]; }
      else cur = kDoneRef;
    }
    if (ri >= 0 && cur == kDoneRef) {
      if (total > 0) {
        if (restart_eps > 0.f) hb.restart_filter(restart_eps, K);
        write_hits<HB>(hb, total, ri, K, out_tri, out_t, out_count, out_total);
      }
      ri = -1;
    }
  }
}

template <class HB, bool COUNT_ALL>
__global__ void __launch_bounds__(128) trace_kernel(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                    const float* __restrict__ origins, const float* __restrict__ dirs,
                                                    int64_t N, int K, int32_t* __restrict__ out_tri,
                                                    float* __restrict__ out_t, int32_t* __restrict__ out_count,
                                                    int32_t* __restrict__ out_total, const int32_t* __restrict__ slot,
                                                    const float4* __restrict__ wnodes, const int32_t* __restrict__ wstate,
                                                    int k_trav, float restart_eps) {
  if (!mostly_coherent(slot)) return;   // trace_refill_kernel handles this list
  __shared__ int s_stack[4][kWideStack];
  if (wnodes && !__ldg(wstate + 3)) wnodes = nullptr;    // the collapse gave up on this mesh: binary tree only
  __shared__ float s_ht[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  __shared__ int s_hi[HB::kSmemSlots ? HB::kSmemSlots * 128 : 1];
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = i < N;
  Ray r = make_ray(origins, dirs, valid ? i : N - 1);
  HB hb(s_ht, s_hi, threadIdx.x);
  int total;
  trace_ray<HB, !COUNT_ALL>(r, valid, nodes, wnodes, tris, k_trav, hb, total, s_stack[threadIdx.x >> 5]);
  if (!valid) return;
  if (restart_eps > 0.f) hb.restart_filter(restart_eps, K);
  write_hits<HB>(hb, total, i, K, out_tri, out_t, out_count, out_total);
}

// ---------------------------------------------------------------- tuple packing (a3)
// thread per ray: plane-hit points, normalised dirs, depth, stable insertion sort by depth, ray-major write
__global__ void hits_pack_kernel(const float4* __restrict__ planes, const float* __restrict__ origins,
                                 const float* __restrict__ dirs, int64_t N, int K, const int32_t* __restrict__ tri,
                                 const int32_t* __restrict__ count, const int64_t* __restrict__ offsets,
                                 float* __restrict__ points, float* __restrict__ vectors, int64_t* __restrict__ index_ray,
                                 float* __restrict__ depth, int64_t* __restrict__ index_tri, float* __restrict__ origins_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  int c = count[i];
  if (c == 0) return;
  Ray r = make_ray(origins, dirs, i);
  float nrm = __fadd_rn(norm3(r.dx, r.dy, r.dz), 1e-7f);
  float vx = __fdiv_rn(r.dx, nrm), vy = __fdiv_rn(r.dy, nrm), vz = __fdiv_rn(r.dz, nrm);
  float px[QF_MAX_HITS], py[QF_MAX_HITS], pz[QF_MAX_HITS], dp[QF_MAX_HITS];
  int id[QF_MAX_HITS];
  for (int s = 0; s < c; ++s) {
    int t = tri[i * K + s];
    float x, y, z;
    plane_hit(r, __ldg(planes + t), x, y, z);
    float d = norm3(__fsub_rn(x, r.ox), __fsub_rn(y, r.oy), __fsub_rn(z, r.oz));
    int q = s;  // stable insertion by depth
    while (q > 0 && dp[q - 1] > d) { px[q] = px[q - 1]; py[q] = py[q - 1]; pz[q] = pz[q - 1]; dp[q] = dp[q - 1]; id[q] = id[q - 1]; --q; }
    px[q] = x; py[q] = y; pz[q] = z; dp[q] = d; id[q] = t;
  }
  int64_t base = offsets[i];
  for (int s = 0; s < c; ++s) {
    int64_t o = base + s;
    points[3 * o] = px[s]; points[3 * o + 1] = py[s]; points[3 * o + 2] = pz[s];
    vectors[3 * o] = vx; vectors[3 * o + 1] = vy; vectors[3 * o + 2] = vz;
    origins_out[3 * o] = r.ox; origins_out[3 * o + 1] = r.oy; origins_out[3 * o + 2] = r.oz;
    index_ray[o] = i; depth[o] = dp[s]; index_tri[o] = id[s];
  }
}

// a4: per-segment stable re-sort by depth of a ray-major tuple; thread per hit finds its rank inside
// its segment (segments are <= K long, so the O(len) scan is a handful of L1 hits).
__global__ void hits_resort_kernel(const int64_t* __restrict__ index_ray, const float* __restrict__ depth, int64_t M,
                                   int64_t* __restrict__ perm, uint8_t* __restrict__ boundary) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  int64_t ray = index_ray[i];
  float d = depth[i];
  int64_t s = i, rank = 0;
  while (s > 0 && index_ray[s - 1] == ray) { --s; if (depth[s] <= d) ++rank; }   // earlier equal keys stay first
  int64_t e = i + 1;
  while (e < M && index_ray[e] == ray) { if (depth[e] < d) ++rank; ++e; }
  perm[s + rank] = i;
  boundary[i] = (i == s) ? 1 : 0;
}

__global__ void widen_count_kernel(const int32_t* __restrict__ c, int64_t n, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = c[i];
  if (i == n) out[n] = 0;
}

// `wnodes` NULL: dry run (counts the wide nodes into d_wstate[2], capacity = the queue's).
static int collapse_wide(qf_mesh* m, float4* wnodes, int capacity, cudaStream_t st) {
  wide_init_kernel<<<1, 1, 0, st>>>(m->d_wstate, m->d_wqueue);
  for (int lvl = 0; lvl < kWideLevels; ++lvl) {
    wide_collapse_kernel<<<kNumSMs * 4, 128, 0, st>>>(m->d_tris, (int)m->n_faces, m->d_scene, m->d_left, m->d_right, m->d_first,
                                                     m->d_last, m->d_ibox, m->d_wqueue, m->d_wstate, capacity, wnodes);
    wide_advance_kernel<<<1, 1, 0, st>>>(m->d_wstate, capacity);
  }
  QF_LAUNCH_CHECK();
  return QF_OK;
}

static int build(qf_mesh* m, cudaStream_t st) {
  const int64_t F = m->n_faces, V = m->n_vertices;
  int* ob = m->d_flags;  // 6 ints of scratch before flags are needed
  bounds_init_kernel<<<1, 32, 0, st>>>(ob);
  bounds_kernel<<<kNumSMs * 2, 256, 0, st>>>(m->d_vertices, V, ob);
  bounds_finish_kernel<<<1, 1, 0, st>>>(ob, m->d_scene);
  QF_LAUNCH_CHECK();
  int blocks = (int)ceil_div(F, 256);
  tri_setup_kernel<<<blocks, 256, 0, st>>>(m->d_vertices, m->d_faces, F, m->d_scene, m->d_keys, m->d_idx, m->d_planes);
  QF_LAUNCH_CHECK();
  size_t tmp = m->sort_tmp_bytes;
  QF_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(m->d_sort_tmp, tmp, m->d_keys, m->d_keys_sorted, m->d_idx,
                                                m->d_idx_sorted, (int)F, 0, 63, st));
  tri_gather_kernel<<<blocks, 256, 0, st>>>(m->d_vertices, m->d_faces, F, m->d_idx_sorted, m->d_tris);
  QF_LAUNCH_CHECK();
  if (F == 1) {
    emit_tiny_root_kernel<<<1, 1, 0, st>>>(m->d_tris, m->d_scene, m->d_nodes);
    QF_LAUNCH_CHECK();
    m->n_nodes = 1;
  } else {
    QF_CUDA_CHECK(cudaMemsetAsync(m->d_flags, 0, sizeof(int) * F, st));
    karras_kernel<<<blocks, 256, 0, st>>>(m->d_keys_sorted, (int)F, m->d_left, m->d_right, m->d_parent,
                                          m->d_leaf_parent, m->d_first, m->d_last);
    fit_kernel<<<blocks, 256, 0, st>>>(m->d_tris, (int)F, m->d_scene, m->d_left, m->d_right, m->d_parent,
                                       m->d_leaf_parent, m->d_flags, m->d_ibox);
    emit_nodes_kernel<<<blocks, 256, 0, st>>>(m->d_tris, (int)F, m->d_scene, m->d_left, m->d_right, m->d_ibox, m->d_nodes);
    QF_LAUNCH_CHECK();
    m->n_nodes = F - 1;
  }
  // wide BVH for coherent packets: level-synchronous collapse, no host synchronisation
  if (m->d_wnodes) return collapse_wide(m, m->d_wnodes, (int)m->wide_capacity, st);
  return QF_OK;
}

template <typename T>
static int dev_alloc(T** p, size_t count, size_t* total) {
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  QF_CUDA_CHECK(cudaMalloc((void**)p, bytes));
  *total += bytes;
  return QF_OK;
}

// Build scratch (Morton keys, Karras hierarchy, per-node boxes, radix-sort and collapse queues): needed by build() only.
static int alloc_build_scratch(qf_mesh* m) {
  if (m->d_keys) return QF_OK;
  const size_t F = (size_t)m->n_faces;
  size_t before = m->bytes;
  int rc = QF_OK;
#define QF_A(call) if (rc == QF_OK) rc = (call)
  QF_A(dev_alloc(&m->d_keys, F, &m->bytes));
  QF_A(dev_alloc(&m->d_keys_sorted, F, &m->bytes));
  QF_A(dev_alloc(&m->d_idx, F, &m->bytes));
  QF_A(dev_alloc(&m->d_idx_sorted, F, &m->bytes));
  QF_A(dev_alloc(&m->d_left, F, &m->bytes));
  QF_A(dev_alloc(&m->d_right, F, &m->bytes));
  QF_A(dev_alloc(&m->d_parent, F, &m->bytes));
  QF_A(dev_alloc(&m->d_leaf_parent, F, &m->bytes));
  QF_A(dev_alloc(&m->d_first, F, &m->bytes));
  QF_A(dev_alloc(&m->d_last, F, &m->bytes));
  QF_A(dev_alloc(&m->d_flags, F + 8, &m->bytes));
  QF_A(dev_alloc(&m->d_ibox, 2 * F, &m->bytes));
  if (m->want_wide) QF_A(dev_alloc(&m->d_wqueue, (size_t)(F / 4 + 64), &m->bytes));
#undef QF_A
  if (rc == QF_OK) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, m->d_keys, m->d_keys_sorted, m->d_idx, m->d_idx_sorted, (int)F, 0, 63, (cudaStream_t)0);
    m->sort_tmp_bytes = tmp;
    if (cudaMalloc(&m->d_sort_tmp, tmp ? tmp : 16) != cudaSuccess) { set_error("qf_mesh: sort scratch alloc failed"); rc = QF_ERR_CUDA; }
    else m->bytes += tmp;
  }
  m->scratch_bytes = m->bytes - before;
  return rc;
}

static void free_build_scratch(qf_mesh* m) {
  void** ptrs[] = {(void**)&m->d_keys, (void**)&m->d_keys_sorted, (void**)&m->d_idx, (void**)&m->d_idx_sorted, (void**)&m->d_left,
                   (void**)&m->d_right, (void**)&m->d_parent, (void**)&m->d_leaf_parent, (void**)&m->d_first, (void**)&m->d_last,
                   (void**)&m->d_flags, (void**)&m->d_ibox, (void**)&m->d_wqueue, &m->d_sort_tmp};
  for (void** p : ptrs) if (*p) { cudaFree(*p); *p = nullptr; }
  m->bytes -= m->scratch_bytes;
  m->scratch_bytes = 0;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_mesh_create(const float* d_vertices, int64_t n_vertices, const int32_t* d_faces, int64_t n_faces,
                              void* stream, qf_mesh** out) {
  QF_REQUIRE(out != nullptr, "qf_mesh_create: out is NULL");
  QF_REQUIRE(d_vertices && d_faces && n_vertices > 0 && n_faces > 0, "qf_mesh_create: empty mesh (V=%lld F=%lld)",
             (long long)n_vertices, (long long)n_faces);
  QF_REQUIRE(n_faces < (1ll << 28), "qf_mesh_create: at most 2^28 faces");
  cudaStream_t st = (cudaStream_t)stream;
  qf_mesh* m = new qf_mesh();
  m->n_vertices = n_vertices;
  m->n_faces = n_faces;
  const size_t F = (size_t)n_faces;
  int rc = QF_OK;
#define QF_A(call) if (rc == QF_OK) rc = (call)
  QF_A(dev_alloc(&m->d_vertices, 3 * (size_t)n_vertices, &m->bytes));
  QF_A(dev_alloc(&m->d_faces, 3 * F, &m->bytes));
  QF_A(dev_alloc(&m->d_tris, 3 * F, &m->bytes));
  QF_A(dev_alloc(&m->d_planes, F, &m->bytes));
  QF_A(dev_alloc(&m->d_nodes, 4 * F, &m->bytes));
  QF_A(dev_alloc(&m->d_scene, 8, &m->bytes));
  QF_A(dev_alloc(&m->d_call_slots, 4 * kCallSlots, &m->bytes));
  // wide nodes: ~F/22 in practice (leaf nodes hold 17-32 triangles); room for F/8 + 64, the collapse falls back to the
  // binary tree if a pathological mesh needs more.  QF_WIDE_BVH=0 disables the wide tree.
  const bool want_wide = !(getenv("QF_WIDE_BVH") && atoi(getenv("QF_WIDE_BVH")) == 0);
  const int64_t queue_capacity = (int64_t)(F / 4 + 64);      // ints only; the node array is sized after a dry run below
  m->want_wide = want_wide;
  if (want_wide) QF_A(dev_alloc(&m->d_wstate, 8, &m->bytes));
  QF_A(alloc_build_scratch(m));
#undef QF_A
  if (rc != QF_OK) { qf_mesh_destroy(m); return rc; }
  cudaError_t e = cudaMemcpyAsync(m->d_faces, d_faces, sizeof(int32_t) * 3 * F, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_vertices, d_vertices, sizeof(float) * 3 * n_vertices, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) { set_error("qf_mesh_create: copy failed: %s", cudaGetErrorString(e)); qf_mesh_destroy(m); return QF_ERR_CUDA; }
  rc = build(m, st);      // binary tree (d_wnodes is still NULL)
  if (rc == QF_OK && want_wide) {
    // dry-run collapse: count the wide nodes, then allocate exactly that (+25 % for vertex updates, which re-collapse
    // into the same array) instead of a worst-case bound — 1 KB per node
    rc = collapse_wide(m, nullptr, (int)queue_capacity, st);
    int32_t h_state[8] = {0};
    if (rc == QF_OK) {
      e = cudaMemcpyAsync(h_state, m->d_wstate, sizeof(int32_t) * 5, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { set_error("qf_mesh_create: wide collapse failed: %s", cudaGetErrorString(e)); rc = QF_ERR_CUDA; }
    }
    if (rc == QF_OK && h_state[3] == 1) {
      m->wide_capacity = (int64_t)h_state[2] + h_state[2] / 4 + 16;
      if (m->wide_capacity > queue_capacity) m->wide_capacity = queue_capacity;
      rc = dev_alloc(&m->d_wnodes, 64 * (size_t)m->wide_capacity, &m->bytes);
      if (rc == QF_OK) rc = collapse_wide(m, m->d_wnodes, (int)m->wide_capacity, st);
    }
  }
  if (rc == QF_OK) {
    e = cudaMemcpyAsync(&m->h_pad, m->d_scene + 6, sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("qf_mesh_create: build failed: %s", cudaGetErrorString(e)); rc = QF_ERR_CUDA; }
  }
  if (rc != QF_OK) { qf_mesh_destroy(m); return rc; }
  // A mesh that is only rendered never rebuilds: give the build scratch back (97 of 302 bytes per triangle: 111 MB of the
  // 346 MB of a 1.15 M-triangle mesh).  The first qf_mesh_update_vertices allocates it again and keeps it from then on, so a
  // training loop that moves the vertices every step allocates once.  QF_MESH_KEEP_SCRATCH=1 keeps it from the start.
  if (!(getenv("QF_MESH_KEEP_SCRATCH") && atoi(getenv("QF_MESH_KEEP_SCRATCH")) != 0)) free_build_scratch(m);
  *out = m;
  return QF_OK;
}

extern "C" int qf_mesh_update_vertices(qf_mesh* m, const float* d_vertices, void* stream) {
  QF_REQUIRE(m && d_vertices, "qf_mesh_update_vertices: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  { int rc0 = alloc_build_scratch(m); if (rc0 != QF_OK) return rc0; }     // no-op while the scratch is held
  QF_CUDA_CHECK(cudaMemcpyAsync(m->d_vertices, d_vertices, sizeof(float) * 3 * m->n_vertices, cudaMemcpyDeviceToDevice, st));
  ++m->geometry_version;      // per-triangle caches derived from the vertices (baked path) are rebuilt on next use
  int rc = build(m, st);
  if (rc != QF_OK) return rc;
  QF_CUDA_CHECK(cudaMemcpyAsync(&m->h_pad, m->d_scene + 6, sizeof(float), cudaMemcpyDeviceToHost, st));
  return QF_OK;
}

extern "C" void qf_mesh_destroy(qf_mesh* m) {
  if (!m) return;
  void* ptrs[] = {m->d_vertices, m->d_faces, m->d_tris, m->d_planes, m->d_nodes, m->d_scene, m->d_keys, m->d_keys_sorted,
                  m->d_idx, m->d_idx_sorted, m->d_left, m->d_right, m->d_parent, m->d_leaf_parent, m->d_first, m->d_last,
                  m->d_flags, m->d_ibox, m->d_sort_tmp, m->d_call_slots, m->d_wnodes, m->d_wqueue, m->d_wstate, m->d_bary};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete m;
}

extern "C" int qf_mesh_set_restart_eps(qf_mesh* m, float eps) {
  QF_REQUIRE(m, "qf_mesh_set_restart_eps: NULL mesh");
  QF_REQUIRE(eps >= 0.f && eps == eps, "qf_mesh_set_restart_eps: eps=%f", eps);
  m->restart_eps = eps;
  return QF_OK;
}

extern "C" int qf_mesh_info(const qf_mesh* m, int64_t* info4, float* box_pad) {
  QF_REQUIRE(m, "qf_mesh_info: NULL mesh");
  if (info4) { info4[0] = m->n_faces; info4[1] = m->n_vertices; info4[2] = m->n_nodes; info4[3] = (int64_t)m->bytes; }
  if (box_pad) *box_pad = m->h_pad;
  return QF_OK;
}

extern "C" size_t qf_trace_workspace_bytes(int64_t n_rays) { return 256 + sizeof(int32_t) * (size_t)(n_rays > 0 ? n_rays : 0) + 256; }

extern "C" int qf_trace_firstk(const qf_mesh* m, const float* d_origins, const float* d_dirs, int64_t n_rays, int K,
                               int32_t* d_tri, float* d_t, int32_t* d_count, int32_t* d_total, void* d_workspace,
                               size_t workspace_bytes, void* stream) {
  QF_REQUIRE(m, "qf_trace_firstk: NULL mesh");
  QF_REQUIRE(K >= 1 && K <= QF_MAX_HITS, "qf_trace_firstk: K=%d outside [1,%d]", K, QF_MAX_HITS);
  QF_REQUIRE(n_rays >= 0, "qf_trace_firstk: n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return QF_OK;   // empty tensors carry NULL data pointers
  QF_REQUIRE(d_origins && d_dirs && d_tri && d_count, "qf_trace_firstk: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)ceil_div(n_rays, 128);
  // caller scratch: [slot: 4 ints | compacted ray ids]; without it the list always takes the thread-per-ray kernel
  const bool have_ws = d_workspace && workspace_bytes >= qf_trace_workspace_bytes(n_rays);
  int32_t* slot = have_ws ? (int32_t*)d_workspace : m->d_call_slots + 4 * (m->call_id++ % kCallSlots);
  int32_t* list = have_ws ? (int32_t*)d_workspace + 64 : nullptr;
  QF_CUDA_CHECK(cudaMemsetAsync(slot, 0, 4 * sizeof(int32_t), st));
  if (have_ws) {   // without scratch slot[0] = slot[1] = 0 reads as "coherent" (2*0 >= 0): thread-per-ray kernel only
    classify_rays_kernel<<<blocks, 128, 0, st>>>(m->d_scene, d_origins, d_dirs, n_rays, slot, list);
    // "no hit" defaults for the rays the refill kernel never writes: tri = -1, count = total = 0, t = +inf
    QF_CUDA_CHECK(cudaMemsetAsync(d_tri, 0xFF, sizeof(int32_t) * n_rays * K, st));
    QF_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(int32_t) * n_rays, st));
    if (d_total) QF_CUDA_CHECK(cudaMemsetAsync(d_total, 0, sizeof(int32_t) * n_rays, st));
    if (d_t) fill_inf_kernel<<<kNumSMs * 4, 256, 0, st>>>(d_t, n_rays * K);
  }
  const int pblocks = blocks < kNumSMs * 8 ? blocks : kNumSMs * 8;   // persistent grid of the refill kernel
#define QF_TRACE(KM, ALL)                                                                                                         \
  do {                                                                                                                            \
    trace_kernel<KM, ALL><<<blocks, 128, 0, st>>>(m->d_nodes, m->d_tris, d_origins, d_dirs, n_rays, K, d_tri, d_t, d_count, d_total, slot, m->d_wnodes, m->d_wstate, k_trav, eps); \
    if (have_ws) trace_refill_kernel<KM, ALL><<<pblocks, 128, 0, st>>>(m->d_nodes, m->d_tris, d_origins, d_dirs, n_rays, K, d_tri, d_t, d_count, d_total, slot, list, k_trav, eps); \
  } while (0)
  // epsilon-restart mode: collect the QF_MAX_HITS nearest raw hits, then keep K of them like the Embree restart loop
  const float eps = m->restart_eps;
  const int k_trav = eps > 0.f ? QF_MAX_HITS : K;
  // the untruncated total needs a traversal without distance culling
  // (an 8-slot shared-memory buffer, the fused frame's choice for K <= 8, is neutral on this path: r2i train step 1.44 vs 1.45 ms)
  if (d_total) { if (k_trav <= 8) QF_TRACE(HitBufReg<8>, true); else QF_TRACE(HitBufSmem, true); }
  else { if (k_trav <= 8) QF_TRACE(HitBufReg<8>, false); else QF_TRACE(HitBufSmem, false); }
#undef QF_TRACE
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" size_t qf_scan_workspace_bytes(int64_t n) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (int64_t*)nullptr, (int64_t*)nullptr, n + 1);
  return tmp + sizeof(int64_t) * (size_t)(n + 1) + 256;
}

extern "C" int qf_hits_offsets(const int32_t* d_count, int64_t n_rays, int64_t* d_offsets, void* d_workspace,
                               size_t workspace_bytes, void* stream) {
  QF_REQUIRE(n_rays >= 0 && d_offsets, "qf_hits_offsets: NULL offsets / negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) {   // empty count tensor (NULL pointer): offsets = [0]
    QF_CUDA_CHECK(cudaMemsetAsync(d_offsets, 0, sizeof(int64_t), st));
    return QF_OK;
  }
  QF_REQUIRE(d_count && d_workspace, "qf_hits_offsets: NULL argument");
  QF_REQUIRE(workspace_bytes >= qf_scan_workspace_bytes(n_rays), "qf_hits_offsets: workspace too small");
  int64_t* wide = (int64_t*)d_workspace;
  size_t wide_bytes = (sizeof(int64_t) * (size_t)(n_rays + 1) + 255) / 256 * 256;
  void* tmp = (char*)d_workspace + wide_bytes;
  size_t tmp_bytes = workspace_bytes - wide_bytes;
  widen_count_kernel<<<(int)ceil_div(n_rays + 1, 256), 256, 0, st>>>(d_count, n_rays, wide);
  QF_LAUNCH_CHECK();
  QF_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, wide, d_offsets, n_rays + 1, st));
  return QF_OK;
}

extern "C" int qf_hits_total(const int64_t* d_offsets, int64_t n_rays, int64_t* h_total, void* stream) {
  QF_REQUIRE(d_offsets && h_total, "qf_hits_total: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  QF_CUDA_CHECK(cudaMemcpyAsync(h_total, d_offsets + n_rays, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  QF_CUDA_CHECK(cudaStreamSynchronize(st));
  return QF_OK;
}

extern "C" int qf_hits_pack(const qf_mesh* m, const float* d_origins, const float* d_dirs, int64_t n_rays, int K,
                            const int32_t* d_tri, const int32_t* d_count, const int64_t* d_offsets, float* d_points,
                            float* d_vectors, int64_t* d_index_ray, float* d_depth, int64_t* d_index_tri,
                            float* d_origins_out, void* stream) {
  QF_REQUIRE(K >= 1 && K <= QF_MAX_HITS, "qf_hits_pack: K=%d outside [1,%d]", K, QF_MAX_HITS);
  if (n_rays == 0) return QF_OK;
  QF_REQUIRE(m && d_origins && d_dirs && d_tri && d_count && d_offsets, "qf_hits_pack: NULL input");
  QF_REQUIRE(d_points && d_vectors && d_index_ray && d_depth && d_index_tri && d_origins_out, "qf_hits_pack: NULL output");
  hits_pack_kernel<<<(int)ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
      m->d_planes, d_origins, d_dirs, n_rays, K, d_tri, d_count, d_offsets, d_points, d_vectors, d_index_ray, d_depth,
      d_index_tri, d_origins_out);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_hits_resort(const int64_t* d_index_ray, const float* d_depth, int64_t n_hits, int64_t* d_perm,
                              uint8_t* d_boundary, void* stream) {
  if (n_hits == 0) return QF_OK;
  QF_REQUIRE(d_index_ray && d_depth && d_perm && d_boundary, "qf_hits_resort: NULL argument");
  hits_resort_kernel<<<(int)ceil_div(n_hits, 256), 256, 0, (cudaStream_t)stream>>>(d_index_ray, d_depth, n_hits, d_perm, d_boundary);
  QF_LAUNCH_CHECK();
  return QF_OK;
}
