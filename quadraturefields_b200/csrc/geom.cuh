// Ray / triangle / box predicates.  fp32, ONE rounding per operation, no FMA contraction
// (explicit __f*_rn intrinsics), operand order fixed — DESIGN.md §3.1.  The brute-force oracle
// evaluates the same expressions in numpy, so hit ids and counts are bit-exact.
#pragma once
#include "common.cuh"

namespace qf {

struct Ray {
  float ox, oy, oz, dx, dy, dz, ix, iy, iz;  // origin, direction, 1/direction (IEEE divide)
};

__device__ __forceinline__ Ray make_ray(const float* __restrict__ o, const float* __restrict__ d, int64_t i) {
  Ray r;
  r.ox = __ldg(o + 3 * i + 0); r.oy = __ldg(o + 3 * i + 1); r.oz = __ldg(o + 3 * i + 2);
  r.dx = __ldg(d + 3 * i + 0); r.dy = __ldg(d + 3 * i + 1); r.dz = __ldg(d + 3 * i + 2);
  r.ix = __fdiv_rn(1.0f, r.dx); r.iy = __fdiv_rn(1.0f, r.dy); r.iz = __fdiv_rn(1.0f, r.dz);
  return r;
}

// Slab test.  tn = max(near_x, near_y, max(near_z, 0)), tf = min(far_x, far_y, far_z) with
// fminf/fmaxf NaN semantics (np.fmin/np.fmax in the oracle).  Monotone in the box: a box that
// contains another passes whenever the inner one does, and its tn is <= the inner tn.
__device__ __forceinline__ bool slab(const Ray& r, float lx, float ly, float lz, float hx, float hy, float hz,
                                     float& tn, float& tf) {
  float ax0 = __fmul_rn(__fsub_rn(lx, r.ox), r.ix), ax1 = __fmul_rn(__fsub_rn(hx, r.ox), r.ix);
  float ay0 = __fmul_rn(__fsub_rn(ly, r.oy), r.iy), ay1 = __fmul_rn(__fsub_rn(hy, r.oy), r.iy);
  float az0 = __fmul_rn(__fsub_rn(lz, r.oz), r.iz), az1 = __fmul_rn(__fsub_rn(hz, r.oz), r.iz);
  tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
  tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
  return tn <= tf;
}

// The same test when the signs of the direction components are known at compile time (bit k of SIGN set: component k
// negative; all three finite and non-zero).  For ix > 0, lo <= hi gives (lo-o)*ix <= (hi-o)*ix because every rounding
// is monotone, so fminf(ax0, ax1) IS ax0 (and ax1 for ix < 0, rounding being sign-symmetric): picking the near / far
// plane up front returns bit-identical tn, tf with 3 min/max per box instead of 9.
template <int SIGN>
__device__ __forceinline__ bool slab_signed(const Ray& r, float lx, float ly, float lz, float hx, float hy, float hz,
                                            float& tn, float& tf) {
  const float nx = (SIGN & 1) ? hx : lx, fx = (SIGN & 1) ? lx : hx;
  const float ny = (SIGN & 2) ? hy : ly, fy = (SIGN & 2) ? ly : hy;
  const float nz = (SIGN & 4) ? hz : lz, fz = (SIGN & 4) ? lz : hz;
  const float anx = __fmul_rn(__fsub_rn(nx, r.ox), r.ix), afx = __fmul_rn(__fsub_rn(fx, r.ox), r.ix);
  const float any_ = __fmul_rn(__fsub_rn(ny, r.oy), r.iy), afy = __fmul_rn(__fsub_rn(fy, r.oy), r.iy);
  const float anz = __fmul_rn(__fsub_rn(nz, r.oz), r.iz), afz = __fmul_rn(__fsub_rn(fz, r.oz), r.iz);
  tn = fmaxf(fmaxf(anx, any_), fmaxf(anz, 0.0f));
  tf = fminf(fminf(afx, afy), afz);
  return tn <= tf;
}

__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Möller–Trumbore.  The full hit predicate is
//     mt_hit  AND  slab(ray, padded box of the triangle) passes  AND  t >= tn of that slab test.
// In the BVH the padded triangle box IS the child box stored in the parent node, so the slab part has already
// been evaluated (bit-identically) when a leaf reference is reached; `tn` travels with the reference.
__device__ __forceinline__ bool ray_triangle_mt(const Ray& r, float v0x, float v0y, float v0z, float v1x, float v1y,
                                                float v1z, float v2x, float v2y, float v2z, float tn, float& t_out) {
  float e1x = __fsub_rn(v1x, v0x), e1y = __fsub_rn(v1y, v0y), e1z = __fsub_rn(v1z, v0z);
  float e2x = __fsub_rn(v2x, v0x), e2y = __fsub_rn(v2y, v0y), e2z = __fsub_rn(v2z, v0z);
  float px = __fsub_rn(__fmul_rn(r.dy, e2z), __fmul_rn(r.dz, e2y));
  float py = __fsub_rn(__fmul_rn(r.dz, e2x), __fmul_rn(r.dx, e2z));
  float pz = __fsub_rn(__fmul_rn(r.dx, e2y), __fmul_rn(r.dy, e2x));
  float det = __fadd_rn(__fadd_rn(__fmul_rn(e1x, px), __fmul_rn(e1y, py)), __fmul_rn(e1z, pz));
  if (det == 0.0f) return false;
  float inv = __fdiv_rn(1.0f, det);
  float tx = __fsub_rn(r.ox, v0x), ty = __fsub_rn(r.oy, v0y), tz = __fsub_rn(r.oz, v0z);
  float u = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, px), __fmul_rn(ty, py)), __fmul_rn(tz, pz)), inv);
  if (!(u >= 0.0f)) return false;
  float qx = __fsub_rn(__fmul_rn(ty, e1z), __fmul_rn(tz, e1y));
  float qy = __fsub_rn(__fmul_rn(tz, e1x), __fmul_rn(tx, e1z));
  float qz = __fsub_rn(__fmul_rn(tx, e1y), __fmul_rn(ty, e1x));
  float v = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.dx, qx), __fmul_rn(r.dy, qy)), __fmul_rn(r.dz, qz)), inv);
  if (!((v >= 0.0f) && (__fadd_rn(u, v) <= 1.0f))) return false;
  float t = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(e2x, qx), __fmul_rn(e2y, qy)), __fmul_rn(e2z, qz)), inv);
  if (!((t > 0.0f) && (t >= tn))) return false;
  t_out = t;
  return true;
}

// mesh_utils.py:33-40: d=-(n.v); t=-((n.o)+d)/(n.r); t=|t|; psi=o+t r.   plane = (n.xyz, d)
__device__ __forceinline__ void plane_hit(const Ray& r, float4 plane, float& px, float& py, float& pz) {
  float no = __fadd_rn(__fadd_rn(__fmul_rn(plane.x, r.ox), __fmul_rn(plane.y, r.oy)), __fmul_rn(plane.z, r.oz));
  float nr = __fadd_rn(__fadd_rn(__fmul_rn(plane.x, r.dx), __fmul_rn(plane.y, r.dy)), __fmul_rn(plane.z, r.dz));
  float t = fabsf(__fdiv_rn(-__fadd_rn(no, plane.w), nr));
  px = __fadd_rn(r.ox, __fmul_rn(t, r.dx));
  py = __fadd_rn(r.oy, __fmul_rn(t, r.dy));
  pz = __fadd_rn(r.oz, __fmul_rn(t, r.dz));
}

__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// ---- BVH node reference encoding -------------------------------------------------------------
//  ref >= 0           internal node index
//  ref <  0           leaf: ~ref = position of ONE triangle in the Morton-sorted triangle array; the child box
//                     stored beside the reference is exactly that triangle's padded box
//  ref == kEmptyRef   nothing (only in the single-triangle mesh's root)
constexpr int kEmptyRef = (int)0x80000000;

}  // namespace qf
