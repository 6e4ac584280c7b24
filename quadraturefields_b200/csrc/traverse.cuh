// First-K all-hits BVH traversal (device templates shared by bvh.cu and render.cu).
#pragma once
#include "geom.cuh"

namespace qf {

// ---------------------------------------------------------------- traversal
template <int KMAX>
struct HitBuf {
  float t[KMAX];
  int id[KMAX];
  int cnt;
};

template <int KMAX>
__device__ __forceinline__ void insert_hit(HitBuf<KMAX>& hb, int K, float t, int id) {
  // carry-insertion with static register indexing; order key is (t, id)
  float ct = t;
  int ci = id;
#pragma unroll
  for (int s = 0; s < KMAX; ++s) {
    if (s < hb.cnt) {
      bool less = (ct < hb.t[s]) || (ct == hb.t[s] && ci < hb.id[s]);
      if (less) {
        float tt = hb.t[s]; int ti = hb.id[s];
        hb.t[s] = ct; hb.id[s] = ci;
        ct = tt; ci = ti;
      }
    } else if (s == hb.cnt && s < K) {
      hb.t[s] = ct; hb.id[s] = ci;
    }
  }
  if (hb.cnt < K) hb.cnt++;
}

template <int KMAX>
__device__ __forceinline__ void leaf_intersect(const Ray& r, const float4* __restrict__ tris, int ref, float pad, int K,
                                               HitBuf<KMAX>& hb, int& total) {
  int inner = ~ref;
  int first = inner & 0x0FFFFFFF, count = ((inner >> 28) & 3) + 1;
  for (int j = 0; j < count; ++j) {
    const float4* p = tris + 3 * (int64_t)(first + j);
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    float t;
    if (ray_triangle(r, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, pad, t)) {
      ++total;
      insert_hit<KMAX>(hb, K, t, __float_as_int(a.w));
    }
  }
}

template <int KMAX, bool CULL = true>
__device__ __forceinline__ void traverse(const Ray& r, const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                         float pad, int K, HitBuf<KMAX>& hb, int& total) {
  int stack[128];  // depth <= 64 key bits + 32 index bits
  int sp = 0;
  int cur = 0;
  hb.cnt = 0;
  total = 0;
#pragma unroll
  for (int s = 0; s < KMAX; ++s) { hb.t[s] = __int_as_float(0x7f800000); hb.id[s] = -1; }
  while (true) {
    const float4* np = nodes + 4 * (int64_t)cur;
    float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
    int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
    float tn0, tf0, tn1, tf1;
    // K-th smallest t so far (+inf until the buffer is full: unused slots stay +inf)
    float tcull = hb.t[KMAX - 1];
    if (!CULL) tcull = __int_as_float(0x7f800000);   // counting every hit: no distance culling
    else if (K < KMAX) {
#pragma unroll
      for (int s = 0; s < KMAX - 1; ++s) if (s == K - 1) tcull = hb.t[s];
    }
    bool h0 = (r0 != kEmptyRef) && slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tn0, tf0) && (tn0 <= tcull);
    bool h1 = (r1 != kEmptyRef) && slab(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tn1, tf1) && (tn1 <= tcull);
    // leaves are intersected immediately (nearer first so the cull distance tightens early)
    if (h0 && h1 && tn1 < tn0) { int tr = r0; r0 = r1; r1 = tr; }
    if (h0 && r0 < 0) { leaf_intersect<KMAX>(r, tris, r0, pad, K, hb, total); h0 = false; }
    if (h1 && r1 < 0) { leaf_intersect<KMAX>(r, tris, r1, pad, K, hb, total); h1 = false; }
    if (h0 && h1) { stack[sp++] = r1; cur = r0; }
    else if (h0) cur = r0;
    else if (h1) cur = r1;
    else {
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
}


}  // namespace qf
