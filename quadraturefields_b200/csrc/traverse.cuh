// First-K all-hits BVH traversal (device templates shared by bvh.cu and render.cu).
//
// Two traversal strategies over the same two-child LBVH, both exact:
//   * traverse_packet — the 32 rays of a warp walk the tree TOGETHER: one warp-uniform stack in shared memory,
//     node and triangle records fetched at one address for the whole warp (a single broadcast transaction),
//     a child is entered when any lane's slab test passes.  Control flow is warp-uniform, so coherent rays
//     (camera rays of neighbouring pixels) run without SIMT divergence.
//   * traverse_single — per-lane while-while traversal (Aila & Laine) with a private stack, for incoherent rays.
// trace kernels pick per warp: packets when all lanes share the origin and lie in a narrow cone.
#pragma once
#include "geom.cuh"

namespace qf {

// ---- per-ray K-buffers.  Both keep the K smallest hits by (t, id) and expose the same interface:
//   init(K) | insert(t,id) | cull_distance() (K-th smallest t, +inf until K hits) | count(K) | for_each(K, f(j,t,id)) in order.
//
// HitBufReg<KMAX> (K <= 8): sorted, entirely in registers.  Every access uses a compile-time slot index (any
// `slot == runtime value` guard lets the compiler fold the access into a dynamically indexed one and the arrays fall
// back to local memory).  The buffer always has KMAX slots; when K < KMAX the first KMAX-K slots hold -inf sentinels
// that sort before every real hit, so the K real slots are [KMAX-K, KMAX) and t[KMAX-1] is the K-th smallest hit.
template <int KMAX>
struct HitBufReg {
  static constexpr int kSmemSlots = 0;
  float t[KMAX];
  int id[KMAX];
  __device__ __forceinline__ HitBufReg(float*, int*, int) {}
  __device__ __forceinline__ void init(int K) {
#pragma unroll
    for (int s = 0; s < KMAX; ++s) {
      const bool real = s >= KMAX - K;
      t[s] = __int_as_float(real ? 0x7f800000 : 0xff800000);
      id[s] = real ? 0x7fffffff : -1;
    }
  }
  __device__ __forceinline__ float cull_distance() const { return t[KMAX - 1]; }
  __device__ __forceinline__ int count(int K) const {
    int c = 0;
#pragma unroll
    for (int s = 0; s < KMAX; ++s) c += (s >= KMAX - K && t[s] != __int_as_float(0x7f800000)) ? 1 : 0;
    return c;
  }
  // carry-insertion: the largest element falls off the end
  __device__ __forceinline__ void insert(float nt, int nid) {
    float ct = nt;
    int ci = nid;
#pragma unroll
    for (int s = 0; s < KMAX; ++s) {
      const bool less = (ct < t[s]) || (ct == t[s] && ci < id[s]);
      const float tt = t[s];
      const int ti = id[s];
      t[s] = less ? ct : tt;
      id[s] = less ? ci : ti;
      ct = less ? tt : ct;
      ci = less ? ti : ci;
    }
  }
  template <typename F>
  __device__ __forceinline__ void for_each(int K, F f) {
#pragma unroll
    for (int s = 0; s < KMAX; ++s)
      if (s >= KMAX - K && t[s] != __int_as_float(0x7f800000)) f(s - (KMAX - K), t[s], id[s]);
  }
  __device__ __forceinline__ void restart_filter(float, int) {}   // the epsilon-restart mode always runs on HitBufSmem
};

// HitBufSmemT<SLOTS> (HitBufSmem = 32 slots for 8 < K <= 32; 8 slots for K <= 8 on the fused frame path since r2i: c2 trace
// 0.114 -> 0.104 ms against HitBufReg<8>, whose carry insertion costs ~70 instructions per hit of any lane of the warp):
// unsorted in shared memory (slot-major, 128 threads per CTA, conflict free), O(1) append
// while fewer than K hits are known — the common case, since K is chosen above the deepest ray — and replace-the-maximum
// once full; sorted once at the end.  Keeps the traversal kernels at ~60 registers instead of 115 for a 32-slot
// register buffer whose insertion costs 32 compare-exchange steps per hit.
template <int SLOTS>
struct HitBufSmemT {
  static constexpr int kSmemSlots = SLOTS;
  float* st;
  int* si;
  int cnt, K, imax, idmax;
  float tmax;
  __device__ __forceinline__ HitBufSmemT(float* t, int* i, int tid) : st(t + tid), si(i + tid) {}
  __device__ __forceinline__ void init(int K_) { cnt = 0; K = K_; imax = 0; idmax = 0x7fffffff; tmax = __int_as_float(0x7f800000); }
  __device__ __forceinline__ float cull_distance() const { return cnt == K ? tmax : __int_as_float(0x7f800000); }
  __device__ __forceinline__ int count(int) const { return cnt; }
  __device__ __forceinline__ void find_max() {
    tmax = st[0]; idmax = si[0]; imax = 0;
    for (int s = 1; s < K; ++s) {
      const float a = st[s * 128];
      const int b = si[s * 128];
      if (a > tmax || (a == tmax && b > idmax)) { tmax = a; idmax = b; imax = s; }
    }
  }
  __device__ __forceinline__ void insert(float nt, int nid) {
    if (cnt < K) {
      st[cnt * 128] = nt; si[cnt * 128] = nid;
      if (++cnt == K) find_max();
    } else if (nt < tmax || (nt == tmax && nid < idmax)) {
      st[imax * 128] = nt; si[imax * 128] = nid;
      find_max();
    }
  }
  template <typename F>
  __device__ __forceinline__ void for_each(int, F f) {
    for (int s = 1; s < cnt; ++s) {   // insertion sort by (t, id)
      const float a = st[s * 128];
      const int b = si[s * 128];
      int q = s;
      while (q > 0 && (st[(q - 1) * 128] > a || (st[(q - 1) * 128] == a && si[(q - 1) * 128] > b))) {
        st[q * 128] = st[(q - 1) * 128]; si[q * 128] = si[(q - 1) * 128];
        --q;
      }
      st[q * 128] = a; si[q * 128] = b;
    }
    for (int s = 0; s < cnt; ++s) f(s, st[s * 128], si[s * 128]);
  }
  // The SHIPPED reference intersector (trimesh + Embree, mesh_utils.py:223,350-354, SURVEY a2'): a first-hit query
  // repeated up to K times, each restarted eps beyond the previous hit — i.e. of the hits in (t, id) order, one is kept
  // iff it lies more than eps (fp32 difference) behind the last KEPT one, until k_out are kept.  Applied to the <= 32
  // nearest raw hits this buffer collected; afterwards count() / for_each() see the kept hits only.
  __device__ __forceinline__ void restart_filter(float eps, int k_out) {
    for_each(0, [](int, float, int) {});          // sort
    float last = __int_as_float(0xff800000);
    int w = 0;
    for (int s = 0; s < cnt && w < k_out; ++s) {
      const float ts = st[s * 128];
      if (__fsub_rn(ts, last) > eps) { st[w * 128] = ts; si[w * 128] = si[s * 128]; last = ts; ++w; }
    }
    cnt = w;
  }
};
using HitBufSmem = HitBufSmemT<QF_MAX_HITS>;

// one triangle (leaf reference) whose padded box passed the slab test with entry distance tn
template <class HB>
__device__ __forceinline__ void leaf_intersect(const Ray& r, const float4* __restrict__ tris, int ref, float tn, HB& hb, int& total) {
  const float4* p = tris + 3 * (int64_t)(~ref);
  float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  float t;
  if (ray_triangle_mt(r, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, tn, t)) {
    ++total;
    hb.insert(t, __float_as_int(a.w));
  }
}

constexpr int kDoneRef = 0x7fffffff;
constexpr int kStackDepth = 96;   // LBVH depth <= 64 Morton bits + 32 index bits

// One internal-node step of the per-lane traversal (shared by traverse_single and the refill kernel).
template <class HB, bool CULL>
__device__ __forceinline__ void single_node_step(const Ray& r, const float4* __restrict__ nodes, const HB& hb,
                                                 int* sref, float* stn, int& sp, int& cur, float& cur_tn) {
  const float inf = __int_as_float(0x7f800000);
  const float4* np = nodes + 4 * (int64_t)cur;
  float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
  int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
  float tn0, tf0, tn1, tf1;
  const float tcull = CULL ? hb.cull_distance() : inf;
  bool h0 = (r0 != kEmptyRef) && slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tn0, tf0) && (tn0 <= tcull);
  bool h1 = (r1 != kEmptyRef) && slab(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tn1, tf1) && (tn1 <= tcull);
  if (h0 && h1) {
    const bool swap = tn1 < tn0;
    sref[sp] = swap ? r0 : r1;
    stn[sp++] = swap ? tn0 : tn1;
    cur = swap ? r1 : r0;
    cur_tn = swap ? tn1 : tn0;
  } else if (h0) { cur = r0; cur_tn = tn0; }
  else if (h1) { cur = r1; cur_tn = tn1; }
  else if (sp) { cur = sref[--sp]; cur_tn = stn[sp]; }
  else cur = kDoneRef;
}

// ---------------------------------------------------------------- per-lane while-while traversal
template <class HB, bool CULL = true>
__device__ __forceinline__ void traverse_single(const Ray& r, const float4* __restrict__ nodes,
                                                const float4* __restrict__ tris, int K, HB& hb, int& total) {
  int sref[kStackDepth];
  float stn[kStackDepth];
  int sp = 0;
  int cur = 0;
  float cur_tn = 0.f;
  total = 0;
  hb.init(K);
  const float inf = __int_as_float(0x7f800000);
  while (cur != kDoneRef) {
    while (cur >= 0 && cur != kDoneRef)     // ---- internal nodes, until this lane holds a leaf
      single_node_step<HB, CULL>(r, nodes, hb, sref, stn, sp, cur, cur_tn);
    while (cur < 0) {                        // ---- leaves, warp reconverged (kDoneRef is positive)
      const float tcull = CULL ? hb.cull_distance() : inf;
      if (cur_tn <= tcull) leaf_intersect<HB>(r, tris, cur, cur_tn, hb, total);
      if (sp) { cur = sref[--sp]; cur_tn = stn[sp]; }
      else cur = kDoneRef;
    }
  }
}

// ---------------------------------------------------------------- warp packet traversal
// `wstack`: kStackDepth ints of shared memory private to the warp.  Lanes with active == false never vote.
// SIGN >= 0: every active lane's direction has finite non-zero components with the sign pattern SIGN (see slab_signed).
template <class HB, bool CULL = true, int SIGN = -1>
__device__ __forceinline__ void traverse_packet(const Ray& r, bool active, const float4* __restrict__ nodes,
                                                const float4* __restrict__ tris, int K, HB& hb, int& total,
                                                int* __restrict__ wstack) {
  int sp = 0;
  int cur = 0;   // warp-uniform
  total = 0;
  hb.init(K);
  const float inf = __int_as_float(0x7f800000);
  while (true) {
    const float4* np = nodes + 4 * (int64_t)cur;
    float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
    const int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
    float tn0, tf0, tn1, tf1;
    const float tcull = CULL ? hb.cull_distance() : inf;
    // both slab tests are evaluated unconditionally (the warp issues them anyway) and masked afterwards: no
    // divergence regions around them
    bool s0, s1;
    if (SIGN < 0) {
      s0 = slab(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tn0, tf0);
      s1 = slab(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tn1, tf1);
    } else {
      s0 = slab_signed<SIGN & 7>(r, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tn0, tf0);
      s1 = slab_signed<SIGN & 7>(r, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, tn1, tf1);
    }
    bool h0 = s0 & active & (r0 != kEmptyRef) & (tn0 <= tcull);
    bool h1 = s1 & active & (r1 != kEmptyRef) & (tn1 <= tcull);
    // leaf children: their box is the triangle's own box, so the lanes that pass go straight to Möller–Trumbore
    if (r0 < 0) { if (h0) leaf_intersect<HB>(r, tris, r0, tn0, hb, total); h0 = false; }
    if (r1 < 0) { if (h1 && tn1 <= (CULL ? hb.cull_distance() : inf)) leaf_intersect<HB>(r, tris, r1, tn1, hb, total); h1 = false; }
    const unsigned b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
    if (b0 && b1) {
      // order by majority preference of the lanes that hit both (nearer child first)
      const unsigned both = b0 & b1;
      const unsigned pref1 = __ballot_sync(0xffffffffu, h0 && h1 && tn1 < tn0);
      const bool first1 = both ? (2 * __popc(pref1) > __popc(both)) : (__popc(b1) > __popc(b0));
      wstack[sp++] = first1 ? r0 : r1;
      cur = first1 ? r1 : r0;
    } else if (b0) cur = r0;
    else if (b1) cur = r1;
    else {
      if (sp == 0) break;
      cur = wstack[--sp];
    }
  }
}

// ---------------------------------------------------------------- warp packet traversal of the WIDE BVH
// A wide node has up to 32 children (bvh.cu: collapse of the binary LBVH); child c is the record pair
//   wnodes[(node * 32 + c) * 2 + 0] = (lo.x, lo.y, lo.z, hi.x)     wnodes[... + 1] = (hi.y, hi.z, ref, -)
// with ref >= 0 a wide node, ref < 0 ONE triangle (~ref = position in the Morton-sorted array; the box is that triangle's
// padded box, exactly), kEmptyRef nothing.  The packet's 32 lanes test the 32 children of a node AT ONCE, lane c testing
// child c against the whole packet: the rays share their origin, so with m = fl(plane - o) (one value for the packet)
// every ray's fl(m * i_r) lies between fl(m * i_min) and fl(m * i_max) — rounding is monotone and the operation is the
// very one of the per-ray slab test — hence
//     tn_lo = max_axes(min(m_near * i_min, m_near * i_max), 0) <= tn_r    and    tf_hi = min_axes(max(m_far * i_min, m_far * i_max)) >= tf_r
// for every ray r of the packet, with NO epsilon: a child is skipped only if no ray's own test could pass, and by
// monotonicity in the box no descendant triangle's either.  Triangles are then tested per lane with the exact predicate
// (slab of the triangle's own box, Möller–Trumbore, t >= tn), so results are bit-identical to the binary traversal and to
// the brute force.  One node visit replaces ~5 levels of two-box visits in which all 32 lanes did the same test.
// Entries of shared memory per warp.  A visit pops one entry and pushes <= 32, so a tree of L wide levels needs at most
// 31 L + 1 entries: 192 hold the kWideLevels = 6 levels the collapse allows (32^5 leaf nodes of ~22 triangles each is far
// beyond the 2^28-face limit of a mesh; deeper, degenerate trees clear the ok flag and are traversed as binary trees).
// 192 rather than 256 (8 levels) lets a sixth K=32 CTA fit on the SM (35.8 KB of shared memory each).
constexpr int kWideStack = 192;

#ifdef QF_TRACE_STATS   // diagnostics build only (tools/diag_trace_stats.py): per-packet work counters of the wide traversal
static __device__ unsigned long long g_trace_stats[8];
#define QF_STAT(i, v) (st_##i += (v))
#else
#define QF_STAT(i, v) ((void)0)
#endif

template <class HB, bool CULL, int SIGN>
__device__ __forceinline__ void traverse_packet_wide(const Ray& r, bool active, const float4* __restrict__ wnodes,
                                                     const float4* __restrict__ tris, int K, HB& hb, int& total,
                                                     int* __restrict__ wstack) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const float inf = __int_as_float(0x7f800000);
  // packet bounds of the reciprocal directions (all active lanes share the sign pattern SIGN and the origin)
  float ixa = active ? r.ix : inf, ixb = active ? r.ix : -inf;
  float iya = active ? r.iy : inf, iyb = active ? r.iy : -inf;
  float iza = active ? r.iz : inf, izb = active ? r.iz : -inf;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ixa = fminf(ixa, __shfl_xor_sync(0xffffffffu, ixa, o)); ixb = fmaxf(ixb, __shfl_xor_sync(0xffffffffu, ixb, o));
    iya = fminf(iya, __shfl_xor_sync(0xffffffffu, iya, o)); iyb = fmaxf(iyb, __shfl_xor_sync(0xffffffffu, iyb, o));
    iza = fminf(iza, __shfl_xor_sync(0xffffffffu, iza, o)); izb = fmaxf(izb, __shfl_xor_sync(0xffffffffu, izb, o));
  }
  const int src = __ffs(__ballot_sync(0xffffffffu, active)) - 1;
  const float ox = __shfl_sync(0xffffffffu, r.ox, src), oy = __shfl_sync(0xffffffffu, r.oy, src), oz = __shfl_sync(0xffffffffu, r.oz, src);
  total = 0;
  hb.init(K);
  int sp = 1;
  if (lane == 0) wstack[0] = 0;
  __syncwarp();
#ifdef QF_TRACE_STATS
  unsigned st_1 = 0, st_2 = 0, st_3 = 0, st_4 = 0, st_6 = 0;
#endif
  while (sp > 0) {
    QF_STAT(1, 1);
    const int node = wstack[--sp];
    __syncwarp();                      // every lane has read the entry before a push may overwrite it
    const float4* rec = wnodes + ((int64_t)node * 32 + lane) * 2;
    const float4 a = __ldg(rec), b = __ldg(rec + 1);
    const int ref = __float_as_int(b.z);
    // child `lane` against the packet
    const float nx = (SIGN & 1) ? a.w : a.x, fx = (SIGN & 1) ? a.x : a.w;
    const float ny = (SIGN & 2) ? b.x : a.y, fy = (SIGN & 2) ? a.y : b.x;
    const float nz = (SIGN & 4) ? b.y : a.z, fz = (SIGN & 4) ? a.z : b.y;
    const float mnx = __fsub_rn(nx, ox), mfx = __fsub_rn(fx, ox);
    const float mny = __fsub_rn(ny, oy), mfy = __fsub_rn(fy, oy);
    const float mnz = __fsub_rn(nz, oz), mfz = __fsub_rn(fz, oz);
    const float tnx = fminf(__fmul_rn(mnx, ixa), __fmul_rn(mnx, ixb)), tfx = fmaxf(__fmul_rn(mfx, ixa), __fmul_rn(mfx, ixb));
    const float tny = fminf(__fmul_rn(mny, iya), __fmul_rn(mny, iyb)), tfy = fmaxf(__fmul_rn(mfy, iya), __fmul_rn(mfy, iyb));
    const float tnz = fminf(__fmul_rn(mnz, iza), __fmul_rn(mnz, izb)), tfz = fmaxf(__fmul_rn(mfz, iza), __fmul_rn(mfz, izb));
    const float tn_lo = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f)), tf_hi = fminf(fminf(tfx, tfy), tfz);
    bool hit = (ref != kEmptyRef) && (tn_lo <= tf_hi);
    if (CULL) {
      // a child further than the K-th hit of EVERY lane cannot contribute (t >= tn_r >= tn_lo > cull distance)
      const float mine = active ? hb.cull_distance() : 0.0f;
      const float tcm = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(mine)));   // distances are >= 0: int order
      hit = hit && (tn_lo <= tcm);
    }
    const unsigned m_int = __ballot_sync(0xffffffffu, hit && ref >= 0);
    unsigned m_tri = __ballot_sync(0xffffffffu, hit && ref < 0);
    if (hit && ref >= 0) wstack[sp + __popc(m_int & lt)] = ref;
    sp += __popc(m_int);
    QF_STAT(6, __popc(m_int));
    __syncwarp();
    while (m_tri) {
      QF_STAT(2, 1);
      const int c = __ffs(m_tri) - 1;
      m_tri &= m_tri - 1;
      const float lx = __shfl_sync(0xffffffffu, a.x, c), ly = __shfl_sync(0xffffffffu, a.y, c), lz = __shfl_sync(0xffffffffu, a.z, c);
      const float hx = __shfl_sync(0xffffffffu, a.w, c), hy = __shfl_sync(0xffffffffu, b.x, c), hz = __shfl_sync(0xffffffffu, b.y, c);
      const int tref = __shfl_sync(0xffffffffu, ref, c);
      float tn, tf;
      const bool s = slab_signed<SIGN & 7>(r, lx, ly, lz, hx, hy, hz, tn, tf);
#ifdef QF_TRACE_STATS
      {
        const unsigned pm = __ballot_sync(0xffffffffu, s && active && tn <= (CULL ? hb.cull_distance() : inf));
        st_3 += pm ? 1 : 0;
        st_4 += __popc(pm);
      }
#endif
      if (s && active && tn <= (CULL ? hb.cull_distance() : inf)) leaf_intersect<HB>(r, tris, tref, tn, hb, total);
    }
  }
#ifdef QF_TRACE_STATS
  {
    const unsigned hits = __reduce_add_sync(0xffffffffu, (unsigned)total);
    if (lane == 0) {
      atomicAdd(&g_trace_stats[0], 1ull); atomicAdd(&g_trace_stats[1], (unsigned long long)st_1);
      atomicAdd(&g_trace_stats[2], (unsigned long long)st_2); atomicAdd(&g_trace_stats[3], (unsigned long long)st_3);
      atomicAdd(&g_trace_stats[4], (unsigned long long)st_4); atomicAdd(&g_trace_stats[5], (unsigned long long)hits);
      atomicAdd(&g_trace_stats[6], (unsigned long long)st_6); atomicAdd(&g_trace_stats[7], st_1 > 1 ? 1ull : 0ull);
    }
  }
#endif
}

// Warp-uniform sign pattern of the direction components (bit k: component k negative) when every active lane has
// finite, non-zero components of the same signs; -1 otherwise.
__device__ __forceinline__ int warp_sign_pattern(const Ray& r, bool active) {
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if (m == 0) return -1;
  const bool fin = isfinite(r.ix) && isfinite(r.iy) && isfinite(r.iz) && r.ix != 0.f && r.iy != 0.f && r.iz != 0.f;
  const int code = (r.ix < 0.f ? 1 : 0) | (r.iy < 0.f ? 2 : 0) | (r.iz < 0.f ? 4 : 0);
  const int first = __shfl_sync(0xffffffffu, code, __ffs(m) - 1);
  const bool ok = !active || (fin && code == first);
  return __all_sync(0xffffffffu, ok) ? first : -1;
}

// All lanes of the warp share one origin and their directions lie within ~2.2 degrees of the first lane's (a
// 32-pixel strip of an 800-wide NeRF-synthetic frame spans 1.6 degrees): the packet then touches few more nodes
// than a single ray.  Warp-uniform result.
__device__ __forceinline__ bool warp_is_coherent(const Ray& r, bool active) {
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if (m == 0) return false;
  const int src = __ffs(m) - 1;
  const float ox = __shfl_sync(0xffffffffu, r.ox, src), oy = __shfl_sync(0xffffffffu, r.oy, src), oz = __shfl_sync(0xffffffffu, r.oz, src);
  const float dx = __shfl_sync(0xffffffffu, r.dx, src), dy = __shfl_sync(0xffffffffu, r.dy, src), dz = __shfl_sync(0xffffffffu, r.dz, src);
  const float dd = r.dx * dx + r.dy * dy + r.dz * dz;
  const float n2 = (r.dx * r.dx + r.dy * r.dy + r.dz * r.dz) * (dx * dx + dy * dy + dz * dz);
  const bool ok = !active || (r.ox == ox && r.oy == oy && r.oz == oz && dd > 0.f && dd * dd >= 0.9985f * n2);
  return __all_sync(0xffffffffu, ok);
}

// Entry used by the trace kernels.  EVERY lane of the warp must call it (invalid lanes pass valid=false).
// mode: 0 = choose per warp, 1 = always per-lane, 2 = always packet, 3 = binary packets only (tuning knobs, results are
// identical).  `wnodes` (may be NULL): the wide BVH; coherent warps with one sign pattern traverse it, every other warp
// the binary tree.  `wstack`: kWideStack ints of shared memory private to the warp.
template <class HB, bool CULL = true>
__device__ __forceinline__ void trace_ray(const Ray& r, bool valid, const float4* __restrict__ nodes,
                                          const float4* __restrict__ wnodes, const float4* __restrict__ tris, int K, HB& hb,
                                          int& total, int* __restrict__ wstack, int mode = 0) {
  const bool coherent = warp_is_coherent(r, valid);
  if (mode == 2 || mode == 3 || (mode == 0 && coherent)) {
    const int sign = warp_sign_pattern(r, valid);   // warp-uniform: one specialisation per octant, the generic test otherwise
    if (wnodes != nullptr && coherent && mode != 3 && sign >= 0) {
      switch (sign) {
        case 0: traverse_packet_wide<HB, CULL, 0>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 1: traverse_packet_wide<HB, CULL, 1>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 2: traverse_packet_wide<HB, CULL, 2>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 3: traverse_packet_wide<HB, CULL, 3>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 4: traverse_packet_wide<HB, CULL, 4>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 5: traverse_packet_wide<HB, CULL, 5>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        case 6: traverse_packet_wide<HB, CULL, 6>(r, valid, wnodes, tris, K, hb, total, wstack); break;
        default: traverse_packet_wide<HB, CULL, 7>(r, valid, wnodes, tris, K, hb, total, wstack); break;
      }
      return;
    }
    // binary packets: the generic slab test only (the octant-specialised instantiations were worth -16 % while this was the
    // main path; now it is the fallback and eight more copies of the loop only cost instruction-cache misses)
    traverse_packet<HB, CULL, -1>(r, valid, nodes, tris, K, hb, total, wstack);
  } else if (valid) {
    traverse_single<HB, CULL>(r, nodes, tris, K, hb, total);
  } else {
    hb.init(K);
    total = 0;
  }
}

}  // namespace qf
