/* TEST INFRASTRUCTURE ONLY — C restatement of oracle/quadfield_oracle.py:intersect_firstk (brute force over all
 * triangles, first K hits by (t, triangle id)).  Same fp32 operation sequence as the numpy oracle and as
 * quadraturefields_b200/csrc/geom.cuh; compile with -ffp-contract=off (no FMA contraction) and without -ffast-math.
 * Used for the large-mesh parity tests where numpy is too slow; pinned to the numpy oracle by
 * tests/test_oracle_golden.py::test_c_bruteforce_matches_numpy.
 *
 *   gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared oracle/bruteforce.c -o oracle/_ref/libqf_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1 to every rank; bench.py's reference arm asks for all host threads explicitly */
void qf_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
static inline float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

/* vertices (V,3) f32, faces (F,3) i32, pad = mesh_box_pad(vertices).  Outputs: tri (N,K) i32 (-1 padded),
 * t (N,K) f32 (+inf padded), count (N) = min(total,K), total (N). */
void qf_oracle_intersect_firstk(const float* origins, const float* dirs, int64_t N, const float* verts, const int32_t* faces,
                                int64_t F, int K, float pad, int32_t* out_tri, float* out_t, int32_t* out_count,
                                int32_t* out_total) {
  float* tv = (float*)malloc(sizeof(float) * 15 * (size_t)F); /* v0(3) e1(3) e2(3) lo(3) hi(3) */
  for (int64_t f = 0; f < F; ++f) {
    const float* a = verts + 3 * (int64_t)faces[3 * f];
    const float* b = verts + 3 * (int64_t)faces[3 * f + 1];
    const float* c = verts + 3 * (int64_t)faces[3 * f + 2];
    float* o = tv + 15 * f;
    for (int k = 0; k < 3; ++k) {
      o[k] = a[k];
      o[3 + k] = b[k] - a[k];
      o[6 + k] = c[k] - a[k];
      o[9 + k] = min3f(a[k], b[k], c[k]) - pad;
      o[12 + k] = max3f(a[k], b[k], c[k]) + pad;
    }
  }
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < N; ++i) {
    const float ox = origins[3 * i], oy = origins[3 * i + 1], oz = origins[3 * i + 2];
    const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    int32_t* tri = out_tri + i * K;
    float* tt = out_t + i * K;
    for (int s = 0; s < K; ++s) { tri[s] = -1; tt[s] = INFINITY; }
    int cnt = 0, total = 0;
    for (int64_t f = 0; f < F; ++f) {
      const float* p = tv + 15 * f;
      const float e1x = p[3], e1y = p[4], e1z = p[5], e2x = p[6], e2y = p[7], e2z = p[8];
      const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
      const float det = (e1x * px + e1y * py) + e1z * pz;
      if (det == 0.0f) continue;
      const float inv = 1.0f / det;
      const float tx = ox - p[0], ty = oy - p[1], tz = oz - p[2];
      const float u = ((tx * px + ty * py) + tz * pz) * inv;
      if (!(u >= 0.0f)) continue;
      const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
      const float v = ((dx * qx + dy * qy) + dz * qz) * inv;
      if (!(v >= 0.0f && (u + v) <= 1.0f)) continue;
      const float t = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
      if (!(t > 0.0f)) continue;
      const float ax0 = (p[9] - ox) * ix, ax1 = (p[12] - ox) * ix;
      const float ay0 = (p[10] - oy) * iy, ay1 = (p[13] - oy) * iy;
      const float az0 = (p[11] - oz) * iz, az1 = (p[14] - oz) * iz;
      const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
      const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
      if (!(tn <= tf && t >= tn)) continue;
      ++total;
      /* insert by (t, id): triangle ids arrive ascending, so equal t keeps the earlier (smaller) id first */
      int pos = cnt < K ? cnt : K;
      while (pos > 0 && tt[pos - 1] > t) --pos;
      if (pos >= K) continue;
      int last = cnt < K ? cnt : K - 1;
      for (int s = last; s > pos; --s) { tt[s] = tt[s - 1]; tri[s] = tri[s - 1]; }
      tt[pos] = t; tri[pos] = (int32_t)f;
      if (cnt < K) ++cnt;
    }
    out_count[i] = cnt;
    out_total[i] = total;
  }
  free(tv);
}

/* ------------------------------------------------------------------------------------------------------------------
 * The same first-K contract through a CPU bounding-volume hierarchy — what the reference's shipped CPU path does
 * (Embree behind trimesh, mesh_utils.py:223,350-354), and therefore the fairer CPU baseline for bench.py than the
 * O(N*F) loop above.  Results are bit-identical to the brute force: the hit predicate is the one above (including the
 * slab test against the triangle's own padded box), a node's box is the exact min/max union of its triangles' padded
 * boxes, and every slab operation is monotone in the box bounds, so a triangle the brute force accepts lies in nodes
 * whose own slab test passes with tn(node) <= tn(tri) <= t; with `want_total` = 0 nodes are also skipped when the K-buffer
 * is full and tn(node) > t_K (ties are visited; they are resolved by triangle id in the buffer).
 * Build: median split of the triangle centroids along the widest axis, <= 4 triangles per leaf.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { float lo[3], hi[3]; int32_t left, right, first, count; } qf_node;   /* count > 0: leaf over order[first..] */

typedef struct { const float* cen; int axis; } qf_sort_ctx;
static qf_sort_ctx g_ctx;   /* qsort has no context argument; the build is single-threaded */
static int qf_cmp_centroid(const void* a, const void* b) {
  const float ca = g_ctx.cen[3 * (int64_t)(*(const int32_t*)a) + g_ctx.axis], cb = g_ctx.cen[3 * (int64_t)(*(const int32_t*)b) + g_ctx.axis];
  if (ca < cb) return -1;
  if (ca > cb) return 1;
  return (*(const int32_t*)a > *(const int32_t*)b) - (*(const int32_t*)a < *(const int32_t*)b);
}

static int32_t qf_build(qf_node* nodes, int32_t* n_nodes, int32_t* order, int32_t first, int32_t count, const float* tv,
                        const float* cen) {
  const int32_t me = (*n_nodes)++;
  qf_node* nd = nodes + me;
  float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = 0; k < 3; ++k) { nd->lo[k] = INFINITY; nd->hi[k] = -INFINITY; }
  for (int32_t j = first; j < first + count; ++j) {
    const float* p = tv + 15 * (int64_t)order[j];
    for (int k = 0; k < 3; ++k) {
      nd->lo[k] = fminf(nd->lo[k], p[9 + k]);
      nd->hi[k] = fmaxf(nd->hi[k], p[12 + k]);
      clo[k] = fminf(clo[k], cen[3 * (int64_t)order[j] + k]);
      chi[k] = fmaxf(chi[k], cen[3 * (int64_t)order[j] + k]);
    }
  }
  if (count <= 4) { nd->first = first; nd->count = count; nd->left = nd->right = -1; return me; }
  int axis = 0;
  if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
  if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
  g_ctx.cen = cen; g_ctx.axis = axis;
  qsort(order + first, (size_t)count, sizeof(int32_t), qf_cmp_centroid);
  const int32_t half = count / 2;
  nd->first = 0; nd->count = 0;
  const int32_t l = qf_build(nodes, n_nodes, order, first, half, tv, cen);
  const int32_t r = qf_build(nodes, n_nodes, order, first + half, count - half, tv, cen);
  nodes[me].left = l; nodes[me].right = r;   /* `nd` may be stale only if nodes were reallocated; they are not */
  return me;
}

static inline int qf_slab(const float* lo, const float* hi, float ox, float oy, float oz, float ix, float iy, float iz, float* tn_out) {
  const float ax0 = (lo[0] - ox) * ix, ax1 = (hi[0] - ox) * ix;
  const float ay0 = (lo[1] - oy) * iy, ay1 = (hi[1] - oy) * iy;
  const float az0 = (lo[2] - oz) * iz, az1 = (hi[2] - oz) * iz;
  const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
  const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
  *tn_out = tn;
  return tn <= tf;
}

void qf_oracle_intersect_firstk_bvh(const float* origins, const float* dirs, int64_t N, const float* verts, const int32_t* faces,
                                    int64_t F, int K, float pad, int want_total, int32_t* out_tri, float* out_t,
                                    int32_t* out_count, int32_t* out_total) {
  float* tv = (float*)malloc(sizeof(float) * 15 * (size_t)F);
  float* cen = (float*)malloc(sizeof(float) * 3 * (size_t)F);
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)F);
  qf_node* nodes = (qf_node*)malloc(sizeof(qf_node) * (size_t)(2 * F + 1));
  for (int64_t f = 0; f < F; ++f) {
    const float* a = verts + 3 * (int64_t)faces[3 * f];
    const float* b = verts + 3 * (int64_t)faces[3 * f + 1];
    const float* c = verts + 3 * (int64_t)faces[3 * f + 2];
    float* o = tv + 15 * f;
    for (int k = 0; k < 3; ++k) {
      o[k] = a[k];
      o[3 + k] = b[k] - a[k];
      o[6 + k] = c[k] - a[k];
      o[9 + k] = min3f(a[k], b[k], c[k]) - pad;
      o[12 + k] = max3f(a[k], b[k], c[k]) + pad;
      cen[3 * f + k] = 0.5f * (o[9 + k] + o[12 + k]);
    }
    order[f] = (int32_t)f;
  }
  int32_t n_nodes = 0;
  if (F > 0) qf_build(nodes, &n_nodes, order, 0, (int32_t)F, tv, cen);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < N; ++i) {
    const float ox = origins[3 * i], oy = origins[3 * i + 1], oz = origins[3 * i + 2];
    const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    int32_t* tri = out_tri + i * K;
    float* tt = out_t + i * K;
    for (int s = 0; s < K; ++s) { tri[s] = -1; tt[s] = INFINITY; }
    int cnt = 0, total = 0, sp = 0;
    int32_t stack[128];
    if (n_nodes > 0) stack[sp++] = 0;
    while (sp > 0) {
      const qf_node* nd = nodes + stack[--sp];
      float ntn;
      if (!qf_slab(nd->lo, nd->hi, ox, oy, oz, ix, iy, iz, &ntn)) continue;
      if (!want_total && cnt == K && ntn > tt[K - 1]) continue;
      if (nd->count == 0) { stack[sp++] = nd->right; stack[sp++] = nd->left; continue; }
      for (int32_t j = nd->first; j < nd->first + nd->count; ++j) {
        const int32_t f = order[j];
        const float* p = tv + 15 * (int64_t)f;
        const float e1x = p[3], e1y = p[4], e1z = p[5], e2x = p[6], e2y = p[7], e2z = p[8];
        const float px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
        const float det = (e1x * px + e1y * py) + e1z * pz;
        if (det == 0.0f) continue;
        const float inv = 1.0f / det;
        const float tx = ox - p[0], ty = oy - p[1], tz = oz - p[2];
        const float u = ((tx * px + ty * py) + tz * pz) * inv;
        if (!(u >= 0.0f)) continue;
        const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
        const float v = ((dx * qx + dy * qy) + dz * qz) * inv;
        if (!(v >= 0.0f && (u + v) <= 1.0f)) continue;
        const float t = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
        if (!(t > 0.0f)) continue;
        float tn;
        if (!(qf_slab(p + 9, p + 12, ox, oy, oz, ix, iy, iz, &tn) && t >= tn)) continue;
        ++total;
        /* insert by (t, id): triangles arrive in tree order, so the id is compared explicitly */
        int pos = cnt < K ? cnt : K;
        while (pos > 0 && (tt[pos - 1] > t || (tt[pos - 1] == t && tri[pos - 1] > f))) --pos;
        if (pos >= K) continue;
        const int last = cnt < K ? cnt : K - 1;
        for (int s = last; s > pos; --s) { tt[s] = tt[s - 1]; tri[s] = tri[s - 1]; }
        tt[pos] = t; tri[pos] = f;
        if (cnt < K) ++cnt;
      }
    }
    out_count[i] = cnt;
    if (out_total) out_total[i] = want_total ? total : -1;
  }
  free(nodes); free(order); free(cen); free(tv);
}

/* ---------------------------------------------------------------------------------------------------------------------
 * C restatement of oracle/quadfield_oracle.py:hashgrid_encode (tcnn kernel_grid forward, recalled): 3-D positions in
 * [0,1], F = 2 features per entry, trilinear interpolation accumulated in HALF precision with one fused multiply-add per
 * corner (weight formed in fp32, cast to half).  The fused half FMA is evaluated exactly: the product of two halves and
 * the sum with a third are exact in double for every non-pathological exponent gap, and the double -> half conversion
 * rounds once.  table: (n_entries, 2) fp32 holding half values; out: (M, 2 L) fp32 holding half values, level-major.
 * Pinned to the numpy/torch version by tests/test_oracle_golden.py::test_c_hashgrid_matches_numpy.
 * ------------------------------------------------------------------------------------------------------------------- */
void qf_oracle_hashgrid_encode(const float* x01, int64_t M, const float* table, int L, const float* scale,
                               const uint32_t* resolution, const int64_t* offset, const int64_t* size, const uint8_t* hashed,
                               float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < M; ++i) {
    const float x[3] = {x01[3 * i], x01[3 * i + 1], x01[3 * i + 2]};
    for (int l = 0; l < L; ++l) {
      uint32_t c[3];
      float fr[3];
      for (int d = 0; d < 3; ++d) {
        const float pos = fmaf(scale[l], x[d], 0.5f);
        const float fl = floorf(pos);
        c[d] = (uint32_t)(int32_t)fl;
        fr[d] = pos - fl;
      }
      const uint32_t res = resolution[l];
      _Float16 acc0 = (_Float16)0.0f, acc1 = (_Float16)0.0f;
      for (int k = 0; k < 8; ++k) {
        float w = 1.0f;
        uint32_t g[3];
        for (int d = 0; d < 3; ++d) {
          if ((k >> d) & 1) { w = w * fr[d]; g[d] = c[d] + 1u; }
          else { w = w * (1.0f - fr[d]); g[d] = c[d]; }
        }
        uint32_t idx;
        if (hashed[l]) idx = (g[0] * 1u) ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
        else idx = g[0] + g[1] * res + g[2] * (res * res);
        const int64_t e = (int64_t)(idx % (uint64_t)size[l]) + offset[l];
        const _Float16 wh = (_Float16)w;
        acc0 = (_Float16)((double)wh * (double)table[2 * e] + (double)acc0);
        acc1 = (_Float16)((double)wh * (double)table[2 * e + 1] + (double)acc1);
      }
      out[(size_t)i * 2 * L + 2 * l] = (float)acc0;
      out[(size_t)i * 2 * L + 2 * l + 1] = (float)acc1;
    }
  }
}
