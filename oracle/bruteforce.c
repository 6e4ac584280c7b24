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
