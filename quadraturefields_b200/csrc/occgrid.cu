// Occupancy-grid ray marcher: nerfacc 0.5.3 `OccGridEstimator.sampling` -> `traverse_grids`
// (called at utils.py:137-148, 241-285, 422-433).  nerfacc is an absent third-party dependency: this is its published
// algorithm as restated in oracle/quadfield_oracle.py::occgrid_march (PARITY UNPINNED, see there) —
//
//   t starts where the ray enters the outermost grid box (clipped to [near, far]); dt = clamp(t * cone_angle, step, 1e10);
//   the sample [t, t + dt] is emitted when its midpoint lies before the exit point and in an occupied cell of the finest
//   level whose box contains it; t += dt either way, so samples of a ray sit on one regular grid.
//
// One thread per ray, two passes over the same deterministic march: pass 0 counts, the caller scans the counts
// (qf_hits_offsets), pass 1 writes (ray_indices, t_starts, t_ends) ray-major.  Compiled without FMA contraction: the
// sample positions and the cell index are bit-identical to the oracle's.
#include "common.cuh"

namespace qf {

struct OccGrid {
  int levels;
  int res[3];
  float lo[QF_OCC_MAX_LEVELS][3], hi[QF_OCC_MAX_LEVELS][3], scale[QF_OCC_MAX_LEVELS][3];
};

template <bool WRITE>
__global__ void __launch_bounds__(128) occgrid_march_kernel(const OccGrid g, const uint8_t* __restrict__ binaries,
                                                            const float* __restrict__ origins, const float* __restrict__ dirs,
                                                            int64_t n, const float* __restrict__ near_planes, float near_plane,
                                                            float far_plane, float step, float cone, int32_t* __restrict__ counts,
                                                            const int64_t* __restrict__ offsets, int64_t* __restrict__ ray_indices,
                                                            float* __restrict__ t_starts, float* __restrict__ t_ends,
                                                            int max_samples, const uint8_t* __restrict__ ray_mask,
                                                            float* __restrict__ termination) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ray_mask && !ray_mask[i]) {       // a ray the caller has retired: no samples, its plane stays where it was
    if (!WRITE) counts[i] = 0;
    if (termination) termination[i] = near_planes ? near_planes[i] : near_plane;
    return;
  }
  const float ox = origins[3 * i], oy = origins[3 * i + 1], oz = origins[3 * i + 2];
  const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
  const float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
  const int L = g.levels;
  const float ax0 = __fmul_rn(__fsub_rn(g.lo[L - 1][0], ox), ix), ax1 = __fmul_rn(__fsub_rn(g.hi[L - 1][0], ox), ix);
  const float ay0 = __fmul_rn(__fsub_rn(g.lo[L - 1][1], oy), iy), ay1 = __fmul_rn(__fsub_rn(g.hi[L - 1][1], oy), iy);
  const float az0 = __fmul_rn(__fsub_rn(g.lo[L - 1][2], oz), iz), az1 = __fmul_rn(__fsub_rn(g.hi[L - 1][2], oz), iz);
  const float tmin = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fminf(az0, az1));
  const float tmax = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fmaxf(az0, az1));
  float t = fmaxf(tmin, near_planes ? near_planes[i] : near_plane);
  const float t_exit = fminf(tmax, far_plane);
  int64_t out = WRITE ? offsets[i] : 0;
  int c = 0;
  if (tmin <= tmax && t < t_exit) {
    const int64_t cells = (int64_t)g.res[0] * g.res[1] * g.res[2];
    while (true) {
      const float dt = fminf(fmaxf(__fmul_rn(t, cone), step), 1e10f);
      const float tm = __fadd_rn(t, __fmul_rn(0.5f, dt));
      if (!(tm < t_exit)) break;
      const float px = __fadd_rn(ox, __fmul_rn(tm, dx)), py = __fadd_rn(oy, __fmul_rn(tm, dy)), pz = __fadd_rn(oz, __fmul_rn(tm, dz));
      bool occ = false;
      for (int l = 0; l < L; ++l) {
        if (px >= g.lo[l][0] && px < g.hi[l][0] && py >= g.lo[l][1] && py < g.hi[l][1] && pz >= g.lo[l][2] && pz < g.hi[l][2]) {
          int cx = (int)floorf(__fmul_rn(__fsub_rn(px, g.lo[l][0]), g.scale[l][0]));
          int cy = (int)floorf(__fmul_rn(__fsub_rn(py, g.lo[l][1]), g.scale[l][1]));
          int cz = (int)floorf(__fmul_rn(__fsub_rn(pz, g.lo[l][2]), g.scale[l][2]));
          cx = min(max(cx, 0), g.res[0] - 1); cy = min(max(cy, 0), g.res[1] - 1); cz = min(max(cz, 0), g.res[2] - 1);
          occ = binaries[l * cells + ((int64_t)cx * g.res[1] + cy) * g.res[2] + cz] != 0;
          break;
        }
      }
      const float te = __fadd_rn(t, dt);
      if (occ) {
        if (WRITE) { ray_indices[out] = i; t_starts[out] = t; t_ends[out] = te; ++out; }
        ++c;
      }
      t = te;
      if (max_samples > 0 && c == max_samples) break;     // the caller continues from `termination` in its next round
    }
  }
  if (!WRITE) counts[i] = c;
  if (termination) termination[i] = t;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_occgrid_march(const qf_occgrid_desc* grid, const uint8_t* d_binaries, const float* d_origins,
                                const float* d_dirs, int64_t n_rays, const float* d_near_planes, float near_plane,
                                float far_plane, float step_size, float cone_angle, int pass, int32_t* d_counts,
                                const int64_t* d_offsets, int64_t* d_ray_indices, float* d_t_starts, float* d_t_ends,
                                void* stream) {
  return qf_occgrid_march_limited(grid, d_binaries, d_origins, d_dirs, n_rays, d_near_planes, near_plane, far_plane, step_size,
                                  cone_angle, pass, d_counts, d_offsets, d_ray_indices, d_t_starts, d_t_ends, 0, nullptr, nullptr,
                                  stream);
}

extern "C" int qf_occgrid_march_limited(const qf_occgrid_desc* grid, const uint8_t* d_binaries, const float* d_origins,
                                        const float* d_dirs, int64_t n_rays, const float* d_near_planes, float near_plane,
                                        float far_plane, float step_size, float cone_angle, int pass, int32_t* d_counts,
                                        const int64_t* d_offsets, int64_t* d_ray_indices, float* d_t_starts, float* d_t_ends,
                                        int max_samples, const uint8_t* d_ray_mask, float* d_termination, void* stream) {
  QF_REQUIRE(max_samples >= 0, "qf_occgrid_march_limited: max_samples=%d", max_samples);
  QF_REQUIRE(grid, "qf_occgrid_march: NULL grid");
  QF_REQUIRE(grid->levels >= 1 && grid->levels <= QF_OCC_MAX_LEVELS, "qf_occgrid_march: levels=%d outside [1,%d]", grid->levels,
             QF_OCC_MAX_LEVELS);
  QF_REQUIRE(grid->resolution[0] >= 1 && grid->resolution[1] >= 1 && grid->resolution[2] >= 1, "qf_occgrid_march: empty grid");
  QF_REQUIRE(step_size > 0.f && cone_angle >= 0.f, "qf_occgrid_march: step_size=%g cone_angle=%g", step_size, cone_angle);
  QF_REQUIRE(pass == 0 || pass == 1, "qf_occgrid_march: pass=%d", pass);
  QF_REQUIRE(n_rays >= 0, "qf_occgrid_march: n_rays=%lld", (long long)n_rays);
  if (n_rays == 0) return QF_OK;
  QF_REQUIRE(d_binaries && d_origins && d_dirs, "qf_occgrid_march: NULL argument");
  OccGrid g;
  g.levels = grid->levels;
  for (int c = 0; c < 3; ++c) g.res[c] = grid->resolution[c];
  for (int l = 0; l < grid->levels; ++l)
    for (int c = 0; c < 3; ++c) {
      g.lo[l][c] = grid->aabbs[l][c];
      g.hi[l][c] = grid->aabbs[l][3 + c];
      QF_REQUIRE(g.hi[l][c] > g.lo[l][c], "qf_occgrid_march: empty box at level %d", l);
      g.scale[l][c] = (float)grid->resolution[c] / (g.hi[l][c] - g.lo[l][c]);   // fp32 divide, as the oracle
    }
  const unsigned blocks = (unsigned)ceil_div(n_rays, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (pass == 0) {
    QF_REQUIRE(d_counts, "qf_occgrid_march: pass 0 needs d_counts");
    occgrid_march_kernel<false><<<blocks, 128, 0, st>>>(g, d_binaries, d_origins, d_dirs, n_rays, d_near_planes, near_plane, far_plane,
                                                        step_size, cone_angle, d_counts, nullptr, nullptr, nullptr, nullptr, max_samples,
                                                        d_ray_mask, d_termination);
  } else {
    QF_REQUIRE(d_offsets, "qf_occgrid_march: pass 1 needs d_offsets");
    occgrid_march_kernel<true><<<blocks, 128, 0, st>>>(g, d_binaries, d_origins, d_dirs, n_rays, d_near_planes, near_plane, far_plane,
                                                       step_size, cone_angle, nullptr, d_offsets, d_ray_indices, d_t_starts, d_t_ends,
                                                       max_samples, d_ray_mask, d_termination);
  }
  QF_LAUNCH_CHECK();
  return QF_OK;
}
