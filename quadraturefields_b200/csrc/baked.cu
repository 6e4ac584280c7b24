// (4) Baked spherical-Gaussian texture path.
//
// The reference keeps 2+2L separate uint8 planes (alpha (S,S), diffuse (S,S,3), per lobe colour and
// [lambda,azimuth,elevation] (S,S,3)) and gathers each one per hit in 32 000-row Python chunks
// (texture_utils.py:149-175, utils.py:1064-1068).  On B200 that is 2+2L scattered 32-byte sectors per
// hit from a 1.5-2.7 GB set, i.e. HBM-bound with 4-10x sector waste.  Here the planes are repacked once
// into ONE interleaved record per texel, 16-byte aligned:
//     [alpha, d0,d1,d2, {lambda, az, el, c0,c1,c2} x L]  = 4+6L bytes -> 32 B (L=3, one sector) / 48 B (L=6)
// so a hit costs one or two sectors, fetched as LDG.128s, and the decode + SG evaluation + sigmoid
// happen in registers.  Compiled with -fmad=false: the texel index must match the oracle bit for bit.
#include "common.cuh"

namespace qf {

__host__ __device__ inline int record_bytes_for(int L) { return (4 + 6 * L + 15) / 16 * 16; }

struct PlanePtrs {
  const uint8_t* alpha;
  const uint8_t* diffuse;
  const uint8_t* colors[QF_MAX_LOBES];
  const uint8_t* lambdas[QF_MAX_LOBES];
};

__global__ void texture_pack_kernel(PlanePtrs p, int L, int64_t n_texels, int rec, uint8_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_texels; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t* o = out + i * rec;
    o[0] = p.alpha[i];
    o[1] = p.diffuse[3 * i]; o[2] = p.diffuse[3 * i + 1]; o[3] = p.diffuse[3 * i + 2];
    for (int l = 0; l < L; ++l) {
      o[4 + 6 * l + 0] = p.lambdas[l][3 * i]; o[4 + 6 * l + 1] = p.lambdas[l][3 * i + 1]; o[4 + 6 * l + 2] = p.lambdas[l][3 * i + 2];
      o[4 + 6 * l + 3] = p.colors[l][3 * i]; o[4 + 6 * l + 4] = p.colors[l][3 * i + 1]; o[4 + 6 * l + 5] = p.colors[l][3 * i + 2];
    }
    for (int b = 4 + 6 * L; b < rec; ++b) o[b] = 0;
  }
}

struct TexParams {
  const uint8_t* records;
  int size, L, rec, colour_logit;
  float lambda_thres;
  const float* tables;   // kTexTables x 256 floats: every dequantiser evaluated once per uint8 code (decode_tables_kernel)
};

// ngp.py:275-281 (quirk Q5: logit only for compress_type == "sigma")
__device__ __forceinline__ float inv_colour(uint8_t q, int logit) {
  float c = (float)q / 255.0f;
  if (logit) return logf(fminf(fmaxf(c / (1.0f - c), 1e-8f), 1e37f));
  return c * 2.0f * 12.0f - 12.0f;
}

// The dequantisers are pure functions of one uint8 code: sigma(alpha), colour(c), lambda(q), and the four trigonometric
// factors of the lobe axis.  Evaluating them inline cost ~450 of the ~900 instructions per hit (precise expf / logf / sinf /
// cosf, 12 of them at L=3; ncu r2f: baked_shade_kernel issue-bound at 68 %, XU pipe 42 %, DRAM 37 %).  They are tabulated
// once per texture set with these very expressions, so a lookup returns the same bits as the inline evaluation.
enum { kTabSigma = 0, kTabColour, kTabLambda, kTabCosAz, kTabSinAz, kTabSinEl, kTabCosEl, kTexTables };

__global__ void decode_tables_kernel(int colour_logit, float lambda_thres, float* __restrict__ tab) {
  const int q = threadIdx.x;   // 256 threads
  const uint8_t b = (uint8_t)q;
  const float a = (float)b / 255.0f;
  tab[kTabSigma * 256 + q] = -logf(fmaxf(1.0f - a, 1e-6f)) / 0.005f;                    // texture_utils.py:61-65
  tab[kTabColour * 256 + q] = inv_colour(b, colour_logit);
  tab[kTabLambda * 256 + q] = expf((float)b * lambda_thres / 255.0f - 2.5f);            // ngp.py:260-262
  const float az = (float)(uint8_t)(b - 128) / 128.0f * 3.14159265358979323846f;         // ngp.py:245-246, uint8 wrap (Q6)
  const float el = (float)b / 256.0f * 3.14159265358979323846f;
  tab[kTabCosAz * 256 + q] = cosf(az);
  tab[kTabSinAz * 256 + q] = sinf(az);
  tab[kTabSinEl * 256 + q] = sinf(el);
  tab[kTabCosEl * 256 + q] = cosf(el);
}

// texel record -> reference feature row [diffuse(3), L x (axis3, lambda, c3), sigma]; `tab`: the tables (global or shared)
__device__ __forceinline__ void decode_record(const TexParams& t, const float* __restrict__ tab, int64_t texel, float* __restrict__ f) {
  const uint4* rp = reinterpret_cast<const uint4*>(t.records + texel * t.rec);
  uint4 raw[4];
  const int nq = t.rec / 16;
#pragma unroll
  for (int q = 0; q < 4; ++q) if (q < nq) raw[q] = __ldg(rp + q);
  const uint8_t* b = reinterpret_cast<const uint8_t*>(raw);
  const float* col = tab + kTabColour * 256;
  f[0] = col[b[1]]; f[1] = col[b[2]]; f[2] = col[b[3]];
  for (int l = 0; l < t.L; ++l) {
    const uint8_t* r = b + 4 + 6 * l;
    const float se = tab[kTabSinEl * 256 + r[2]];
    float* o = f + 3 + 7 * l;
    o[0] = tab[kTabCosAz * 256 + r[1]] * se; o[1] = tab[kTabSinAz * 256 + r[1]] * se; o[2] = tab[kTabCosEl * 256 + r[2]];
    o[3] = tab[kTabLambda * 256 + r[0]];
    o[4] = col[r[3]]; o[5] = col[r[4]]; o[6] = col[r[5]];
  }
  f[3 + 7 * t.L] = tab[kTabSigma * 256 + b[0]];
}

// Compile-time lobe count: the record bytes are picked out of registers with constant shifts and the feature row stays
// in registers (with a run-time L both live in local memory: 304 bytes of stack per thread in the shading kernel).
template <int L>
__device__ __forceinline__ void decode_record_static(const TexParams& t, const float* __restrict__ tab, int64_t texel,
                                                     float* __restrict__ f) {
  constexpr int NQ = (4 + 6 * L + 15) / 16;
  const uint4* rp = reinterpret_cast<const uint4*>(t.records + texel * (NQ * 16));
  uint32_t w[NQ * 4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) { const uint4 v = __ldg(rp + q); w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
  auto byte = [&](int k) { return (w[k >> 2] >> (8 * (k & 3))) & 255u; };
  const float* col = tab + kTabColour * 256;
  f[0] = col[byte(1)]; f[1] = col[byte(2)]; f[2] = col[byte(3)];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int r = 4 + 6 * l;
    const float se = tab[kTabSinEl * 256 + byte(r + 2)];
    float* o = f + 3 + 7 * l;
    o[0] = tab[kTabCosAz * 256 + byte(r + 1)] * se; o[1] = tab[kTabSinAz * 256 + byte(r + 1)] * se; o[2] = tab[kTabCosEl * 256 + byte(r + 2)];
    o[3] = tab[kTabLambda * 256 + byte(r)];
    o[4] = col[byte(r + 3)]; o[5] = col[byte(r + 4)]; o[6] = col[byte(r + 5)];
  }
  f[3 + 7 * L] = tab[kTabSigma * 256 + byte(0)];
}

// ngp.py:371-393,456-461: rgb = sigmoid(diffuse + sum_l c_l exp(|lambda_l| (a_l/|a_l| . d - 1)))
__device__ __forceinline__ void sg_rgb(const float* __restrict__ f, int L, float dx, float dy, float dz, float* rgb) {
  float r = f[0], g = f[1], b = f[2];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float* o = f + 3 + 7 * l;
    float n = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    float ax = o[0] / n, ay = o[1] / n, az = o[2] / n;
    float e = expf(fabsf(o[3]) * ((ax * dx + ay * dy + az * dz) - 1.0f));
    r += o[4] * e; g += o[5] * e; b += o[6] * e;
  }
  rgb[0] = 1.0f / (1.0f + expf(-r)); rgb[1] = 1.0f / (1.0f + expf(-g)); rgb[2] = 1.0f / (1.0f + expf(-b));
}

// utils.py:1055-1063: fp64 Cramer barycentrics (trimesh.triangles.points_to_barycentric), clamp, renormalise,
// uv = sum b_k uv_k, floor, clip.  Returns (row, col) = (uv.x, uv.y) floored — the reference indexes
// texture[uv[:,0], uv[:,1]].
__device__ __forceinline__ void hit_texel(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                          const float* __restrict__ uv, int64_t tri, float px, float py, float pz, int S,
                                          int64_t& t0, int64_t& t1) {
  int ia = faces[3 * tri], ib = faces[3 * tri + 1], ic = faces[3 * tri + 2];
  double v0[3], e0[3], e1[3], w[3];
  const double p[3] = {(double)px, (double)py, (double)pz};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v0[k] = (double)verts[3 * (int64_t)ia + k];
    e0[k] = (double)verts[3 * (int64_t)ib + k] - v0[k];
    e1[k] = (double)verts[3 * (int64_t)ic + k] - v0[k];
    w[k] = p[k] - v0[k];
  }
  auto dot = [](const double* a, const double* b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
  };
  double d00 = dot(e0, e0), d01 = dot(e0, e1), d02 = dot(e0, w), d11 = dot(e1, e1), d12 = dot(e1, w);
  double inv = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(d00, d11), __dmul_rn(d01, d01)));
  double b2 = __dmul_rn(__dsub_rn(__dmul_rn(d00, d12), __dmul_rn(d01, d02)), inv);
  double b1 = __dmul_rn(__dsub_rn(__dmul_rn(d11, d02), __dmul_rn(d01, d12)), inv);
  double b0 = __dsub_rn(__dsub_rn(1.0, b1), b2);
  float c0 = fminf(fmaxf((float)b0, 0.f), 1.f), c1 = fminf(fmaxf((float)b1, 0.f), 1.f), c2 = fminf(fmaxf((float)b2, 0.f), 1.f);
  float s = __fadd_rn(__fadd_rn(c0, c1), c2);
  c0 = __fdiv_rn(c0, s); c1 = __fdiv_rn(c1, s); c2 = __fdiv_rn(c2, s);
  float u = __fadd_rn(__fadd_rn(__fmul_rn(uv[2 * (int64_t)ia], c0), __fmul_rn(uv[2 * (int64_t)ib], c1)), __fmul_rn(uv[2 * (int64_t)ic], c2));
  float v = __fadd_rn(__fadd_rn(__fmul_rn(uv[2 * (int64_t)ia + 1], c0), __fmul_rn(uv[2 * (int64_t)ib + 1], c1)), __fmul_rn(uv[2 * (int64_t)ic + 1], c2));
  int64_t i0 = (int64_t)floorf(u), i1 = (int64_t)floorf(v);
  t0 = i0 < 0 ? 0 : (i0 > S - 1 ? S - 1 : i0);
  t1 = i1 < 0 ? 0 : (i1 > S - 1 ? S - 1 : i1);
}

// The triangle-only half of hit_texel, done once per triangle: 16 doubles = 128 bytes per record
//   [v0(3) | e0(3) | e1(3) | d00 d01 d11 | 1/(d00 d11 - d01^2) | uv_a, uv_b, uv_c as 6 floats (3 doubles' worth)]
// One aligned 128-byte gather per hit replaces seven scattered ones (face, 3 vertices, 3 uv pairs) and ~110 of the ~200
// fp64 instructions; the per-hit half below repeats hit_texel's remaining operations in the same order: same bits.
__global__ void bary_setup_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, const float* __restrict__ uv,
                                  int64_t F, double* __restrict__ out) {
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
  double v0[3], e0[3], e1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v0[k] = (double)verts[3 * (int64_t)ia + k];
    e0[k] = (double)verts[3 * (int64_t)ib + k] - v0[k];
    e1[k] = (double)verts[3 * (int64_t)ic + k] - v0[k];
  }
  auto dot = [](const double* a, const double* b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
  };
  const double d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1);
  const double inv = __ddiv_rn(1.0, __dsub_rn(__dmul_rn(d00, d11), __dmul_rn(d01, d01)));
  double* o = out + 16 * f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { o[k] = v0[k]; o[3 + k] = e0[k]; o[6 + k] = e1[k]; }
  o[9] = d00; o[10] = d01; o[11] = d11; o[12] = inv;
  float* u = reinterpret_cast<float*>(o + 13);
  u[0] = uv[2 * (int64_t)ia]; u[1] = uv[2 * (int64_t)ia + 1]; u[2] = uv[2 * (int64_t)ib]; u[3] = uv[2 * (int64_t)ib + 1];
  u[4] = uv[2 * (int64_t)ic]; u[5] = uv[2 * (int64_t)ic + 1];
}

__device__ __forceinline__ void hit_texel_cached(const double* __restrict__ bary, int64_t tri, float px, float py, float pz, int S,
                                                 int64_t& t0, int64_t& t1) {
  const double2* rp = reinterpret_cast<const double2*>(bary + 16 * tri);
  double q[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) { const double2 v = __ldg(rp + k); q[2 * k] = v.x; q[2 * k + 1] = v.y; }
  const double p[3] = {(double)px, (double)py, (double)pz};
  double w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = p[k] - q[k];
  auto dot = [](const double* a, const double* b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
  };
  const double d00 = q[9], d01 = q[10], d11 = q[11], inv = q[12];
  const double d02 = dot(q + 3, w), d12 = dot(q + 6, w);
  double b2 = __dmul_rn(__dsub_rn(__dmul_rn(d00, d12), __dmul_rn(d01, d02)), inv);
  double b1 = __dmul_rn(__dsub_rn(__dmul_rn(d11, d02), __dmul_rn(d01, d12)), inv);
  double b0 = __dsub_rn(__dsub_rn(1.0, b1), b2);
  float c0 = fminf(fmaxf((float)b0, 0.f), 1.f), c1 = fminf(fmaxf((float)b1, 0.f), 1.f), c2 = fminf(fmaxf((float)b2, 0.f), 1.f);
  float s = __fadd_rn(__fadd_rn(c0, c1), c2);
  c0 = __fdiv_rn(c0, s); c1 = __fdiv_rn(c1, s); c2 = __fdiv_rn(c2, s);
  const float* uvp = reinterpret_cast<const float*>(q + 13);
  float u = __fadd_rn(__fadd_rn(__fmul_rn(uvp[0], c0), __fmul_rn(uvp[2], c1)), __fmul_rn(uvp[4], c2));
  float v = __fadd_rn(__fadd_rn(__fmul_rn(uvp[1], c0), __fmul_rn(uvp[3], c1)), __fmul_rn(uvp[5], c2));
  int64_t i0 = (int64_t)floorf(u), i1 = (int64_t)floorf(v);
  t0 = i0 < 0 ? 0 : (i0 > S - 1 ? S - 1 : i0);
  t1 = i1 < 0 ? 0 : (i1 > S - 1 ? S - 1 : i1);
}

__global__ void texture_decode_kernel(TexParams t, const int64_t* __restrict__ idx, int64_t M, float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float f[3 + 7 * QF_MAX_LOBES + 1];
  decode_record(t, t.tables, idx[2 * i] * t.size + idx[2 * i + 1], f);
  const int W = 3 + 7 * t.L + 1;
  for (int k = 0; k < W; ++k) out[i * W + k] = f[k];
}

__global__ void sg_rgb_kernel(const float* __restrict__ feats, int64_t stride, int L, const float* __restrict__ dirs,
                              int64_t M, float* __restrict__ rgb) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float f[3 + 7 * QF_MAX_LOBES];
  for (int k = 0; k < 3 + 7 * L; ++k) f[k] = feats[i * stride + k];
  float o[3];
  sg_rgb(f, L, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], o);
  rgb[3 * i] = o[0]; rgb[3 * i + 1] = o[1]; rgb[3 * i + 2] = o[2];
}

// backward of sg_rgb with respect to the feature row (diffuse, per lobe axis / lambda / colour); thread per sample
__global__ void sg_rgb_backward_kernel(const float* __restrict__ feats, int64_t stride, int L, const float* __restrict__ dirs,
                                       int64_t M, const float* __restrict__ g_rgb, float* __restrict__ g_feats, int64_t g_stride) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float* f = feats + i * stride;
  const float dx = dirs[3 * i], dy = dirs[3 * i + 1], dz = dirs[3 * i + 2];
  float u[3] = {f[0], f[1], f[2]};
  for (int l = 0; l < L; ++l) {
    const float* o = f + 3 + 7 * l;
    const float n = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    const float e = expf(fabsf(o[3]) * ((o[0] * dx + o[1] * dy + o[2] * dz) / n - 1.0f));
    u[0] += o[4] * e; u[1] += o[5] * e; u[2] += o[6] * e;
  }
  float gu[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float sgm = 1.0f / (1.0f + expf(-u[c]));
    gu[c] = g_rgb[3 * i + c] * sgm * (1.0f - sgm);
  }
  float* go = g_feats + i * g_stride;
  go[0] = gu[0]; go[1] = gu[1]; go[2] = gu[2];
  for (int l = 0; l < L; ++l) {
    const float* o = f + 3 + 7 * l;
    float* g = go + 3 + 7 * l;
    const float n = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    const float ax = o[0] / n, ay = o[1] / n, az = o[2] / n;
    const float cosv = ax * dx + ay * dy + az * dz;
    const float lam = fabsf(o[3]);
    const float e = expf(lam * (cosv - 1.0f));
    g[4] = gu[0] * e; g[5] = gu[1] * e; g[6] = gu[2] * e;                    // d colour
    const float se = (gu[0] * o[4] + gu[1] * o[5] + gu[2] * o[6]) * e;         // dL/d(exponent)
    g[3] = se * (cosv - 1.0f) * (o[3] > 0.f ? 1.0f : (o[3] < 0.f ? -1.0f : 0.0f));   // d lambda through |.|
    const float gc = se * lam;                                                  // dL/d(cos)
    // d(a/|a| . d)/da = (d - a_hat (a_hat . d)) / |a|
    g[0] = gc * (dx - ax * cosv) / n; g[1] = gc * (dy - ay * cosv) / n; g[2] = gc * (dz - az * cosv) / n;
  }
}

__global__ void hit_texels_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                  const float* __restrict__ uv, const float* __restrict__ points,
                                  const int64_t* __restrict__ tri, int64_t M, int S, int64_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  int64_t t0, t1;
  hit_texel(verts, faces, uv, tri[i], points[3 * i], points[3 * i + 1], points[3 * i + 2], S, t0, t1);
  out[2 * i] = t0; out[2 * i + 1] = t1;
}

// SURVEY §8 row f-4 — the bake writer FeatureCompression.compress (texture_utils.py:67-98): quantise feature rows
// [diffuse(3), L x (axis3, lambda, c3), sigma] into the uint8 planes, at row i or at texel (idx[i,0], idx[i,1]).
struct PlaneOut {
  uint8_t* alpha;
  uint8_t* diffuse;
  uint8_t* colors[QF_MAX_LOBES];
  uint8_t* lambdas[QF_MAX_LOBES];
};
__device__ __forceinline__ uint8_t to_u8(float v) { return (uint8_t)(int)v; }   // torch .to(uint8): truncate, low 8 bits
__device__ __forceinline__ uint8_t quant_colour(float c, int logit) {
  if (logit) c = 1.0f / (1.0f + expf(-c));                                   // ngp.py:264-273
  else c = (fminf(fmaxf(c, -12.0f), 12.0f) + 12.0f) / 2.0f / 12.0f;
  return to_u8(c * 255.0f);
}
__global__ void texture_compress_kernel(const float* __restrict__ feats, int64_t M, int L, int colour_logit, float lambda_thres,
                                        const int64_t* __restrict__ idx, int S, PlaneOut p) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int W = 3 + 7 * L + 1;
  const float* f = feats + i * W;
  const int64_t o = idx ? idx[2 * i] * S + idx[2 * i + 1] : i;
  const float a = (1.0f - expf(-f[W - 1] * 0.005f)) * 255.0f;                 // texture_utils.py:51-55
  p.alpha[o] = to_u8(fminf(fmaxf(a, 0.0f), 255.0f));
  for (int c = 0; c < 3; ++c) p.diffuse[3 * o + c] = quant_colour(f[c], colour_logit);
  for (int l = 0; l < L; ++l) {
    const float* q = f + 3 + 7 * l;
    float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]) + 1e-6f;           // ngp.py:239-243
    float vx = q[0] / n, vy = q[1] / n, vz = q[2] / n;
    float az = atan2f(vy, vx) * 128.0f / 3.14159265358979323846f + 128.0f;
    float el = acosf(vz) * 256.0f / 3.14159265358979323846f;
    float lam = (logf(fmaxf(fabsf(q[3]), 1e-5f)) + 2.5f) / lambda_thres;       // ngp.py:254-258
    lam = 255.0f * fminf(fmaxf(lam, 0.0f), 1.0f);
    p.lambdas[l][3 * o] = to_u8(lam); p.lambdas[l][3 * o + 1] = to_u8(az); p.lambdas[l][3 * o + 2] = to_u8(el);
    for (int c = 0; c < 3; ++c) p.colors[l][3 * o + c] = quant_colour(q[4 + c], colour_logit);
  }
}

// fused shading of compact hit records for qf_render_mesh_baked (render.cu): texel lookup + decode + SG
template <int LS>   // LS > 0: lobe count known at compile time; 0: generic
__global__ void __launch_bounds__(256) baked_shade_kernel(TexParams t, const double* __restrict__ bary,
                                                          const float4* __restrict__ hit_pd, const int2* __restrict__ hit_rt,
                                                          const float* __restrict__ viewdirs, const int32_t* __restrict__ d_M,
                                                          float4* __restrict__ out4) {
  __shared__ float s_tab[kTexTables * 256];
  for (int k = threadIdx.x; k < kTexTables * 256; k += blockDim.x) s_tab[k] = __ldg(t.tables + k);
  __syncthreads();
  const int64_t M = *d_M;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pd = hit_pd[i];
    int2 rt = hit_rt[i];
    int64_t t0, t1;
    hit_texel_cached(bary, rt.y, pd.x, pd.y, pd.z, t.size, t0, t1);
    float f[3 + 7 * (LS ? LS : QF_MAX_LOBES) + 1];
    if (LS) decode_record_static<LS ? LS : 1>(t, s_tab, t0 * t.size + t1, f);
    else decode_record(t, s_tab, t0 * t.size + t1, f);
    const int L = LS ? LS : t.L;
    // tuple dirs: d / (|d| + 1e-7)   (mesh_utils.py:369-370, quirk Q7)
    float dx = viewdirs[3 * (int64_t)rt.x], dy = viewdirs[3 * (int64_t)rt.x + 1], dz = viewdirs[3 * (int64_t)rt.x + 2];
    float n = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-7f);
    float o[3];
    sg_rgb(f, L, __fdiv_rn(dx, n), __fdiv_rn(dy, n), __fdiv_rn(dz, n), o);
    out4[i] = make_float4(o[0], o[1], o[2], f[3 + 7 * L]);
  }
}

int launch_baked_shade(const qf_texture* tex, const qf_mesh* mesh, const float* d_uv, const float4* hit_pd,
                       const int2* hit_rt, const float* d_viewdirs, const int32_t* d_M, float4* out4, cudaStream_t st) {
  TexParams t{tex->d_records, tex->size, tex->num_lobes, tex->record_bytes, tex->colour_logit, tex->lambda_thres, tex->d_tables};
  // per-triangle barycentric records for this uv array (rebuilt when the array or the vertices change)
  if (mesh->d_bary == nullptr) {
    if (cudaMalloc((void**)&mesh->d_bary, sizeof(double) * 16 * (size_t)mesh->n_faces) != cudaSuccess) {
      set_error("qf_render_mesh_baked: cannot allocate %lld bytes of per-triangle records", (long long)(128 * mesh->n_faces));
      return QF_ERR_CUDA;
    }
    mesh->bary_uv = nullptr;
  }
  if (mesh->bary_uv != d_uv || mesh->bary_version != mesh->geometry_version) {
    bary_setup_kernel<<<(int)ceil_div(mesh->n_faces, 256), 256, 0, st>>>(mesh->d_vertices, mesh->d_faces, d_uv, mesh->n_faces, mesh->d_bary);
    mesh->bary_uv = d_uv;
    mesh->bary_version = mesh->geometry_version;
  }
  switch (tex->num_lobes) {
    case 3: baked_shade_kernel<3><<<kNumSMs * 8, 256, 0, st>>>(t, mesh->d_bary, hit_pd, hit_rt, d_viewdirs, d_M, out4); break;
    case 6: baked_shade_kernel<6><<<kNumSMs * 8, 256, 0, st>>>(t, mesh->d_bary, hit_pd, hit_rt, d_viewdirs, d_M, out4); break;
    default: baked_shade_kernel<0><<<kNumSMs * 8, 256, 0, st>>>(t, mesh->d_bary, hit_pd, hit_rt, d_viewdirs, d_M, out4); break;
  }
  QF_LAUNCH_CHECK();
  return QF_OK;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_texture_create(int size, int num_lobes, const uint8_t* d_alpha, const uint8_t* d_diffuse,
                                 const uint8_t* const* h_d_colors, const uint8_t* const* h_d_lambdas, int colour_logit,
                                 float lambda_thres, void* stream, qf_texture** out) {
  QF_REQUIRE(out && d_alpha && d_diffuse && h_d_colors && h_d_lambdas, "qf_texture_create: NULL argument");
  QF_REQUIRE(size > 0 && num_lobes >= 0 && num_lobes <= QF_MAX_LOBES, "qf_texture_create: size=%d lobes=%d", size, num_lobes);
  qf_texture* t = new qf_texture();
  t->size = size; t->num_lobes = num_lobes; t->colour_logit = colour_logit; t->lambda_thres = lambda_thres;
  t->record_bytes = record_bytes_for(num_lobes);
  QF_REQUIRE(t->record_bytes <= 64, "qf_texture_create: record too large");
  int64_t n = (int64_t)size * size;
  if (cudaMalloc((void**)&t->d_records, (size_t)n * t->record_bytes) != cudaSuccess) {
    set_error("qf_texture_create: cannot allocate %lld bytes", (long long)n * t->record_bytes);
    delete t;
    return QF_ERR_CUDA;
  }
  PlanePtrs p{};
  p.alpha = d_alpha; p.diffuse = d_diffuse;
  for (int l = 0; l < num_lobes; ++l) { p.colors[l] = h_d_colors[l]; p.lambdas[l] = h_d_lambdas[l]; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMalloc((void**)&t->d_tables, sizeof(float) * kTexTables * 256) != cudaSuccess) {
    set_error("qf_texture_create: cannot allocate the dequantiser tables");
    qf_texture_destroy(t);
    return QF_ERR_CUDA;
  }
  decode_tables_kernel<<<1, 256, 0, st>>>(colour_logit, lambda_thres, t->d_tables);
  texture_pack_kernel<<<kNumSMs * 8, 256, 0, st>>>(p, num_lobes, n, t->record_bytes, t->d_records);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_error("qf_texture_create: pack failed: %s", cudaGetErrorString(e)); qf_texture_destroy(t); return QF_ERR_CUDA; }
  *out = t;
  return QF_OK;
}

extern "C" void qf_texture_destroy(qf_texture* t) {
  if (!t) return;
  if (t->d_records) cudaFree(t->d_records);
  if (t->d_tables) cudaFree(t->d_tables);
  delete t;
}

extern "C" int qf_texture_decode(const qf_texture* t, const int64_t* d_indices, int64_t M, float* d_features, void* stream) {
  QF_REQUIRE(t && d_indices && d_features, "qf_texture_decode: NULL argument");
  if (M == 0) return QF_OK;
  TexParams p{t->d_records, t->size, t->num_lobes, t->record_bytes, t->colour_logit, t->lambda_thres, t->d_tables};
  texture_decode_kernel<<<(int)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(p, d_indices, M, d_features);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_texture_compress(const float* d_features, int64_t M, int num_lobes, int colour_logit, float lambda_thres,
                                   const int64_t* d_indices, int texture_size, uint8_t* d_alpha, uint8_t* d_diffuse,
                                   uint8_t* const* h_d_colors, uint8_t* const* h_d_lambdas, void* stream) {
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_features && d_alpha && d_diffuse && h_d_colors && h_d_lambdas, "qf_texture_compress: NULL argument");
  QF_REQUIRE(num_lobes >= 0 && num_lobes <= QF_MAX_LOBES, "qf_texture_compress: lobes=%d", num_lobes);
  PlaneOut p{};
  p.alpha = d_alpha; p.diffuse = d_diffuse;
  for (int l = 0; l < num_lobes; ++l) { p.colors[l] = h_d_colors[l]; p.lambdas[l] = h_d_lambdas[l]; }
  texture_compress_kernel<<<(int)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d_features, M, num_lobes, colour_logit, lambda_thres,
                                                                                   d_indices, texture_size, p);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_sg_features_to_rgb(const float* d_features, int64_t stride, int num_lobes, const float* d_dirs, int64_t M,
                                     float* d_rgb, void* stream) {
  QF_REQUIRE(d_features && d_dirs && d_rgb, "qf_sg_features_to_rgb: NULL argument");
  QF_REQUIRE(num_lobes >= 0 && num_lobes <= QF_MAX_LOBES && stride >= 3 + 7 * num_lobes, "qf_sg_features_to_rgb: lobes=%d stride=%lld",
             num_lobes, (long long)stride);
  if (M == 0) return QF_OK;
  sg_rgb_kernel<<<(int)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d_features, stride, num_lobes, d_dirs, M, d_rgb);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_hit_texels(const qf_mesh* mesh, const float* d_points, const int64_t* d_index_tri, int64_t M,
                             const float* d_uv_scaled, int texture_size, int64_t* d_texels, void* stream) {
  QF_REQUIRE(mesh && d_points && d_index_tri && d_uv_scaled && d_texels, "qf_hit_texels: NULL argument");
  if (M == 0) return QF_OK;
  hit_texels_kernel<<<(int)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(mesh->d_vertices, mesh->d_faces, d_uv_scaled,
                                                                             d_points, d_index_tri, M, texture_size, d_texels);
  QF_LAUNCH_CHECK();
  return QF_OK;
}


extern "C" int qf_sg_features_to_rgb_backward(const float* d_features, int64_t stride, int num_lobes, const float* d_dirs,
                                              int64_t M, const float* d_grad_rgb, float* d_grad_features, int64_t grad_stride,
                                              void* stream) {
  QF_REQUIRE(num_lobes >= 0 && num_lobes <= QF_MAX_LOBES && stride >= 3 + 7 * num_lobes && grad_stride >= 3 + 7 * num_lobes,
             "qf_sg_features_to_rgb_backward: lobes=%d stride=%lld grad_stride=%lld", num_lobes, (long long)stride, (long long)grad_stride);
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_features && d_dirs && d_grad_rgb && d_grad_features, "qf_sg_features_to_rgb_backward: NULL argument");
  sg_rgb_backward_kernel<<<(int)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(d_features, stride, num_lobes, d_dirs, M, d_grad_rgb,
                                                                                  d_grad_features, grad_stride);
  QF_LAUNCH_CHECK();
  return QF_OK;
}
