// Device helpers shared by the forward (field.cu) and backward (field_bwd.cu) field kernels.
#pragma once
#include "common.cuh"

namespace qf {

// ---- shared-memory weight image (halves); strides padded for conflict-free fragment loads
constexpr int kS32 = 40;   // row stride of a (out x 32) matrix
constexpr int kS64 = 72;   // row stride of a (out x 64) matrix
constexpr int kW1 = 0;                    // base  L1  64 x 32
constexpr int kW2 = kW1 + 64 * kS32;      // base  L2  16 x 64
constexpr int kW3 = kW2 + 16 * kS64;      // head  L1  64 x 32 (columns permuted, see prep)
constexpr int kW4 = kW3 + 64 * kS32;      // head  L2  64 x 64
constexpr int kW5 = kW4 + 64 * kS64;      // head  L3  16 x 64
constexpr int kWTotal = kW5 + 16 * kS64;  // 12032 halves = 24064 B
constexpr int kTileStride = 56;           // per-sample staging row: 32 enc + 16 SH + 8 pad (halves)

// ---- hash-grid encode of one point: 16 levels, trilinear, fp32 interpolation, fp16 result
// The level loop is kept ROLLED (two levels per trip, the gathers of both issued before either is consumed):
// fully unrolled it is ~13k instructions and the kernel stalls on instruction fetch.
struct Corner8 { uint32_t idx[8]; float fx, fy, fz; };

__device__ __forceinline__ void level_indices(const qf_grid_desc& d, int l, float x, float y, float z, Corner8& c) {
  const float scale = d.scale[l];
  const uint32_t res = d.resolution[l], size = d.size[l];
  float px = fmaf(scale, x, 0.5f), py = fmaf(scale, y, 0.5f), pz = fmaf(scale, z, 0.5f);
  const float flx = floorf(px), fly = floorf(py), flz = floorf(pz);
  const uint32_t cx = (uint32_t)(int)flx, cy = (uint32_t)(int)fly, cz = (uint32_t)(int)flz;
  c.fx = px - flx; c.fy = py - fly; c.fz = pz - flz;
  if (d.hashed[l]) {   // hashed levels always have size == 2^log2_hashmap_size
    const uint32_t mask = size - 1;
    const uint32_t y0 = cy * 2654435761u, y1 = (cy + 1) * 2654435761u, z0 = cz * 805459861u, z1 = (cz + 1) * 805459861u;
#pragma unroll
    for (int k = 0; k < 8; ++k) c.idx[k] = ((cx + (k & 1)) ^ ((k & 2) ? y1 : y0) ^ ((k & 4) ? z1 : z0)) & mask;
  } else {
    const uint32_t sy = res, sz = res * res;
    bool wrap = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      c.idx[k] = (cx + (k & 1)) + (cy + ((k >> 1) & 1)) * sy + (cz + ((k >> 2) & 1)) * sz;
      wrap |= c.idx[k] >= size;
    }
    if (wrap) {   // upper box faces and points outside the aabb wrap exactly like tcnn (index % level size)
#pragma unroll
      for (int k = 0; k < 8; ++k) c.idx[k] %= size;
    }
  }
}

// Trilinear blend of one level, tcnn's arithmetic (kernel_grid, recalled: `result = fma((T)weight, grid_val(...), result)`
// with T = __half): the corner weight ((wx*wy)*wz) is formed in fp32, rounded to half and accumulated with one half2 FMA
// per corner, corners in tcnn's order (bit d of k selects the upper neighbour along dimension d).  The oracle
// (quadfield_oracle.hashgrid_encode) restates exactly this, so the 32 encoded features agree bit for bit.
__device__ __forceinline__ uint32_t level_blend(const Corner8& c, const __half2* v) {
  const float wx[2] = {1.f - c.fx, c.fx}, wy[2] = {1.f - c.fy, c.fy}, wz[2] = {1.f - c.fz, c.fz};
  __half2 acc = __float2half2_rn(0.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float w = __fmul_rn(__fmul_rn(wx[k & 1], wy[(k >> 1) & 1]), wz[k >> 2]);
    acc = __hfma2(__float2half2_rn(w), v[k], acc);
  }
  return *reinterpret_cast<uint32_t*>(&acc);
}

// The 8 corner entries of one level.
// Dense levels: eight 4-byte gathers.  Hashed levels: the hash leaves x un-multiplied (prime 1), so the two x-neighbours
// of a (y,z) corner pair are entries i0 = (cx ^ h) & mask and i1 = ((cx+1) ^ h) & mask; for an EVEN cell x they differ
// in bit 0 only and sit in one aligned 8-byte word.  Every lane fetches the aligned pair around i0 with one 8-byte load;
// lanes with an odd cell x add a predicated 4-byte load for i1.  An L1TEX wavefront is spent per distinct 128-byte line
// per instruction, so at the fine levels (32 lanes, 32 lines) the x-pair costs 1.5 instead of 2 instructions' worth of
// wavefronts; the gather warps of the field kernel are bound by exactly that pipe (profiles/r2c).  Bit-identical output.
#ifndef QF_PAIRED_GATHER
#define QF_PAIRED_GATHER 1
#endif
__device__ __forceinline__ uint32_t ldg_entry(const __half2* t, uint32_t idx) {
  uint32_t r;
  asm("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 4, %2;\n\tld.global.nc.b32 %0, [a];\n\t}" : "=r"(r) : "r"(idx), "l"(t));
  return r;
}
__device__ __forceinline__ void load_corners(const qf_grid_desc& d, int l, const __half2* __restrict__ table, const Corner8& c,
                                             __half2* v) {
  const __half2* t = table + d.offset[l];
#if QF_PAIRED_GATHER
  if (d.hashed[l]) {
    // x-corner parity: idx[0] and idx[1] differ in bit 0 only  <=>  the cell's x is even (all four pairs alike)
    const bool even = ((c.idx[0] ^ c.idx[1]) == 1u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t i0 = c.idx[2 * j], i1 = c.idx[2 * j + 1];
      uint32_t wx, wy, lone = 0u;
      asm("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %2, 8, %3;\n\tld.global.nc.v2.b32 {%0, %1}, [a];\n\t}"
          : "=r"(wx), "=r"(wy) : "r"(i0 >> 1), "l"(t));
      if (!even) lone = ldg_entry(t, i1);
      const bool hi0 = i0 & 1u;
      const uint32_t e0 = hi0 ? wy : wx;
      const uint32_t e1 = even ? (hi0 ? wx : wy) : lone;
      v[2 * j] = *reinterpret_cast<const __half2*>(&e0);
      v[2 * j + 1] = *reinterpret_cast<const __half2*>(&e1);
    }
    return;
  }
#endif
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t r = ldg_entry(t, c.idx[k]);
    v[k] = *reinterpret_cast<const __half2*>(&r);
  }
}

template <typename Store>
__device__ __forceinline__ void encode_point(const qf_grid_desc& d, const __half2* __restrict__ table, float x, float y,
                                             float z, Store store) {
  const int L = d.n_levels;
#pragma unroll 1
  for (int l = 0; l < L; l += 2) {
    Corner8 c0, c1;
    __half2 v0[8], v1[8];
    level_indices(d, l, x, y, z, c0);
    load_corners(d, l, table, c0, v0);
    const bool two = l + 1 < L;
    if (two) {
      level_indices(d, l + 1, x, y, z, c1);
      load_corners(d, l + 1, table, c1, v1);
    }
    store(l, level_blend(c0, v0));
    if (two) store(l + 1, level_blend(c1, v1));
  }
}
#ifndef QF_BWD_PAIR_ATOMICS
#define QF_BWD_PAIR_ATOMICS 1
#endif
// backward of one level: scatter the gradient of its two features to the 8 corners (fp32 atomics)
__device__ __forceinline__ void scatter_level(const qf_grid_desc& d, float2* __restrict__ g_table, int l, float x, float y,
                                              float z, float g0, float g1) {
  if (g0 == 0.f && g1 == 0.f) return;
  Corner8 c;
  level_indices(d, l, x, y, z, c);
  float2* lvl = g_table + d.offset[l];
#if QF_BWD_PAIR_ATOMICS
  // The two x-neighbours of a (y,z) corner pair are adjacent entries whenever their indices differ in bit 0 only (hashed
  // levels: even cell x; dense levels: even linear index): ONE 16-byte vector reduction (REDG.E.ADD.F32x4) instead of two
  // 8-byte ones — 6 instead of 8 L2 atomics per level on average.  Level offsets are multiples of 8 entries and the entry
  // points require a 16-byte aligned table gradient, so the pair is aligned.  Same weights ((wx * wy) * wz), same sums.
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t i0 = c.idx[2 * j], i1 = c.idx[2 * j + 1];
    const float wy = (j & 1) ? c.fy : 1.f - c.fy, wz = (j & 2) ? c.fz : 1.f - c.fz;
    const float w0 = ((1.f - c.fx) * wy) * wz, w1 = (c.fx * wy) * wz;
    if ((i0 ^ i1) == 1u) {
      const bool lo0 = !(i0 & 1u);                     // i0 is the lower entry of the aligned pair
      const float a0 = lo0 ? w0 : w1, a1 = lo0 ? w1 : w0;
      float* p = reinterpret_cast<float*>(lvl + (i0 & ~1u));
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a0 * g0), "f"(a0 * g1), "f"(a1 * g0), "f"(a1 * g1) : "memory");
    } else {
      atomicAdd(lvl + i0, make_float2(w0 * g0, w0 * g1));
      atomicAdd(lvl + i1, make_float2(w1 * g0, w1 * g1));
    }
  }
#else
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float w = ((k & 1) ? c.fx : 1.f - c.fx) * ((k & 2) ? c.fy : 1.f - c.fy);
    w *= (k & 4) ? c.fz : 1.f - c.fz;
    atomicAdd(lvl + c.idx[k], make_float2(w * g0, w * g1));
  }
#endif
}

// scatter_level plus the gradient with respect to the normalised position (tcnn's grid input gradient): with
// enc_f = sum_k w_k v_k[f] and w_k the trilinear weight, d enc_f / d x = scale * sum_k (+-1)(w_y w_z)_k v_k[f], etc.
__device__ __forceinline__ void scatter_level_pos(const qf_grid_desc& d, float2* __restrict__ g_table,
                                                  const __half2* __restrict__ table, int l, float x, float y, float z, float g0,
                                                  float g1, float* gx) {
  if (g0 == 0.f && g1 == 0.f) return;
  Corner8 c;
  level_indices(d, l, x, y, z, c);
  float2* lvl = g_table + d.offset[l];
  const __half2* tl = table + d.offset[l];
  float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float wx = (k & 1) ? c.fx : 1.f - c.fx, wy = (k & 2) ? c.fy : 1.f - c.fy, wz = (k & 4) ? c.fz : 1.f - c.fz;
    const float w = (wx * wy) * wz;
    atomicAdd(lvl + c.idx[k], make_float2(w * g0, w * g1));
    const float2 v = __half22float2(__ldg(tl + c.idx[k]));
    const float gv = g0 * v.x + g1 * v.y;
    ax += ((k & 1) ? gv : -gv) * (wy * wz);
    ay += ((k & 2) ? gv : -gv) * (wx * wz);
    az += ((k & 4) ? gv : -gv) * (wx * wy);
  }
  const float sc = d.scale[l];
  gx[0] += sc * ax; gx[1] += sc * ay; gx[2] += sc * az;
}

__device__ __forceinline__ void sh4(float x, float y, float z, float* o) {
  float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// ---- tensor-core helpers (legacy HMMA path; the tcgen05 variant lives in field_tc.cu)
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }

// D[16 x 8*NT] += A[16 x 16*KT] * W^T, W row-major (n, k) in shared memory with row stride `stride`
template <int KT, int NT>
__device__ __forceinline__ void layer(float (*acc)[4], const uint32_t (*a)[4], const __half* __restrict__ w, int stride,
                                      int g, int t) {
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const __half* wr = w + (n * 8 + g) * stride + t * 2;
#pragma unroll
    for (int k = 0; k < KT; ++k) mma16816(acc[n], a[k], lds32(wr + k * 16), lds32(wr + k * 16 + 8));
  }
}


}  // namespace qf
