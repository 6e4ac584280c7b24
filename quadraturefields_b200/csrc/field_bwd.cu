// Training mode of the Instant-NGP field (NGPRadianceField under autograd: train_finetune.py:494-531,
// train_fit_sg.py): gradients of the hash table and of the five MLP matrices given dL/drgb and dL/dsigma at the
// samples.  Replaces tinycudann's grid backward (atomicAdd into the table gradient) and its MLP backward GEMMs.
//
//   ngp_backward_kernel   per warp 32 samples: RECOMPUTES the forward (re-gathers the table — cheaper than storing
//                         512 B of activations per sample in the forward pass), then walks the layers backwards on
//                         tensor cores (mma.sync, transposed weight images in shared memory, the accumulator fragment
//                         of one layer re-packed as the A fragment of the next).  The encoding gradient lands in
//                         exactly the fragment layout where lane (g,t) owns levels {t, t+4, t+8, t+12} of samples
//                         g and g+8, so the table scatter (8 float2 atomics per level) needs no shuffles.
//                         Layer inputs and output-gradients are written once as fp16 rows (960 B / sample).
//   weight_grad_kernel    dW_l = G_l^T A_l as a split-K tensor-core GEMM over the sample dimension, all five layers in
//                         one pass over the rows (cp.async ring, accumulators in registers)
//                         (ldmatrix.trans fragments, fp32 accumulate, one atomicAdd per element per chunk).
//   unpack_weight_grads   padded kernel image -> tinycudann flat layout (row-major (out,in), head columns un-permuted).
#include "field_common.cuh"

namespace qf {

// transposed weight images (halves): WT[n][k] = W[k][n]
constexpr int kT24 = 24;   // row stride for 16-wide rows
constexpr int kT5 = 0;                      // W5^T  64 x 16
constexpr int kT4 = kT5 + 64 * kT24;        // W4^T  64 x 64
constexpr int kT3 = kT4 + 64 * kS64;        // W3p^T 32 x 64   (kernel column order [SH | pad | feat])
constexpr int kT2 = kT3 + 32 * kS64;        // W2^T  64 x 16
constexpr int kT1 = kT2 + 64 * kT24;        // W1^T  32 x 64
constexpr int kTTotal = kT1 + 32 * kS64;    // 12288 halves

// per-sample rows written for the weight-gradient GEMMs
constexpr int kActRow = 256;  // [0,32) enc | [32,96) relu(h1) | [96,128) head input (kernel order) | [128,192) relu(h2) | [192,256) relu(h3)
constexpr int kGrdRow = 224;  // [0,64) dL/dh1pre | [64,80) dL/d base out | [80,144) dL/dh2pre | [144,208) dL/dh3pre | [208,224) dL/d rgb logits

__global__ void transpose_weights_kernel(const __half* __restrict__ img, __half* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTTotal) return;
  __half v = __float2half_rn(0.f);
  if (i < kT4) { int n = i / kT24, k = i % kT24; if (k < 16) v = img[kW5 + k * kS64 + n]; }
  else if (i < kT3) { int j = i - kT4, n = j / kS64, k = j % kS64; if (k < 64) v = img[kW4 + k * kS64 + n]; }
  else if (i < kT2) { int j = i - kT3, n = j / kS64, k = j % kS64; if (k < 64) v = img[kW3 + k * kS32 + n]; }
  else if (i < kT1) { int j = i - kT2, n = j / kT24, k = j % kT24; if (k < 16) v = img[kW2 + k * kS64 + n]; }
  else { int j = i - kT1, n = j / kS64, k = j % kS64; if (k < 64) v = img[kW1 + k * kS32 + n]; }
  out[i] = v;
}

struct BwdArgs {
  qf_grid_desc desc;
  const __half2* table;
  const __half* weights;    // forward image (kWTotal)
  const __half* weights_t;  // transposed image (kTTotal)
  const float* pos;
  int pos_stride;
  const float* dirs;
  const int64_t* ray64;
  int64_t M;
  const float* g_rgb;       // (M,3)
  const float* g_sigma;     // (M) or NULL
  const float* g_feat;      // (M,15) dL/d(geo features), FEAT mode only
  const float* g_absmax;    // device scalar: max |g_rgb| (for the fp16 dynamic-range scale)
  float2* g_table;          // (n_entries) float2, atomically accumulated
  __half* act;              // (M, kActRow)
  __half* grd;              // (M, kGrdRow)
  float* g_pos;             // (M,3) dL/dposition or NULL
  float inv_ext[3];         // 1 / (aabb_max - aabb_min)
};

__device__ __forceinline__ unsigned relu_mask8(float (*acc)[4]) {
  unsigned m = 0;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int q = 0; q < 4; ++q) m |= (acc[n][q] > 0.f ? 1u : 0u) << (n * 4 + q);
  return m;
}
__device__ __forceinline__ void apply_mask8(float (*acc)[4], unsigned m) {
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[n][q] = ((m >> (n * 4 + q)) & 1u) ? acc[n][q] : 0.f;
}
// accumulator fragments (16 x 64, C layout) -> A fragments of a 64-deep contraction; optional ReLU
template <bool RELU>
__device__ __forceinline__ void c_to_a(uint32_t (*a)[4], float (*acc)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float* c = acc[2 * k + h];
      a[k][2 * h] = RELU ? pack_h2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f)) : pack_h2(c[0], c[1]);
      a[k][2 * h + 1] = RELU ? pack_h2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f)) : pack_h2(c[2], c[3]);
    }
}
// The activation / gradient rows (960 B per sample, 424 MB per step at 442 K samples) are written once and read once by
// weight_grad_kernel: stored with the streaming (evict-first) policy so they do not push the 50 MB gradient table, which
// the scatter's atomics keep hitting, out of L2.
#ifndef QF_BWD_STREAM_ROWS
#define QF_BWD_STREAM_ROWS 1
#endif
__device__ __forceinline__ void st_row32(__half* p, uint32_t v) {
#if QF_BWD_STREAM_ROWS
  __stcs(reinterpret_cast<unsigned int*>(p), v);
#else
  *reinterpret_cast<uint32_t*>(p) = v;
#endif
}
// A fragments (rows g / g+8) -> fp16 row-major global rows: cols [col0, col0 + 16*KT)
template <int KT>
__device__ __forceinline__ void store_a(__half* __restrict__ base, int row_len, int col0, int64_t r_lo, int64_t r_hi, int64_t M,
                                        const uint32_t (*a)[4], int t) {
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int c = col0 + k * 16 + t * 2;
    if (r_lo < M) {
      st_row32(base + r_lo * row_len + c, a[k][0]);
      st_row32(base + r_lo * row_len + c + 8, a[k][2]);
    }
    if (r_hi < M) {
      st_row32(base + r_hi * row_len + c, a[k][1]);
      st_row32(base + r_hi * row_len + c + 8, a[k][3]);
    }
  }
}
// accumulator fragments (C layout, NT n-tiles) -> fp16 rows
template <int NT>
__device__ __forceinline__ void store_c(__half* __restrict__ base, int row_len, int col0, int64_t r_lo, int64_t r_hi, int64_t M,
                                        float (*acc)[4], int t) {
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int c = col0 + n * 8 + t * 2;
    if (r_lo < M) st_row32(base + r_lo * row_len + c, pack_h2(acc[n][0], acc[n][1]));
    if (r_hi < M) st_row32(base + r_hi * row_len + c, pack_h2(acc[n][2], acc[n][3]));
  }
}

// Gradients enter the fp16 tensor-core path scaled by a power of two that brings max|dL/drgb| to [0.5, 1) — what the
// reference gets from torch GradScaler(2**10) (train_finetune.py) — and are unscaled in fp32 before they are accumulated.
__device__ __forceinline__ float grad_scale_from(float absmax) {
  if (!(absmax > 0.f) || !isfinite(absmax)) return 1.0f;
  int e;
  frexpf(absmax, &e);            // absmax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.0f, -e);
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

#ifndef QF_BWD_WARPS
#define QF_BWD_WARPS 4
#endif
constexpr int kBwdWarps = QF_BWD_WARPS, kBwdThreads = kBwdWarps * 32;
constexpr int kBwdSmemBytes = (kWTotal + kTTotal + kBwdWarps * 32 * kTileStride) * 2 + kBwdWarps * 32 * 3 * 4;

// FEAT = false: backward of the full forward (rgb, sigma).  FEAT = true: backward of `query_density(return_feat=True)`
// — the upstream gradient arrives at the 16 outputs of the base MLP (sigma and the 15 geo features), the tcnn head is
// not part of the graph (the spherical-Gaussian field feeds the features to a torch decoder instead).
#ifndef QF_BWD_MIN_CTAS
#define QF_BWD_MIN_CTAS 3
#endif
template <bool FEAT>
__global__ void __launch_bounds__(kBwdThreads, QF_BWD_MIN_CTAS) ngp_backward_kernel(const BwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* s_w = reinterpret_cast<__half*>(smem_raw);
  __half* s_wt = s_w + kWTotal;
  __half* s_tile_all = s_wt + kTTotal;
  float* s_x_all = reinterpret_cast<float*>(s_tile_all + kBwdWarps * 32 * kTileStride);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < kWTotal / 8; i += kBwdThreads) dst[i] = __ldg(src + i);
    const uint4* srct = reinterpret_cast<const uint4*>(a.weights_t);
    uint4* dstt = reinterpret_cast<uint4*>(s_wt);
    for (int i = tid; i < kTTotal / 8; i += kBwdThreads) dstt[i] = __ldg(srct + i);
  }
  __syncthreads();
  const int64_t M = a.M;
  const float gscale = grad_scale_from(__ldg(a.g_absmax)), ginv = 1.0f / gscale;
  __half* tile = s_tile_all + warp * 32 * kTileStride;
  float* sx = s_x_all + warp * 32 * 3;
  const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
  const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};

  for (int64_t base = ((int64_t)blockIdx.x * kBwdWarps + warp) * 32; base < M; base += (int64_t)gridDim.x * kBwdThreads) {
    const int64_t i = base + lane;
    const bool valid = i < M;
    float x = 0.5f, y = 0.5f, z = 0.5f;
    bool sel = false;
    if (valid) {
      const float* p = a.pos + i * a.pos_stride;
      x = __fdiv_rn(__ldg(p) - amin[0], aext[0]);
      y = __fdiv_rn(__ldg(p + 1) - amin[1], aext[1]);
      z = __fdiv_rn(__ldg(p + 2) - amin[2], aext[2]);
      sel = (x > 0.f) && (x < 1.f) && (y > 0.f) && (y < 1.f) && (z > 0.f) && (z < 1.f);
    }
    sx[lane * 3] = x; sx[lane * 3 + 1] = y; sx[lane * 3 + 2] = z;
    uint32_t* row32 = reinterpret_cast<uint32_t*>(tile + lane * kTileStride);
    encode_point(a.desc, a.table, x, y, z, [&](int l, uint32_t h2) { row32[l] = h2; });
    uint4* row = reinterpret_cast<uint4*>(row32);
    if (!FEAT) {
      float dx = 0.f, dy = 0.f, dz = 1.f;
      if (valid) {
        int64_t r = a.ray64 ? __ldg(a.ray64 + i) : i;
        const float* dp = a.dirs + 3 * r;
        dx = ((__ldg(dp) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dy = ((__ldg(dp + 1) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dz = ((__ldg(dp + 2) + 1.0f) / 2.0f) * 2.0f - 1.0f;
      }
      float sh[16];
      sh4(dx, dy, dz, sh);
      row[4] = make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sh[3]), pack_h2(sh[4], sh[5]), pack_h2(sh[6], sh[7]));
      row[5] = make_uint4(pack_h2(sh[8], sh[9]), pack_h2(sh[10], sh[11]), pack_h2(sh[12], sh[13]), pack_h2(sh[14], sh[15]));
    }
    if (valid) {  // layer-1 input rows (the encoding), one 64-byte row per lane
      uint4* dst = reinterpret_cast<uint4*>(a.act + i * kActRow);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#if QF_BWD_STREAM_ROWS
        __stcs(dst + q, row[q]);
#else
        dst[q] = row[q];
#endif
      }
    }
    const unsigned selmask = __ballot_sync(0xffffffffu, sel);
    __syncwarp();

#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      if (base + mt * 16 >= M) break;
      const int64_t r_lo = base + mt * 16 + g, r_hi = r_lo + 8;
      const __half* ta = tile + (mt * 16 + g) * kTileStride + t * 2;
      const __half* tb = ta + 8 * kTileStride;
      // ================= forward recompute =================
      uint32_t a1[2][4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        a1[k][0] = lds32(ta + k * 16); a1[k][1] = lds32(tb + k * 16);
        a1[k][2] = lds32(ta + k * 16 + 8); a1[k][3] = lds32(tb + k * 16 + 8);
      }
      float acc[8][4];
      uint32_t af[4][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      layer<2, 8>(acc, a1, s_w + kW1, kS32, g, t);
      const unsigned mask1 = relu_mask8(acc);
      c_to_a<true>(af, acc);
      store_a<4>(a.act, kActRow, 32, r_lo, r_hi, M, af, t);
      float acc2[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) acc2[n][0] = acc2[n][1] = acc2[n][2] = acc2[n][3] = 0.f;
      layer<4, 2>(acc2, af, s_w + kW2, kS64, g, t);
      const float h0_lo = acc2[0][0], h0_hi = acc2[0][2];   // meaningful on t == 0
      float gin[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) gin[n][0] = gin[n][1] = gin[n][2] = gin[n][3] = 0.f;
      if (!FEAT) {
        uint32_t a3[2][4];
        a3[0][0] = lds32(ta + 32); a3[0][1] = lds32(tb + 32); a3[0][2] = lds32(ta + 40); a3[0][3] = lds32(tb + 40);
        a3[1][0] = pack_h2(t == 0 ? 1.0f : acc2[0][0], acc2[0][1]);
        a3[1][1] = pack_h2(t == 0 ? 1.0f : acc2[0][2], acc2[0][3]);
        a3[1][2] = pack_h2(acc2[1][0], acc2[1][1]);
        a3[1][3] = pack_h2(acc2[1][2], acc2[1][3]);
        store_a<2>(a.act, kActRow, 96, r_lo, r_hi, M, a3, t);
  #pragma unroll
        for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        layer<2, 8>(acc, a3, s_w + kW3, kS32, g, t);
        const unsigned mask3 = relu_mask8(acc);
        c_to_a<true>(af, acc);
        store_a<4>(a.act, kActRow, 128, r_lo, r_hi, M, af, t);
  #pragma unroll
        for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        layer<4, 8>(acc, af, s_w + kW4, kS64, g, t);
        const unsigned mask4 = relu_mask8(acc);
        c_to_a<true>(af, acc);
        store_a<4>(a.act, kActRow, 192, r_lo, r_hi, M, af, t);
        float acc5[1][4] = {{0.f, 0.f, 0.f, 0.f}};
        layer<4, 1>(acc5, af, s_w + kW5, kS64, g, t);
        // ================= backward =================
        // dL/d rgb logits = g_rgb * rgb (1 - rgb); lane t owns logit columns 2t, 2t+1 of rows g and g+8
        uint32_t ga[1][4] = {{0u, 0u, 0u, 0u}};
        if (t < 2) {
          auto dsig = [gscale](float v) { float s = 1.0f / (1.0f + __expf(-v)); return gscale * s * (1.0f - s); };
          float gl0 = 0.f, gl1 = 0.f, gh0 = 0.f, gh1 = 0.f;
          if (r_lo < M) { gl0 = a.g_rgb[3 * r_lo + 2 * t] * dsig(acc5[0][0]); if (t == 0) gl1 = a.g_rgb[3 * r_lo + 1] * dsig(acc5[0][1]); }
          if (r_hi < M) { gh0 = a.g_rgb[3 * r_hi + 2 * t] * dsig(acc5[0][2]); if (t == 0) gh1 = a.g_rgb[3 * r_hi + 1] * dsig(acc5[0][3]); }
          ga[0][0] = pack_h2(gl0, gl1);
          ga[0][1] = pack_h2(gh0, gh1);
        }
        store_a<1>(a.grd, kGrdRow, 208, r_lo, r_hi, M, ga, t);
        // head L3:  dL/dh3 = g_o W5, masked
  #pragma unroll
        for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        layer<1, 8>(acc, ga, s_wt + kT5, kT24, g, t);
        apply_mask8(acc, mask4);
        store_c<8>(a.grd, kGrdRow, 144, r_lo, r_hi, M, acc, t);
        c_to_a<false>(af, acc);
        // head L2:  dL/dh2 = g_h3 W4, masked
  #pragma unroll
        for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        layer<4, 8>(acc, af, s_wt + kT4, kS64, g, t);
        apply_mask8(acc, mask3);
        store_c<8>(a.grd, kGrdRow, 80, r_lo, r_hi, M, acc, t);
        c_to_a<false>(af, acc);
        // head L1: only the [pad | feat] half of the input carries gradient on to the base MLP
        layer<4, 2>(gin, af, s_wt + kT3 + 16 * kS64, kS64, g, t);
      } else {
        // upstream gradient of the 15 geo features: output column c = n*8 + 2t + q holds feature c-1 (column 0 = logit)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int col = n * 8 + 2 * t + q;
            if (col >= 1) {
              if (r_lo < M) gin[n][q] = gscale * a.g_feat[r_lo * 15 + col - 1];
              if (r_hi < M) gin[n][2 + q] = gscale * a.g_feat[r_hi * 15 + col - 1];
            }
          }
        }
      }
      if (t == 0) {  // column 0 of the base output is the density logit: dL/dh0 = dL/dsigma * exp(min(h0-1, 15)) * selector
        float gs_lo = (a.g_sigma && r_lo < M) ? a.g_sigma[r_lo] : 0.f, gs_hi = (a.g_sigma && r_hi < M) ? a.g_sigma[r_hi] : 0.f;
        gin[0][0] = ((selmask >> (mt * 16 + g)) & 1u) ? gscale * gs_lo * expf(fminf(h0_lo - 1.0f, 15.0f)) : 0.f;
        gin[0][2] = ((selmask >> (mt * 16 + g + 8)) & 1u) ? gscale * gs_hi * expf(fminf(h0_hi - 1.0f, 15.0f)) : 0.f;
      }
      store_c<2>(a.grd, kGrdRow, 64, r_lo, r_hi, M, gin, t);
      uint32_t g2[1][4];
      g2[0][0] = pack_h2(gin[0][0], gin[0][1]); g2[0][1] = pack_h2(gin[0][2], gin[0][3]);
      g2[0][2] = pack_h2(gin[1][0], gin[1][1]); g2[0][3] = pack_h2(gin[1][2], gin[1][3]);
      // base L2:  dL/dh1 = g_out2 W2, masked
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      layer<1, 8>(acc, g2, s_wt + kT2, kT24, g, t);
      apply_mask8(acc, mask1);
      store_c<8>(a.grd, kGrdRow, 0, r_lo, r_hi, M, acc, t);
      c_to_a<false>(af, acc);
      // base L1:  dL/denc = g_h1 W1 -> table scatter
      float ge[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) ge[n][0] = ge[n][1] = ge[n][2] = ge[n][3] = 0.f;
      layer<4, 4>(ge, af, s_wt + kT1, kS64, g, t);
      const float* xl = sx + (mt * 16 + g) * 3;
      const float* xh = xl + 24;
      if (a.g_pos == nullptr) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int l = n * 4 + t;   // accumulator columns n*8 + 2t, +1 are the two features of level n*4 + t
          if (r_lo < M) scatter_level(a.desc, a.g_table, l, xl[0], xl[1], xl[2], ge[n][0] * ginv, ge[n][1] * ginv);
          if (r_hi < M) scatter_level(a.desc, a.g_table, l, xh[0], xh[1], xh[2], ge[n][2] * ginv, ge[n][3] * ginv);
        }
      } else {
        // also dL/dposition (tcnn input gradient): each lane sums its 4 levels, the 4 lanes of a row are reduced
        float gp_lo[3] = {0.f, 0.f, 0.f}, gp_hi[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int n = 0; n < 4; ++n) {
          const int l = n * 4 + t;
          if (r_lo < M) scatter_level_pos(a.desc, a.g_table, a.table, l, xl[0], xl[1], xl[2], ge[n][0] * ginv, ge[n][1] * ginv, gp_lo);
          if (r_hi < M) scatter_level_pos(a.desc, a.g_table, a.table, l, xh[0], xh[1], xh[2], ge[n][2] * ginv, ge[n][3] * ginv, gp_hi);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          gp_lo[c] += __shfl_xor_sync(0xffffffffu, gp_lo[c], 1); gp_lo[c] += __shfl_xor_sync(0xffffffffu, gp_lo[c], 2);
          gp_hi[c] += __shfl_xor_sync(0xffffffffu, gp_hi[c], 1); gp_hi[c] += __shfl_xor_sync(0xffffffffu, gp_hi[c], 2);
        }
        if (t == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (r_lo < M) a.g_pos[3 * r_lo + c] = gp_lo[c] * a.inv_ext[c];
            if (r_hi < M) a.g_pos[3 * r_hi + c] = gp_hi[c] * a.inv_ext[c];
          }
        }
      }
    }
    __syncwarp();
  }
}

// tcnn grid backward alone: dL/dtable += scatter of dL/denc (M, 2L) at x01 (M,3); thread per (sample, level)
__global__ void hashgrid_backward_kernel(const qf_grid_desc d, const float* __restrict__ x01, const float* __restrict__ g_enc,
                                         int64_t M, float2* __restrict__ g_table) {
  const int L = d.n_levels;
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= M * L) return;
  const int64_t i = id / L;
  const int l = (int)(id % L);
  scatter_level(d, g_table, l, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], g_enc[i * 2 * L + 2 * l], g_enc[i * 2 * L + 2 * l + 1]);
}

// ---------------------------------------------------------------- dW_l = G_l^T A_l
struct LayerGemm { int N, K, act_off, grd_off, img_off, img_stride; };
__constant__ LayerGemm c_layers[5] = {
    {64, 32, 0, 0, kW1, kS32}, {16, 64, 32, 64, kW2, kS64}, {64, 32, 96, 80, kW3, kS32},
    {64, 64, 128, 144, kW4, kS64}, {16, 64, 192, 208, kW5, kS64}};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* p) {
  unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// One CTA streams whole sample rows (activations 512 B + gradients 448 B, read exactly once, fully coalesced) through a
// 3-stage cp.async ring of 16-sample tiles; its four warps own disjoint sets of the 14 (layer, 16-row m-tile) items and
// keep their fp32 accumulators in registers for the CTA's whole share of the samples:
//   warp 0: base L1 (4 m-tiles, K=32) + base L2 (K=64)     96 accumulators, 24 MMAs per tile
//   warp 1: head L1 (4 m-tiles, K=32) + head L3 (K=64)     96 accumulators, 24 MMAs per tile
//   warp 2: head L2 m-tiles 0,1 (K=64)                     64 accumulators, 16 MMAs per tile
//   warp 3: head L2 m-tiles 2,3 (K=64)                     64 accumulators, 16 MMAs per tile
// (The first version launched one CTA per item and re-read every row once per item through un-pipelined loads: 0.27 ms
// for 442 K samples, 1.2 KB of DRAM reads per sample.)
constexpr int kWgTile = 16;          // samples per pipeline stage
constexpr int kWgStages = 3;
constexpr int kWgActStride = 264;    // halves; 528 B rows -> the 8 rows of an ldmatrix hit 8 distinct 16-byte bank groups
constexpr int kWgGrdStride = 232;    // halves; 464 B rows, same property
constexpr int kWgStageHalves = kWgTile * (kWgActStride + kWgGrdStride);
constexpr int kWgSmemBytes = kWgStages * kWgStageHalves * 2;   // 47616 B

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// acc[KT8][4] += G^T (16 gradient columns at grd_col) x A (KT8*8 input columns at act_col) over the 16 samples of a stage
template <int KT8>
__device__ __forceinline__ void wg_item(float (&acc)[KT8][4], const __half* sa, const __half* sg, int act_col, int grd_col, int lane) {
  uint32_t af[4];
  {
    const int j = lane >> 3, r = lane & 7;
    ldmatrix_x4_trans(af, sg + ((j >> 1) * 8 + r) * kWgGrdStride + grd_col + (j & 1) * 8);
  }
#pragma unroll
  for (int n = 0; n < KT8; ++n) {
    uint32_t b0, b1;
    ldmatrix_x2_trans(b0, b1, sa + (lane & 15) * kWgActStride + act_col + n * 8);
    mma16816(acc[n], af, b0, b1);
  }
}

template <int KT8>
__device__ __forceinline__ void wg_flush(const float (&acc)[KT8][4], float* __restrict__ stage, int layer, int mtile, int lane,
                                         int col0 = 0) {
  const LayerGemm L = c_layers[layer];
  const int g = lane >> 2, t = lane & 3;
  float* out = stage + L.img_off + (mtile * 16) * L.img_stride + col0;
#pragma unroll
  for (int n = 0; n < KT8; ++n) {
    const int c = n * 8 + t * 2;
    atomicAdd(out + g * L.img_stride + c, acc[n][0]);
    atomicAdd(out + g * L.img_stride + c + 1, acc[n][1]);
    atomicAdd(out + (g + 8) * L.img_stride + c, acc[n][2]);
    atomicAdd(out + (g + 8) * L.img_stride + c + 1, acc[n][3]);
  }
}

template <int KT8, int NI>
__device__ __forceinline__ void wg_zero(float (&acc)[NI][KT8][4]) {
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int n = 0; n < KT8; ++n) acc[i][n][0] = acc[i][n][1] = acc[i][n][2] = acc[i][n][3] = 0.f;
}

// base_only: the geo-feature backward (qf_ngp_backward_features) wrote only the base MLP's columns
__global__ void __launch_bounds__(128, 3) weight_grad_kernel(const __half* __restrict__ act, const __half* __restrict__ grd,
                                                             int64_t M, float* __restrict__ stage, int base_only) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  __half* smem = reinterpret_cast<__half*>(wg_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (M + kWgTile - 1) / kWgTile;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto issue = [&](int64_t it) {          // tile `it` of this CTA -> stage it % kWgStages (an empty group past the end)
    if (it < my_tiles) {
      const int64_t s0 = ((int64_t)blockIdx.x + it * gridDim.x) * kWgTile;
      __half* sa = smem + (it % kWgStages) * kWgStageHalves;
      __half* sg = sa + kWgTile * kWgActStride;
      constexpr int kA = kWgTile * (kActRow / 8), kG = kWgTile * (kGrdRow / 8);     // 16-byte pieces: 512 + 448
      for (int q = tid; q < kA + kG; q += 128) {
        if (q < kA) {
          const int r = q / (kActRow / 8), c = q % (kActRow / 8);
          const bool ok = s0 + r < M;
          cp_async16_zfill(sa + r * kWgActStride + c * 8, act + (ok ? (s0 + r) : 0) * kActRow + c * 8, ok);
        } else {
          const int q2 = q - kA, r = q2 / (kGrdRow / 8), c = q2 % (kGrdRow / 8);
          const bool ok = s0 + r < M;
          cp_async16_zfill(sg + r * kWgGrdStride + c * 8, grd + (ok ? (s0 + r) : 0) * kGrdRow + c * 8, ok);
        }
      }
    }
    cp_async_commit();
  };

  float acc32[4][4][4];      // warps 0,1: four K=32 m-tiles; warps 2,3: [0],[1] = the two column halves of their second K=64 m-tile
  float acc64[1][8][4];      // warps 0,1: the 16-row K=64 layer; warps 2,3: their first K=64 m-tile of head L2
  wg_zero<4, 4>(acc32);
  wg_zero<8, 1>(acc64);
  for (int i = 0; i < kWgStages - 1; ++i) issue(i);
  for (int64_t it = 0; it < my_tiles; ++it) {
    cp_async_wait<kWgStages - 2>();
    __syncthreads();                       // tile `it` has landed for every thread; everyone is done with tile it-1
    issue(it + kWgStages - 1);             // refills the stage of tile it-1
    const __half* sa = smem + (it % kWgStages) * kWgStageHalves;
    const __half* sg = sa + kWgTile * kWgActStride;
    if (warp == 0) {
#pragma unroll
      for (int m = 0; m < 4; ++m) wg_item<4>(acc32[m], sa, sg, 0, m * 16, lane);            // base L1: enc -> dL/dh1pre
      wg_item<8>(acc64[0], sa, sg, 32, 64, lane);                                           // base L2: relu(h1) -> dL/d base out
    } else if (!base_only) {
      if (warp == 1) {
#pragma unroll
        for (int m = 0; m < 4; ++m) wg_item<4>(acc32[m], sa, sg, 96, 80 + m * 16, lane);    // head L1: head input -> dL/dh2pre
        wg_item<8>(acc64[0], sa, sg, 192, 208, lane);                                       // head L3: relu(h3) -> dL/d logits
      } else {
        const int m0 = (warp - 2) * 2;
        wg_item<8>(acc64[0], sa, sg, 128, 144 + m0 * 16, lane);                             // head L2: relu(h2) -> dL/dh3pre
        wg_item<4>(acc32[0], sa, sg, 128, 144 + (m0 + 1) * 16, lane);
        wg_item<4>(acc32[1], sa, sg, 160, 144 + (m0 + 1) * 16, lane);
      }
    }
  }
  cp_async_wait<0>();
  if (my_tiles == 0) return;
  if (warp == 0) {
#pragma unroll
    for (int m = 0; m < 4; ++m) wg_flush<4>(acc32[m], stage, 0, m, lane);
    wg_flush<8>(acc64[0], stage, 1, 0, lane);
  } else if (!base_only) {
    if (warp == 1) {
#pragma unroll
      for (int m = 0; m < 4; ++m) wg_flush<4>(acc32[m], stage, 2, m, lane);
      wg_flush<8>(acc64[0], stage, 4, 0, lane);
    } else {
      const int m0 = (warp - 2) * 2;
      wg_flush<8>(acc64[0], stage, 3, m0, lane);
      wg_flush<4>(acc32[0], stage, 3, m0 + 1, lane);
      wg_flush<4>(acc32[1], stage, 3, m0 + 1, lane, 32);
    }
  }
}

static int launch_weight_grad(const __half* act, const __half* grd, int64_t M, float* stage, bool base_only, cudaStream_t st) {
  QF_ENSURE_DYNAMIC_SMEM(weight_grad_kernel, kWgSmemBytes);
  const int64_t tiles = ceil_div(M, kWgTile);
  const int blocks = (int)(tiles < (int64_t)kNumSMs * 3 ? tiles : (int64_t)kNumSMs * 3);
  weight_grad_kernel<<<blocks, 128, kWgSmemBytes, st>>>(act, grd, M, stage, base_only ? 1 : 0);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

// kernel image (fp32) -> tinycudann flat layouts, accumulated into the caller's gradient buffers
__global__ void unpack_weight_grads_kernel(const float* __restrict__ stage, const float* __restrict__ g_absmax,
                                           float* __restrict__ g_base, float* __restrict__ g_head) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float ginv = 1.0f / grad_scale_from(__ldg(g_absmax));
  const int n_base = 64 * 32 + 16 * 64, n_head = 64 * 32 + 64 * 64 + 16 * 64;
  if (i < n_base) {
    float v;
    if (i < 2048) v = stage[kW1 + (i / 32) * kS32 + (i % 32)];
    else { int j = i - 2048; v = stage[kW2 + (j / 64) * kS64 + (j % 64)]; }
    g_base[i] += v * ginv;
  } else if (i < n_base + n_head) {
    int j = i - n_base;
    float v;
    if (j < 2048) {
      int r = j / 32, c = j % 32;
      int kc = c < 16 ? c : (c == 31 ? 16 : c + 1);   // tcnn [SH | feat | pad] -> kernel [SH | pad | feat]
      v = stage[kW3 + r * kS32 + kc];
    } else if (j < 2048 + 4096) { int q = j - 2048; v = stage[kW4 + (q / 64) * kS64 + (q % 64)]; }
    else { int q = j - 2048 - 4096; v = stage[kW5 + (q / 64) * kS64 + (q % 64)]; }
    if (g_head) g_head[j] += v * ginv;
  }
}

}  // namespace qf

using namespace qf;

static size_t align256b(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int qf_hashgrid_backward(const qf_ngp* f, const float* d_x01, const float* d_grad_enc, int64_t M, float* d_grad_table,
                                    void* stream) {
  if (M == 0) return QF_OK;
  QF_REQUIRE(f && d_x01 && d_grad_enc && d_grad_table, "qf_hashgrid_backward: NULL argument");
  QF_REQUIRE((reinterpret_cast<uintptr_t>(d_grad_table) & 15) == 0, "qf_hashgrid_backward: d_grad_table must be 16-byte aligned");
  const int64_t n = M * f->desc.n_levels;
  hashgrid_backward_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(f->desc, d_x01, d_grad_enc, M,
                                                                                       reinterpret_cast<float2*>(d_grad_table));
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" size_t qf_ngp_backward_workspace_bytes(int64_t M) {
  return 256 + align256b(sizeof(__half) * kTTotal) + align256b(sizeof(float) * kWTotal) + align256b(sizeof(__half) * (size_t)M * kActRow) +
         align256b(sizeof(__half) * (size_t)M * kGrdRow) + 1024;
}

extern "C" int qf_ngp_backward(const qf_ngp* f, const float* d_positions, const float* d_directions, const int64_t* d_ray_index,
                               int64_t M, const float* d_grad_rgb, const float* d_grad_density, float* d_grad_table,
                               float* d_grad_base_w, float* d_grad_head_w, void* d_workspace, size_t workspace_bytes, void* stream) {
  return qf_ngp_backward_inputs(f, d_positions, d_directions, d_ray_index, M, d_grad_rgb, d_grad_density, d_grad_table,
                                d_grad_base_w, d_grad_head_w, nullptr, d_workspace, workspace_bytes, stream);
}

extern "C" int qf_ngp_backward_inputs(const qf_ngp* f, const float* d_positions, const float* d_directions,
                                      const int64_t* d_ray_index, int64_t M, const float* d_grad_rgb,
                                      const float* d_grad_density, float* d_grad_table, float* d_grad_base_w,
                                      float* d_grad_head_w, float* d_grad_positions, void* d_workspace, size_t workspace_bytes,
                                      void* stream) {
  if (M == 0) return QF_OK;
  QF_REQUIRE(f && d_positions && d_directions && d_grad_rgb && d_grad_table && d_grad_base_w && d_grad_head_w && d_workspace,
             "qf_ngp_backward: NULL argument");
  QF_REQUIRE(f->d_weights, "qf_ngp_backward: this field handle holds a grid only (qf_grid_create)");
  QF_REQUIRE(workspace_bytes >= qf_ngp_backward_workspace_bytes(M), "qf_ngp_backward: workspace too small");
  QF_REQUIRE((reinterpret_cast<uintptr_t>(d_grad_table) & 15) == 0, "qf_ngp_backward: d_grad_table must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  float* gmax = (float*)ws; ws += 256;
  QF_CUDA_CHECK(cudaMemsetAsync(gmax, 0, sizeof(float), st));
  absmax_kernel<<<kNumSMs * 2, 256, 0, st>>>(d_grad_rgb, 3 * M, gmax);
  __half* wt = (__half*)ws; ws += align256b(sizeof(__half) * kTTotal);
  float* stage = (float*)ws; ws += align256b(sizeof(float) * kWTotal);
  __half* act = (__half*)ws; ws += align256b(sizeof(__half) * (size_t)M * kActRow);
  __half* grd = (__half*)ws;
  transpose_weights_kernel<<<(int)ceil_div(kTTotal, 256), 256, 0, st>>>(f->d_weights, wt);
  QF_CUDA_CHECK(cudaMemsetAsync(stage, 0, sizeof(float) * kWTotal, st));
  BwdArgs a = {};
  a.desc = f->desc; a.table = f->d_table; a.weights = f->d_weights; a.weights_t = wt;
  a.pos = d_positions; a.pos_stride = 3; a.dirs = d_directions; a.ray64 = d_ray_index; a.M = M;
  a.g_rgb = d_grad_rgb; a.g_sigma = d_grad_density; a.g_absmax = gmax; a.g_table = reinterpret_cast<float2*>(d_grad_table);
  a.act = act; a.grd = grd;
  a.g_pos = d_grad_positions;
  for (int c = 0; c < 3; ++c) a.inv_ext[c] = 1.0f / (f->desc.aabb[3 + c] - f->desc.aabb[c]);
  QF_ENSURE_DYNAMIC_SMEM(ngp_backward_kernel<false>, kBwdSmemBytes);
  int64_t tiles = ceil_div(M, kBwdThreads);
  int blocks = (int)(tiles < (int64_t)kNumSMs * QF_BWD_MIN_CTAS ? tiles : (int64_t)kNumSMs * QF_BWD_MIN_CTAS);   // 3 CTAs/SM: 64.5 KB smem, <= 168 registers
  ngp_backward_kernel<false><<<blocks, kBwdThreads, kBwdSmemBytes, st>>>(a);
  QF_LAUNCH_CHECK();
  { int rc = launch_weight_grad(act, grd, M, stage, false, st); if (rc != QF_OK) return rc; }
  unpack_weight_grads_kernel<<<(int)ceil_div(3072 + 7168, 256), 256, 0, st>>>(stage, gmax, d_grad_base_w, d_grad_head_w);
  QF_LAUNCH_CHECK();
  return QF_OK;
}


// Backward of `query_density(x, return_feat=True)` (ngp.py:402-428): gradients of sigma (M) and of the 15 geo features
// (M,15) into the hash table and the base MLP (and optionally the positions).  The spherical-Gaussian field trains through
// this (train_fit_sg.py: features -> torch decoder -> SG mixture -> composite).  Same workspace as qf_ngp_backward.
extern "C" int qf_ngp_backward_features(const qf_ngp* f, const float* d_positions, int64_t M, const float* d_grad_density,
                                        const float* d_grad_feat, float* d_grad_table, float* d_grad_base_w,
                                        float* d_grad_positions, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (M == 0) return QF_OK;
  QF_REQUIRE(f && d_positions && d_grad_feat && d_grad_table && d_grad_base_w && d_workspace, "qf_ngp_backward_features: NULL argument");
  QF_REQUIRE(f->d_weights, "qf_ngp_backward_features: this field handle holds a grid only (qf_grid_create)");
  QF_REQUIRE(workspace_bytes >= qf_ngp_backward_workspace_bytes(M), "qf_ngp_backward_features: workspace too small");
  QF_REQUIRE((reinterpret_cast<uintptr_t>(d_grad_table) & 15) == 0, "qf_ngp_backward_features: d_grad_table must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)d_workspace;
  float* gmax = (float*)ws; ws += 256;
  QF_CUDA_CHECK(cudaMemsetAsync(gmax, 0, sizeof(float), st));
  absmax_kernel<<<kNumSMs * 2, 256, 0, st>>>(d_grad_feat, 15 * M, gmax);
  __half* wt = (__half*)ws; ws += align256b(sizeof(__half) * kTTotal);
  float* stage = (float*)ws; ws += align256b(sizeof(float) * kWTotal);
  __half* act = (__half*)ws; ws += align256b(sizeof(__half) * (size_t)M * kActRow);
  __half* grd = (__half*)ws;
  transpose_weights_kernel<<<(int)ceil_div(kTTotal, 256), 256, 0, st>>>(f->d_weights, wt);
  QF_CUDA_CHECK(cudaMemsetAsync(stage, 0, sizeof(float) * kWTotal, st));
  BwdArgs a = {};
  a.desc = f->desc; a.table = f->d_table; a.weights = f->d_weights; a.weights_t = wt;
  a.pos = d_positions; a.pos_stride = 3; a.M = M;
  a.g_sigma = d_grad_density; a.g_feat = d_grad_feat; a.g_absmax = gmax; a.g_table = reinterpret_cast<float2*>(d_grad_table);
  a.act = act; a.grd = grd; a.g_pos = d_grad_positions;
  for (int c = 0; c < 3; ++c) a.inv_ext[c] = 1.0f / (f->desc.aabb[3 + c] - f->desc.aabb[c]);
  QF_ENSURE_DYNAMIC_SMEM(ngp_backward_kernel<true>, kBwdSmemBytes);
  int64_t tiles = ceil_div(M, kBwdThreads);
  int blocks = (int)(tiles < (int64_t)kNumSMs * QF_BWD_MIN_CTAS ? tiles : (int64_t)kNumSMs * QF_BWD_MIN_CTAS);
  ngp_backward_kernel<true><<<blocks, kBwdThreads, kBwdSmemBytes, st>>>(a);
  QF_LAUNCH_CHECK();
  { int rc = launch_weight_grad(act, grd, M, stage, true, st); if (rc != QF_OK) return rc; }      // the two base layers only
  unpack_weight_grads_kernel<<<(int)ceil_div(3072, 256), 256, 0, st>>>(stage, gmax, d_grad_base_w, nullptr);
  QF_LAUNCH_CHECK();
  return QF_OK;
}
