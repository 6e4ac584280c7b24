// (2)+(3) Instant-NGP field: multiresolution hash-grid gather fused with the 64-wide density and
// colour MLPs on tensor cores.  Replaces tinycudann HashGrid / FullyFusedMLP / SphericalHarmonics
// behind NGPRadianceField (radiance_fields/ngp.py:657-809).
//
// One lane encodes one hit sample (16 levels x 8 corner gathers of one half2 = 512 B algorithmic,
// served from L2: the 25 MB fp16 table of T=2^19 is L2-resident on B200).  A warp then owns 32
// samples = two m16 row tiles; the five weight matrices (24 KB fp16, padded so B-fragment loads are
// bank-conflict free) are staged once per CTA in shared memory, and the layers are chained in
// registers: the fp32 accumulator fragment of layer n is re-packed as the fp16 A fragment of layer
// n+1 (mma.sync m16n8k16, fp32 accumulate), so activations never touch shared or global memory.
// The density logit uses a hi+lo fp16 split of the hidden activations, which keeps sigma within
// ~1e-6 of the fp32 oracle.
#include <stdlib.h>

#include "field_common.cuh"

namespace qf {

// tcnn layout (row-major (out,in), fp32) -> padded fp16 image.  Head L1 input order in tcnn is
// [SH(16) | feat(15) | pad(1)]; the kernel feeds [SH(16) | pad | feat(15)] so that the base MLP's
// output fragment (sigma_raw, feat0..14) can be reused in place with sigma_raw replaced by the pad value.
__global__ void prep_weights_kernel(const float* __restrict__ base_w, const float* __restrict__ head_w,
                                    __half* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kWTotal) return;
  float v = 0.f;
  if (i < kW2) { int r = i / kS32, c = i % kS32; if (c < 32) v = base_w[r * 32 + c]; }
  else if (i < kW3) { int j = i - kW2, r = j / kS64, c = j % kS64; if (c < 64) v = base_w[64 * 32 + r * 64 + c]; }
  else if (i < kW4) {
    int j = i - kW3, r = j / kS32, c = j % kS32;
    if (c < 16) v = head_w[r * 32 + c];
    else if (c == 16) v = head_w[r * 32 + 31];
    else if (c < 32) v = head_w[r * 32 + c - 1];
  }
  else if (i < kW5) { int j = i - kW4, r = j / kS64, c = j % kS64; if (c < 64) v = head_w[64 * 32 + r * 64 + c]; }
  else { int j = i - kW5, r = j / kS64, c = j % kS64; if (c < 64) v = head_w[64 * 32 + 64 * 64 + r * 64 + c]; }
  out[i] = __float2half_rn(v);
}

__global__ void prep_table_kernel(const float* __restrict__ t, int64_t n_entries, __half2* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_entries; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __floats2half2_rn(t[2 * i], t[2 * i + 1]);
}

struct FieldArgs {
  qf_grid_desc desc;
  const __half2* table;
  const __half* weights;
  const float* pos;          // sample positions, `pos_stride` floats apart
  int pos_stride;
  const float* dirs;         // directions: per sample, or per ray when a ray index is given
  const int64_t* ray64;      // optional int64 ray index per sample
  const int32_t* ray32;      // optional int32 ray index per sample, `ray32_stride` ints apart
  int ray32_stride;
  int64_t M;                 // number of samples (used when d_M == NULL)
  const int32_t* d_M;        // optional device-side count
  float4* out4;              // (rgb, sigma) per sample            [mode 0, fused path]
  float* rgb;                // (M,3)                               [mode 0]
  float* density;            // (M,1)                               [mode 0/1]
  float* feat;               // (M,15)                              [mode 1]
};

template <int MODE>  // 0: full forward (rgb + sigma); 1: density + geo features only
__global__ void __launch_bounds__(128, 5) ngp_forward_kernel(const FieldArgs a) {
  __shared__ __align__(16) __half s_w[kWTotal];
  __shared__ __align__(16) __half s_tile[4][32 * kTileStride];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    constexpr int n16 = (MODE == 0 ? kWTotal : kW3) / 8;
    for (int i = tid; i < n16; i += 128) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int64_t M = a.d_M ? (int64_t)__ldg(a.d_M) : a.M;
  __half* tile = s_tile[warp];
  const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
  const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};

  for (int64_t base = ((int64_t)blockIdx.x * 4 + warp) * 32; base < M; base += (int64_t)gridDim.x * 128) {
    const int64_t i = base + lane;
    const bool valid = i < M;
    float x = 0.5f, y = 0.5f, z = 0.5f;
    bool sel = false;
    if (valid) {
      const float* p = a.pos + i * a.pos_stride;
      x = __fdiv_rn(__ldg(p) - amin[0], aext[0]);       // ngp.py:761-763
      y = __fdiv_rn(__ldg(p + 1) - amin[1], aext[1]);
      z = __fdiv_rn(__ldg(p + 2) - amin[2], aext[2]);
      sel = (x > 0.f) && (x < 1.f) && (y > 0.f) && (y < 1.f) && (z > 0.f) && (z < 1.f);
    }
    uint32_t* row32 = reinterpret_cast<uint32_t*>(tile + lane * kTileStride);
    encode_point(a.desc, a.table, x, y, z, [&](int l, uint32_t h2) { row32[l] = h2; });
    uint4* row = reinterpret_cast<uint4*>(row32);
    if (MODE == 0) {
      float dx = 0.f, dy = 0.f, dz = 1.f;
      if (valid) {
        int64_t r = a.ray64 ? __ldg(a.ray64 + i) : (a.ray32 ? (int64_t)__ldg(a.ray32 + i * a.ray32_stride) : i);
        const float* dp = a.dirs + 3 * r;
        // (d+1)/2 then tcnn maps back with *2-1 (ngp.py:784; tcnn SH kernel)
        dx = ((__ldg(dp) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dy = ((__ldg(dp + 1) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dz = ((__ldg(dp + 2) + 1.0f) / 2.0f) * 2.0f - 1.0f;
      }
      float sh[16];
      sh4(dx, dy, dz, sh);
      row[4] = make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sh[3]), pack_h2(sh[4], sh[5]), pack_h2(sh[6], sh[7]));
      row[5] = make_uint4(pack_h2(sh[8], sh[9]), pack_h2(sh[10], sh[11]), pack_h2(sh[12], sh[13]), pack_h2(sh[14], sh[15]));
    }
    const unsigned selmask = __ballot_sync(0xffffffffu, sel);
    __syncwarp();

#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      if (base + mt * 16 >= M) break;
      const __half* ta = tile + (mt * 16 + g) * kTileStride + t * 2;
      const __half* tb = ta + 8 * kTileStride;
      // ---- base layer 1: 32 -> 64, ReLU
      uint32_t a1[2][4];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        a1[k][0] = lds32(ta + k * 16); a1[k][1] = lds32(tb + k * 16);
        a1[k][2] = lds32(ta + k * 16 + 8); a1[k][3] = lds32(tb + k * 16 + 8);
      }
      float acc[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      layer<2, 8>(acc, a1, s_w + kW1, kS32, g, t);
      // ---- base layer 2: 64 -> 16 with hi+lo split of the hidden activations
      uint32_t ahi[4][4], alo[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* c = acc[2 * k + h];
          float r0 = fmaxf(c[0], 0.f), r1 = fmaxf(c[1], 0.f), r2 = fmaxf(c[2], 0.f), r3 = fmaxf(c[3], 0.f);
          __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
          float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          ahi[k][2 * h] = *reinterpret_cast<uint32_t*>(&h01);
          ahi[k][2 * h + 1] = *reinterpret_cast<uint32_t*>(&h23);
          alo[k][2 * h] = pack_h2(r0 - f01.x, r1 - f01.y);
          alo[k][2 * h + 1] = pack_h2(r2 - f23.x, r3 - f23.y);
        }
      }
      float acc2[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) acc2[n][0] = acc2[n][1] = acc2[n][2] = acc2[n][3] = 0.f;
      layer<4, 2>(acc2, alo, s_w + kW2, kS64, g, t);
      layer<4, 2>(acc2, ahi, s_w + kW2, kS64, g, t);
      // sigma = trunc_exp(h0 - 1) * selector   (ngp.py:772-775; _TruncExp forward is plain exp)
      const int64_t r_lo = base + mt * 16 + g, r_hi = r_lo + 8;
      const float s_lo = ((selmask >> (mt * 16 + g)) & 1u) ? expf(acc2[0][0] - 1.0f) : 0.f;
      const float s_hi = ((selmask >> (mt * 16 + g + 8)) & 1u) ? expf(acc2[0][2] - 1.0f) : 0.f;
      if (MODE == 1) {
        if (t == 0) {
          if (r_lo < M) a.density[r_lo] = s_lo;
          if (r_hi < M) a.density[r_hi] = s_hi;
        }
        if (a.feat) {
#pragma unroll
          for (int n = 0; n < 2; ++n) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              int col = n * 8 + t * 2 + q - 1;  // output column 0 is the density logit
              if (col >= 0) {
                if (r_lo < M) a.feat[r_lo * 15 + col] = acc2[n][q];
                if (r_hi < M) a.feat[r_hi * 15 + col] = acc2[n][2 + q];
              }
            }
          }
        }
        continue;
      }
      // ---- head layer 1: [SH(16) | pad, feat(15)] -> 64, ReLU
      uint32_t a3[2][4];
      a3[0][0] = lds32(ta + 32); a3[0][1] = lds32(tb + 32); a3[0][2] = lds32(ta + 40); a3[0][3] = lds32(tb + 40);
      a3[1][0] = pack_h2(t == 0 ? 1.0f : acc2[0][0], acc2[0][1]);
      a3[1][1] = pack_h2(t == 0 ? 1.0f : acc2[0][2], acc2[0][3]);
      a3[1][2] = pack_h2(acc2[1][0], acc2[1][1]);
      a3[1][3] = pack_h2(acc2[1][2], acc2[1][3]);
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      layer<2, 8>(acc, a3, s_w + kW3, kS32, g, t);
      // ---- head layer 2: 64 -> 64, ReLU
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* c = acc[2 * k + h];
          ahi[k][2 * h] = pack_h2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
          ahi[k][2 * h + 1] = pack_h2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
        }
      }
#pragma unroll
      for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
      layer<4, 8>(acc, ahi, s_w + kW4, kS64, g, t);
      // ---- head layer 3: 64 -> 3 (first n-tile only), sigmoid
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float* c = acc[2 * k + h];
          ahi[k][2 * h] = pack_h2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
          ahi[k][2 * h + 1] = pack_h2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
        }
      }
      float acc5[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      layer<4, 1>(acc5, ahi, s_w + kW5, kS64, g, t);
      const float b_lo = __shfl_down_sync(0xffffffffu, acc5[0][0], 1);
      const float b_hi = __shfl_down_sync(0xffffffffu, acc5[0][2], 1);
      if (t == 0) {
        auto sig = [](float v) { return 1.0f / (1.0f + __expf(-v)); };
        float4 o_lo = make_float4(sig(acc5[0][0]), sig(acc5[0][1]), sig(b_lo), s_lo);
        float4 o_hi = make_float4(sig(acc5[0][2]), sig(acc5[0][3]), sig(b_hi), s_hi);
        if (a.out4) {
          if (r_lo < M) a.out4[r_lo] = o_lo;
          if (r_hi < M) a.out4[r_hi] = o_hi;
        } else {
          if (r_lo < M) { a.rgb[3 * r_lo] = o_lo.x; a.rgb[3 * r_lo + 1] = o_lo.y; a.rgb[3 * r_lo + 2] = o_lo.z; a.density[r_lo] = o_lo.w; }
          if (r_hi < M) { a.rgb[3 * r_hi] = o_hi.x; a.rgb[3 * r_hi + 1] = o_hi.y; a.rgb[3 * r_hi + 2] = o_hi.z; a.density[r_hi] = o_hi.w; }
        }
      }
    }
    __syncwarp();
  }
}

__global__ void hashgrid_forward_kernel(const qf_grid_desc desc, const __half2* __restrict__ table,
                                        const float* __restrict__ x01, int64_t M, float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int L = desc.n_levels;
  float* o = out + i * 2 * L;
  encode_point(desc, table, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], [&](int l, uint32_t h2) {
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&h2));
    o[2 * l] = f.x;
    o[2 * l + 1] = f.y;
  });
}

// field_tc.cu
struct FieldTcArgs {
  qf_grid_desc desc; const __half2* table; const unsigned char* weights_tc; const float* pos; int pos_stride; const float* dirs;
  const int64_t* ray64; const int32_t* ray32; int ray32_stride; int64_t M; const int32_t* d_M; float4* out4; float* rgb; float* density;
};
int launch_ngp_forward_tc(const qf_ngp* f, FieldTcArgs& a, cudaStream_t st);
int prep_weights_tc(qf_ngp* f, cudaStream_t st);

// QF_FIELD_TC=0 selects the mma.sync kernel, 1 the tcgen05/TMEM kernel for the full forward
static int field_tc_mode() {
  static const int m = getenv("QF_FIELD_TC") ? atoi(getenv("QF_FIELD_TC")) : 0;
  return m;
}

// diagnostic (tools/diag_shade.py): the gather alone — every level encoded, 4 bytes written per sample
__global__ void __launch_bounds__(128) encode_sum_kernel(const qf_grid_desc desc, const __half2* __restrict__ table,
                                                         const float* __restrict__ x01, int64_t M, float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float acc = 0.f;
  encode_point(desc, table, x01[3 * i], x01[3 * i + 1], x01[3 * i + 2], [&](int l, uint32_t h2) {
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&h2));
    acc += f.x + f.y;
  });
  out[i] = acc;
}

int launch_ngp_forward(const qf_ngp* f, FieldArgs& a, int mode, cudaStream_t st) {
  if (mode == 0 && field_tc_mode() == 1) {
    FieldTcArgs t = {};
    t.pos = a.pos; t.pos_stride = a.pos_stride; t.dirs = a.dirs; t.ray64 = a.ray64; t.ray32 = a.ray32; t.ray32_stride = a.ray32_stride;
    t.M = a.M; t.d_M = a.d_M; t.out4 = a.out4; t.rgb = a.rgb; t.density = a.density;
    return launch_ngp_forward_tc(f, t, st);
  }
  a.desc = f->desc;
  a.table = f->d_table;
  a.weights = f->d_weights;
  static const int per_sm = getenv("QF_SHADE_BLOCKS_PER_SM") ? atoi(getenv("QF_SHADE_BLOCKS_PER_SM")) : 5;   // one resident wave: 5 CTAs/SM at 96 registers
  int64_t tiles = a.d_M ? (int64_t)kNumSMs * per_sm : ceil_div(a.M, 128);
  int blocks = (int)(tiles < (int64_t)kNumSMs * per_sm ? tiles : (int64_t)kNumSMs * per_sm);
  if (blocks < 1) blocks = 1;
  if (mode == 0) ngp_forward_kernel<0><<<blocks, 128, 0, st>>>(a);
  else ngp_forward_kernel<1><<<blocks, 128, 0, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

// shading entry for the fused render (render.cu): compact hit records in, (rgb, sigma) out
int launch_ngp_forward_hits(const qf_ngp* f, const float4* hit_pd, const int2* hit_rt, const float* d_viewdirs,
                            const int32_t* d_M, float4* out4, cudaStream_t st) {
  FieldArgs a = {};
  a.pos = reinterpret_cast<const float*>(hit_pd); a.pos_stride = 4;
  a.dirs = d_viewdirs;                               // original unit viewdirs gathered by ray id (utils.py:517, quirk Q7)
  a.ray32 = reinterpret_cast<const int32_t*>(hit_rt); a.ray32_stride = 2;
  a.d_M = d_M; a.out4 = out4;
  return launch_ngp_forward(f, a, 0, st);
}

static int upload(qf_ngp* f, const float* d_table, const float* d_base_w, const float* d_head_w, cudaStream_t st) {
  if (d_table) prep_table_kernel<<<kNumSMs * 8, 256, 0, st>>>(d_table, f->n_entries, f->d_table);
  if (d_base_w && d_head_w) {
    prep_weights_kernel<<<(int)ceil_div(kWTotal, 256), 256, 0, st>>>(d_base_w, d_head_w, f->d_weights);
    QF_LAUNCH_CHECK();
    return prep_weights_tc(f, st);
  }
  QF_LAUNCH_CHECK();
  return QF_OK;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_ngp_create(const qf_grid_desc* desc, const float* d_table, int64_t n_entries, const float* d_base_w,
                             const float* d_head_w, void* stream, qf_ngp** out) {
  QF_REQUIRE(desc && d_table && d_base_w && d_head_w && out, "qf_ngp_create: NULL argument");
  QF_REQUIRE(desc->n_levels == 16, "qf_ngp_create: n_levels=%d; the fused MLP takes the 32-wide encoding (16 levels x 2)",
             desc->n_levels);
  int64_t need = (int64_t)desc->offset[desc->n_levels - 1] + desc->size[desc->n_levels - 1];
  QF_REQUIRE(n_entries >= need, "qf_ngp_create: table has %lld entries, level table needs %lld", (long long)n_entries,
             (long long)need);
  qf_ngp* f = new qf_ngp();
  f->desc = *desc;
  f->n_entries = n_entries;
  if (cudaMalloc((void**)&f->d_table, sizeof(__half2) * (size_t)n_entries) != cudaSuccess ||
      cudaMalloc((void**)&f->d_weights, sizeof(__half) * kWTotal) != cudaSuccess ||
      cudaMalloc((void**)&f->d_weights_tc, 20480) != cudaSuccess) {
    set_error("qf_ngp_create: device allocation failed");
    qf_ngp_destroy(f);
    return QF_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = upload(f, d_table, d_base_w, d_head_w, st);
  if (rc == QF_OK && cudaStreamSynchronize(st) != cudaSuccess) { set_error("qf_ngp_create: upload failed"); rc = QF_ERR_CUDA; }
  if (rc != QF_OK) { qf_ngp_destroy(f); return rc; }
  *out = f;
  return QF_OK;
}

extern "C" int qf_ngp_update(qf_ngp* f, const float* d_table, const float* d_base_w, const float* d_head_w, void* stream) {
  QF_REQUIRE(f, "qf_ngp_update: NULL field");
  return upload(f, d_table, d_base_w, d_head_w, (cudaStream_t)stream);
}

extern "C" void qf_ngp_destroy(qf_ngp* f) {
  if (!f) return;
  if (f->d_table) cudaFree(f->d_table);
  if (f->d_weights) cudaFree(f->d_weights);
  if (f->d_weights_tc) cudaFree(f->d_weights_tc);
  delete f;
}

extern "C" int qf_debug_encode_sum(const qf_ngp* f, const float* d_x01, int64_t M, float* d_out, void* stream) {
  if (M == 0) return QF_OK;
  encode_sum_kernel<<<(int)ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(f->desc, f->d_table, d_x01, M, d_out);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_hashgrid_forward(const qf_ngp* f, const float* d_x01, int64_t M, float* d_enc, void* stream) {
  QF_REQUIRE(f && d_x01 && d_enc, "qf_hashgrid_forward: NULL argument");
  if (M == 0) return QF_OK;
  hashgrid_forward_kernel<<<(int)ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(f->desc, f->d_table, d_x01, M, d_enc);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_ngp_query_density(const qf_ngp* f, const float* d_positions, int64_t M, float* d_density, float* d_feat,
                                    void* stream) {
  QF_REQUIRE(f && d_positions && d_density, "qf_ngp_query_density: NULL argument");
  if (M == 0) return QF_OK;
  FieldArgs a = {};
  a.pos = d_positions; a.pos_stride = 3; a.M = M; a.density = d_density; a.feat = d_feat;
  return launch_ngp_forward(f, a, 1, (cudaStream_t)stream);
}

extern "C" int qf_ngp_forward(const qf_ngp* f, const float* d_positions, const float* d_directions,
                              const int64_t* d_ray_index, int64_t M, float* d_rgb, float* d_density, void* stream) {
  QF_REQUIRE(f && d_positions && d_directions && d_rgb && d_density, "qf_ngp_forward: NULL argument");
  if (M == 0) return QF_OK;
  FieldArgs a = {};
  a.pos = d_positions; a.pos_stride = 3; a.dirs = d_directions; a.ray64 = d_ray_index; a.M = M;
  a.rgb = d_rgb; a.density = d_density;
  return launch_ngp_forward(f, a, 0, (cudaStream_t)stream);
}
