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

// producer half of one 32-sample tile (one warp, lane = sample): normalise, encode the 16 levels and (MODE 0) the SH
// basis into the lane's row of `tile`; returns the warp's selector ballot
template <int MODE>
__device__ __forceinline__ unsigned produce_tile(const FieldArgs& a, const float* amin, const float* aext, int64_t base,
                                                 int64_t M, int lane, __half* tile) {
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
  return __ballot_sync(0xffffffffu, sel);
}

// consumer half: the two 16-row m-tiles of `tile` through the MLPs (mma.sync, fragments chained in registers)
template <int MODE>
__device__ __forceinline__ void consume_tile(const FieldArgs& a, const __half* s_w, const __half* tile, unsigned selmask,
                                             int64_t base, int64_t M, int g, int t) {
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
}

// WARPS warps per CTA share one weight image; CTAs per SM are chosen so that 20 warps are resident (96 registers each).
// Fewer, larger CTAs replicate the 24 KB weight image less often, which leaves more of the 256 KB L1/shared array as
// L1 for the table gathers.
template <int MODE, int WARPS>  // MODE 0: full forward (rgb + sigma); 1: density + geo features only
__global__ void __launch_bounds__(WARPS * 32, 20 / WARPS) ngp_forward_kernel(const FieldArgs a) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  __half* s_w = reinterpret_cast<__half*>(fused_smem);
  __half* s_tiles = s_w + kWTotal;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    constexpr int n16 = (MODE == 0 ? kWTotal : kW3) / 8;
    for (int i = tid; i < n16; i += WARPS * 32) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int64_t M = a.d_M ? (int64_t)__ldg(a.d_M) : a.M;
  __half* tile = s_tiles + warp * (32 * kTileStride);
  const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
  const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};

  for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * 32; base < M; base += (int64_t)gridDim.x * (WARPS * 32)) {
    const unsigned selmask = produce_tile<MODE>(a, amin, aext, base, M, lane, tile);
    __syncwarp();
    consume_tile<MODE>(a, s_w, tile, selmask, base, M, g, t);
    __syncwarp();
  }
}

// ---- warp-specialised variant of the full forward (QF_FIELD_TC=2) ------------------------------------------------------
// One persistent CTA per SM: PROD producer warps only gather (48 registers each after setmaxnreg.dec) and CONS consumer
// warps only run the MLPs (setmaxnreg.inc).  Producer p owns two tile slots in shared memory; a full/empty mbarrier
// pair per slot hands tiles over.  Consumer c serves producers p = c (mod CONS).  Same arithmetic as the fused kernel
// (produce_tile / consume_tile), so results are bit-identical.  Measured (16 producers + 4 consumers): c2 0.255 ms,
// c4 3.54 ms against 0.251 / 3.87 ms for the fused kernel: the gather is bound by L1TEX wavefronts, not by the MLP
// phases sharing its warps, so the split only pays where the larger L1 share (less shared memory) does.
constexpr int kWsSlots = 2;
constexpr int kWsTileHalfs = 32 * kTileStride;
constexpr int kWsProdRegs = 48;
template <int PROD, int CONS>
struct WsCfg {
  static constexpr int kThreads = (PROD + CONS) * 32;
  // ptxas allocates the launch_bounds maximum per thread; setmaxnreg.inc blocks forever if the CTA's pool
  // (threads x launch registers) cannot cover producers x 48 + consumers x kConsRegs
  static constexpr int kLaunchRegs = 65536 / kThreads / 8 * 8;
  static constexpr int kConsRegsRaw = (kThreads * kLaunchRegs - PROD * 32 * kWsProdRegs) / (CONS * 32) / 8 * 8;
  static constexpr int kConsRegs = kConsRegsRaw > 200 ? 200 : kConsRegsRaw;
  static constexpr int kSmTiles = kWTotal * 2;
  static constexpr int kSmSel = kSmTiles + PROD * kWsSlots * kWsTileHalfs * 2;
  static constexpr int kSmBar = (kSmSel + PROD * kWsSlots * 4 + 7) / 8 * 8;
  static constexpr int kSmemBytes = kSmBar + 2 * PROD * kWsSlots * 8;
  static_assert(PROD % 4 == 0 && CONS % 4 == 0 && PROD % CONS == 0 && kThreads <= 1024, "warpgroup-aligned roles");
  static_assert(kSmTiles % 16 == 0 && kSmemBytes <= 227 * 1024 && kConsRegs >= 96, "warp-specialised smem / register budget");
};

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void ws_mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(mbar) : "memory");
}

template <int PROD, int CONS>
__global__ void __launch_bounds__((PROD + CONS) * 32, 1) ngp_forward_ws_kernel(const FieldArgs a) {
  using Cfg = WsCfg<PROD, CONS>;
  extern __shared__ __align__(16) unsigned char ws_smem[];
  __half* s_w = reinterpret_cast<__half*>(ws_smem);
  __half* s_tiles = reinterpret_cast<__half*>(ws_smem + Cfg::kSmTiles);
  uint32_t* s_sel = reinterpret_cast<uint32_t*>(ws_smem + Cfg::kSmSel);
  const uint32_t bar_full = ws_smem_u32(ws_smem + Cfg::kSmBar), bar_empty = bar_full + PROD * kWsSlots * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < kWTotal / 8; i += Cfg::kThreads) dst[i] = __ldg(src + i);
    if (tid < 2 * PROD * kWsSlots)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar_full + tid * 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const int64_t M = a.d_M ? (int64_t)__ldg(a.d_M) : a.M;
  const int64_t n_tiles = (M + 31) >> 5;
  // round `it` of this CTA takes PROD consecutive tiles, one per producer warp
  const int64_t round_stride = (int64_t)gridDim.x * PROD;

  if (warp < PROD) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" :: "n"(kWsProdRegs));
    const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
    const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};
    int it = 0;
    for (int64_t tile_id = (int64_t)blockIdx.x * PROD + warp; tile_id < n_tiles; tile_id += round_stride, ++it) {
      const int idx = warp * kWsSlots + (it & 1), use = it >> 1;
      if (use > 0) ws_mbar_wait(bar_empty + idx * 8, (use - 1) & 1);
      const unsigned selmask = produce_tile<0>(a, amin, aext, tile_id * 32, M, lane, s_tiles + idx * kWsTileHalfs);
      if (lane == 0) s_sel[idx] = selmask;
      __syncwarp();
      if (lane == 0) ws_mbar_arrive(bar_full + idx * 8);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" :: "n"(Cfg::kConsRegs));
    const int c = warp - PROD, g = lane >> 2, t = lane & 3;
    for (int it = 0;; ++it) {
      const int64_t round_base = (int64_t)blockIdx.x * PROD + it * round_stride;
      if (round_base + c >= n_tiles) break;
      const int use = it >> 1;
#pragma unroll 1
      for (int p = c; p < PROD; p += CONS) {
        const int64_t tile_id = round_base + p;
        if (tile_id >= n_tiles) break;
        const int idx = p * kWsSlots + (it & 1);
        ws_mbar_wait(bar_full + idx * 8, use & 1);
        const unsigned selmask = s_sel[idx];
        consume_tile<0>(a, s_w, s_tiles + idx * kWsTileHalfs, selmask, tile_id * 32, M, g, t);
        __syncwarp();
        if (lane == 0) ws_mbar_arrive(bar_empty + idx * 8);
      }
    }
  }
}

template <int PROD, int CONS>
static int launch_ws(FieldArgs& a, cudaStream_t st) {
  using Cfg = WsCfg<PROD, CONS>;
  QF_ENSURE_DYNAMIC_SMEM((ngp_forward_ws_kernel<PROD, CONS>), Cfg::kSmemBytes);
  int64_t rounds = a.d_M ? kNumSMs : ceil_div(a.M, 32 * PROD);
  int blocks = (int)(rounds < kNumSMs ? (rounds < 1 ? 1 : rounds) : kNumSMs);
  ngp_forward_ws_kernel<PROD, CONS><<<blocks, Cfg::kThreads, Cfg::kSmemBytes, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
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

// The full forward runs on the warp-specialised tcgen05 / TMEM kernel (field_tc.cu) by default.  QF_FIELD_TC=0 selects the
// fused mma.sync kernel (kept as the comparison baseline and for the density-only mode), 2 the warp-specialised mma.sync one.
static int field_tc_mode() {
  static const int m = getenv("QF_FIELD_TC") ? atoi(getenv("QF_FIELD_TC")) : 1;
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

// diagnostic (tools/diag_gather_occ.py): the same gather as a persistent grid of `blocks_per_sm` x 148 CTAs — occupancy /
// CTA-shape / shared-memory carve-out sweeps
__global__ void __launch_bounds__(1024) encode_sum_persist_kernel(const qf_grid_desc desc, const __half2* __restrict__ table,
                                                                  const float* __restrict__ x01, int64_t M, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ((M + 31) & ~31ll); i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i < M ? i : M - 1;
    float acc = 0.f;
    encode_point(desc, table, x01[3 * j], x01[3 * j + 1], x01[3 * j + 2], [&](int l, uint32_t h2) {
      float2 f = __half22float2(*reinterpret_cast<__half2*>(&h2));
      acc += f.x + f.y;
    });
    if (i < M) out[i] = acc;
  }
}

template <int WARPS>
static int launch_fused(FieldArgs& a, int mode, cudaStream_t st) {
  constexpr int smem = kWTotal * 2 + WARPS * 32 * kTileStride * 2, per_sm = 20 / WARPS;   // one resident wave
  QF_ENSURE_DYNAMIC_SMEM((ngp_forward_kernel<0, WARPS>), smem);
  QF_ENSURE_DYNAMIC_SMEM((ngp_forward_kernel<1, WARPS>), smem);
  int64_t tiles = a.d_M ? (int64_t)kNumSMs * per_sm : ceil_div(a.M, WARPS * 32);
  int blocks = (int)(tiles < (int64_t)kNumSMs * per_sm ? tiles : (int64_t)kNumSMs * per_sm);
  if (blocks < 1) blocks = 1;
  if (mode == 0) ngp_forward_kernel<0, WARPS><<<blocks, WARPS * 32, smem, st>>>(a);
  else ngp_forward_kernel<1, WARPS><<<blocks, WARPS * 32, smem, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

int launch_ngp_forward(const qf_ngp* f, FieldArgs& a, int mode, cudaStream_t st) {
  QF_REQUIRE(f->d_weights, "this field handle holds a grid only (qf_grid_create); the NGP MLPs need qf_ngp_create");
  if (mode == 0 && field_tc_mode() == 1) {
    FieldTcArgs t = {};
    t.pos = a.pos; t.pos_stride = a.pos_stride; t.dirs = a.dirs; t.ray64 = a.ray64; t.ray32 = a.ray32; t.ray32_stride = a.ray32_stride;
    t.M = a.M; t.d_M = a.d_M; t.out4 = a.out4; t.rgb = a.rgb; t.density = a.density;
    return launch_ngp_forward_tc(f, t, st);
  }
  a.desc = f->desc;
  a.table = f->d_table;
  a.weights = f->d_weights;
  if (mode == 0 && field_tc_mode() == 2) return launch_ws<16, 4>(a, st);
  // 10 warps per CTA, 2 CTAs per SM measured best (c2 0.251 ms, c4 3.87 ms; 4 warps x 5 CTAs: 0.252 / 4.67 ms)
  static const int warps = getenv("QF_SHADE_WARPS") ? atoi(getenv("QF_SHADE_WARPS")) : 10;
  switch (warps) {
    case 4: return launch_fused<4>(a, mode, st);
    case 20: return launch_fused<20>(a, mode, st);
    default: return launch_fused<10>(a, mode, st);
  }
}

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
  const int bps = getenv("QF_DEBUG_ENC_BLOCKS_PER_SM") ? atoi(getenv("QF_DEBUG_ENC_BLOCKS_PER_SM")) : 0;
  if (bps > 0) {
    const int thr = getenv("QF_DEBUG_ENC_THREADS") ? atoi(getenv("QF_DEBUG_ENC_THREADS")) : 128;
    const int smem = getenv("QF_DEBUG_ENC_SMEM") ? atoi(getenv("QF_DEBUG_ENC_SMEM")) : 0;
    encode_sum_persist_kernel<<<kNumSMs * bps, thr, smem, (cudaStream_t)stream>>>(f->desc, f->d_table, d_x01, M, d_out);
    QF_LAUNCH_CHECK();
    return QF_OK;
  }
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
