// (2)+(3) Blackwell-native fused field kernel — the DEFAULT full forward (radiance_fields/ngp.py:757-809):
// warp-specialised CTA, hash-grid gathers on producer warps, the five MLP layers on the 5th-generation tensor cores
// (tcgen05.mma, SASS UTCHMMA) with every activation living in tensor memory.
//
//   CTA = Q gather "quads" (4 warps each) + G MLP warpgroups; CPS CTAs per SM, persistent grid of CPS x 148.
//   Default Q=3, G=1, CPS=2 (512 threads, 256 TMEM columns per CTA); the A/B of r2 is in DESIGN §4a.
//
//   GATHER quads: a quad owns one 128-sample tile at a time; lane = sample, warp w of the quad = rows 32w..32w+31 = TMEM
//                 lanes 32w..  (a warp can only touch the TMEM lane quarter warp_id % 4, which is why roles are
//                 quad-aligned).  Each lane gathers its 16 levels x 8 corners from the table and writes the 32 encoded
//                 features (16 packed columns) and the 16 SH values (8 columns) of its sample STRAIGHT INTO TENSOR MEMORY
//                 with tcgen05.st — the layer-1 / layer-3 A operands never exist in shared memory.  Two A1 slots and one
//                 SH slot per quad, full/empty mbarriers.
//   MLP warpgroup: thread r <-> row r of the tile <-> TMEM lane r.  Its first thread issues the tcgen05.mma chain (A from
//                 TMEM, B = weight images in shared memory through UMMA descriptors, fp32 accumulators in TMEM) and
//                 commits to an mbarrier; all 128 threads then read their accumulator row back with tcgen05.ld (LDTM),
//                 apply ReLU / exp / sigmoid in fp32 and write the next layer's A row back to TMEM (tcgen05.st): no LDS /
//                 STS at all in the steady state, no CTA-wide barrier — a 128-thread named barrier orders the warpgroup's
//                 TMEM writes before its leader issues the next layer.  With G > 1 the warpgroups take alternate tiles.
//
// Why: the mma.sync kernel (field.cu) is bound by the L1TEX data pipe (67 % busy, profiles/r1h): 204 K wavefronts per SM
// of table gathers plus 93 K wavefronts of shared-memory fragment loads / tile stores for the MLPs, issued from the same
// warps that gather.  Here the tensor core reads A from TMEM and only the 20 KB weight image from shared memory, the
// gather warps never run MLP code, and the 20 KB of shared memory per CTA leave ~200 KB of the SM's array as L1.
//
// TMEM columns per CTA (32-bit cells, lane = tile row); per MLP warpgroup 96 columns:
//   [  0, 64)  R   D1 -> (in place) A2 hi|lo interleaved per 16-wide k-chunk -> D3 -> D4 -> D5 (rgb logits, 16 cols)
//   [ 64, 96)  X   D2 (density logit + 15 geo features) in [80,96) -> [1, feat] chunk of A3 in [64,72) -> A4 -> A5
//                  (packed fp16 pairs, 32 cols = K 64)
// then per gather quad 40 columns: A1 slot 0 | A1 slot 1 (16 cols each = 32 encoded features) | SH (8 cols = 16 values)
//
// Numerics are those of field.cu: fp16 operands, fp32 accumulation, hi+lo split of the hidden activations for the
// density logit (here hi = the activation truncated to fp16's mantissa, lo = the exact remainder); the
// encoding is bit-identical (same encode_point).
#include "field_common.cuh"

namespace qf {

// ---- operand images: K-major canonical layout, byte offset of element (row r, column k) of an (R x K) fp16 matrix
__host__ __device__ constexpr int canon_off_bytes(int r, int k, int K) {
  return (r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
}
// weight image offsets (bytes)
constexpr int kTcW1 = 0;                       // 64 x 32
constexpr int kTcW2 = kTcW1 + 64 * 32 * 2;     // 16 x 64
constexpr int kTcW3 = kTcW2 + 16 * 64 * 2;     // 64 x 32  (kernel column order [SH | pad | feat])
constexpr int kTcW4 = kTcW3 + 64 * 32 * 2;     // 64 x 64
constexpr int kTcW5 = kTcW4 + 64 * 64 * 2;     // 16 x 64
constexpr int kTcWBytes = kTcW5 + 16 * 64 * 2; // 20480

__global__ void prep_weights_tc_kernel(const __half* __restrict__ img, unsigned char* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per weight element
  const int n1 = 64 * 32, n2 = 16 * 64, n3 = 64 * 32, n4 = 64 * 64, n5 = 16 * 64;
  int off, r, k, K, base;
  const __half* src;
  if (i < n1) { r = i / 32; k = i % 32; K = 32; base = kTcW1; src = img + kW1 + r * kS32 + k; }
  else if ((i -= n1) < n2) { r = i / 64; k = i % 64; K = 64; base = kTcW2; src = img + kW2 + r * kS64 + k; }
  else if ((i -= n2) < n3) { r = i / 32; k = i % 32; K = 32; base = kTcW3; src = img + kW3 + r * kS32 + k; }
  else if ((i -= n3) < n4) { r = i / 64; k = i % 64; K = 64; base = kTcW4; src = img + kW4 + r * kS64 + k; }
  else if ((i -= n4) < n5) { r = i / 64; k = i % 64; K = 64; base = kTcW5; src = img + kW5 + r * kS64 + k; }
  else return;
  off = base + canon_off_bytes(r, k, K);
  *reinterpret_cast<__half*>(out + off) = *src;
}

// ---- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
  // layout_type [61,64) = 0 (no swizzle)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// InstrDescriptor for kind::f16: D=F32 (bits 4-5 = 1), A=B=F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A is read from tensor memory (row i = lane i, two K-consecutive halves per cell)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (!done) {
    // the suspend-time hint lets the hardware park the thread until the phase flips (or the hint runs out) instead of
    // returning at once: without it the waits of r2h polled 4.8 M times per launch — 5 % of all instructions and an extra
    // shared-memory wavefront each, in a kernel bound by the L1TEX wavefront pipe
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity), "r"(0x989680u) : "memory");
    if (!done && ++spins > (1u << 22)) __trap();   // never hang the GPU on a protocol error
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(threads) : "memory"); }

// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};\n" :: "r"(taddr), "r"(r0), "r"(r1) : "memory");
}

struct FieldTcArgs {
  qf_grid_desc desc;
  const __half2* table;
  const unsigned char* weights_tc;  // kTcWBytes canonical images
  const float* pos;
  int pos_stride;
  const float* dirs;
  const int64_t* ray64;
  const int32_t* ray32;
  int ray32_stride;
  int64_t M;
  const int32_t* d_M;
  float4* out4;
  float* rgb;
  float* density;
};

// ---- roles and resources
// Q gather quads (4 warps each) + G MLP warpgroups per CTA, CPS CTAs per SM.  TMEM: G x 96 pipeline columns, then Q x 40.
#ifndef QF_TC_QUADS
#define QF_TC_QUADS 3
#endif
#ifndef QF_TC_GROUPS
#define QF_TC_GROUPS 1
#endif
#ifndef QF_TC_CTAS_PER_SM
#define QF_TC_CTAS_PER_SM 2
#endif
template <int Q, int G, int CPS>
struct TcCfg {
  static constexpr int kGatherWarps = 4 * Q, kThreads = (4 * Q + 4 * G) * 32;
  static constexpr int kPipeCols = 96, kQuadCols = 40, kColSlot0 = G * kPipeCols;
  static constexpr int kColsNeeded = kColSlot0 + Q * kQuadCols;
  static constexpr int kTmemCols = kColsNeeded <= 128 ? 128 : (kColsNeeded <= 256 ? 256 : 512);
  // shared memory map (bytes)
  static constexpr int kSmW = 0;                                   // weights, 20480
  static constexpr int kSmFull = kSmW + kTcWBytes;                 // full[quad][slot] mbarriers (count 4: one per gather warp)
  static constexpr int kSmEmpty = kSmFull + Q * 2 * 8;             // empty[quad] mbarriers (count 1: tcgen05.commit of layer 3)
  static constexpr int kSmMma = kSmEmpty + Q * 8;                  // one MLP chain mbarrier per warpgroup
  static constexpr int kSmTmem = kSmMma + G * 8;                   // TMEM base address
  static constexpr int kSmSel = kSmTmem + 8;                       // selector ballots [quad][slot][4 warps]
  static constexpr int kSmemBytes = kSmSel + Q * 2 * 4 * 4;
  static_assert(kColsNeeded <= 512 && kTmemCols * CPS <= 512, "TMEM column budget");
  static_assert(kThreads <= 1024 && kThreads * CPS <= 2048, "thread budget");
};
// pipeline-relative TMEM columns
constexpr int kColR = 0, kColX = 64, kColS = 80;

// two floats -> packed halves with ReLU folded into the conversion (lo -> low half)
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int Q, int G, int CPS>
__global__ void __launch_bounds__(TcCfg<Q, G, CPS>::kThreads, CPS) ngp_forward_tc_kernel(const FieldTcArgs a) {
  using Cfg = TcCfg<Q, G, CPS>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::kSmTmem);
  uint32_t* s_sel = reinterpret_cast<uint32_t*>(smem + Cfg::kSmSel);
  const uint32_t bar_full = smem_u32(smem + Cfg::kSmFull), bar_empty = smem_u32(smem + Cfg::kSmEmpty);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights_tc);
    uint4* dst = reinterpret_cast<uint4*>(smem + Cfg::kSmW);
    for (int i = tid; i < kTcWBytes / 16; i += Cfg::kThreads) dst[i] = __ldg(src + i);
  }
  if (tid < Q * 2) mbar_init(bar_full + tid * 8, 4);
  if (tid < Q) mbar_init(bar_empty + tid * 8, 1);
  if (tid < G) mbar_init(smem_u32(smem + Cfg::kSmMma) + tid * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "n"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_async_smem();          // weights written through the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t M = a.d_M ? (int64_t)__ldg(a.d_M) : a.M;
  const int64_t n_tiles = (M + 127) >> 7;
  // the CTA's n-th tile is tile (blockIdx.x + n * gridDim.x); quad (n % Q) gathers it — its k-th tile, k = n / Q, into A1
  // slot (k % 2) — and MLP warpgroup (n % G) consumes it
  const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
  auto a1_col = [](int quad, int slot) { return (uint32_t)(Cfg::kColSlot0 + quad * Cfg::kQuadCols + slot * 16); };
  auto sh_col = [](int quad) { return (uint32_t)(Cfg::kColSlot0 + quad * Cfg::kQuadCols + 32); };

  if (warp < Cfg::kGatherWarps) {
    // ===================== GATHER =====================
    const int quad = warp >> 2, sub = warp & 3;
    const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
    const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};
    int k = 0;   // this quad's k-th tile
    for (int64_t tile = blockIdx.x + (int64_t)quad * gridDim.x; tile < n_tiles; tile += (int64_t)Q * gridDim.x, ++k) {
      const int slot = k & 1, idx = quad * 2 + slot;
      // A1 slot `slot` was last read by layer 1 of this quad's tile k-2: complete, because before writing tile k-1's SH
      // this warp waited for layer 3 of tile k-2 (below)
      const uint32_t t_a1 = tmem_base + lane_bits + a1_col(quad, slot);
      const int64_t i = tile * 128 + sub * 32 + lane;
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
      float dx = 0.f, dy = 0.f, dz = 1.f;
      if (valid) {
        int64_t r = a.ray64 ? __ldg(a.ray64 + i) : (a.ray32 ? (int64_t)__ldg(a.ray32 + i * a.ray32_stride) : i);
        const float* dp = a.dirs + 3 * r;
        // (d+1)/2 then tcnn maps back with *2-1 (ngp.py:784; tcnn SH kernel)
        dx = ((__ldg(dp) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dy = ((__ldg(dp + 1) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dz = ((__ldg(dp + 2) + 1.0f) / 2.0f) * 2.0f - 1.0f;
      }
      // 16 levels, two per trip: the pair's gathers are in flight together; each trip writes 2 TMEM columns of the row
      {
        uint32_t even = 0;
        encode_point(a.desc, a.table, x, y, z, [&](int l, uint32_t h2) {
          if (l & 1) tmem_st2(t_a1 + (l - 1), even, h2);
          else even = h2;
        });
      }
      {
        float sh[16];
        sh4(dx, dy, dz, sh);
        uint32_t shp[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) shp[q] = pack_h2(sh[2 * q], sh[2 * q + 1]);
        if (k > 0) {
          mbar_wait(bar_empty + quad * 8, (k - 1) & 1);     // layer 3 of this quad's previous tile has read the SH columns
          tc_fence_after();
        }
        tmem_st8(tmem_base + lane_bits + sh_col(quad), shp);
      }
      const unsigned selmask = __ballot_sync(0xffffffffu, sel);
      if (lane == 0) s_sel[idx * 4 + sub] = selmask;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + idx * 8);
    }
  } else {
    // ===================== MLP warpgroups =====================
    const int group = (warp - Cfg::kGatherWarps) >> 2;
    const int row = (tid - Cfg::kGatherWarps * 32) & 127;          // 0..127 = TMEM lane
    const bool leader = row == 0;
    const uint32_t pipe = tmem_base + group * Cfg::kPipeCols;      // this warpgroup's R | X columns
    const uint32_t t_row = pipe + lane_bits;                       // this warp's lane quarter
    const uint32_t bar_mma = smem_u32(smem + Cfg::kSmMma) + group * 8;
    const int bar_id = 1 + group;
    const uint32_t sW = smem_u32(smem + Cfg::kSmW);
    const uint32_t id64 = umma_idesc(128, 64), id16 = umma_idesc(128, 16);
    // B descriptors: chunk j (16 k-values = two 8x8 core matrices) of a (N x K) K-major canonical image
    auto bdesc = [&](int w_off, int K, int j) { return umma_desc(sW + w_off + j * 256, 128, (K >> 3) * 128); };
    uint32_t phase = 0;
    for (int64_t n = group; ; n += G) {
      const int64_t tile = blockIdx.x + n * gridDim.x;
      if (tile >= n_tiles) break;
      const int quad = (int)(n % Q), kq = (int)(n / Q), slot = kq & 1, idx = quad * 2 + slot;
      mbar_wait(bar_full + idx * 8, (kq >> 1) & 1);
      tc_fence_after();
      const bool sel = (s_sel[idx * 4 + (row >> 5)] >> (row & 31)) & 1u;
      const uint32_t c_a1 = tmem_base + a1_col(quad, slot), c_sh = tmem_base + sh_col(quad);
      // ---- base L1: R = A1 (K=32) * W1^T
      if (leader) {
        umma_f16_ts(pipe + kColR, c_a1, bdesc(kTcW1, 32, 0), id64, 0u);
        umma_f16_ts(pipe + kColR, c_a1 + 8, bdesc(kTcW1, 32, 1), id64, 1u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      {
        // 64 hidden units: ReLU and hi | lo split, written in place over the 16 accumulator columns of each k-chunk.
        // hi = the value truncated to fp16's 10 mantissa bits (exactly representable), lo = the exact remainder (same sign
        // as the value), both converted with ReLU folded in: no max, no unpack.
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tmem_ld16(t_row + kColR + q * 16, v);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t0 = __uint_as_float(__float_as_uint(v[2 * e]) & 0xFFFFE000u), t1 = __uint_as_float(__float_as_uint(v[2 * e + 1]) & 0xFFFFE000u);
            hi[e] = pack_relu_h2(t0, t1);
            lo[e] = pack_relu_h2(v[2 * e] - t0, v[2 * e + 1] - t1);
          }
          tmem_st8(t_row + kColR + q * 16, hi);
          tmem_st8(t_row + kColR + q * 16 + 8, lo);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      named_bar(bar_id, 128);
      // ---- base L2: S = A2lo * W2^T + A2hi * W2^T  (K = 64 each)
      if (leader) {
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_f16_ts(pipe + kColS, pipe + kColR + j * 16 + 8, bdesc(kTcW2, 64, j), id16, j > 0 ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_f16_ts(pipe + kColS, pipe + kColR + j * 16, bdesc(kTcW2, 64, j), id16, 1u);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      float sigma;
      {
        float v[16];
        tmem_ld16(t_row + kColS, v);
        sigma = sel ? expf(v[0] - 1.0f) : 0.f;                                  // ngp.py:772-775
        // second k-chunk of the head input row (kernel order [SH(16) | 1 | feat(15)])
        uint32_t f[8] = {pack_h2(1.0f, v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]),
                         pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15])};
        tmem_st8(t_row + kColX, f);
      }
      tmem_wait_st();
      tc_fence_before();
      named_bar(bar_id, 128);
      // ---- head L1: R = [SH | 1, feat] (K=32) * W3p^T ; its completion also frees the quad's SH columns
      if (leader) {
        tc_fence_after();
        umma_f16_ts(pipe + kColR, c_sh, bdesc(kTcW3, 32, 0), id64, 0u);
        umma_f16_ts(pipe + kColR, pipe + kColX, bdesc(kTcW3, 32, 1), id64, 1u);
        umma_commit(bar_empty + quad * 8);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int stage = 0; stage < 2; ++stage) {
        // ReLU(R) -> X (A4 / A5)
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tmem_ld16(t_row + kColR + q * 16, v);
          uint32_t h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = pack_relu_h2(v[2 * e], v[2 * e + 1]);
          tmem_st8(t_row + kColX + q * 8, h);
        }
        tmem_wait_st();
        tc_fence_before();
        named_bar(bar_id, 128);
        if (leader) {
          tc_fence_after();
          if (stage == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16_ts(pipe + kColR, pipe + kColX + j * 8, bdesc(kTcW4, 64, j), id64, j > 0 ? 1u : 0u);   // head L2
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16_ts(pipe + kColR, pipe + kColX + j * 8, bdesc(kTcW5, 64, j), id16, j > 0 ? 1u : 0u);   // head L3 -> R[0,16)
          }
          umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase); phase ^= 1;
        tc_fence_after();
      }
      {
        float v[16];
        tmem_ld16(t_row + kColR, v);
        auto sig = [](float q) { return 1.0f / (1.0f + __expf(-q)); };
        const int64_t i = tile * 128 + row;
        if (i < M) {
          float4 o = make_float4(sig(v[0]), sig(v[1]), sig(v[2]), sigma);
          if (a.out4) a.out4[i] = o;
          else { a.rgb[3 * i] = o.x; a.rgb[3 * i + 1] = o.y; a.rgb[3 * i + 2] = o.z; a.density[i] = o.w; }
        }
      }
      // the next tile's L1 overwrites R, which every thread of the warpgroup must have finished reading
      tc_fence_before();
      named_bar(bar_id, 128);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
}

int launch_ngp_forward_tc(const qf_ngp* f, FieldTcArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<QF_TC_QUADS, QF_TC_GROUPS, QF_TC_CTAS_PER_SM>;
  a.desc = f->desc;
  a.table = f->d_table;
  a.weights_tc = f->d_weights_tc;
  constexpr int max_blocks = kNumSMs * QF_TC_CTAS_PER_SM;
  int64_t tiles = a.d_M ? (int64_t)max_blocks : ceil_div(a.M, 128);
  int blocks = (int)(tiles < max_blocks ? tiles : max_blocks);
  if (blocks < 1) blocks = 1;
  ngp_forward_tc_kernel<QF_TC_QUADS, QF_TC_GROUPS, QF_TC_CTAS_PER_SM><<<blocks, Cfg::kThreads, Cfg::kSmemBytes, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

int prep_weights_tc(qf_ngp* f, cudaStream_t st) {
  prep_weights_tc_kernel<<<(int)ceil_div(10240, 256), 256, 0, st>>>(f->d_weights, f->d_weights_tc);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

}  // namespace qf
