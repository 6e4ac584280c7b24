// (3) Blackwell-native variant of the fused field kernel: the five MLP layers run on the 5th-generation tensor
// cores with tcgen05.mma (operands in shared memory through UMMA descriptors, accumulators in TMEM), one CTA of
// 128 threads per 128-sample tile.
//
//   thread i  <->  sample i of the tile  <->  TMEM lane i
//
// Each thread gathers the hash-grid features of ITS sample and stores them as one row of the layer-1 A operand
// (K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices).  One elected thread issues
// tcgen05.mma (M=128, N=64 or 16, K=16 per instruction, fp16 operands, fp32 accumulate) and commits to an
// mbarrier; every thread then reads its own accumulator row back with tcgen05.ld (32x32b: lane = row), applies
// ReLU / exp / sigmoid in fp32 and writes the next layer's A row.  Compared with the mma.sync kernel (field.cu) the
// weight fragments are never loaded through the LSU — the tensor core reads them from shared memory through the
// async proxy — which leaves the L1TEX pipe, the limiter of this gather-bound kernel, to the table gathers.
//
// Numerics are those of field.cu: fp16 operands, fp32 accumulation, hi+lo split of the hidden activations for the
// density logit.
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

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // never hang the GPU on a protocol error
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane
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

// shared memory map (bytes)
constexpr int kSmW = 0;                        // weights, 20480
constexpr int kSmX = kSmW + kTcWBytes;         // region X, 16 KB: A1 (128x32) -> A2lo (128x64) -> A3 (128x32)
constexpr int kSmY = kSmX + 128 * 64 * 2;      // region Y, 16 KB: A2hi -> A4 -> A5 (128x64)
constexpr int kSmBar = kSmY + 128 * 64 * 2;    // mbarrier (8 B) + tmem base (4 B)
constexpr int kTcSmemBytes = kSmBar + 64;
constexpr int kTmemCols = 128;                 // D wide at columns [0,64), D narrow at [64,80)

// store 8 halves (one 16-byte k-chunk) of this thread's row
__device__ __forceinline__ void st_chunk(unsigned char* tile, int row, int chunk, int K, uint4 v) {
  *reinterpret_cast<uint4*>(tile + (row >> 3) * (K >> 3) * 128 + chunk * 128 + (row & 7) * 16) = v;
}

// D[128 x N] (+)= A[128 x K] * W[N x K]^T, K/16 instructions
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_addr, int K_a, uint32_t w_addr, int K, int N, bool accumulate) {
  const uint32_t idesc = umma_idesc(128, N);
  for (int j = 0; j < K / 16; ++j) {
    const uint64_t da = umma_desc(a_addr + j * 256, 128, (K_a >> 3) * 128);
    const uint64_t db = umma_desc(w_addr + j * 256, 128, (K >> 3) * 128);
    umma_f16(tmem_d, da, db, idesc, (accumulate || j > 0) ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(128, 4) ngp_forward_tc_kernel(const FieldTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint64_t* mbar_ptr = reinterpret_cast<uint64_t*>(smem + kSmBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 8);
  const uint32_t mbar = smem_u32(mbar_ptr);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights_tc);
    uint4* dst = reinterpret_cast<uint4*>(smem + kSmW);
    for (int i = tid; i < kTcWBytes / 16; i += 128) dst[i] = __ldg(src + i);
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_async_smem();          // weights written through the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
  const uint32_t sW = smem_u32(smem + kSmW), sX = smem_u32(smem + kSmX), sY = smem_u32(smem + kSmY);
  unsigned char* X = smem + kSmX;
  unsigned char* Y = smem + kSmY;
  uint32_t phase = 0;

  const int64_t M = a.d_M ? (int64_t)__ldg(a.d_M) : a.M;
  const float amin[3] = {a.desc.aabb[0], a.desc.aabb[1], a.desc.aabb[2]};
  const float aext[3] = {a.desc.aabb[3] - a.desc.aabb[0], a.desc.aabb[4] - a.desc.aabb[1], a.desc.aabb[5] - a.desc.aabb[2]};

  for (int64_t base = (int64_t)blockIdx.x * 128; base < M; base += (int64_t)gridDim.x * 128) {
    const int64_t i = base + tid;
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
    // ---- encode: 32 features = 4 k-chunks of the layer-1 A row
    {
      unsigned char* xrow = X + (tid >> 3) * 512 + (tid & 7) * 16;   // K=32 canonical row: chunk c at +128*c, level l at chunk l/4
      encode_point(a.desc, a.table, x, y, z,
                   [&](int l, uint32_t h2) { *reinterpret_cast<uint32_t*>(xrow + (l >> 2) * 128 + (l & 3) * 4) = h2; });
    }
    uint32_t shp[8];
    {
      float dx = 0.f, dy = 0.f, dz = 1.f;
      if (valid) {
        int64_t r = a.ray64 ? __ldg(a.ray64 + i) : (a.ray32 ? (int64_t)__ldg(a.ray32 + i * a.ray32_stride) : i);
        const float* dp = a.dirs + 3 * r;
        dx = ((__ldg(dp) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dy = ((__ldg(dp + 1) + 1.0f) / 2.0f) * 2.0f - 1.0f;
        dz = ((__ldg(dp + 2) + 1.0f) / 2.0f) * 2.0f - 1.0f;
      }
      float sh[16];
      sh4(dx, dy, dz, sh);
#pragma unroll
      for (int q = 0; q < 8; ++q) shp[q] = pack_h2(sh[2 * q], sh[2 * q + 1]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- base L1: D[.,0:64) = A1 (X, K=32) * W1^T
    if (tid == 0) { tc_fence_after(); issue_layer(tmem_base, sX, 32, sW + kTcW1, 32, 64, false); umma_commit(mbar); }
    mbar_wait(mbar, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // 64 hidden units: ReLU, hi -> Y, lo -> X
        tmem_ld16(t_row + q * 16, v);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float r0 = fmaxf(v[2 * e], 0.f), r1 = fmaxf(v[2 * e + 1], 0.f);
          __half2 h = __floats2half2_rn(r0, r1);
          float2 f = __half22float2(h);
          hi[e] = *reinterpret_cast<uint32_t*>(&h);
          lo[e] = pack_h2(r0 - f.x, r1 - f.y);
        }
        st_chunk(Y, tid, 2 * q, 64, make_uint4(hi[0], hi[1], hi[2], hi[3]));
        st_chunk(Y, tid, 2 * q + 1, 64, make_uint4(hi[4], hi[5], hi[6], hi[7]));
        st_chunk(X, tid, 2 * q, 64, make_uint4(lo[0], lo[1], lo[2], lo[3]));
        st_chunk(X, tid, 2 * q + 1, 64, make_uint4(lo[4], lo[5], lo[6], lo[7]));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- base L2: D[.,64:80) = A2lo (X) * W2^T + A2hi (Y) * W2^T
    if (tid == 0) {
      tc_fence_after();
      issue_layer(tmem_base + 64, sX, 64, sW + kTcW2, 64, 16, false);
      issue_layer(tmem_base + 64, sY, 64, sW + kTcW2, 64, 16, true);
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    tc_fence_after();
    float sigma;
    {
      float v[16];
      tmem_ld16(t_row + 64, v);
      sigma = sel ? expf(v[0] - 1.0f) : 0.f;                                  // ngp.py:772-775
      // head input row (kernel order): [SH(16) | 1 | feat(15)]
      st_chunk(X, tid, 0, 32, make_uint4(shp[0], shp[1], shp[2], shp[3]));
      st_chunk(X, tid, 1, 32, make_uint4(shp[4], shp[5], shp[6], shp[7]));
      st_chunk(X, tid, 2, 32, make_uint4(pack_h2(1.0f, v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
      st_chunk(X, tid, 3, 32, make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15])));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- head L1: D[.,0:64) = A3 (X, K=32) * W3p^T
    if (tid == 0) { tc_fence_after(); issue_layer(tmem_base, sX, 32, sW + kTcW3, 32, 64, false); umma_commit(mbar); }
    mbar_wait(mbar, phase); phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int stage = 0; stage < 2; ++stage) {
      // ReLU(D) -> Y (A4 / A5)
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld16(t_row + q * 16, v);
        uint32_t h[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) h[e] = pack_h2(fmaxf(v[2 * e], 0.f), fmaxf(v[2 * e + 1], 0.f));
        st_chunk(Y, tid, 2 * q, 64, make_uint4(h[0], h[1], h[2], h[3]));
        st_chunk(Y, tid, 2 * q + 1, 64, make_uint4(h[4], h[5], h[6], h[7]));
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        if (stage == 0) issue_layer(tmem_base, sY, 64, sW + kTcW4, 64, 64, false);        // head L2 -> D[.,0:64)
        else issue_layer(tmem_base + 64, sY, 64, sW + kTcW5, 64, 16, false);              // head L3 -> D[.,64:80)
        umma_commit(mbar);
      }
      mbar_wait(mbar, phase); phase ^= 1;
      tc_fence_after();
    }
    {
      float v[16];
      tmem_ld16(t_row + 64, v);
      auto sig = [](float q) { return 1.0f / (1.0f + __expf(-q)); };
      if (valid) {
        float4 o = make_float4(sig(v[0]), sig(v[1]), sig(v[2]), sigma);
        if (a.out4) a.out4[i] = o;
        else { a.rgb[3 * i] = o.x; a.rgb[3 * i + 1] = o.y; a.rgb[3 * i + 2] = o.z; a.density[i] = o.w; }
      }
    }
    tc_fence_before();
    __syncthreads();   // region X is rewritten by the next tile's encode; its last reader (head L1) has completed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
}

int launch_ngp_forward_tc(const qf_ngp* f, FieldTcArgs& a, cudaStream_t st) {
  a.desc = f->desc;
  a.table = f->d_table;
  a.weights_tc = f->d_weights_tc;
  QF_ENSURE_DYNAMIC_SMEM(ngp_forward_tc_kernel, kTcSmemBytes);
  int64_t tiles = a.d_M ? (int64_t)kNumSMs * 4 : ceil_div(a.M, 128);
  int blocks = (int)(tiles < (int64_t)kNumSMs * 4 ? tiles : (int64_t)kNumSMs * 4);
  if (blocks < 1) blocks = 1;
  ngp_forward_tc_kernel<<<blocks, 128, kTcSmemBytes, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

int prep_weights_tc(qf_ngp* f, cudaStream_t st) {
  prep_weights_tc_kernel<<<(int)ceil_div(10240, 256), 256, 0, st>>>(f->d_weights, f->d_weights_tc);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

}  // namespace qf
