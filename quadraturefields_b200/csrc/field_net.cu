// The quadrature `Field` net (field.py:130-259) with back_prop=False, the setting of both reference
// call sites (train_field.py:248, train_finetune.py:397):
//
//   x01 = (x - xyz_min) / (xyz_max - xyz_min)                                   field.py:188
//   h   = hash grid(x01.detach())          fp16, 16 levels x 2 features          field.py:192
//   out = W3 act(W2 act(W1 [x01, h] + b1) + b2) + b3     torch fp32 Linear x3    field.py:193, BasicDecoder :91-104
//   field_grad = d(sum out)/dx  through the raw-xyz inputs of the MLP only       field.py:229-238
//
// field_grad is produced analytically in the forward kernel (reverse sweep through the two hidden layers), and the
// backward kernel differentiates BOTH outputs with respect to the MLP parameters and the grid table: the part through
// field_grad is the reference's double backward (autograd.grad(create_graph=True) followed by loss.backward()), which
// needs act'' of the hidden layers but only the FIRST-order grid backward.
//
// One thread = one sample; the three weight matrices live in shared memory (every lane reads the same weight:
// broadcast), per-sample hidden vectors in local memory; parameter gradients of a warp's 32 samples are formed as outer-product sums over two staged (32 x 36)
// shared-memory panels, accumulated per CTA in shared memory and flushed with one atomicAdd per entry per CTA.
#include "field_common.cuh"

namespace qf {

constexpr int kFnIn = 35;        // [x01 (3), encoding (32)]
constexpr int kFnPanel = 36;     // row stride of the staging panels

struct FieldNetArgs {
  qf_grid_desc desc;
  const __half2* table;
  const float *w1, *b1, *w2, *b2, *w3, *b3;   // biases may be NULL
  int out_dim;
  float lo[3], ext[3];
  const float* x;
  int64_t M;
  float* field;        // (M, out_dim)
  float* field_grad;   // (M, 3) or NULL
  // backward
  const float* g_field;   // (M, out_dim) or NULL
  const float* g_fgrad;   // (M, 3) or NULL
  float2* g_table;        // (n_entries) float2, accumulated
  float *g_w1, *g_b1, *g_w2, *g_b2, *g_w3, *g_b3;   // accumulated; bias grads may be NULL
};

template <int H>
struct FnSmem {
  static constexpr int kW1 = 0, kB1 = kW1 + H * kFnIn, kW2 = kB1 + H, kB2 = kW2 + H * H, kW3 = kB2 + H, kB3 = kW3 + 3 * H,
                       kTotal = kB3 + 4;
};

template <int H>
__device__ __forceinline__ void fn_load_weights(const FieldNetArgs& a, float* s, int tid, int nthreads) {
  using S = FnSmem<H>;
  for (int i = tid; i < H * kFnIn; i += nthreads) s[S::kW1 + i] = a.w1[i];
  for (int i = tid; i < H * H; i += nthreads) s[S::kW2 + i] = a.w2[i];
  for (int i = tid; i < 3 * H; i += nthreads) s[S::kW3 + i] = i < a.out_dim * H ? a.w3[i] : 0.f;
  for (int i = tid; i < H; i += nthreads) {
    s[S::kB1 + i] = a.b1 ? a.b1[i] : 0.f;
    s[S::kB2 + i] = a.b2 ? a.b2[i] : 0.f;
  }
  if (tid < 4) s[S::kB3 + tid] = (a.b3 && tid < a.out_dim) ? a.b3[tid] : 0.f;
}

template <int ACT>  // 0: ELU(alpha=1), 1: ReLU
__device__ __forceinline__ void fn_act(float z, float& a, float& d) {
  if (ACT == 0) {
    const float e = expf(z);
    a = z > 0.f ? z : e - 1.0f;
    d = z > 0.f ? 1.0f : e;
  } else {
    a = fmaxf(z, 0.f);
    d = z > 0.f ? 1.0f : 0.f;
  }
}
// second derivative from the activation and its first derivative: ELU'' = e^z = ELU' on z <= 0 (where ELU <= 0), else 0
template <int ACT>
__device__ __forceinline__ float fn_dd(float a, float d) { return (ACT == 0 && !(a > 0.f)) ? d : 0.f; }

// H = 16 (train_field.py) is fully unrolled: the per-sample vectors stay in registers (255 registers, 2 CTAs per SM;
// measured 0.76 ms against 1.05 ms for the rolled form on 442 k samples, forcing 3 CTAs/SM: 1.01 ms).  H = 32 keeps
// the outer loop of every matrix-vector product rolled with the vectors in local memory (fully unrolled it needs
// ~500 live floats and ptxas falls back to a 10 KB stack frame); weight rows are read from shared memory (broadcast).
template <int H>
__device__ __forceinline__ void fn_hidden(const float* s, const float* inp, float* z1) {
  using S = FnSmem<H>;
#pragma unroll (H <= 16 ? H : 1)
  for (int j = 0; j < H; ++j) {
    const float* w = s + S::kW1 + j * kFnIn;
    float acc = s[S::kB1 + j];
#pragma unroll (H <= 16 ? kFnIn : 7)
    for (int i = 0; i < kFnIn; ++i) acc = fmaf(w[i], inp[i], acc);
    z1[j] = acc;
  }
}

// out[j] = bias[j] + sum_k W[j][k] v[k]      (W row-major H x H)
template <int H>
__device__ __forceinline__ void fn_matvec(const float* W, const float* bias, const float* v, float* out) {
#pragma unroll (H <= 16 ? H : 1)
  for (int j = 0; j < H; ++j) {
    const float* w = W + j * H;
    float acc = bias ? bias[j] : 0.f;
#pragma unroll (H <= 16 ? H : 8)
    for (int k = 0; k < H; ++k) acc = fmaf(w[k], v[k], acc);
    out[j] = acc;
  }
}

// out[k] = sum_j W[j][k] v[j]                (transposed product)
template <int H>
__device__ __forceinline__ void fn_matvec_t(const float* W, const float* v, float* out) {
#pragma unroll (H <= 16 ? H : 1)
  for (int k = 0; k < H; ++k) {
    float acc = 0.f;
#pragma unroll (H <= 16 ? H : 8)
    for (int j = 0; j < H; ++j) acc = fmaf(W[j * H + k], v[j], acc);
    out[k] = acc;
  }
}

template <int H>
__device__ __forceinline__ void fn_encode(const FieldNetArgs& a, int64_t i, bool valid, float* inp) {
  float x01[3] = {0.5f, 0.5f, 0.5f};
  if (valid) {
#pragma unroll
    for (int c = 0; c < 3; ++c) x01[c] = __fdiv_rn(a.x[3 * i + c] - a.lo[c], a.ext[c]);
  }
  inp[0] = x01[0]; inp[1] = x01[1]; inp[2] = x01[2];
  encode_point(a.desc, a.table, x01[0], x01[1], x01[2], [&](int l, uint32_t h2) {
    float2 f = __half22float2(*reinterpret_cast<__half2*>(&h2));
    inp[3 + 2 * l] = f.x;
    inp[3 + 2 * l + 1] = f.y;
  });
}

template <int H, int ACT>
__global__ void __launch_bounds__(128) field_net_forward_kernel(const FieldNetArgs a) {
  using S = FnSmem<H>;
  __shared__ float s[S::kTotal];
  fn_load_weights<H>(a, s, threadIdx.x, blockDim.x);
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.M; i += (int64_t)gridDim.x * blockDim.x) {
    float inp[kFnIn];
    fn_encode<H>(a, i, true, inp);
    float z[H], a1[H], d1[H], a2[H], q2[H];
    fn_hidden<H>(s, inp, z);
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) fn_act<ACT>(z[j], a1[j], d1[j]);
    fn_matvec<H>(s + S::kW2, s + S::kB2, a1, z);
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) {
      float d2;
      fn_act<ACT>(z[j], a2[j], d2);
      q2[j] = d2 * (s[S::kW3 + j] + s[S::kW3 + H + j] + s[S::kW3 + 2 * H + j]);   // rows >= out_dim are zero
    }
    for (int o = 0; o < a.out_dim; ++o) {
      float acc = s[S::kB3 + o];
#pragma unroll 8
      for (int k = 0; k < H; ++k) acc = fmaf(s[S::kW3 + o * H + k], a2[k], acc);
      a.field[i * a.out_dim + o] = acc;
    }
    if (a.field_grad) {
      fn_matvec_t<H>(s + S::kW2, q2, z);          // p1
      float g[3] = {0.f, 0.f, 0.f};
#pragma unroll (H <= 16 ? H : 1)
      for (int k = 0; k < H; ++k) {
        const float q1 = d1[k] * z[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = fmaf(s[S::kW1 + k * kFnIn + c], q1, g[c]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) a.field_grad[3 * i + c] = __fdiv_rn(g[c], a.ext[c]);
    }
  }
}

// Outer-product sum over the warp's 32 samples: acc[r*ldc + c] += sum_s U[s][r] * V[s][c] for r < R, c < Cn.
// U, V are the warp's staging panels (row = sample, stride kFnPanel); acc is the CTA's shared accumulator.
__device__ __forceinline__ void fn_outer(const float* U, const float* V, int R, int Cn, float* acc, int ldc, int lane) {
  for (int e = lane; e < R * Cn; e += 32) {
    const int r = e / Cn, c = e - r * Cn;
    float sum = 0.f;
#pragma unroll 8
    for (int sidx = 0; sidx < 32; ++sidx) sum = fmaf(U[sidx * kFnPanel + r], V[sidx * kFnPanel + c], sum);
    atomicAdd(acc + r * ldc + c, sum);
  }
}

template <int H>
__device__ __forceinline__ void fn_stage(float* row, const float* v) {
#pragma unroll (H <= 16 ? H : 8)
  for (int j = 0; j < H; ++j) row[j] = v[j];
}

template <int H, int ACT>
__global__ void __launch_bounds__(128) field_net_backward_kernel(const FieldNetArgs a) {
  using S = FnSmem<H>;
  extern __shared__ __align__(16) float fn_smem[];
  float* s = fn_smem;                             // weights
  float* s_acc = s + S::kTotal;                   // gradient accumulators, same layout as the weights
  float* s_u = s_acc + S::kTotal;                 // [4][32 * kFnPanel]
  float* s_v = s_u + 4 * 32 * kFnPanel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  fn_load_weights<H>(a, s, tid, blockDim.x);
  for (int i = tid; i < S::kTotal; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float* U = s_u + warp * 32 * kFnPanel;
  float* V = s_v + warp * 32 * kFnPanel;
  float* urow = U + lane * kFnPanel;
  float* vrow = V + lane * kFnPanel;

  for (int64_t base = ((int64_t)blockIdx.x * 4 + warp) * 32; base < a.M; base += (int64_t)gridDim.x * 128) {
    const int64_t i = base + lane;
    const bool valid = i < a.M;
    float inp[kFnIn];
    fn_encode<H>(a, i, valid, inp);
    float t[H], a1[H], d1[H], a2[H], d2[H];
    fn_hidden<H>(s, inp, t);
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) fn_act<ACT>(t[j], a1[j], d1[j]);
    fn_matvec<H>(s + S::kW2, s + S::kB2, a1, t);
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) fn_act<ACT>(t[j], a2[j], d2[j]);
    // upstream gradients (zero for padding lanes)
    float gF[3] = {0.f, 0.f, 0.f}, g01[3] = {0.f, 0.f, 0.f};
    if (valid) {
      if (a.g_field) for (int o = 0; o < a.out_dim; ++o) gF[o] = a.g_field[i * a.out_dim + o];
      if (a.g_fgrad) {
#pragma unroll
        for (int c = 0; c < 3; ++c) g01[c] = __fdiv_rn(a.g_fgrad[3 * i + c], a.ext[c]);
      }
    }
    // ---- reverse sweep that produced field_grad (q2, p1, q1), and its adjoint (q1b, p1b, q2b)
    float q2[H], p1[H], p1b[H], q1b[H];
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) q2[j] = d2[j] * (s[S::kW3 + j] + s[S::kW3 + H + j] + s[S::kW3 + 2 * H + j]);
    fn_matvec_t<H>(s + S::kW2, q2, p1);
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) {
      const float* w = s + S::kW1 + j * kFnIn;
      q1b[j] = w[0] * g01[0] + w[1] * g01[1] + w[2] * g01[2];
      p1b[j] = q1b[j] * d1[j];
    }
    float z2b[H], w3sb[H];
    fn_matvec<H>(s + S::kW2, nullptr, p1b, t);    // q2b
#pragma unroll (H <= 16 ? H : 1)
    for (int j = 0; j < H; ++j) {
      const float w3s = s[S::kW3 + j] + s[S::kW3 + H + j] + s[S::kW3 + 2 * H + j];
      const float a2b = gF[0] * s[S::kW3 + j] + gF[1] * s[S::kW3 + H + j] + gF[2] * s[S::kW3 + 2 * H + j];
      w3sb[j] = t[j] * d2[j];
      z2b[j] = a2b * d2[j] + (t[j] * w3s) * fn_dd<ACT>(a2[j], d2[j]);
    }
    float z1b[H];
    fn_matvec_t<H>(s + S::kW2, z2b, t);           // a1b
#pragma unroll (H <= 16 ? H : 1)
    for (int k = 0; k < H; ++k) z1b[k] = t[k] * d1[k] + (q1b[k] * p1[k]) * fn_dd<ACT>(a1[k], d1[k]);
    // ---- grid: d loss / d encoding -> table scatter (first-order grid backward)
    if (valid && a.g_table) {
#pragma unroll 1
      for (int l = 0; l < 16; ++l) {
        float g0 = 0.f, g1 = 0.f;
#pragma unroll (H <= 16 ? H : 8)
        for (int j = 0; j < H; ++j) {
          g0 = fmaf(s[S::kW1 + j * kFnIn + 3 + 2 * l], z1b[j], g0);
          g1 = fmaf(s[S::kW1 + j * kFnIn + 4 + 2 * l], z1b[j], g1);
        }
        scatter_level(a.desc, a.g_table, l, inp[0], inp[1], inp[2], g0, g1);
      }
    }
    // ---- parameter gradients: outer-product sums over the warp's samples
    // (1) dW1 += z1b (x) inp ; db1 += z1b (column 35 of V is the constant 1)
    fn_stage<H>(urow, z1b);
#pragma unroll 7
    for (int k = 0; k < kFnIn; ++k) vrow[k] = inp[k];
    vrow[kFnIn] = 1.0f;
    __syncwarp();
    fn_outer(U, V, H, kFnIn, s_acc + S::kW1, kFnIn, lane);
    fn_outer(U, V + kFnIn, H, 1, s_acc + S::kB1, 1, lane);
    __syncwarp();
    // (2) dW1[:, 0:3] += q1 (x) g01
#pragma unroll 8
    for (int j = 0; j < H; ++j) urow[j] = d1[j] * p1[j];
    vrow[0] = g01[0]; vrow[1] = g01[1]; vrow[2] = g01[2];
    __syncwarp();
    fn_outer(U, V, H, 3, s_acc + S::kW1, kFnIn, lane);
    __syncwarp();
    // (3) dW2 += z2b (x) a1 ; db2 += z2b
    fn_stage<H>(urow, z2b);
    fn_stage<H>(vrow, a1);
    vrow[H] = 1.0f;
    __syncwarp();
    fn_outer(U, V, H, H, s_acc + S::kW2, H, lane);
    fn_outer(U, V + H, H, 1, s_acc + S::kB2, 1, lane);
    __syncwarp();
    // (4) dW2 += q2 (x) p1b
    fn_stage<H>(urow, q2);
    fn_stage<H>(vrow, p1b);
    __syncwarp();
    fn_outer(U, V, H, H, s_acc + S::kW2, H, lane);
    __syncwarp();
    // (5) dW3[o] += gF[o] * a2 ; db3 += gF
    urow[0] = gF[0]; urow[1] = gF[1]; urow[2] = gF[2];
    fn_stage<H>(vrow, a2);
    vrow[H] = 1.0f;
    __syncwarp();
    fn_outer(U, V, 3, H, s_acc + S::kW3, H, lane);
    fn_outer(U, V + H, 3, 1, s_acc + S::kB3, 1, lane);
    __syncwarp();
    // (6) the field_grad path reaches every row o < out_dim of W3 with the same value (sum over outputs)
    fn_stage<H>(vrow, w3sb);
    __syncwarp();
    for (int e = lane; e < a.out_dim * H; e += 32) {
      const int c = e % H;
      float sum = 0.f;
#pragma unroll 8
      for (int sidx = 0; sidx < 32; ++sidx) sum += V[sidx * kFnPanel + c];
      atomicAdd(s_acc + S::kW3 + e, sum);
    }
    __syncwarp();
  }
  __syncthreads();
  // flush: one global atomicAdd per entry per CTA
  for (int i = tid; i < H * kFnIn; i += blockDim.x) atomicAdd(a.g_w1 + i, s_acc[S::kW1 + i]);
  for (int i = tid; i < H * H; i += blockDim.x) atomicAdd(a.g_w2 + i, s_acc[S::kW2 + i]);
  for (int i = tid; i < a.out_dim * H; i += blockDim.x) atomicAdd(a.g_w3 + i, s_acc[S::kW3 + i]);
  for (int i = tid; i < H; i += blockDim.x) {
    if (a.g_b1) atomicAdd(a.g_b1 + i, s_acc[S::kB1 + i]);
    if (a.g_b2) atomicAdd(a.g_b2 + i, s_acc[S::kB2 + i]);
  }
  if (a.g_b3 && tid < a.out_dim) atomicAdd(a.g_b3 + tid, s_acc[S::kB3 + tid]);
}

template <int H, int ACT>
static int launch_bwd(const FieldNetArgs& a, int blocks, cudaStream_t st) {
  constexpr int smem = (2 * FnSmem<H>::kTotal + 2 * 4 * 32 * kFnPanel) * (int)sizeof(float);
  QF_ENSURE_DYNAMIC_SMEM((field_net_backward_kernel<H, ACT>), smem);
  field_net_backward_kernel<H, ACT><<<blocks, 128, smem, st>>>(a);
  QF_LAUNCH_CHECK();
  return QF_OK;
}

static int check_desc(const qf_ngp* grid, const qf_field_desc* fd, const char* who) {
  QF_REQUIRE(grid && fd, "%s: NULL grid / descriptor", who);
  QF_REQUIRE(grid->desc.n_levels == 16, "%s: the Field MLP takes the 32-wide encoding (16 levels x 2)", who);
  QF_REQUIRE(fd->hidden == 16 || fd->hidden == 32, "%s: hidden_size=%d (16 or 32 supported)", who, fd->hidden);
  QF_REQUIRE(fd->out_dim >= 1 && fd->out_dim <= 3, "%s: output_dim=%d outside [1,3]", who, fd->out_dim);
  QF_REQUIRE(fd->activation == QF_ACT_ELU || fd->activation == QF_ACT_RELU, "%s: activation=%d", who, fd->activation);
  for (int c = 0; c < 3; ++c) QF_REQUIRE(fd->xyz_max[c] > fd->xyz_min[c], "%s: empty bounding box", who);
  return QF_OK;
}

static void fill_args(FieldNetArgs& a, const qf_ngp* grid, const qf_field_desc* fd, const float* w1, const float* b1,
                      const float* w2, const float* b2, const float* w3, const float* b3, const float* x, int64_t M) {
  a.desc = grid->desc; a.table = grid->d_table;
  a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.w3 = w3; a.b3 = b3;
  a.out_dim = fd->out_dim;
  for (int c = 0; c < 3; ++c) { a.lo[c] = fd->xyz_min[c]; a.ext[c] = fd->xyz_max[c] - fd->xyz_min[c]; }
  a.x = x; a.M = M;
}

}  // namespace qf

using namespace qf;

extern "C" int qf_grid_create(const qf_grid_desc* desc, const float* d_table, int64_t n_entries, void* stream, qf_ngp** out) {
  QF_REQUIRE(desc && d_table && out, "qf_grid_create: NULL argument");
  QF_REQUIRE(desc->n_levels >= 1 && desc->n_levels <= QF_MAX_LEVELS, "qf_grid_create: n_levels=%d", desc->n_levels);
  int64_t need = (int64_t)desc->offset[desc->n_levels - 1] + desc->size[desc->n_levels - 1];
  QF_REQUIRE(n_entries >= need, "qf_grid_create: table has %lld entries, level table needs %lld", (long long)n_entries,
             (long long)need);
  qf_ngp* f = new qf_ngp();
  f->desc = *desc;
  f->n_entries = n_entries;
  if (cudaMalloc((void**)&f->d_table, sizeof(__half2) * (size_t)n_entries) != cudaSuccess) {
    set_error("qf_grid_create: device allocation failed");
    qf_ngp_destroy(f);
    return QF_ERR_CUDA;
  }
  int rc = qf_ngp_update(f, d_table, nullptr, nullptr, stream);
  if (rc == QF_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) { set_error("qf_grid_create: upload failed"); rc = QF_ERR_CUDA; }
  if (rc != QF_OK) { qf_ngp_destroy(f); return rc; }
  *out = f;
  return QF_OK;
}

extern "C" int qf_field_forward(const qf_ngp* grid, const qf_field_desc* fd, const float* d_w1, const float* d_b1,
                                const float* d_w2, const float* d_b2, const float* d_w3, const float* d_b3, const float* d_x,
                                int64_t M, float* d_field, float* d_field_grad, void* stream) {
  int rc = check_desc(grid, fd, "qf_field_forward");
  if (rc != QF_OK) return rc;
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_w1 && d_w2 && d_w3 && d_x && d_field, "qf_field_forward: NULL argument");
  FieldNetArgs a = {};
  fill_args(a, grid, fd, d_w1, d_b1, d_w2, d_b2, d_w3, d_b3, d_x, M);
  a.field = d_field; a.field_grad = d_field_grad;
  int64_t want = ceil_div(M, 128);
  const int blocks = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
  const int key = fd->hidden * 2 + fd->activation;
  switch (key) {
    case 32 + QF_ACT_ELU: field_net_forward_kernel<16, 0><<<blocks, 128, 0, st>>>(a); break;
    case 32 + QF_ACT_RELU: field_net_forward_kernel<16, 1><<<blocks, 128, 0, st>>>(a); break;
    case 64 + QF_ACT_ELU: field_net_forward_kernel<32, 0><<<blocks, 128, 0, st>>>(a); break;
    default: field_net_forward_kernel<32, 1><<<blocks, 128, 0, st>>>(a); break;
  }
  QF_LAUNCH_CHECK();
  return QF_OK;
}

extern "C" int qf_field_backward(const qf_ngp* grid, const qf_field_desc* fd, const float* d_w1, const float* d_b1,
                                 const float* d_w2, const float* d_b2, const float* d_w3, const float* d_b3, const float* d_x,
                                 int64_t M, const float* d_grad_field, const float* d_grad_field_grad, float* d_grad_table,
                                 float* d_grad_w1, float* d_grad_b1, float* d_grad_w2, float* d_grad_b2, float* d_grad_w3,
                                 float* d_grad_b3, void* stream) {
  int rc = check_desc(grid, fd, "qf_field_backward");
  if (rc != QF_OK) return rc;
  if (M == 0) return QF_OK;
  QF_REQUIRE(d_w1 && d_w2 && d_w3 && d_x && d_grad_w1 && d_grad_w2 && d_grad_w3, "qf_field_backward: NULL argument");
  QF_REQUIRE(d_grad_field || d_grad_field_grad, "qf_field_backward: no upstream gradient");
  QF_REQUIRE((reinterpret_cast<uintptr_t>(d_grad_table) & 15) == 0, "qf_field_backward: d_grad_table must be 16-byte aligned");
  FieldNetArgs a = {};
  fill_args(a, grid, fd, d_w1, d_b1, d_w2, d_b2, d_w3, d_b3, d_x, M);
  a.g_field = d_grad_field; a.g_fgrad = d_grad_field_grad;
  a.g_table = reinterpret_cast<float2*>(d_grad_table);
  a.g_w1 = d_grad_w1; a.g_b1 = d_grad_b1; a.g_w2 = d_grad_w2; a.g_b2 = d_grad_b2; a.g_w3 = d_grad_w3; a.g_b3 = d_grad_b3;
  int64_t want = ceil_div(M, 128);
  // 4 CTAs per SM fit (44-55 KB of shared memory each); the per-CTA flush costs ~2 k atomics, so no more than that
  const int blocks = (int)(want < (int64_t)kNumSMs * 4 ? want : (int64_t)kNumSMs * 4);
  cudaStream_t st = (cudaStream_t)stream;
  const int key = fd->hidden * 2 + fd->activation;
  switch (key) {
    case 32 + QF_ACT_ELU: return launch_bwd<16, 0>(a, blocks, st);
    case 32 + QF_ACT_RELU: return launch_bwd<16, 1>(a, blocks, st);
    case 64 + QF_ACT_ELU: return launch_bwd<32, 0>(a, blocks, st);
    default: return launch_bwd<32, 1>(a, blocks, st);
  }
}
