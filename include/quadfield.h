/*
 * quadfield.h — C ABI of libquadfield.so, the B200 (sm_100a) drop-in for the Quadfield
 * (ubc-vision/quadraturefields) render hot path.
 *
 * Conventions
 *   - every pointer whose name starts with d_ is DEVICE memory on the current CUDA device;
 *     h_ is host memory.  The caller owns all buffers it passes (outputs come from torch.empty).
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised, unless the
 *     function is documented as synchronous (the *_create functions and qf_hits_total).
 *   - return value 0 = ok; anything else = error, text via qf_last_error() (thread local).
 *   - no exceptions cross this boundary; no torch types appear in it.
 *
 * Each entry point names the reference interface it replaces; file:line is into
 * /root/reference/examples/ (see SURVEY.md §8b).  The reference-side binding a maintainer would
 * add is shown in INTEGRATION.md.
 */
#ifndef QUADFIELD_H
#define QUADFIELD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QF_OK 0
#define QF_ERR_INVALID 1
#define QF_ERR_CUDA 2
#define QF_ERR_UNSUPPORTED 3

#define QF_MAX_HITS 32       /* largest K (max_hits / num_intersections) supported */
#define QF_MAX_LEVELS 16     /* hash-grid levels */
#define QF_MAX_LOBES 8       /* spherical-Gaussian lobes in the baked texture set */

#define QF_BG_WHITE 0
#define QF_BG_BLACK 1
#define QF_BG_RANDOM 2       /* "random": hit rays mix with render_bkgd, missed rays stay white (quirk Q2) */

const char* qf_last_error(void);
int qf_version(void);

/* ------------------------------------------------------------------------------------------
 * (1) Geometry: quadrature mesh + BVH.
 * Replaces the private OptiX pybind object `intersector.Intersector(vertices_flat, max_hits, device)`
 * with `.find_intersections(rays_flat)` / `.update_vertices(vertices)` (mesh_utils.py:77-96) and the
 * Embree `RayMeshIntersector(mesh).intersects_id(...)` default (mesh_utils.py:223,350-354).
 * ------------------------------------------------------------------------------------------ */
typedef struct qf_mesh qf_mesh;

/* Build: padded triangle boxes, fp64 face normals, 63-bit Morton LBVH.  Synchronous. */
int qf_mesh_create(const float* d_vertices, int64_t n_vertices, const int32_t* d_faces, int64_t n_faces,
                   void* stream, qf_mesh** out);
/* `Intersector.update_vertices` (mesh_utils.py:83-84; train_finetune.py:714-718): same topology, new
 * positions; rebuilds the hierarchy on `stream`. */
int qf_mesh_update_vertices(qf_mesh* mesh, const float* d_vertices, void* stream);
void qf_mesh_destroy(qf_mesh* mesh);
/* Hit-set semantics.  eps = 0 (default): ALL hits with t > 0 in (t, triangle id) order, first K — what the reference's
 * OptiX intersector returns (mesh_utils.py:77-96, `optix=True`).  eps > 0: the SHIPPED default intersector, trimesh +
 * Embree (`optix=False`; mesh_utils.py:223, 350-354 `intersects_id(multiple_hits=True, max_hits=K)`): a first-hit query
 * repeated up to K times, each restarted eps beyond the previous hit — of the hits in (t, id) order one is kept iff it lies
 * more than eps (fp32 difference of t) behind the last kept one, until K are kept.  The reference's eps is
 * clip(1e-4 * 100 / mesh.scale, 1e-8, inf) world units (mesh.scale = bounding-box diagonal).  The filter sees the
 * QF_MAX_HITS nearest raw hits of a ray.  Applies to qf_trace_firstk and the fused renders of this mesh. */
int qf_mesh_set_restart_eps(qf_mesh* mesh, float eps);
/* info[0]=n_faces, [1]=n_vertices, [2]=n_nodes, [3]=device bytes held */
int qf_mesh_info(const qf_mesh* mesh, int64_t* info4, float* box_pad);

/* `find_intersections`: first K hits of every ray ordered by (t, triangle id); d_tri is the
 * reference's int[N*max_hits] with -1 padding (mesh_utils.py:91-96).  d_t (may be NULL) gets the
 * Möller–Trumbore t (+inf padded), d_count min(total,K), d_total (may be NULL) the untruncated count. */
/* d_workspace (>= qf_trace_workspace_bytes(n_rays), may be NULL): scratch for the incoherent-ray path — rays that miss
 * the scene box are dropped up front and persistent warps refill finished lanes; results are identical without it. */
size_t qf_trace_workspace_bytes(int64_t n_rays);
int qf_trace_firstk(const qf_mesh* mesh, const float* d_origins, const float* d_dirs, int64_t n_rays, int K,
                    int32_t* d_tri, float* d_t, int32_t* d_count, int32_t* d_total, void* d_workspace,
                    size_t workspace_bytes, void* stream);

/* Exclusive scan of d_count into d_offsets[n_rays+1] (ray-major hit layout).  workspace >= qf_scan_workspace_bytes. */
size_t qf_scan_workspace_bytes(int64_t n);
int qf_hits_offsets(const int32_t* d_count, int64_t n_rays, int64_t* d_offsets, void* d_workspace,
                    size_t workspace_bytes, void* stream);
/* Synchronous read of d_offsets[n_rays] (the reference syncs here too: it returns numpy arrays). */
int qf_hits_total(const int64_t* d_offsets, int64_t n_rays, int64_t* h_total, void* stream);

/* `MeshIntersection.sampling_raytrace_numpy` (mesh_utils.py:343-387) after the intersector call:
 * plane-hit points (mesh_utils.py:33-40), dirs/(|d|+1e-7), depth=|p-o|, per-ray stable sort by depth.
 * Outputs are the reference's data tuple (nerf_synthetic.py:256-257) in ray-major order. */
int qf_hits_pack(const qf_mesh* mesh, const float* d_origins, const float* d_dirs, int64_t n_rays, int K,
                 const int32_t* d_tri, const int32_t* d_count, const int64_t* d_offsets,
                 float* d_points, float* d_vectors, int64_t* d_index_ray, float* d_depth, int64_t* d_index_tri,
                 float* d_origins_out, void* stream);

/* `sampling_indexing` (mesh_utils.py:389-412): per-ray stable re-sort by depth of a ray-major tuple
 * (the reference does a GPU->CPU lexsort here), pack boundaries (kaolin mark_pack_boundaries) and the
 * constant quadrature step (find_deltas, :225-231).  d_perm receives the permutation applied. */
int qf_hits_resort(const int64_t* d_index_ray, const float* d_depth, int64_t n_hits, int64_t* d_perm,
                   uint8_t* d_boundary, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2)+(3) Instant-NGP radiance field: multiresolution hash grid + fully fused 64-wide MLPs.
 * Replaces tinycudann HashGrid / FullyFusedMLP / SphericalHarmonics as used by
 * NGPRadianceField (radiance_fields/ngp.py:657-809).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_levels;                    /* <= QF_MAX_LEVELS, 2 features per level */
  float scale[QF_MAX_LEVELS];          /* tcnn grid_scale(level) */
  uint32_t resolution[QF_MAX_LEVELS];  /* ceil(scale)+1 */
  uint32_t offset[QF_MAX_LEVELS];      /* first entry of the level in the table */
  uint32_t size[QF_MAX_LEVELS];        /* entries in the level */
  uint32_t hashed[QF_MAX_LEVELS];      /* 1: spatial hash, 0: dense index */
  float aabb[6];                       /* NGPRadianceField.aabb buffer */
} qf_grid_desc;

typedef struct qf_ngp qf_ngp;

/* Parameters in tinycudann's layout, fp32 master copies on the device (converted to fp16 inside):
 *   d_table   (n_entries, 2)
 *   d_base_w  [64x32 | 16x64]            mlp_base  32 -> 64 -> 16   (row-major (out,in), no bias)
 *   d_head_w  [64x32 | 64x64 | 16x64]    mlp_head  31(+pad) -> 64 -> 64 -> 3(+pad)
 * Synchronous. */
int qf_ngp_create(const qf_grid_desc* desc, const float* d_table, int64_t n_entries, const float* d_base_w,
                  const float* d_head_w, void* stream, qf_ngp** out);
/* refresh the fp16 working copies after an optimizer step */
int qf_ngp_update(qf_ngp* f, const float* d_table, const float* d_base_w, const float* d_head_w, void* stream);
void qf_ngp_destroy(qf_ngp* f);

/* tcnn HashGrid forward alone: x01 (M,3) in [0,1] -> (M, 2*n_levels) fp32 holding fp16-rounded features. */
int qf_hashgrid_forward(const qf_ngp* f, const float* d_x01, int64_t M, float* d_enc, void* stream);
/* tcnn HashGrid backward alone: ACCUMULATES dL/dtable (n_entries,2) from dL/denc (M, 2*n_levels) at x01 (M,3).
 * Every d_grad_table of this header must be 16-byte aligned (pairs of adjacent entries take one 16-byte reduction). */
int qf_hashgrid_backward(const qf_ngp* f, const float* d_x01, const float* d_grad_enc, int64_t M, float* d_grad_table,
                         void* stream);
/* `NGPRadianceField.query_density(x, return_feat)` (ngp.py:757-779): d_feat (M,15) may be NULL. */
int qf_ngp_query_density(const qf_ngp* f, const float* d_positions, int64_t M, float* d_density, float* d_feat,
                         void* stream);
/* `NGPRadianceField.forward(positions, directions)` (ngp.py:798-809) -> rgb (M,3), density (M,1).
 * If d_ray_index != NULL directions are gathered as d_directions[d_ray_index[i]] (utils.py:515-529). */
int qf_ngp_forward(const qf_ngp* f, const float* d_positions, const float* d_directions,
                   const int64_t* d_ray_index, int64_t M, float* d_rgb, float* d_density, void* stream);

/* Training mode (tinycudann grid + MLP backward under autograd: train_finetune.py:494-531, train_fit_sg.py):
 * given dL/drgb (M,3) and dL/ddensity (M) [may be NULL], ACCUMULATES into caller-initialised fp32 buffers in
 * tinycudann's layout: d_grad_table (n_entries,2), d_grad_base_w (64x32 | 16x64), d_grad_head_w (64x32 | 64x64 | 16x64).
 * The forward is recomputed inside; positions/directions get no gradient.  workspace >= qf_ngp_backward_workspace_bytes(M). */
size_t qf_ngp_backward_workspace_bytes(int64_t M);
int qf_ngp_backward(const qf_ngp* f, const float* d_positions, const float* d_directions, const int64_t* d_ray_index,
                    int64_t M, const float* d_grad_rgb, const float* d_grad_density, float* d_grad_table,
                    float* d_grad_base_w, float* d_grad_head_w, void* d_workspace, size_t workspace_bytes, void* stream);
/* The same, plus dL/dpositions (M,3) — tinycudann's grid input gradient, which the finetune step relies on when the
 * deformation field moves the quadrature points (utils.py:566-583: points = xyzs + dh feed radiance_field). The
 * gradient flows through the trilinear weights of the hash grid and the 1/(aabb_max-aabb_min) normalisation
 * (ngp.py:761-763); directions get no gradient. d_grad_positions may be NULL. */
int qf_ngp_backward_inputs(const qf_ngp* f, const float* d_positions, const float* d_directions,
                           const int64_t* d_ray_index, int64_t M, const float* d_grad_rgb,
                           const float* d_grad_density, float* d_grad_table, float* d_grad_base_w,
                           float* d_grad_head_w, float* d_grad_positions, void* d_workspace, size_t workspace_bytes,
                           void* stream);
/* Backward of `query_density(x, return_feat=True)` (ngp.py:402-428): upstream gradients of sigma (M, may be NULL) and of
 * the 15 geo features (M,15) into the hash table and the base MLP (d_grad_base_w: 64x32 + 16x64), optionally the
 * positions.  Workspace as qf_ngp_backward.  The spherical-Gaussian field (ngp.py:284-470) trains through this. */
int qf_ngp_backward_features(const qf_ngp* f, const float* d_positions, int64_t M, const float* d_grad_density,
                             const float* d_grad_feat, float* d_grad_table, float* d_grad_base_w,
                             float* d_grad_positions, void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4) Baked spherical-Gaussian texture path.
 * Replaces FeatureCompression.get_features_from_texture_map (texture_utils.py:149-175), the
 * barycentric texel lookup (utils.py:1055-1063) and features_to_rgb (ngp.py:371-393,456-461).
 * ------------------------------------------------------------------------------------------ */
typedef struct qf_texture qf_texture;
/* planes as the reference holds them: alpha (S,S) u8, diffuse (S,S,3) u8, per lobe colour (S,S,3) u8 and
 * [lambda,azimuth,elevation] (S,S,3) u8.  Repacked into one interleaved record per texel.
 * colour_logit = 1 only for compression_type == "sigma" (quirk Q5).  Synchronous. */
int qf_texture_create(int size, int num_lobes, const uint8_t* d_alpha, const uint8_t* d_diffuse,
                      const uint8_t* const* h_d_colors, const uint8_t* const* h_d_lambdas, int colour_logit,
                      float lambda_thres, void* stream, qf_texture** out);
void qf_texture_destroy(qf_texture* t);
/* get_features_from_texture_map: indices (M,2) int64 -> features (M, 3+7L+1) fp32 */
int qf_texture_decode(const qf_texture* t, const int64_t* d_indices, int64_t M, float* d_features, void* stream);
/* The bake writer `FeatureCompression.compress` / `assign_values_to_texture_map` (texture_utils.py:67-106):
 * quantise feature rows (M, 3+7L+1) into the uint8 planes, at row i (d_indices NULL) or at texel d_indices[i] of SxS planes. */
int qf_texture_compress(const float* d_features, int64_t M, int num_lobes, int colour_logit, float lambda_thres,
                        const int64_t* d_indices, int texture_size, uint8_t* d_alpha, uint8_t* d_diffuse,
                        uint8_t* const* h_d_colors, uint8_t* const* h_d_lambdas, void* stream);
/* features_to_rgb: features (M, 3+7L[+1]) with row stride `stride` floats, dirs (M,3) -> rgb (M,3) */
int qf_sg_features_to_rgb(const float* d_features, int64_t stride, int num_lobes, const float* d_dirs, int64_t M,
                          float* d_rgb, void* stream);
/* its backward with respect to the feature rows (the SG field is fitted through it, train_fit_sg.py):
 * grad_features (M, >= 3+7L) with row stride `grad_stride`; directions get no gradient. */
int qf_sg_features_to_rgb_backward(const float* d_features, int64_t stride, int num_lobes, const float* d_dirs,
                                   int64_t M, const float* d_grad_rgb, float* d_grad_features, int64_t grad_stride,
                                   void* stream);
/* utils.py:1055-1063: hit points + triangle ids -> texel (M,2) int64; uv_scaled (V,2) fp32 */
int qf_hit_texels(const qf_mesh* mesh, const float* d_points, const int64_t* d_index_tri, int64_t M,
                  const float* d_uv_scaled, int texture_size, int64_t* d_texels, void* stream);

/* ------------------------------------------------------------------------------------------
 * (5) Compositing.
 * ------------------------------------------------------------------------------------------ */
/* `utils.derive_properties` (utils.py:863-898) on ray-major hits described by offsets (n_rays+1):
 * rgb (N,3), alpha (N,1), depth (N,1), weights (M) [may be NULL].  d_bkgd (3) only for QF_BG_RANDOM. */
int qf_derive_properties(const float* d_color, const float* d_density, const float* d_depths, float delta,
                         const int64_t* d_offsets, int64_t n_rays, int bg_mode, const float* d_bkgd,
                         float* d_rgb, float* d_alpha, float* d_depth_out, float* d_weights, void* stream);

/* autograd of derive_properties w.r.t. color (M,3) and density (M); grad inputs (N,3)/(N,1)/(N,1) may be NULL */
int qf_derive_properties_backward(const float* d_color, const float* d_density, const float* d_depths, float delta,
                                  const int64_t* d_offsets, int64_t n_rays, int64_t n_hits, int bg_mode, const float* d_bkgd,
                                  const float* d_grad_rgb, const float* d_grad_alpha, const float* d_grad_depth,
                                  float* d_grad_color, float* d_grad_density, void* stream);

/* nerfacc-style segmented scans behind field_rendering.py (exclusive_prod / exclusive_sum, :203,:261).
 * packed_info (n_rays,2) int64 [start,count].  mode 0: from alphas; 1: from sigmas*(t_ends-t_starts).
 * Writes weights, trans, alphas (any may be NULL).  Warp-per-ray segmented scan. */
int qf_render_weights(int mode, const float* d_alphas_or_sigmas, const float* d_t_starts, const float* d_t_ends,
                      const int64_t* d_packed_info, int64_t n_rays, int64_t n_samples, const float* d_prefix_trans,
                      float* d_weights, float* d_trans, float* d_alphas_out, void* stream);
/* backward of the above w.r.t. alphas / sigmas given dL/dweights and dL/dtrans (either may be NULL) */
int qf_render_weights_backward(int mode, const float* d_alphas_or_sigmas, const float* d_t_starts,
                               const float* d_t_ends, const int64_t* d_packed_info, int64_t n_rays,
                               int64_t n_samples, const float* d_prefix_trans, const float* d_grad_weights,
                               const float* d_grad_trans, float* d_grad_in, void* stream);
/* `accumulate_along_rays` (field_rendering.py:483-547): out (n_rays,D) = sum_i w_i * v_i, deterministic
 * (segment order) when packed_info is given; values may be NULL (D=1). */
int qf_accumulate_along_rays(const float* d_weights, const float* d_values, int D, const int64_t* d_packed_info,
                             int64_t n_rays, float* d_out, int accumulate_into, void* stream);
/* same for arbitrary (unsorted) ray_indices — `index_add_` semantics (field_rendering.py:544,571), atomics;
 * d_out must be pre-initialised (zeros, or the tensor of accumulate_along_rays_). */
int qf_accumulate_along_rays_indexed(const float* d_weights, const float* d_values, int D, const int64_t* d_ray_indices,
                                     int64_t n_samples, float* d_out, void* stream);
/* autograd of accumulate_along_rays: grad_weights (M) and grad_values (M,D), either may be NULL */
int qf_accumulate_along_rays_backward(const float* d_weights, const float* d_values, int D, const int64_t* d_ray_indices,
                                      int64_t n_samples, const float* d_grad_out, float* d_grad_weights,
                                      float* d_grad_values, void* stream);
/* nerfacc pack.pack_info: counts by index, starts by exclusive scan (so a *descending* index array
 * reproduces quirk Q3).  workspace >= qf_pack_info_workspace_bytes(n_rays). */
size_t qf_pack_info_workspace_bytes(int64_t n_rays);
int qf_pack_info(const int64_t* d_ray_indices, int64_t n_samples, int64_t n_rays, int64_t* d_packed_info,
                 void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused frame render: rays -> image, everything resident (utils.py:465-607 with scaling=0, and
 * utils.py:998-1095).  Outputs rgb (N,3), alpha (N,1), depth (N,1); d_hits_total (1 x int32, may be NULL)
 * receives the number of hit samples shaded.  image_width > 0 declares the rays row-major image-ordered (the
 * reference passes (H,W,3) ray tensors, utils.py:489-495): warps then take 8x4 pixel tiles; 0 = plain ray list.
 * Results do not depend on it.
 * ------------------------------------------------------------------------------------------ */
size_t qf_render_workspace_bytes(int64_t n_rays, int K);
int qf_render_mesh_ngp(const qf_mesh* mesh, const qf_ngp* field, const float* d_origins, const float* d_viewdirs,
                       int64_t n_rays, int image_width, int K, float delta, int bg_mode, const float* d_bkgd, float* d_rgb,
                       float* d_alpha, float* d_depth, int32_t* d_hits_total, void* d_workspace,
                       size_t workspace_bytes, void* stream);
/* Baked variant.  d_uv_scaled (V,2): the per-vertex atlas coordinates already multiplied by the texture size
 * (test_baking_texture_images.py:325-328).  The library caches a 128-byte barycentric record per triangle for the uv array
 * of the last call: the array is taken to be IMMUTABLE while its address is unchanged (pass a new buffer after editing it;
 * vertex updates through qf_mesh_update_vertices invalidate the cache by themselves). */
int qf_render_mesh_baked(const qf_mesh* mesh, const qf_texture* tex, const float* d_uv_scaled,
                         const float* d_origins, const float* d_viewdirs, int64_t n_rays, int image_width, int K, float delta,
                         int bg_mode, const float* d_bkgd, float* d_rgb, float* d_alpha, float* d_depth,
                         int32_t* d_hits_total, void* d_workspace, size_t workspace_bytes, void* stream);

/* Ray-sharded frames (BASELINE configs[3] / [4]: one frame over 1/2/4/8 GPUs).  The rays are a rank's band-cyclic share of
 * a frame `frame_width` pixels wide (qf_generate_rays_banded with the same band_rows / band_stride / band_offset; n_rays =
 * qf_band_rows(...) * frame_width) and the three outputs address the WHOLE frame (H*W pixels): every pixel is stored where it
 * sits in the frame.  With the outputs in a buffer of the gathering rank (qf_peer_open below) the final image gather
 * is the composite kernel's own stores over NVLink — no collective, no reassembly pass.  The reference renders such frames
 * in 160 000-ray splits on one GPU and scatters them into the frame with rgb[split] = color[split]
 * (train_finetune.py:590-617, test_baking_texture_images.py:355-371). */
int qf_render_mesh_ngp_to_frame(const qf_mesh* mesh, const qf_ngp* field, const float* d_origins, const float* d_viewdirs,
                                int64_t n_rays, int K, float delta, int bg_mode, const float* d_bkgd, int band_rows,
                                int band_stride, int band_offset, int frame_width, float* d_frame_rgb, float* d_frame_alpha,
                                float* d_frame_depth, int32_t* d_hits_total, void* d_workspace, size_t workspace_bytes,
                                void* stream);
int qf_render_mesh_baked_to_frame(const qf_mesh* mesh, const qf_texture* tex, const float* d_uv_scaled,
                                  const float* d_origins, const float* d_viewdirs, int64_t n_rays, int K, float delta,
                                  int bg_mode, const float* d_bkgd, int band_rows, int band_stride, int band_offset,
                                  int frame_width, float* d_frame_rgb, float* d_frame_alpha, float* d_frame_depth,
                                  int32_t* d_hits_total, void* d_workspace, size_t workspace_bytes, void* stream);
/* The PNG-ready images of the reference's eval loops (train_finetune.py:639-646: rgb8 = uint8(clamp(rgb,0,1)*255),
 * depth8 = uint8(depth / depth.max() * 255), fp32 arithmetic, truncation), made on the device: 4 instead of 16 bytes per
 * pixel to copy to the host.  d_depth8 may be NULL (colour only); d_scratch: 4 bytes of device memory. */
int qf_frame_to_u8(const float* d_rgb, const float* d_depth, int64_t n_pixels, unsigned char* d_rgb8,
                   unsigned char* d_depth8, int32_t* d_scratch, void* stream);
/* Peer memory for the ray-sharded frames (one process per GPU of one node): device memory allocated by one rank (cudaMalloc), exported
 * as a 64-byte CUDA IPC handle that travels to the peers over any host channel, and mapped there (peer access over NVLink is
 * enabled on first open).  A mapping is closed by the rank that opened it, the allocation freed by its owner afterwards. */
int qf_peer_alloc(size_t bytes, void** d_ptr);
int qf_peer_free(void* d_ptr);
int qf_peer_export(const void* d_ptr, unsigned char* handle64);
int qf_peer_open(const unsigned char* handle64, void** d_ptr);
int qf_peer_close(void* d_ptr);

/* ------------------------------------------------------------------------------------------
 * (5b) The quadrature `Field` net (field.py:130-259) with back_prop=False (both reference
 * call sites): out = MLP([x01, grid(x01)]) with a 16-level fp16 hash grid (tcnn.Encoding "Grid"/"Hash") and a
 * torch fp32 BasicDecoder (2 hidden layers of 16 or 32 units, ELU or ReLU, output_dim <= 3);
 * field_grad = d(sum out)/dx through the raw-xyz inputs of the MLP only (field.py:196-199, 229-238).
 * Weights are torch nn.Linear layouts: w1 (H,35), w2 (H,H), w3 (out_dim,H); biases may be NULL.
 * ------------------------------------------------------------------------------------------ */
enum { QF_ACT_ELU = 0, QF_ACT_RELU = 1 };
typedef struct {
  int32_t hidden;      /* 16 or 32 */
  int32_t out_dim;     /* 1..3 */
  int32_t activation;  /* QF_ACT_* */
  float xyz_min[3];    /* field.py:141-142: -scale, +scale */
  float xyz_max[3];
} qf_field_desc;
/* A grid-only field handle (no NGP MLPs): usable with qf_hashgrid_forward/backward and qf_field_*; free with qf_ngp_destroy,
 * refresh the table with qf_ngp_update(f, d_table, NULL, NULL, stream). */
int qf_grid_create(const qf_grid_desc* desc, const float* d_table, int64_t n_entries, void* stream, qf_ngp** out);
/* Field.forward (field.py:203-221): field (M,out_dim); field_grad (M,3) or NULL (return_grad=False). */
int qf_field_forward(const qf_ngp* grid, const qf_field_desc* fd, const float* d_w1, const float* d_b1,
                     const float* d_w2, const float* d_b2, const float* d_w3, const float* d_b3, const float* d_x,
                     int64_t M, float* d_field, float* d_field_grad, void* stream);
/* Backward of BOTH outputs (the field_grad part is the reference's double backward through the torch MLP plus the
 * first-order grid backward).  Upstream gradients may be NULL (= zero); parameter gradients are ACCUMULATED
 * (d_grad_table (n_entries,2) fp32 may be NULL to skip the grid; bias gradients may be NULL). */
int qf_field_backward(const qf_ngp* grid, const qf_field_desc* fd, const float* d_w1, const float* d_b1,
                      const float* d_w2, const float* d_b2, const float* d_w3, const float* d_b3, const float* d_x,
                      int64_t M, const float* d_grad_field, const float* d_grad_field_grad, float* d_grad_table,
                      float* d_grad_w1, float* d_grad_b1, float* d_grad_w2, float* d_grad_b2, float* d_grad_w3,
                      float* d_grad_b3, void* stream);

/* ------------------------------------------------------------------------------------------
 * (6) Mesh finetuning accumulators: replace the torch_scatter calls of
 * MeshFinetune (mesh_utils.py:112-156) and of the prune pass (prune_mesh_after_finetuning.py:354-357).
 * ------------------------------------------------------------------------------------------ */
/* MeshFinetune.update_d (mesh_utils.py:126-133): cache_d[tri] += d*w (F,3), cache_w[tri] += w (F,) over M samples. */
int qf_triangle_accumulate(const float* d_disp, const float* d_w, const int64_t* d_index_tri, int64_t M,
                           int64_t n_faces, float* d_cache_d, float* d_cache_w, void* stream);
/* MeshFinetune.update_faces (mesh_utils.py:135-144): per-triangle clip(cache_d / cache_w, +-scaling), scatter_mean over
 * the face corners onto the vertices, d_vertices (V,3) += mean (in place; pass the result to qf_mesh_update_vertices). */
size_t qf_vertex_displace_workspace_bytes(int64_t n_vertices);
int qf_vertex_displace(const float* d_cache_d, const float* d_cache_w, const int32_t* d_faces, int64_t n_faces,
                       int64_t n_vertices, float scaling, float* d_vertices, void* d_workspace, size_t workspace_bytes,
                       void* stream);
/* prune pass: tri_w[tri] = max(tri_w[tri], w) with the running maximum starting at 0
 * (scatter_max into zeros + torch.maximum, prune_mesh_after_finetuning.py:354-357); weights read `stride` floats apart. */
int qf_triangle_weight_max(const float* d_weights, int64_t stride, const int64_t* d_index_tri, int64_t M,
                           int64_t n_faces, float* d_tri_w, void* stream);

/* ------------------------------------------------------------------------------------------
 * (7) Occupancy-grid ray marcher: nerfacc 0.5.3 `OccGridEstimator.sampling` ->
 * `traverse_grids` as called at utils.py:137-148, 241-285, 422-433.  Two passes over the same march:
 * pass 0 writes the per-ray sample counts, the caller scans them (qf_hits_offsets) and allocates,
 * pass 1 writes (ray_indices int64, t_starts, t_ends) ray-major.  binaries: (levels, rx, ry, rz) bytes.
 * d_near_planes (N) overrides near_plane per ray (stratified sampling jitters it).
 * ------------------------------------------------------------------------------------------ */
#define QF_OCC_MAX_LEVELS 8
typedef struct {
  int32_t levels;
  int32_t resolution[3];
  float aabbs[QF_OCC_MAX_LEVELS][6]; /* level l: xmin ymin zmin xmax ymax zmax */
} qf_occgrid_desc;
int qf_occgrid_march(const qf_occgrid_desc* grid, const uint8_t* d_binaries, const float* d_origins,
                     const float* d_dirs, int64_t n_rays, const float* d_near_planes, float near_plane,
                     float far_plane, float step_size, float cone_angle, int pass, int32_t* d_counts,
                     const int64_t* d_offsets, int64_t* d_ray_indices, float* d_t_starts, float* d_t_ends,
                     void* stream);
/* The chunked marcher of `render_image_with_occgrid_test` (utils.py:175-350; nerfacc `traverse_grids` with a step limit):
 * a ray emits at most max_samples samples (0 = no limit) and reports in d_termination (N, may be NULL) the plane where it
 * stopped — the near plane of the caller's next round; rays with d_ray_mask[i] == 0 (N bytes, NULL = all alive) emit
 * nothing and keep their plane. */
int qf_occgrid_march_limited(const qf_occgrid_desc* grid, const uint8_t* d_binaries, const float* d_origins,
                             const float* d_dirs, int64_t n_rays, const float* d_near_planes, float near_plane,
                             float far_plane, float step_size, float cone_angle, int pass, int32_t* d_counts,
                             const int64_t* d_offsets, int64_t* d_ray_indices, float* d_t_starts, float* d_t_ends,
                             int max_samples, const uint8_t* d_ray_mask, float* d_termination, void* stream);

/* Optional per-stage CUDA-event timing of the fused render on its own stream (off by default).
 * qf_profile_read sums {trace, shade, composite} milliseconds recorded since the last read (synchronises). */
int qf_profile_enable(int on);
int qf_profile_read(double* ms3, int64_t* n_chunks);

/* a1: eval-mode pinhole rays of one camera (datasets/nerf_synthetic.py:310-360), c2w (3,4) row-major on host */
int qf_generate_rays(const float* h_c2w, int W, int H, float focal, float cx, float cy, int opengl,
                     float* d_origins, float* d_viewdirs, void* stream);
/* The same for the band-cyclic share of one rank of a ray-sharded frame: the image is cut into bands of band_rows rows
 * (H % band_rows == 0) and the rays of the bands b with b % band_stride == band_offset are written compactly, in order
 * (qf_band_rows(...) rows of W rays).  Bands dealt round-robin balance the ranks: hit counts vary a lot over an image. */
int64_t qf_band_rows(int H, int band_rows, int band_stride, int band_offset);
int qf_generate_rays_banded(const float* c2w_3x4, int W, int H, float focal, float cx, float cy, int opengl, int band_rows,
                            int band_stride, int band_offset, float* d_origins, float* d_viewdirs, void* stream);
/* a1, training branch (datasets/nerf_synthetic.py:293-309, 341-370): ray i looks through pixel (x[i], y[i]) of camera
 * image_id[i] (NULL = camera 0 for all).  c2w (n_views,3,4) row-major ON THE DEVICE; x, y fp32 (pixel indices, or index +
 * U[0,1) for add_ray_direction_noise).  An id outside [0, n_views) yields a NaN ray. */
int qf_generate_rays_indexed(const float* d_c2w, int64_t n_views, const int64_t* d_image_id, const float* d_x,
                             const float* d_y, int64_t n, float focal, float cx, float cy, int opengl,
                             float* d_origins, float* d_viewdirs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUADFIELD_H */
