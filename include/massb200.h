/*
 * massb200.h -- C ABI of libmassb200.so: hand-written sm_100a CUDA kernels for the
 * mapping-and-matching hot path of brandontrabucco/mass (MaSS).
 *
 * The reference has no FFI: its boundary for this path is a Python call surface
 * (SURVEY.md 8b).  Each entry point below names the reference function it
 * replaces (paths relative to /root/reference); mass_b200/ binds them with ctypes
 * behind classes/functions that carry the reference's names and arguments.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; mb_last_error()
 *     returns a thread-local message for the last failure on the calling thread;
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued asynchronously on it, nothing synchronises unless stated;
 *   - the library owns no memory: scratch space is a caller-provided workspace whose
 *     size comes from the matching *_workspace_bytes() function;
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 *
 * Map layout (reference: mass/nn/base_projection_layer.py:158-160, 339):
 *   float map[S0 = map_height (y, flipped)][S1 = map_width (x)][S2 = map_depth (z)][F], F contiguous.
 * Pose layout: 12 floats per frame = rotation R row-major (R[i][0] = (eye x up)[i],
 *   R[i][1] = up[i], R[i][2] = -eye[i]; mass/utils/projection.py:104-105) followed by
 *   the camera position (x, y, z).  R is produced on the host with the reference's own
 *   ATen ops so that it is bit-identical to the reference CPU path.
 */
#ifndef MASSB200_H
#define MASSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK 0
#define MB_ERR_ARG 1
#define MB_ERR_CUDA 2
#define MB_ERR_WORKSPACE 3

/* arithmetic mode of the voxel reduce */
#define MB_MODE_EXACT 0 /* reference operation order, no FMA: map values bitwise equal to the CPU path;
                           frames are processed one at a time (sort by voxel + one warp per voxel segment)  */
#define MB_MODE_FAST 1  /* per-voxel affine form new = a*old + b: <= 1e-5 relative, occupancy
                           still bit-exact; all frames of the call go through the batched cell pipeline     */

const char *mb_last_error(void);
int mb_version(void);
/* number of kernels this library has launched in the calling process (bench.py: gpu_launches) */
uint64_t mb_launch_count(void);

/* ---- a3: mass/utils/projection.py:77-110 transform_rays ----------------------------------
 * out[p][i] = (rays[p][0]*R[i][0] + rays[p][1]*R[i][1]) + rays[p][2]*R[i][2], each product and sum
 * rounded separately.  pose points at 12 floats (only R is read). */
int mb_transform_rays(void *stream, const float *rays, int64_t npix, const float *pose, float *out);

/* ---- a4: mass/utils/projection.py:113-230 bin_rays -----------------------------------------
 * rays are ORIENTED rays [npix][3]; origin = 3 floats.  Valid pixels are compacted in row-major
 * order.  Outputs sized npix; *count (device int64) receives N.  pix receives flat pixel ids
 * (the reference's `indices`).  Axis 1 is flipped exactly as the reference does. */
size_t mb_bin_rays_workspace_bytes(int64_t npix);
int mb_bin_rays(void *stream, const float *bins0, int n0, const float *bins1, int n1,
                const float *bins2, int n2, const float *origin, const float *rays,
                const float *depth, int64_t npix, float min_ray_depth, float max_ray_depth,
                int64_t *ind0, int64_t *ind1, int64_t *ind2, float *ratio0, float *ratio1,
                float *ratio2, int64_t *pix, int64_t *count, void *workspace, size_t workspace_bytes);

/* ---- a5: mass/utils/projection.py:233-351 update_feature_map -------------------------------
 * Trilinear 8-neighbour splat + per-voxel weighted-average recurrence, in place on `map`
 * [S0][S1][S2][F].  features is [npts][F].  Deterministic: contributions are radix-sorted by
 * voxel key (stable, so the reference's slot-major/point order is kept) and reduced per voxel
 * segment by one warp; no float atomics. */
size_t mb_update_feature_map_workspace_bytes(int64_t npts, int S0, int S1, int S2);
int mb_update_feature_map(void *stream, const int64_t *ind0, const int64_t *ind1, const int64_t *ind2,
                          const float *ratio0, const float *ratio1, const float *ratio2,
                          const float *features, int64_t npts, int F, float *map, int S0, int S1,
                          int S2, float interpolation_weight, int mode, void *workspace,
                          size_t workspace_bytes);

/* ---- a6..a9: BaseProjectionLayer.update (mass/nn/base_projection_layer.py:282-343) ---------
 * Fused unproject + voxelise + deterministic voxel reduce for T consecutive frames, applied in
 * frame order (frames do not commute).  MB_MODE_FAST composes the T per-frame affine
 * updates of every voxel and touches each map row once; nothing synchronises and
 * nothing is allocated, so the call can be captured in a CUDA graph.
 *   rays      [H*W][3]  camera-frame ray table (the layer's `rays` buffer)
 *   depth     [T][H*W]
 *   features  [T][fh*fw][F], nearest up-sampled to H x W by integer factors H/fh, W/fw
 *             (repeat_interleave of lines 322-325); or NULL with class_ids != NULL
 *   class_ids [T][H*W] int64 semantic ids: one-hot features without materialising them
 *             (mass/nn/applications/semantic_projection_layer.py:203-214); else NULL
 *   pose      [T][12]
 *   bins_x/y/z edge tables with nx/ny/nz entries (map is [ny-1][nx-1][nz-1][F]) */
size_t mb_layer_update_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F, int mode);
/* smallest workspace with which T frames still go through as ONE chunk (MB_MODE_FAST: the feature pass may
 * then take several rounds; 0 if T frames can never form one chunk).  With less, mb_layer_update splits the
 * call into chunks of fewer frames; with less than the T = 1 size it fails. */
size_t mb_layer_update_min_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F, int mode);
int mb_layer_update(void *stream, const float *rays, const float *depth, const float *features,
                    const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw,
                    int F, const float *bins_x, int nx, const float *bins_y, int ny,
                    const float *bins_z, int nz, float *map, float interpolation_weight,
                    float min_ray_depth, float max_ray_depth, int mode, void *workspace,
                    size_t workspace_bytes);

/* ---- frame-sharded scenes: ordered affine combine of partial maps (no counterpart in the reference, which is
 * single-GPU; the per-voxel update it composes is mass/utils/projection.py:335-351) ----
 * mb_layer_fold: like mb_layer_update in MB_MODE_FAST, but also folds the frames' per-voxel coefficients into
 *   partial_a [S0*S1*S2]: entries equal to 2.0f mark untouched voxels on entry and stay 2.0f if the call does
 *   not touch them; a touched voxel ends as the product of its frames' a (times its previous value unless that
 *   was 2.0f).  With partial_b zeroed and partial_a filled with 2.0f beforehand, the T frames act on any map M
 *   as  M[v] <- partial_a[v] * M[v] + partial_b[v]  for every touched v.
 * mb_affine_apply_rows: map[voxel_index[i]] = a[i] * map[voxel_index[i]] + b[i] (rows of F floats, distinct
 *   indices), the in-order application of one rank's partial. */
int mb_layer_fold(void *stream, const float *rays, const float *depth, const float *features,
                  const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                  const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                  float *partial_b, float *partial_a, float interpolation_weight, float min_ray_depth,
                  float max_ray_depth, void *workspace, size_t workspace_bytes);
int mb_affine_apply_rows(void *stream, float *map, int F, const int64_t *voxel_index, const float *a,
                         const float *b, int64_t n);

/* Sparse partials: the same fold without a second dense map.  A partial buffer holds
 *   {count u32 | index int64[capacity] | a f32[capacity] | b f32[capacity][F]}   (offsets: mb_partial_buffer_layout,
 *   in that order) -- ONE allocation, so that a peer GPU can map it and read it over NVLink;
 * slot_table int32[S0*S1*S2] (-1 = no row yet) finds a voxel's row while the chunks of a rank are folded in, in frame
 * order: mb_layer_fold_sparse composes each chunk onto the rows (a <- a_chunk * a, b <- a_chunk * b + b_chunk).
 * mb_partial_reset empties table and buffer with memsets (first use); mb_partial_clear empties them by walking the
 * rows in use.  mb_affine_apply_partial: map[index[i]] = a[i] * map[index[i]] + b[i] for i < count, with count read
 * ON THE DEVICE (no host round trip) and `partial_buffer` allowed to live on a peer GPU: transfer and application
 * are then one kernel.  A full partial sets error bit 2 (value 4) of mb_layer_update_status and drops the row. */
size_t mb_partial_buffer_bytes(uint32_t capacity, int F);
int mb_partial_buffer_layout(uint32_t capacity, int F, size_t *offsets_host /* [4]: count, index, a, b */);
int mb_partial_reset(void *stream, int32_t *slot_table, int64_t voxels, void *partial_buffer);
int mb_partial_clear(void *stream, int32_t *slot_table, void *partial_buffer, uint32_t capacity, int F);
int mb_layer_fold_sparse(void *stream, const float *rays, const float *depth, const float *features,
                         const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                         const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                         int32_t *slot_table, void *partial_buffer, uint32_t capacity, float interpolation_weight,
                         float min_ray_depth, float max_ray_depth, void *workspace, size_t workspace_bytes);
int mb_affine_apply_partial(void *stream, float *map, int F, const void *partial_buffer, uint32_t capacity);
/* All peers' partials into local staging slots (each mb_partial_buffer_bytes large, same layout) in ONE kernel that
 * reads from every peer at once -- rank `self` itself is skipped -- with the row counts read on the device.
 * peer_buffers_host / staging_slots_host: HOST arrays of `world` (<= 16) device pointers, indexed by rank. */
int mb_partial_pull(void *stream, const void *const *peer_buffers_host, void *const *staging_slots_host, int world,
                    int self, uint32_t capacity, int F);

/* Peer memory for partial buffers (CUDA IPC: another process of the box -- another GPU over NVLink -- maps the
 * allocation).  The one exception to "the library owns no memory": an IPC handle names a whole allocation. */
int mb_peer_alloc(size_t bytes, void **ptr_host);
int mb_peer_free(void *ptr);
int mb_peer_export(const void *ptr, void *handle64_host /* 64 bytes */);
int mb_peer_open(const void *handle64_host, void **ptr_host);
int mb_peer_close(void *ptr);

/* ---- a11: SemanticProjectionLayer.find (mass/nn/applications/semantic_projection_layer.py:257-362) ----
 * Step 1, lines 309-317: image[y][x] = any over z of (box mean of map[..., category] with kernel
 * 2*contour_padding+1, zero padded, divisor k^3) > contour_threshold; uint8 [S0][S1].
 * Step 2 (host, as in the reference): cv2.findContours + cv2.boundingRect on that image.
 * Step 3, lines 329-357: per bounding box (x, y, w, h) over the full map depth,
 *   out[box] = {confidence, coord_x, coord_y, coord_z, size, feature[FF]} with
 *   weights = mask / (sum(mask) + 1e-9), confidence = sum(mask*weights), coord = sum(centre*weights),
 *   size = sum(mask), feature = sum(feat_map[box]*weights) (feat_map may be NULL, then FF is ignored and
 *   rows are 5 floats).  centres_* are the per-axis cell-centre tables (y already flipped).
 *   semantic_category == -1: boxes are 5 ints (x, y, w, h, class): the boxes of all classes in one launch. */
size_t mb_class_presence_workspace_bytes(int S0, int S1, int S2, int contour_padding);
int mb_class_presence(void *stream, const float *map, int S0, int S1, int S2, int F, int semantic_category,
                      int contour_padding, float contour_threshold, uint8_t *image, void *workspace,
                      size_t workspace_bytes);
int mb_instance_pool(void *stream, const int32_t *boxes, int nboxes, const float *sem_map, int S0, int S1, int S2,
                     int F, int semantic_category, const float *feat_map, int FF, const float *centres_x,
                     const float *centres_y, const float *centres_z, float *out);

/* ---- next to the path: the two readers of the whole map ----------------------------
 * amax [S0][S1][F]    = max over z of map[y][x][z][f]   (agent.py:330-331, 391-392: data.amax(dim=2), the input of
 *                       the semantic search policy); may be NULL
 * blocked [S0][S1] u8 = any z in [z_lo, z_hi): sum_f |map[y][x][z][f]| > obstacle_threshold
 *                       (mass/navigation_policy.py:207-216: torch.norm(data, p=1, dim=3) > thr, sliced, any(dim=2));
 *                       may be NULL.  Exact for the reference's default threshold 0; for other thresholds the
 *                       fp32 summation order over f may differ from ATen's.
 * One pass over the map. */
int mb_column_summary(void *stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi,
                      float obstacle_threshold, float *amax, uint8_t *blocked);

/* Perception hand-off (mass/thor/segmentation_config.py:314-334, the Mask R-CNN branch of
 * SemanticRearrangeSensor.get_segmentation): masks [n][npix] (non-zero = inside), classes [n], scores [n] ->
 * ids [npix] = arg-max over classes of the number of instances of that class (score >= detection_threshold) whose
 * mask covers the pixel; first maximum, so 0 where nothing was detected.  The result feeds mb_layer_update's
 * class_ids directly: no [H][W][54] buffer, no device-host round trip. */
int mb_masks_to_ids(void *stream, const uint8_t *masks, const int64_t *classes, const float *scores, int n,
                    int64_t npix, int num_classes, float detection_threshold, int64_t *ids);

/* Top-down rendering (mass/nn/base_projection_layer.py:345-379, BaseProjectionLayer.top_down):
 * out [S0][S1][F] = the feature row of the top-most voxel of [z_lo, z_hi) with any non-zero channel, zeros if the
 * column is empty in the slice.  A pure selection: bit-exact. */
int mb_top_down(void *stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float *out);

/* ---- next to the path: coordinate transforms and the navigation graph's traversability tests ----
 * mb_world_to_map  mass/nn/base_projection_layer.py:513-547 (with clamp_to_world, :381-413): coords [n][k] (k = 2 or 3,
 *                  xyz order) -> int64 map cells [n][k]; clamp to the span of the voxel mid-points,
 *                  bucketize(right=True) - 1, y index flipped.
 * mb_map_to_world  mass/nn/base_projection_layer.py:452-511 (with clamp_to_map, :415-450): float map coordinates
 *                  [n][k] -> world [n][k]: left + (right - left) * frac between neighbouring voxel mid-points, every
 *                  operation rounded on its own as the ATen CPU ops are.
 * mb_navigable_area  mass/navigation_policy.py:219-221: 1 - max_pool2d(blocked, 2*padding+1, stride 1, padding) as float
 *                  [S0][S1] from the uint8 `blocked` image mb_column_summary produces.
 * mb_nav_graph_lattice  mass/navigation_policy.py:253-285 (reset_navigation_graph): nodes at rows offset_y + a*step,
 *                  columns offset_x + b*step; node_ok [ny][nx] (the node's cell is navigable), edge_ok [ny][nx][2]
 *                  (the segment to the next node down / right lies inside the map and every cell of it == 1);
 *                  ny = ceil((S0 - offset_y) / step), nx likewise.
 * mb_nav_rects_clear  mass/navigation_policy.py:315-341 (update_navigation_graph): for rectangles {row0,row1,col0,col1}
 *                  (inclusive), clear[i] = every cell == 1: the test of every node and edge of an existing graph. */
int mb_world_to_map(void *stream, const float *coords, int64_t n, int k, const float *bins_x, int nx,
                    const float *bins_y, int ny, const float *bins_z, int nz, int64_t *out);
int mb_map_to_world(void *stream, const float *coords, int64_t n, int k, const float *bins_x, int nx,
                    const float *bins_y, int ny, const float *bins_z, int nz, float *out);
int mb_navigable_area(void *stream, const uint8_t *blocked, int S0, int S1, int padding, float *navigable);
int mb_nav_graph_lattice(void *stream, const float *navigable, int S0, int S1, int offset_y, int offset_x,
                         int step_size, uint8_t *node_ok, uint8_t *edge_ok);
int mb_nav_rects_clear(void *stream, const float *navigable, int S0, int S1, const int32_t *rects, int m,
                       uint8_t *clear);

/* ---- a12: predict_scene_differences (mass/utils/experimentation.py:261-287) ------------------------------
 * mb_pairwise_l2: out[i][j] = ||a[i] - b[j]||_2 from direct differences (torch.linalg.norm of the
 *   broadcast difference), a [n][d], b [m][d], out [n][m].
 * mb_lsap: scipy.optimize.linear_sum_assignment on an n x m cost matrix (float32 as the reference hands
 *   it over, or float64): same row order, scan order and tie rule.  Writes
 *   min(n, m) (row, col) pairs sorted by row; *status = 1 if the matrix is infeasible. */
int mb_pairwise_l2(void *stream, const float *a, int n, const float *b, int m, int d, float *out);
/* Not on the reference's path (it matches by L2 distance + assignment); an additional op:
 * best[i] = argmax_j cos(a_i, b_j) (first maximum, torch.argmax's tie rule; -1 if m == 0), best_sim[i] = that
 * cosine; a zero vector has similarity 0 to everything. */
int mb_cosine_best_match(void *stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                         float *best_sim);
/* The same result (float64-exact decision, first maximum on ties) for LARGE instance matrices -- thousands of rows, a
 * real dense contraction -- on the tcgen05 tensor cores: A B^T in 3 x TF32 with fp32 accumulators in TMEM ranks the
 * candidates, every column within 2e-3 of a row's largest approximate cosine is re-evaluated exactly.  Workspace from
 * mb_cosine_best_match_tc_workspace_bytes.  (No reference counterpart: SURVEY.md F3.) */
size_t mb_cosine_best_match_tc_workspace_bytes(int n, int m, int d);
int mb_cosine_best_match_tc(void *stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                            float *best_sim, void *workspace, size_t workspace_bytes);
size_t mb_lsap_workspace_bytes(int n, int m);
int mb_lsap(void *stream, const float *cost32, const double *cost64, int n, int m, int64_t *rows, int64_t *cols,
            int32_t *status, void *workspace, size_t workspace_bytes);

/* Synchronises `stream` and returns the sticky error bits the batched kernels left in the workspace of
 * the last MB_MODE_FAST call (the bits of ALL internal chunks of that call: a later chunk does not clear what an
 * earlier one set): 0 = fine; bit 0 = more accumulate runs than the planned rounds hold (an
 * internal invariant: never expected, the map is then not trustworthy); bit 1 = a class id outside
 * [0, F) in `class_ids` (torch.nn.functional.one_hot raises on it in the reference,
 * mass/nn/applications/semantic_projection_layer.py:203-214; the kernel adds nothing for that pixel). */
int mb_layer_update_status(void *stream, const void *workspace, uint32_t *error_bits_host);

/* Synchronises `stream` and copies the counters the last MB_MODE_FAST chunk left in its workspace to
 * counters_host[0 .. min(capacity, 16)): measurement aid (bench.py derives SURVEY.md 8d's bytes per frame from
 * them instead of hard-coding it).  Indices: */
#define MB_COUNTER_ITEMS 1      /* (cell, tile) items sorted */
#define MB_COUNTER_CELLS 2      /* distinct cells */
#define MB_COUNTER_SEGMENTS 3   /* (cell, frame) segments */
#define MB_COUNTER_RUNS 4       /* accumulate runs */
#define MB_COUNTER_ERROR 5      /* the bits of mb_layer_update_status */
#define MB_COUNTER_VOXELS 6     /* distinct voxels touched by the chunk */
#define MB_COUNTER_VOXEL_FRAMES 7 /* (voxel, frame) pairs: the sum over the chunk's frames of U_f, the voxels one
                                     frame touches (the reference rewrites each of them once per frame,
                                     mass/utils/projection.py:335-351) */
int mb_layer_update_counters(void *stream, const void *workspace, uint32_t *counters_host, int capacity);

/* ---- measurement aid (bench.py) -------------------------------------------------------------------------
 * mb_profile_stages(1) makes MB_MODE_FAST calls record CUDA events on their stream between the stages of
 * the batched pipeline; mb_profile_read waits for the last recorded call and writes the duration in ms of
 * {voxelise, sort, index, scalar pass, feature accumulate (round 0), apply + later rounds}; returns the
 * number of values written (0 if nothing was recorded).  Off by default. */
int mb_profile_stages(int enable);
int mb_profile_read(float *ms_host, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* MASSB200_H */
