// Internal launcher interfaces between the translation units of libmassb200.
#pragma once
#include "common.cuh"

// device counters at the start of the workspace, zeroed by K1 of every call
constexpr int MB_CNT_HEADS = 0;      // per-frame path: number of voxel segments (= voxels touched by the frame)
constexpr int MB_CNT_NVALID = 1;     // batched path: valid pixels (sorted positions before the invalid tail)
constexpr int MB_CNT_CELLS = 2;      // batched path: distinct cells
constexpr int MB_CNT_SEGS = 3;       // batched path: (cell, frame) segments
constexpr int MB_CNT_RUNS = 4;       // batched path: accumulate runs
constexpr int MB_CNT_ERROR = 5;      // sticky error bits (read by mb_layer_update_status)
constexpr int MB_CNT_VOX = 6;        // batched path: touched voxels
constexpr int MB_CNT_VOXFRAMES = 7;  // batched path: (touched voxel, frame) pairs = sum over the chunk's frames of U_f
constexpr int MB_CNT_TASKQ = 8;      // batched path: work-queue heads: +0/+1 accumulate (used in turn by its launches), +2 voxel scalars
constexpr int MB_CNT_LONGSEGS = 11; // batched path: segments of more than SEG_LONG_ITEMS items, summed by whole warps
constexpr int MB_NUM_COUNTERS = 16;
constexpr int MB_MAX_CHUNK_FRAMES = 1024;   // frames fused per batched chunk (per-voxel frame table in shared memory)

// How a contribution's point id maps to its feature row.
//   dense:   row = point id (upsample == 0), or the nearest-upsampled source pixel
//   one-hot: class_ids[point id] is the only non-zero channel
struct MbFeatIndex {
    uint32_t np;        // points per slot (contribution id = slot * np + point)
    uint32_t W;         // camera width (for the up-sampling map)
    uint32_t ky, kx;    // integer up-sampling factors (1 = none)
    uint32_t fw;        // feature image width
};

// voxelise.cu
int mbk_transform_rays(cudaStream_t stream, const float *rays, int64_t npix, const float *pose, float *out);
int mbk_bin_flags(cudaStream_t stream, const float *bins0, int n0, const float *bins1, int n1,
                  const float *bins2, int n2, const float *origin, const float *rays, const float *depth,
                  int64_t npix, float min_d, float max_d, uint32_t *flags);
int mbk_bin_write(cudaStream_t stream, const float *bins0, int n0, const float *bins1, int n1,
                  const float *bins2, int n2, const float *origin, const float *rays, const float *depth,
                  int64_t npix, float min_d, float max_d, const uint32_t *offsets, int64_t *ind0,
                  int64_t *ind1, int64_t *ind2, float *ratio0, float *ratio1, float *ratio2, int64_t *pix,
                  int64_t *count);
int mbk_unproject_voxelise(cudaStream_t stream, const float *rays, const float *depth, const float *pose,
                           uint32_t npix, const float *bins_x, int nx, const float *bins_y, int ny,
                           const float *bins_z, int nz, const MbGrid &g, float min_d, float max_d,
                           uint32_t *keys, float4 *pt_ratio, uint32_t *counters);
int mbk_points_to_keys(cudaStream_t stream, const int64_t *ind0, const int64_t *ind1, const int64_t *ind2,
                       const float *ratio0, const float *ratio1, const float *ratio2, uint32_t npts,
                       const MbGrid &g, uint32_t *keys, float4 *pt_ratio, uint32_t *counters);

// voxel_reduce.cu
int mbk_segment_heads(cudaStream_t stream, const uint32_t *keys, uint32_t n, const MbGrid &g,
                      uint32_t *heads, uint32_t *counters);
int mbk_voxel_reduce(cudaStream_t stream, const uint32_t *keys, const uint32_t *vals, uint32_t n,
                     const uint32_t *heads, const uint32_t *counters, const float4 *pt_ratio,
                     const MbFeatIndex &fi, const float *features, const int64_t *class_ids, int F,
                     float *map, const MbGrid &g, float alpha, int mode);

// cells.cu: batched cell-sorted pipeline (affine form of the update; see the file header)
int mbk_batch_frames_that_fit(int H, int W, int nx, int ny, int nz, int F, size_t workspace_bytes, int T);
int mbk_batch_max_chunk_frames(int H, int W);
size_t mbk_batch_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F);
size_t mbk_batch_min_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F);
// fold target of frame-sharded scenes: a sparse partial instead of the map (cells.cu, "sparse partials")
struct MbSparseFold {
    int32_t *slot_table;        // [voxels] row of a voxel, -1 = none yet
    void *buffer;               // {count | index | a | b}, mbk_partial_buffer_layout
    uint32_t capacity;          // rows the buffer holds
};
int mbk_batch_update(cudaStream_t stream, const float *rays, const float *depth, const float *features,
                     const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                     const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                     float *map, float *affine_a, float alpha, float min_d, float max_d, void *workspace,
                     size_t workspace_bytes, const MbSparseFold *sparse = nullptr, bool keep_error_bits = false);
size_t mbk_partial_buffer_layout(uint32_t capacity, int F, size_t *offsets);
int mbk_partial_reset(cudaStream_t stream, int32_t *slot_table, int64_t voxels, void *buffer);
int mbk_partial_clear(cudaStream_t stream, int32_t *slot_table, void *buffer, uint32_t capacity, int F);
int mbk_affine_apply_partial(cudaStream_t stream, float *map, int F, const void *buffer, uint32_t capacity);
int mbk_partial_pull(cudaStream_t stream, const void *const *peer_buffers_host, void *const *slots_host, int world,
                     int self, uint32_t capacity, int F);

int mbk_affine_apply_rows(cudaStream_t stream, float *map, int F, const int64_t *idx, const float *a, const float *b,
                          int64_t n);
int mbk_profile_enable(int enable);
int mbk_profile_read(float *ms_host, int capacity);

// instances.cu: instance extraction (find) and matching
size_t mbk_class_presence_workspace_bytes(int S0, int S1, int S2, int pad);
int mbk_class_presence(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int c, int pad, float thr,
                       uint8_t *image, void *workspace, size_t workspace_bytes);
int mbk_column_summary(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float thr,
                       float *amax, uint8_t *blocked);
int mbk_masks_to_ids(cudaStream_t stream, const uint8_t *masks, const int64_t *classes, const float *scores, int n,
                     size_t npix, int num_classes, float threshold, int64_t *ids);
int mbk_top_down(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float *out);
int mbk_instance_pool(cudaStream_t stream, const int *boxes, int nboxes, const float *sem, int S0, int S1, int S2, int F,
                      int c, const float *feat, int FF, const float *mx, const float *my, const float *mz, float *out);
int mbk_pairwise_l2(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, float *out);
int mbk_cosine_best_match(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                          float *best_sim);
size_t mbk_lsap_workspace_bytes(int n, int m);
int mbk_lsap(cudaStream_t stream, const float *cost32, const double *cost64, int n, int m, int64_t *rows, int64_t *cols,
             int *status, void *workspace, size_t workspace_bytes);

// navigation.cu: coordinate transforms and navigation-graph tests next to the path
int mbk_world_to_map(cudaStream_t stream, const float *coords, int64_t n, int k, const float *bx, int nx, const float *by,
                     int ny, const float *bz, int nz, int64_t *out);
int mbk_map_to_world(cudaStream_t stream, const float *coords, int64_t n, int k, const float *bx, int nx, const float *by,
                     int ny, const float *bz, int nz, float *out);
int mbk_navigable_area(cudaStream_t stream, const uint8_t *blocked, int S0, int S1, int padding, float *out);
int mbk_nav_edges(cudaStream_t stream, const float *navigable, int S0, int S1, int off_y, int off_x, int step, int ny,
                  int nx, uint8_t *node_ok, uint8_t *edge_ok);
int mbk_nav_rects(cudaStream_t stream, const float *navigable, int S1, const int32_t *rects, int m, uint8_t *clear);

// tc_match.cu: cosine similarity + best match on tcgen05 tensor cores (large instance matrices)
size_t mbk_cosine_tc_workspace_bytes(int n, int m, int d);
int mbk_cosine_best_match_tc(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                             float *best_sim, void *workspace, size_t workspace_bytes);
