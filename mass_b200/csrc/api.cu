// extern "C" entry points of libmassb200.so (see include/massb200.h for the contract).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <string.h>
#include "common.cuh"
#include "kernels.cuh"

namespace {
thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void mb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void mb_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

MB_API const char *mb_last_error(void) { return g_error; }
MB_API int mb_version(void) { return 100; }
MB_API uint64_t mb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- a3 -------------------------------------------------------------------------------------
MB_API int mb_transform_rays(void *stream, const float *rays, int64_t npix, const float *pose, float *out)
{
    MB_REQUIRE(rays && pose && out && npix >= 0, "mb_transform_rays: null pointer or negative size");
    return mbk_transform_rays((cudaStream_t)stream, rays, npix, pose, out);
}

// ---- a4 -------------------------------------------------------------------------------------
MB_API size_t mb_bin_rays_workspace_bytes(int64_t npix)
{
    if (npix <= 0) return 256;
    return mb_align_up((size_t)npix * sizeof(uint32_t)) + mb_scan_workspace_bytes((uint32_t)npix) + 512;
}

MB_API int mb_bin_rays(void *stream_, const float *bins0, int n0, const float *bins1, int n1,
                       const float *bins2, int n2, const float *origin, const float *rays,
                       const float *depth, int64_t npix, float min_ray_depth, float max_ray_depth,
                       int64_t *ind0, int64_t *ind1, int64_t *ind2, float *ratio0, float *ratio1,
                       float *ratio2, int64_t *pix, int64_t *count, void *workspace, size_t workspace_bytes)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_REQUIRE(bins0 && bins1 && bins2 && origin && count, "mb_bin_rays: null pointer");
    MB_REQUIRE(n0 >= 2 && n1 >= 2 && n2 >= 2, "mb_bin_rays: every edge table needs >= 2 entries");
    MB_REQUIRE(npix >= 0 && npix < (int64_t)1 << 31, "mb_bin_rays: npix out of range");
    if (npix == 0) {
        MB_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t), stream));
        return MB_OK;
    }
    MB_REQUIRE(rays && depth && ind0 && ind1 && ind2 && ratio0 && ratio1 && ratio2 && pix,
               "mb_bin_rays: null pointer");
    MB_REQUIRE(workspace && workspace_bytes >= mb_bin_rays_workspace_bytes(npix),
               "mb_bin_rays: workspace too small (%zu < %zu)", workspace_bytes,
               mb_bin_rays_workspace_bytes(npix));
    MbArena arena(workspace, workspace_bytes);
    uint32_t *flags = arena.take<uint32_t>((size_t)npix);
    const size_t scan_bytes = mb_scan_workspace_bytes((uint32_t)npix);
    char *scan_ws = arena.take<char>(scan_bytes);
    int rc = mbk_bin_flags(stream, bins0, n0, bins1, n1, bins2, n2, origin, rays, depth, npix,
                           min_ray_depth, max_ray_depth, flags);
    if (rc) return rc;
    rc = mb_exclusive_scan_u32(stream, flags, flags, (uint32_t)npix, scan_ws, scan_bytes);
    if (rc) return rc;
    return mbk_bin_write(stream, bins0, n0, bins1, n1, bins2, n2, origin, rays, depth, npix, min_ray_depth,
                         max_ray_depth, flags, ind0, ind1, ind2, ratio0, ratio1, ratio2, pix, count);
}

// ---- shared back half: sort contributions by voxel, find segments, reduce -----------------------
namespace {

struct SplatBuffers {
    uint32_t *keys_a, *keys_b, *vals_a, *vals_b, *heads, *counters;
    float4 *pt_ratio;
    char *sort_ws;
    size_t sort_bytes;
};

size_t splat_workspace_bytes(uint32_t npts)
{
    MbArena a(nullptr, 0);
    const size_t n = (size_t)npts * 8;
    a.take<uint32_t>(n); a.take<uint32_t>(n); a.take<uint32_t>(n); a.take<uint32_t>(n);
    a.take<uint32_t>(n);
    a.take<uint32_t>(MB_NUM_COUNTERS);
    a.take<float4>(npts);
    a.take<char>(mb_sort_workspace_bytes((uint32_t)n));
    return a.used + 256;
}

bool carve(SplatBuffers &b, void *workspace, size_t bytes, uint32_t npts)
{
    MbArena a(workspace, bytes);
    const size_t n = (size_t)npts * 8;
    b.keys_a = a.take<uint32_t>(n);
    b.keys_b = a.take<uint32_t>(n);
    b.vals_a = a.take<uint32_t>(n);
    b.vals_b = a.take<uint32_t>(n);
    b.heads = a.take<uint32_t>(n);
    b.counters = a.take<uint32_t>(MB_NUM_COUNTERS);
    b.pt_ratio = a.take<float4>(npts);
    b.sort_bytes = mb_sort_workspace_bytes((uint32_t)n);
    b.sort_ws = a.take<char>(b.sort_bytes);
    return a.ok();
}

// keys_a / pt_ratio / counters have been produced by a K1 variant
int sort_and_reduce(cudaStream_t stream, SplatBuffers &b, uint32_t npts, const MbGrid &g,
                    const MbFeatIndex &fi, const float *features, const int64_t *class_ids, int F,
                    float *map, float alpha, int mode)
{
    const uint32_t n = npts * 8;
    uint32_t *keys, *vals;
    int rc = mb_sort_pairs(stream, b.keys_a, b.vals_a, b.keys_b, b.vals_b, n, nullptr, mb_key_bits(g), true,
                           b.sort_ws, b.sort_bytes, &keys, &vals);
    if (rc) return rc;
    rc = mbk_segment_heads(stream, keys, n, g, b.heads, b.counters);
    if (rc) return rc;
    return mbk_voxel_reduce(stream, keys, vals, n, b.heads, b.counters, b.pt_ratio, fi, features, class_ids,
                            F, map, g, alpha, mode);
}

}  // namespace

// ---- a5 -------------------------------------------------------------------------------------
MB_API size_t mb_update_feature_map_workspace_bytes(int64_t npts, int S0, int S1, int S2)
{
    (void)S0; (void)S1; (void)S2;
    if (npts <= 0) return 256;
    return splat_workspace_bytes((uint32_t)npts);
}

MB_API int mb_update_feature_map(void *stream_, const int64_t *ind0, const int64_t *ind1, const int64_t *ind2,
                                 const float *ratio0, const float *ratio1, const float *ratio2,
                                 const float *features, int64_t npts, int F, float *map, int S0, int S1,
                                 int S2, float interpolation_weight, int mode, void *workspace,
                                 size_t workspace_bytes)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_REQUIRE(npts >= 0 && npts < ((int64_t)1 << 28), "mb_update_feature_map: npts out of range");
    if (npts == 0) return MB_OK;
    MB_REQUIRE(ind0 && ind1 && ind2 && ratio0 && ratio1 && ratio2 && features && map,
               "mb_update_feature_map: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && S2 > 0 && F > 0, "mb_update_feature_map: bad map shape");
    MB_REQUIRE(mode == MB_MODE_EXACT || mode == MB_MODE_FAST, "mb_update_feature_map: bad mode %d", mode);
    const MbGrid g = mb_make_grid(S0, S1, S2);
    MB_REQUIRE((uint64_t)g.B0 * g.B1 * g.B2 * MB_BRICK_VOX < 0xffffffffull, "map too large for 32-bit keys");
    SplatBuffers b;
    MB_REQUIRE(workspace && carve(b, workspace, workspace_bytes, (uint32_t)npts),
               "mb_update_feature_map: workspace too small (%zu < %zu)", workspace_bytes,
               splat_workspace_bytes((uint32_t)npts));
    int rc = mbk_points_to_keys(stream, ind0, ind1, ind2, ratio0, ratio1, ratio2, (uint32_t)npts, g, b.keys_a,
                                b.pt_ratio, b.counters);
    if (rc) return rc;
    MbFeatIndex fi = { (uint32_t)npts, 1u, 1u, 1u, 1u };
    return sort_and_reduce(stream, b, (uint32_t)npts, g, fi, features, nullptr, F, map, interpolation_weight,
                           mode);
}

// Waits for `stream` and reports the sticky error bits the batched kernels left in the workspace
// of the last MB_MODE_FAST call (0 = fine; bit 0: more accumulate runs than the planned rounds hold -- an internal
// invariant; bit 1: a class id outside [0, F) in the class_ids image, where functional.one_hot would raise).
MB_API int mb_layer_update_status(void *stream_, const void *workspace, uint32_t *error_bits_host)
{
    MB_REQUIRE(workspace && error_bits_host, "mb_layer_update_status: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_CHECK_CUDA(cudaMemcpyAsync(error_bits_host, (const uint32_t *)workspace + MB_CNT_ERROR, sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, stream));
    MB_CHECK_CUDA(cudaStreamSynchronize(stream));
    return MB_OK;
}

// Waits for `stream` and copies the device counters the last MB_MODE_FAST chunk left in its workspace
// (MB_COUNTER_* indices of massb200.h) to the host: what bench.py derives its bytes-per-frame figure from.
MB_API int mb_layer_update_counters(void *stream_, const void *workspace, uint32_t *counters_host, int capacity)
{
    MB_REQUIRE(workspace && counters_host && capacity > 0, "mb_layer_update_counters: null pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = capacity < MB_NUM_COUNTERS ? capacity : MB_NUM_COUNTERS;
    MB_CHECK_CUDA(cudaMemcpyAsync(counters_host, workspace, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    MB_CHECK_CUDA(cudaStreamSynchronize(stream));
    return MB_OK;
}

// ---- a6..a9 ---------------------------------------------------------------------------------
// frames the batched path fuses per internal chunk when the caller sizes the workspace with
// mb_layer_update_workspace_bytes (a larger workspace is used if given)
static const int MB_DEFAULT_CHUNK_FRAMES = 512;

MB_API size_t mb_layer_update_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F, int mode)
{
    if (H <= 0 || W <= 0 || T <= 0 || F <= 0 || nx < 2 || ny < 2 || nz < 2) return 256;
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    if (mode == MB_MODE_EXACT) return splat_workspace_bytes(npix);
    int limit = MB_DEFAULT_CHUNK_FRAMES;
    if (const char *e = getenv("MASSB200_CHUNK_FRAMES")) {      // tuning aid
        const int v = atoi(e);
        if (v >= 1 && v <= MB_MAX_CHUNK_FRAMES) limit = v;
    }
    int chunk = T < limit ? T : limit;
    const int most = mbk_batch_max_chunk_frames(H, W);
    if (chunk > most) chunk = most;
    if (chunk < 1) return 0;
    return mbk_batch_workspace_bytes(H, W, nx, ny, nz, chunk, F);
}

MB_API size_t mb_layer_update_min_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F, int mode)
{
    if (H <= 0 || W <= 0 || T <= 0 || F <= 0 || nx < 2 || ny < 2 || nz < 2) return 256;
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    if (mode == MB_MODE_EXACT) return splat_workspace_bytes(npix);
    if (T > mbk_batch_max_chunk_frames(H, W)) return 0;          // never fits one chunk
    return mbk_batch_min_workspace_bytes(H, W, nx, ny, nz, T, F);
}

static int layer_update(void *stream_, const float *rays, const float *depth, const float *features,
                        const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw,
                        int F, const float *bins_x, int nx, const float *bins_y, int ny,
                        const float *bins_z, int nz, float *map, float *affine_a, float interpolation_weight,
                        float min_ray_depth, float max_ray_depth, int mode, void *workspace,
                        size_t workspace_bytes, const MbSparseFold *sparse = nullptr)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    MB_REQUIRE(T >= 0 && H > 0 && W > 0 && F > 0, "mb_layer_update: bad sizes");
    if (T == 0) return MB_OK;
    MB_REQUIRE(rays && depth && pose && bins_x && bins_y && bins_z && (map || sparse), "mb_layer_update: null pointer");
    MB_REQUIRE((features != nullptr) != (class_ids != nullptr),
               "mb_layer_update: pass exactly one of features / class_ids");
    MB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "mb_layer_update: every edge table needs >= 2 entries");
    MB_REQUIRE(mode == MB_MODE_EXACT || mode == MB_MODE_FAST, "mb_layer_update: bad mode %d", mode);
    MB_REQUIRE((int64_t)H * W < ((int64_t)1 << 28), "mb_layer_update: frame too large");
    if (features) {
        MB_REQUIRE(fh > 0 && fw > 0 && H % fh == 0 && W % fw == 0,
                   "mb_layer_update: feature image %dx%d does not divide the camera %dx%d", fh, fw, H, W);
    } else {
        fh = H; fw = W;
    }
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    const size_t feat_stride = (size_t)fh * fw * F;
    if (mode == MB_MODE_FAST) {
        // batched cell pipeline, as many frames per chunk as the workspace holds
        const int chunk = workspace ? mbk_batch_frames_that_fit(H, W, nx, ny, nz, F, workspace_bytes, T) : 0;
        MB_REQUIRE(chunk >= 1, "mb_layer_update: workspace too small (%zu < %zu)", workspace_bytes,
                   mbk_batch_workspace_bytes(H, W, nx, ny, nz, 1, F));
        for (int t = 0; t < T; t += chunk) {
            const int n = T - t < chunk ? T - t : chunk;
            int rc = mbk_batch_update(stream, rays, depth + (size_t)t * npix,
                                      features ? features + (size_t)t * feat_stride : nullptr,
                                      class_ids ? class_ids + (size_t)t * npix : nullptr, pose + (size_t)t * 12, n,
                                      H, W, fh, fw, F, bins_x, nx, bins_y, ny, bins_z, nz, map, affine_a,
                                      interpolation_weight, min_ray_depth, max_ray_depth, workspace, workspace_bytes,
                                      sparse, t > 0);      // (the error bits of a call's earlier chunks stay set)
            if (rc) return rc;
        }
        return MB_OK;
    }
    const MbGrid g = mb_make_grid(ny - 1, nx - 1, nz - 1);
    MB_REQUIRE((uint64_t)g.B0 * g.B1 * g.B2 * MB_BRICK_VOX < 0xffffffffull, "map too large for 32-bit keys");
    SplatBuffers b;
    MB_REQUIRE(workspace && carve(b, workspace, workspace_bytes, npix),
               "mb_layer_update: workspace too small (%zu < %zu)", workspace_bytes,
               splat_workspace_bytes(npix));
    MbFeatIndex fi = { npix, (uint32_t)W, (uint32_t)(H / fh), (uint32_t)(W / fw), (uint32_t)fw };
    for (int t = 0; t < T; ++t) {
        int rc = mbk_unproject_voxelise(stream, rays, depth + (size_t)t * npix, pose + (size_t)t * 12, npix,
                                        bins_x, nx, bins_y, ny, bins_z, nz, g, min_ray_depth, max_ray_depth,
                                        b.keys_a, b.pt_ratio, b.counters);
        if (rc) return rc;
        rc = sort_and_reduce(stream, b, npix, g, fi, features ? features + (size_t)t * feat_stride : nullptr,
                             class_ids ? class_ids + (size_t)t * npix : nullptr, F, map,
                             interpolation_weight, mode);
        if (rc) return rc;
    }
    return MB_OK;
}

MB_API int mb_layer_update(void *stream, const float *rays, const float *depth, const float *features,
                           const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw,
                           int F, const float *bins_x, int nx, const float *bins_y, int ny,
                           const float *bins_z, int nz, float *map, float interpolation_weight,
                           float min_ray_depth, float max_ray_depth, int mode, void *workspace,
                           size_t workspace_bytes)
{
    return layer_update(stream, rays, depth, features, class_ids, pose, T, H, W, fh, fw, F, bins_x, nx, bins_y, ny,
                        bins_z, nz, map, nullptr, interpolation_weight, min_ray_depth, max_ray_depth, mode, workspace,
                        workspace_bytes);
}

// ---- frame-sharded scenes ---------------------------------------------------------------
MB_API int mb_layer_fold(void *stream, const float *rays, const float *depth, const float *features,
                         const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                         const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                         float *partial_b, float *partial_a, float interpolation_weight, float min_ray_depth,
                         float max_ray_depth, void *workspace, size_t workspace_bytes)
{
    MB_REQUIRE(partial_a != nullptr, "mb_layer_fold: null pointer");
    return layer_update(stream, rays, depth, features, class_ids, pose, T, H, W, fh, fw, F, bins_x, nx, bins_y, ny,
                        bins_z, nz, partial_b, partial_a, interpolation_weight, min_ray_depth, max_ray_depth,
                        MB_MODE_FAST, workspace, workspace_bytes);
}

// ---- sparse partials + peer memory (frame-sharded scenes over NVLink) -------------------------------------------------
MB_API size_t mb_partial_buffer_bytes(uint32_t capacity, int F)
{
    return F > 0 ? mbk_partial_buffer_layout(capacity, F, nullptr) : 0;
}

MB_API int mb_partial_buffer_layout(uint32_t capacity, int F, size_t *offsets_host)
{
    MB_REQUIRE(offsets_host && F > 0, "mb_partial_buffer_layout: bad arguments");
    mbk_partial_buffer_layout(capacity, F, offsets_host);
    return MB_OK;
}

MB_API int mb_partial_reset(void *stream, int32_t *slot_table, int64_t voxels, void *partial_buffer)
{
    MB_REQUIRE(slot_table && partial_buffer && voxels > 0, "mb_partial_reset: bad arguments");
    return mbk_partial_reset((cudaStream_t)stream, slot_table, voxels, partial_buffer);
}

MB_API int mb_partial_clear(void *stream, int32_t *slot_table, void *partial_buffer, uint32_t capacity, int F)
{
    MB_REQUIRE(slot_table && partial_buffer && F > 0, "mb_partial_clear: bad arguments");
    return mbk_partial_clear((cudaStream_t)stream, slot_table, partial_buffer, capacity, F);
}

MB_API int mb_layer_fold_sparse(void *stream, const float *rays, const float *depth, const float *features,
                                const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                                const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                                int32_t *slot_table, void *partial_buffer, uint32_t capacity, float interpolation_weight,
                                float min_ray_depth, float max_ray_depth, void *workspace, size_t workspace_bytes)
{
    MB_REQUIRE(slot_table && partial_buffer && capacity > 0, "mb_layer_fold_sparse: null pointer");
    MB_REQUIRE(capacity < 0x7fffffffu, "mb_layer_fold_sparse: capacity too large");
    const MbSparseFold sp = { slot_table, partial_buffer, capacity };
    return layer_update(stream, rays, depth, features, class_ids, pose, T, H, W, fh, fw, F, bins_x, nx, bins_y, ny,
                        bins_z, nz, nullptr, nullptr, interpolation_weight, min_ray_depth, max_ray_depth, MB_MODE_FAST,
                        workspace, workspace_bytes, &sp);
}

MB_API int mb_partial_pull(void *stream, const void *const *peer_buffers_host, void *const *staging_slots_host, int world,
                           int self, uint32_t capacity, int F)
{
    MB_REQUIRE(peer_buffers_host && staging_slots_host && F > 0, "mb_partial_pull: bad arguments");
    return mbk_partial_pull((cudaStream_t)stream, peer_buffers_host, staging_slots_host, world, self, capacity, F);
}

MB_API int mb_affine_apply_partial(void *stream, float *map, int F, const void *partial_buffer, uint32_t capacity)
{
    MB_REQUIRE(map && partial_buffer && F > 0, "mb_affine_apply_partial: bad arguments");
    return mbk_affine_apply_partial((cudaStream_t)stream, map, F, partial_buffer, capacity);
}

// Peer memory: a partial buffer another process / GPU of the box can map (CUDA IPC; over NVLink between GPUs).
// This is the one place where the library allocates: an IPC handle names a whole cudaMalloc allocation, so the
// buffer must be its own allocation rather than a slice of the caller's caching allocator.
MB_API int mb_peer_alloc(size_t bytes, void **ptr_host)
{
    MB_REQUIRE(ptr_host && bytes > 0, "mb_peer_alloc: bad arguments");
    MB_CHECK_CUDA(cudaMalloc(ptr_host, bytes));
    MB_CHECK_CUDA(cudaMemset(*ptr_host, 0, bytes));
    return MB_OK;
}

MB_API int mb_peer_free(void *ptr)
{
    if (ptr) MB_CHECK_CUDA(cudaFree(ptr));
    return MB_OK;
}

MB_API int mb_peer_export(const void *ptr, void *handle64_host)
{
    MB_REQUIRE(ptr && handle64_host, "mb_peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    MB_CHECK_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64_host, const_cast<void *>(ptr)));
    return MB_OK;
}

MB_API int mb_peer_open(const void *handle64_host, void **ptr_host)
{
    MB_REQUIRE(handle64_host && ptr_host, "mb_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    MB_CHECK_CUDA(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return MB_OK;
}

MB_API int mb_peer_close(void *ptr)
{
    if (ptr) MB_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return MB_OK;
}

MB_API int mb_affine_apply_rows(void *stream, float *map, int F, const int64_t *voxel_index, const float *a,
                                const float *b, int64_t n)
{
    MB_REQUIRE(n >= 0 && F > 0, "mb_affine_apply_rows: bad sizes");
    if (n == 0) return MB_OK;
    MB_REQUIRE(map && voxel_index && a && b, "mb_affine_apply_rows: null pointer");
    return mbk_affine_apply_rows((cudaStream_t)stream, map, F, voxel_index, a, b, n);
}

// ---- measurement aid ---------------------------------------------------------------------------------
MB_API int mb_profile_stages(int enable) { return mbk_profile_enable(enable); }

MB_API int mb_profile_read(float *ms_host, int capacity)
{
    MB_REQUIRE(ms_host && capacity > 0, "mb_profile_read: null pointer");
    return mbk_profile_read(ms_host, capacity);
}

// ---- a11: SemanticProjectionLayer.find -------------------------------------------------------------
MB_API size_t mb_class_presence_workspace_bytes(int S0, int S1, int S2, int contour_padding)
{
    if (S0 <= 0 || S1 <= 0 || S2 <= 0) return 256;
    return mbk_class_presence_workspace_bytes(S0, S1, S2, contour_padding);
}

MB_API int mb_class_presence(void *stream, const float *map, int S0, int S1, int S2, int F, int semantic_category,
                             int contour_padding, float contour_threshold, uint8_t *image, void *workspace,
                             size_t workspace_bytes)
{
    MB_REQUIRE(map && image, "mb_class_presence: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && S2 > 0 && F > 0, "mb_class_presence: bad map shape");
    MB_REQUIRE(semantic_category >= 0 && semantic_category < F, "mb_class_presence: category %d outside [0, %d)",
               semantic_category, F);
    MB_REQUIRE(contour_padding >= 0 && contour_padding <= 64, "mb_class_presence: bad contour_padding %d", contour_padding);
    return mbk_class_presence((cudaStream_t)stream, map, S0, S1, S2, F, semantic_category, contour_padding,
                              contour_threshold, image, workspace, workspace_bytes);
}

MB_API int mb_column_summary(void *stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi,
                             float obstacle_threshold, float *amax, uint8_t *blocked)
{
    MB_REQUIRE(map && (amax || blocked), "mb_column_summary: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && S2 > 0 && F > 0, "mb_column_summary: bad map shape");
    MB_REQUIRE(0 <= z_lo && z_lo <= z_hi && z_hi <= S2, "mb_column_summary: depth slice [%d, %d) outside [0, %d]", z_lo, z_hi, S2);
    return mbk_column_summary((cudaStream_t)stream, map, S0, S1, S2, F, z_lo, z_hi, obstacle_threshold, amax, blocked);
}

MB_API int mb_masks_to_ids(void *stream, const uint8_t *masks, const int64_t *classes, const float *scores, int n,
                           int64_t npix, int num_classes, float detection_threshold, int64_t *ids)
{
    MB_REQUIRE(n >= 0 && npix >= 0, "mb_masks_to_ids: negative size");
    if (npix == 0) return MB_OK;
    MB_REQUIRE(ids && (n == 0 || (masks && classes && scores)), "mb_masks_to_ids: null pointer");
    return mbk_masks_to_ids((cudaStream_t)stream, masks, classes, scores, n, (size_t)npix, num_classes, detection_threshold, ids);
}

MB_API int mb_top_down(void *stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float *out)
{
    MB_REQUIRE(map && out, "mb_top_down: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && S2 > 0 && F > 0, "mb_top_down: bad map shape");
    MB_REQUIRE(0 <= z_lo && z_lo <= z_hi && z_hi <= S2, "mb_top_down: depth slice [%d, %d) outside [0, %d]", z_lo, z_hi, S2);
    return mbk_top_down((cudaStream_t)stream, map, S0, S1, S2, F, z_lo, z_hi, out);
}

MB_API int mb_instance_pool(void *stream, const int32_t *boxes, int nboxes, const float *sem_map, int S0, int S1, int S2,
                            int F, int semantic_category, const float *feat_map, int FF, const float *centres_x,
                            const float *centres_y, const float *centres_z, float *out)
{
    MB_REQUIRE(nboxes >= 0, "mb_instance_pool: negative box count");
    if (nboxes == 0) return MB_OK;
    MB_REQUIRE(boxes && sem_map && centres_x && centres_y && centres_z && out, "mb_instance_pool: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && S2 > 0 && F > 0 && semantic_category >= -1 && semantic_category < F,
               "mb_instance_pool: bad map shape or category");
    MB_REQUIRE(feat_map == nullptr || FF > 0, "mb_instance_pool: feature map without a feature size");
    return mbk_instance_pool((cudaStream_t)stream, boxes, nboxes, sem_map, S0, S1, S2, F, semantic_category, feat_map,
                             FF, centres_x, centres_y, centres_z, out);
}

// ---- next to the path: coordinate transforms + navigation graph ---------------------------------
MB_API int mb_world_to_map(void *stream, const float *coords, int64_t n, int k, const float *bins_x, int nx,
                           const float *bins_y, int ny, const float *bins_z, int nz, int64_t *out)
{
    MB_REQUIRE(n >= 0 && (k == 2 || k == 3), "mb_world_to_map: coordinates must be [n, 2] or [n, 3]");
    if (n == 0) return MB_OK;
    MB_REQUIRE(coords && out && bins_x && bins_y && (k == 2 || bins_z), "mb_world_to_map: null pointer");
    MB_REQUIRE(nx >= 2 && ny >= 2 && (k == 2 || nz >= 2), "mb_world_to_map: edge tables need two entries");
    return mbk_world_to_map((cudaStream_t)stream, coords, n, k, bins_x, nx, bins_y, ny, bins_z, nz, out);
}

MB_API int mb_map_to_world(void *stream, const float *coords, int64_t n, int k, const float *bins_x, int nx,
                           const float *bins_y, int ny, const float *bins_z, int nz, float *out)
{
    MB_REQUIRE(n >= 0 && (k == 2 || k == 3), "mb_map_to_world: coordinates must be [n, 2] or [n, 3]");
    if (n == 0) return MB_OK;
    MB_REQUIRE(coords && out && bins_x && bins_y && (k == 2 || bins_z), "mb_map_to_world: null pointer");
    MB_REQUIRE(nx >= 2 && ny >= 2 && (k == 2 || nz >= 2), "mb_map_to_world: edge tables need two entries");
    return mbk_map_to_world((cudaStream_t)stream, coords, n, k, bins_x, nx, bins_y, ny, bins_z, nz, out);
}

MB_API int mb_navigable_area(void *stream, const uint8_t *blocked, int S0, int S1, int padding, float *navigable)
{
    MB_REQUIRE(blocked && navigable && S0 > 0 && S1 > 0 && padding >= 0, "mb_navigable_area: bad arguments");
    return mbk_navigable_area((cudaStream_t)stream, blocked, S0, S1, padding, navigable);
}

MB_API int mb_nav_graph_lattice(void *stream, const float *navigable, int S0, int S1, int offset_y, int offset_x,
                                int step_size, uint8_t *node_ok, uint8_t *edge_ok)
{
    MB_REQUIRE(navigable && node_ok && edge_ok, "mb_nav_graph_lattice: null pointer");
    MB_REQUIRE(S0 > 0 && S1 > 0 && step_size > 0 && offset_y >= 0 && offset_x >= 0 && offset_y < step_size &&
               offset_x < step_size, "mb_nav_graph_lattice: bad lattice");
    const int ny = (S0 - offset_y + step_size - 1) / step_size, nx = (S1 - offset_x + step_size - 1) / step_size;
    if (ny <= 0 || nx <= 0) return MB_OK;
    return mbk_nav_edges((cudaStream_t)stream, navigable, S0, S1, offset_y, offset_x, step_size, ny, nx, node_ok, edge_ok);
}

MB_API int mb_nav_rects_clear(void *stream, const float *navigable, int S0, int S1, const int32_t *rects, int m,
                              uint8_t *clear)
{
    MB_REQUIRE(m >= 0 && S0 > 0 && S1 > 0, "mb_nav_rects_clear: bad sizes");
    if (m == 0) return MB_OK;
    MB_REQUIRE(navigable && rects && clear, "mb_nav_rects_clear: null pointer");
    return mbk_nav_rects((cudaStream_t)stream, navigable, S1, rects, m, clear);
}

// ---- a12: predict_scene_differences --------------------------------------------------------------------
MB_API int mb_pairwise_l2(void *stream, const float *a, int n, const float *b, int m, int d, float *out)
{
    MB_REQUIRE(n >= 0 && m >= 0 && d >= 0, "mb_pairwise_l2: negative size");
    if (n == 0 || m == 0) return MB_OK;
    MB_REQUIRE(a && b && out, "mb_pairwise_l2: null pointer");
    return mbk_pairwise_l2((cudaStream_t)stream, a, n, b, m, d, out);
}

MB_API int mb_cosine_best_match(void *stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                                float *best_sim)
{
    MB_REQUIRE(n >= 0 && m >= 0 && d >= 0, "mb_cosine_best_match: negative size");
    if (n == 0) return MB_OK;
    MB_REQUIRE(a && best && best_sim && (m == 0 || b), "mb_cosine_best_match: null pointer");
    return mbk_cosine_best_match((cudaStream_t)stream, a, n, b, m, d, best, best_sim);
}

MB_API size_t mb_cosine_best_match_tc_workspace_bytes(int n, int m, int d)
{
    if (n <= 0 || d <= 0) return 256;
    return mbk_cosine_tc_workspace_bytes(n, m > 0 ? m : 1, d);
}

MB_API int mb_cosine_best_match_tc(void *stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                                   float *best_sim, void *workspace, size_t workspace_bytes)
{
    MB_REQUIRE(n >= 0 && m >= 0 && d > 0, "mb_cosine_best_match_tc: bad sizes");
    if (n == 0) return MB_OK;
    MB_REQUIRE(a && best && best_sim && (m == 0 || b), "mb_cosine_best_match_tc: null pointer");
    return mbk_cosine_best_match_tc((cudaStream_t)stream, a, n, b, m, d, best, best_sim, workspace, workspace_bytes);
}

MB_API size_t mb_lsap_workspace_bytes(int n, int m)
{
    if (n <= 0 || m <= 0) return 256;
    return mbk_lsap_workspace_bytes(n, m);
}

MB_API int mb_lsap(void *stream, const float *cost32, const double *cost64, int n, int m, int64_t *rows, int64_t *cols,
                   int32_t *status, void *workspace, size_t workspace_bytes)
{
    MB_REQUIRE(n >= 0 && m >= 0, "mb_lsap: negative size");
    if (n == 0 || m == 0) return MB_OK;
    MB_REQUIRE((cost32 != nullptr) != (cost64 != nullptr), "mb_lsap: pass exactly one of cost32 / cost64");
    MB_REQUIRE(rows && cols && status, "mb_lsap: null pointer");
    return mbk_lsap((cudaStream_t)stream, cost32, cost64, n, m, rows, cols, status, workspace, workspace_bytes);
}
