// K1: depth -> world unprojection and voxelisation (sm_100a).
//
// Restates, with the reference's fp32 operation order and no FMA contraction:
//   transform_rays   /root/reference/mass/utils/projection.py:104-110
//   bin_rays         /root/reference/mass/utils/projection.py:182-230
//   splat indices    /root/reference/mass/utils/projection.py:280-298
// Every multiply/add below is an explicit __f*_rn intrinsic so that the result does not depend
// on compiler contraction flags.
#include "common.cuh"
#include "kernels.cuh"
#include "geometry.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_transform_rays(const float *__restrict__ rays, int64_t npix, const float *__restrict__ pose,
                 float *__restrict__ out)
{
    __shared__ float R[9];
    if (threadIdx.x < 9) R[threadIdx.x] = pose[threadIdx.x];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * blockDim.x) {
        float o0, o1, o2;
        orient(R, rays[3 * p], rays[3 * p + 1], rays[3 * p + 2], o0, o1, o2);
        out[3 * p] = o0; out[3 * p + 1] = o1; out[3 * p + 2] = o2;
    }
}

// bin_rays stage 1: per-pixel validity flag (u32) for the order-preserving compaction
__global__ void __launch_bounds__(256)
k_bin_flags(const float *__restrict__ bins0, int n0, const float *__restrict__ bins1, int n1,
            const float *__restrict__ bins2, int n2, const float *__restrict__ origin,
            const float *__restrict__ rays, const float *__restrict__ depth, int64_t npix, float min_d,
            float max_d, uint32_t *__restrict__ flags)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    BinResult b = bin_point(bins0, n0, bins1, n1, bins2, n2, origin[0], origin[1], origin[2],
                            rays[3 * p], rays[3 * p + 1], rays[3 * p + 2], depth[p], min_d, max_d);
    flags[p] = b.ok ? 1u : 0u;
}

// bin_rays stage 2: recompute and write each valid pixel at its scanned position
__global__ void __launch_bounds__(256)
k_bin_write(const float *__restrict__ bins0, int n0, const float *__restrict__ bins1, int n1,
            const float *__restrict__ bins2, int n2, const float *__restrict__ origin,
            const float *__restrict__ rays, const float *__restrict__ depth, int64_t npix, float min_d,
            float max_d, const uint32_t *__restrict__ offsets, int64_t *__restrict__ ind0,
            int64_t *__restrict__ ind1, int64_t *__restrict__ ind2, float *__restrict__ ratio0,
            float *__restrict__ ratio1, float *__restrict__ ratio2, int64_t *__restrict__ pix,
            int64_t *__restrict__ count)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    BinResult b = bin_point(bins0, n0, bins1, n1, bins2, n2, origin[0], origin[1], origin[2],
                            rays[3 * p], rays[3 * p + 1], rays[3 * p + 2], depth[p], min_d, max_d);
    const uint32_t pos = offsets[p];
    if (b.ok) {
        ind0[pos] = b.i0; ind1[pos] = b.i1; ind2[pos] = b.i2;
        ratio0[pos] = b.q0; ratio1[pos] = b.q1; ratio2[pos] = b.q2;
        pix[pos] = p;
    }
    if (p == npix - 1) *count = (int64_t)pos + (b.ok ? 1 : 0);
}

__device__ __forceinline__ void emit_keys(const MbGrid &g, bool ok, int a0, int a1, int a2, float q0,
                                          float q1, float q2, uint32_t p, uint32_t np,
                                          uint32_t *__restrict__ keys, float4 *__restrict__ pt_ratio)
{
    int lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
    if (ok) {
        axis_pair(a0, q0, g.S0, lo0, hi0);
        axis_pair(a1, q1, g.S1, lo1, hi1);
        axis_pair(a2, q2, g.S2, lo2, hi2);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int i0 = (s & 4) ? hi0 : lo0, i1 = (s & 2) ? hi1 : lo1, i2 = (s & 1) ? hi2 : lo2;
        keys[(size_t)s * np + p] = ok ? mb_voxel_key(g, i0, i1, i2) : g.invalid;
    }
    pt_ratio[p] = make_float4(q0, q1, q2, 0.f);
}

// Fused K1 for one frame: orient ray, unproject, bin, emit the 8 splat keys per pixel
// (slot-major: contribution id = slot * npix + pixel) and the per-pixel ratios.  The map axes
// are (y flipped, x, z) = input axes (1, 0, 2): base_projection_layer.py:339.
__global__ void __launch_bounds__(256)
k_unproject_voxelise(const float *__restrict__ rays, const float *__restrict__ depth,
                     const float *__restrict__ pose, uint32_t npix, const float *__restrict__ bins_x,
                     int nx, const float *__restrict__ bins_y, int ny, const float *__restrict__ bins_z,
                     int nz, MbGrid g, float min_d, float max_d, uint32_t *__restrict__ keys,
                     float4 *__restrict__ pt_ratio, uint32_t *__restrict__ counters)
{
    __shared__ float P[12];
    if (threadIdx.x < 12) P[threadIdx.x] = pose[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x < MB_NUM_COUNTERS) counters[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float r0, r1, r2;
    orient(P, rays[3 * (size_t)p], rays[3 * (size_t)p + 1], rays[3 * (size_t)p + 2], r0, r1, r2);
    BinResult b = bin_point(bins_x, nx, bins_y, ny, bins_z, nz, P[9], P[10], P[11], r0, r1, r2,
                            depth[p], min_d, max_d);
    emit_keys(g, b.ok, b.i1, b.i0, b.i2, b.q1, b.q0, b.q2, p, npix, keys, pt_ratio);
}

// K1 for explicit point lists (update_feature_map): indices/ratios are given.
__global__ void __launch_bounds__(256)
k_points_to_keys(const int64_t *__restrict__ ind0, const int64_t *__restrict__ ind1,
                 const int64_t *__restrict__ ind2, const float *__restrict__ ratio0,
                 const float *__restrict__ ratio1, const float *__restrict__ ratio2, uint32_t npts,
                 MbGrid g, uint32_t *__restrict__ keys, float4 *__restrict__ pt_ratio,
                 uint32_t *__restrict__ counters)
{
    if (blockIdx.x == 0 && threadIdx.x < MB_NUM_COUNTERS) counters[threadIdx.x] = 0;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npts) return;
    const int64_t a0 = ind0[p], a1 = ind1[p], a2 = ind2[p];
    // indices outside the map would be an out-of-bounds scatter in the reference; drop them
    const bool ok = a0 >= 0 && a0 < g.S0 && a1 >= 0 && a1 < g.S1 && a2 >= 0 && a2 < g.S2;
    emit_keys(g, ok, (int)a0, (int)a1, (int)a2, ratio0[p], ratio1[p], ratio2[p], p, npts, keys, pt_ratio);
}

}  // namespace

int mbk_transform_rays(cudaStream_t stream, const float *rays, int64_t npix, const float *pose, float *out)
{
    if (npix <= 0) return MB_OK;
    int64_t blocks = (npix + 255) / 256;
    if (blocks > MB_NUM_SMS * 16) blocks = MB_NUM_SMS * 16;
    k_transform_rays<<<(unsigned)blocks, 256, 0, stream>>>(rays, npix, pose, out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_bin_flags(cudaStream_t stream, const float *bins0, int n0, const float *bins1, int n1,
                  const float *bins2, int n2, const float *origin, const float *rays, const float *depth,
                  int64_t npix, float min_d, float max_d, uint32_t *flags)
{
    k_bin_flags<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(bins0, n0, bins1, n1, bins2, n2, origin,
                                                                   rays, depth, npix, min_d, max_d, flags);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_bin_write(cudaStream_t stream, const float *bins0, int n0, const float *bins1, int n1,
                  const float *bins2, int n2, const float *origin, const float *rays, const float *depth,
                  int64_t npix, float min_d, float max_d, const uint32_t *offsets, int64_t *ind0,
                  int64_t *ind1, int64_t *ind2, float *ratio0, float *ratio1, float *ratio2, int64_t *pix,
                  int64_t *count)
{
    k_bin_write<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(
        bins0, n0, bins1, n1, bins2, n2, origin, rays, depth, npix, min_d, max_d, offsets, ind0, ind1, ind2,
        ratio0, ratio1, ratio2, pix, count);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_unproject_voxelise(cudaStream_t stream, const float *rays, const float *depth, const float *pose,
                           uint32_t npix, const float *bins_x, int nx, const float *bins_y, int ny,
                           const float *bins_z, int nz, const MbGrid &g, float min_d, float max_d,
                           uint32_t *keys, float4 *pt_ratio, uint32_t *counters)
{
    k_unproject_voxelise<<<(npix + 255) / 256, 256, 0, stream>>>(rays, depth, pose, npix, bins_x, nx, bins_y,
                                                                 ny, bins_z, nz, g, min_d, max_d, keys,
                                                                 pt_ratio, counters);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_points_to_keys(cudaStream_t stream, const int64_t *ind0, const int64_t *ind1, const int64_t *ind2,
                       const float *ratio0, const float *ratio1, const float *ratio2, uint32_t npts,
                       const MbGrid &g, uint32_t *keys, float4 *pt_ratio, uint32_t *counters)
{
    k_points_to_keys<<<(npts + 255) / 256, 256, 0, stream>>>(ind0, ind1, ind2, ratio0, ratio1, ratio2, npts, g,
                                                             keys, pt_ratio, counters);
    MB_LAUNCHED();
    return MB_OK;
}
