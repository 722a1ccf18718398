// Next to the mapping path: the coordinate transforms between world and map and the
// traversability tests behind the navigation graph, as kernels over whole batches / the whole node lattice.
//   world_to_map   /root/reference/mass/nn/base_projection_layer.py:513-547 (+ clamp_to_world :381-413)
//   map_to_world   /root/reference/mass/nn/base_projection_layer.py:452-511 (+ clamp_to_map :415-450)
//   navigable area /root/reference/mass/navigation_policy.py:205-221 (the 2-D part: obstacle padding)
//   graph edges    /root/reference/mass/navigation_policy.py:253-285 (reset) and :315-341 (update)
// All of it is exact integer / fp32 work in the reference's operation order (explicit round-to-nearest intrinsics:
// the library is built without FMA contraction, and the interpolation must not be contracted either).
#include "common.cuh"
#include "kernels.cuh"
#include "geometry.cuh"

namespace {

struct Axis3 {
    const float *bins[3];
    int n[3];
};

// mid-point of the i-th voxel along an axis; the y axis is stored flipped (line 476-477)
__device__ __forceinline__ float voxel_mid(const Axis3 &A, int axis, int i)
{
    if (axis == 1) i = A.n[1] - 2 - i;
    return __fdiv_rn(__fadd_rn(__ldg(A.bins[axis] + i), __ldg(A.bins[axis] + i + 1)), 2.0f);
}

// world -> map: clamp to the span of the voxel mid-points, bucketize(right=True) - 1, y flipped
__global__ void __launch_bounds__(256)
k_world_to_map(const float *__restrict__ coords, int64_t n, int k, Axis3 A, int64_t *__restrict__ out)
{
    const int64_t total = n * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int axis = (int)(e % k);
        const float *b = A.bins[axis];
        const int nb = A.n[axis];
        const float lo = __fdiv_rn(__fadd_rn(__ldg(b), __ldg(b + 1)), 2.0f);
        const float hi = __fdiv_rn(__fadd_rn(__ldg(b + nb - 1), __ldg(b + nb - 2)), 2.0f);
        float x = coords[e];
        // torch.clamp(min, max) = min(max(x, lo), hi) with NaN passed through
        if (x == x) x = fminf(fmaxf(x, lo), hi);
        const int r = bucket_right(b, nb, x);                      // bucketize(x, right=True) - 1
        out[e] = axis == 1 ? (int64_t)(nb - 2 - r) : (int64_t)r;
    }
}

// map -> world: clamp to [0, size - 1], split into cell + fraction, interpolate between neighbouring mid-points
__global__ void __launch_bounds__(256)
k_map_to_world(const float *__restrict__ coords, int64_t n, int k, Axis3 A, float *__restrict__ out)
{
    const int64_t total = n * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int axis = (int)(e % k);
        const int size = A.n[axis] - 1;
        float x = coords[e];
        if (x == x) x = fminf(fmaxf(x, 0.0f), (float)(size - 1));
        const float fl = floorf(x);
        int i = (int)fl;
        if (!(x == x)) i = 0;                                     // (NaN has no cell; the result is NaN either way)
        const int ir = min(max(i + 1, 0), size - 1);
        const float l = voxel_mid(A, axis, i), r = voxel_mid(A, axis, ir);
        out[e] = __fadd_rn(l, __fmul_rn(__fsub_rn(r, l), __fsub_rn(x, fl)));
    }
}

// 1 - max_pool2d(1 - navigable, 2p+1, stride 1, padding p): a cell is navigable iff no cell of its (2p+1)^2
// neighbourhood (clipped to the image) is blocked
__global__ void __launch_bounds__(256)
k_navigable_area(const uint8_t *__restrict__ blocked, int S0, int S1, int pad, float *__restrict__ out)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= S1 || y >= S0) return;
    bool any = false;
    for (int yy = max(y - pad, 0); yy <= min(y + pad, S0 - 1) && !any; ++yy)
        for (int xx = max(x - pad, 0); xx <= min(x + pad, S1 - 1); ++xx)
            if (blocked[(size_t)yy * S1 + xx]) { any = true; break; }
    out[(size_t)y * S1 + x] = any ? 0.0f : 1.0f;
}

// node lattice of the navigation graph: node (a, b) sits at row off_y + a*step, column off_x + b*step.
// node_ok: the node's own cell is navigable; edge_ok[.][0]: the segment down to the next node (rows i .. i+step,
// column j) is inside the map and entirely navigable; edge_ok[.][1]: the segment to the right likewise.
__global__ void __launch_bounds__(256)
k_nav_edges(const float *__restrict__ nav, int S0, int S1, int off_y, int off_x, int step, int ny, int nx,
            uint8_t *__restrict__ node_ok, uint8_t *__restrict__ edge_ok)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ny * nx) return;
    const int a = idx / nx, b = idx % nx;
    const int i = off_y + a * step, j = off_x + b * step;
    node_ok[idx] = nav[(size_t)i * S1 + j] != 0.0f;
    bool down = i + step < S0, right = j + step < S1;
    for (int d = 0; d <= step && down; ++d) down = nav[(size_t)(i + d) * S1 + j] == 1.0f;
    for (int d = 0; d <= step && right; ++d) right = nav[(size_t)i * S1 + j + d] == 1.0f;
    edge_ok[2 * idx] = down;
    edge_ok[2 * idx + 1] = right;
}

// rects[m][4] = {row0, row1, col0, col1} inclusive (clipped to the map by the caller): clear[m] = every cell == 1,
// the test update_navigation_graph applies to each node (a 1 x 1 rectangle) and each edge of an existing graph
__global__ void __launch_bounds__(256)
k_nav_rects(const float *__restrict__ nav, int S1, const int32_t *__restrict__ rects, int m, uint8_t *__restrict__ clear)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m) return;
    const int4 r = *(const int4 *)(rects + 4 * (size_t)idx);
    bool ok = true;
    for (int i = r.x; i <= r.y && ok; ++i)
        for (int j = r.z; j <= r.w; ++j)
            if (nav[(size_t)i * S1 + j] != 1.0f) { ok = false; break; }
    clear[idx] = ok;
}

Axis3 make_axes(const float *bx, int nx, const float *by, int ny, const float *bz, int nz)
{
    Axis3 A;
    A.bins[0] = bx; A.bins[1] = by; A.bins[2] = bz;
    A.n[0] = nx; A.n[1] = ny; A.n[2] = nz;
    return A;
}

int grid_for(int64_t total)
{
    const int64_t g = (total + 255) / 256;
    return (int)(g < MB_NUM_SMS * 16 ? (g < 1 ? 1 : g) : MB_NUM_SMS * 16);
}

}  // namespace

int mbk_world_to_map(cudaStream_t stream, const float *coords, int64_t n, int k, const float *bx, int nx, const float *by,
                     int ny, const float *bz, int nz, int64_t *out)
{
    k_world_to_map<<<grid_for(n * k), 256, 0, stream>>>(coords, n, k, make_axes(bx, nx, by, ny, bz, nz), out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_map_to_world(cudaStream_t stream, const float *coords, int64_t n, int k, const float *bx, int nx, const float *by,
                     int ny, const float *bz, int nz, float *out)
{
    k_map_to_world<<<grid_for(n * k), 256, 0, stream>>>(coords, n, k, make_axes(bx, nx, by, ny, bz, nz), out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_navigable_area(cudaStream_t stream, const uint8_t *blocked, int S0, int S1, int padding, float *out)
{
    dim3 grid((S1 + 31) / 32, (S0 + 7) / 8);
    k_navigable_area<<<grid, 256, 0, stream>>>(blocked, S0, S1, padding, out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_nav_edges(cudaStream_t stream, const float *navigable, int S0, int S1, int off_y, int off_x, int step, int ny,
                  int nx, uint8_t *node_ok, uint8_t *edge_ok)
{
    k_nav_edges<<<(ny * nx + 255) / 256, 256, 0, stream>>>(navigable, S0, S1, off_y, off_x, step, ny, nx, node_ok, edge_ok);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_nav_rects(cudaStream_t stream, const float *navigable, int S1, const int32_t *rects, int m, uint8_t *clear)
{
    k_nav_rects<<<(m + 255) / 256, 256, 0, stream>>>(navigable, S1, rects, m, clear);
    MB_LAUNCHED();
    return MB_OK;
}
