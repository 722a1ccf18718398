// Stable LSD radix sort of (key, value) u32 pairs and a device-wide exclusive scan.
//
// Used to order splat contributions by voxel key.  Stability is what makes the voxel reduce
// reproduce the reference's summation order (slot-major, then pixel order:
// /root/reference/mass/utils/projection.py:294-298, 319-323, 349-351) without float atomics.
//
// One pass = per-tile digit histogram -> exclusive scan over (digit, tile) -> stable scatter.
// A tile is 2048 (8-bit digits) or 4096 (9-bit digits) elements; inside a tile each warp owns 256
// consecutive elements and ranks them in 8 rounds of 32 with match.any, so the order
// (tile, warp, round, lane) is the input order.
#include "common.cuh"

namespace {

constexpr int SORT_ITEMS = 8;      // keys per thread and pass

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---------------------------------------------------------------------------------------------
// exclusive scan
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread over a CTA of SCAN_THREADS threads; returns the
// exclusive prefix and (in *total) the CTA sum.  Safe to call repeatedly (ends on a barrier).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
    constexpr int NW = SCAN_THREADS / 32;
    __shared__ uint32_t warp_off_s[NW];
    __shared__ uint32_t total_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) warp_off_s[warp] = inc;          // warp totals
    __syncthreads();
    if (warp == 0) {
        const uint32_t s = lane < NW ? warp_off_s[lane] : 0u;
        const uint32_t si = warp_inclusive_scan(s, lane);
        if (lane < NW) warp_off_s[lane] = si - s;    // exclusive warp offsets
        if (lane == NW - 1) total_s = si;
    }
    __syncthreads();
    const uint32_t excl = warp_off_s[warp] + inc - v;
    *total = total_s;
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tile_sums(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ sums)
{
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += in[base + i];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of m block sums in place (chunked with a running carry)
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(uint32_t *sums, uint32_t m)
{
    uint32_t carry = 0;
    for (uint32_t base = 0; base < m; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < m ? sums[i] : 0;
        uint32_t total;
        uint32_t e = block_exclusive_scan(v, &total);
        if (i < m) sums[i] = carry + e;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const uint32_t *in, uint32_t *out, uint32_t n, const uint32_t *__restrict__ sums)
{
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, &total) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// radix pass, BITS bits per digit with one thread per digit (256 or 512 threads per CTA).
// n_dev (optional) overrides n with a count produced on the device; the grid is sized for the
// host-side upper bound n and tiles past the device count write empty histograms.
template <int BITS>
__global__ void __launch_bounds__(1 << BITS)
k_radix_hist(const uint32_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ n_dev, int shift,
             uint32_t *__restrict__ tile_hist, uint32_t ntiles)
{
    constexpr int R = 1 << BITS, TILE = R * SORT_ITEMS;
    __shared__ uint32_t h[R];
    if (n_dev) n = min(n, *n_dev);
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * TILE;
    if (base < n) {
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            uint32_t idx = base + i * R + threadIdx.x;
            if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (R - 1)], 1u);
        }
    }
    __syncthreads();
    tile_hist[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];   // digit-major
}

template <int BITS>
__global__ void __launch_bounds__(1 << BITS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                const uint32_t *__restrict__ n_dev, int shift, const uint32_t *__restrict__ tile_offs,
                uint32_t ntiles, int vals_iota)
{
    constexpr int R = 1 << BITS, WARPS = R / 32, TILE = R * SORT_ITEMS;
    __shared__ uint32_t wcnt[WARPS][R];
    if (n_dev) n = min(n, *n_dev);
    if (blockIdx.x * TILE >= n) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < WARPS * R; i += R) (&wcnt[0][0])[i] = 0;
    __syncthreads();

    const uint32_t wbase = blockIdx.x * TILE + warp * (32 * SORT_ITEMS);
    uint32_t k[SORT_ITEMS], v[SORT_ITEMS], rk[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : 0xffffffffu;
        v[r] = valid ? (vals_iota ? idx : vals_in[idx]) : 0u;
        const uint32_t d = (k[r] >> shift) & (R - 1);
        // out-of-range lanes get a private pseudo-digit so they never join a real group
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (R + lane));
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        const uint32_t prev = valid ? wcnt[warp][d] : 0u;
        __syncwarp();
        if (valid && rank == 0) wcnt[warp][d] = prev + __popc(m);
        __syncwarp();
        rk[r] = prev + rank;
    }
    __syncthreads();
    {   // thread d: exclusive scan of digit d over the warps, seeded with this tile's global offset
        uint32_t run = tile_offs[threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = (k[r] >> shift) & (R - 1);
            const uint32_t pos = wcnt[warp][d] + rk[r];
            keys_out[pos] = k[r];
            vals_out[pos] = v[r];
        }
    }
}

template <int BITS>
int radix_pass(cudaStream_t stream, const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
               uint32_t n, const uint32_t *n_dev, int shift, bool iota, uint32_t *hist, void *scan_ws,
               size_t scan_bytes)
{
    constexpr int R = 1 << BITS, TILE = R * SORT_ITEMS;
    const uint32_t ntiles = (n + TILE - 1) / TILE;
    k_radix_hist<BITS><<<ntiles, R, 0, stream>>>(kin, n, n_dev, shift, hist, ntiles);
    MB_LAUNCHED();
    int rc = mb_exclusive_scan_u32(stream, hist, hist, ntiles * R, scan_ws, scan_bytes);
    if (rc) return rc;
    k_radix_scatter<BITS><<<ntiles, R, 0, stream>>>(kin, vin, kout, vout, n, n_dev, shift, hist, ntiles,
                                                    iota ? 1 : 0);
    MB_LAUNCHED();
    return MB_OK;
}

}  // namespace

size_t mb_scan_workspace_bytes(uint32_t n)
{
    return mb_align_up(((size_t)n + SCAN_TILE - 1) / SCAN_TILE * sizeof(uint32_t)) + 256;
}

int mb_exclusive_scan_u32(cudaStream_t stream, const uint32_t *in, uint32_t *out, uint32_t n,
                          void *workspace, size_t workspace_bytes)
{
    if (n == 0) return MB_OK;
    MB_REQUIRE(workspace_bytes >= mb_scan_workspace_bytes(n), "scan workspace too small");
    uint32_t *sums = (uint32_t *)workspace;
    const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tile_sums<<<tiles, SCAN_THREADS, 0, stream>>>(in, n, sums);
    MB_LAUNCHED();
    k_scan_sums<<<1, SCAN_THREADS, 0, stream>>>(sums, tiles);
    MB_LAUNCHED();
    k_scan_apply<<<tiles, SCAN_THREADS, 0, stream>>>(in, out, n, sums);
    MB_LAUNCHED();
    return MB_OK;
}

// 9-bit digits only where they save a pass: 18-bit brick keys sort in two passes, 24-bit voxel keys
// in three 8-bit ones
static int digit_bits(int key_bits) { return (key_bits + 8) / 9 < (key_bits + 7) / 8 ? 9 : 8; }

size_t mb_sort_workspace_bytes(uint32_t n)
{
    const size_t ntiles = ((size_t)n + 256 * SORT_ITEMS - 1) / (256 * SORT_ITEMS);   // the smaller tile
    const size_t hist = mb_align_up(ntiles * 512 * sizeof(uint32_t));
    return hist + mb_scan_workspace_bytes((uint32_t)(ntiles * 512)) + 256;
}

int mb_sort_pairs(cudaStream_t stream, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b,
                  uint32_t *vals_b, uint32_t n, const uint32_t *n_dev, int key_bits, bool vals_a_is_iota,
                  void *workspace, size_t workspace_bytes, uint32_t **keys_out, uint32_t **vals_out)
{
    *keys_out = keys_a;
    *vals_out = vals_a;
    if (n == 0) return MB_OK;
    MB_REQUIRE(workspace_bytes >= mb_sort_workspace_bytes(n), "sort workspace too small");
    MbArena arena(workspace, workspace_bytes);
    const size_t ntiles = ((size_t)n + 256 * SORT_ITEMS - 1) / (256 * SORT_ITEMS);
    uint32_t *hist = arena.take<uint32_t>(ntiles * 512);
    const size_t scan_bytes = mb_scan_workspace_bytes((uint32_t)(ntiles * 512));
    char *scan_ws = arena.take<char>(scan_bytes);

    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    bool iota = vals_a_is_iota;
    const int bits = digit_bits(key_bits);
    const int passes = (key_bits + bits - 1) / bits;
    for (int p = 0; p < passes; ++p) {
        int rc = bits == 9 ? radix_pass<9>(stream, kin, vin, kout, vout, n, n_dev, p * 9, iota, hist, scan_ws, scan_bytes)
                           : radix_pass<8>(stream, kin, vin, kout, vout, n, n_dev, p * 8, iota, hist, scan_ws, scan_bytes);
        if (rc) return rc;
        iota = false;
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *keys_out = kin;
    *vals_out = vin;
    return MB_OK;
}
