// Stable LSD radix sort of (key, value) u32 pairs and a device-wide exclusive scan.
//
// Used to order splat contributions by voxel key.  Stability is what makes the voxel reduce
// reproduce the reference's summation order (slot-major, then pixel order:
// /root/reference/mass/utils/projection.py:294-298, 319-323, 349-351) without float atomics.
//
// 8-bit digits, one sweep over the data per pass (see k_radix_onesweep below).  Inside a tile each warp
// owns 512 consecutive elements and ranks them in 16 rounds of 32 with match.any, so the order
// (tile, warp, round, lane) is the input order.
#include "common.cuh"

namespace {

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
// Radix sort, 8 bits per pass, one sweep over the data per pass ("onesweep"):
//   k_radix_hist_all   one read of the keys -> global digit histograms of ALL passes
//   k_radix_bases      exclusive scan of each pass's 256 counts
//   k_radix_onesweep   per pass: a tile of 4096 pairs is ranked inside the CTA (match.any per warp round,
//                      per-warp digit counters), the tile's position inside every digit comes from a
//                      decoupled look-back over the preceding tiles' digit counts (tiles take their index
//                      from a ticket, so a tile only ever waits for tiles that started before it), the
//                      pairs are staged through shared memory in sorted order and written out coalesced.
// Stable: the global order inside a digit is (tile, warp, round, lane) = the input order.
constexpr int OS_ITEMS = 16;
constexpr int OS_THREADS = 256;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VALUE_MASK = (1u << 30) - 1u;

// Each thread takes 16 consecutive keys and merges runs of equal digits before touching the shared
// histogram: neighbouring keys mostly share their digits (neighbouring pixels fall into the same cell),
// and the lanes of a warp are 16 keys apart, so same-address conflicts are rare.
__global__ void __launch_bounds__(256)
k_radix_hist_all(const uint32_t *__restrict__ keys, uint32_t n, int passes, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t h[4][256];
    for (int i = threadIdx.x; i < 4 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t ngroups = (n + 15u) / 16u;
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += gridDim.x * blockDim.x) {
        uint32_t k[16];
        const uint32_t base = gi * 16u;
        if (base + 16u <= n) {
            const uint4 *p = (const uint4 *)(keys + base);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = __ldg(p + q);
                k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) k[q] = base + q < n ? keys[base + q] : 0xffffffffu;
        }
        const int cnt = (int)min(16u, n - base);
        for (int p = 0; p < passes; ++p) {
            uint32_t cur = (k[0] >> (8 * p)) & 255u, run = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (q < cnt) {
                    const uint32_t d = (k[q] >> (8 * p)) & 255u;
                    if (d != cur) { atomicAdd(&h[p][cur], run); cur = d; run = 0; }
                    ++run;
                }
            }
            atomicAdd(&h[p][cur], run);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) {
        const uint32_t c = h[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * 256 + threadIdx.x], c);
    }
}

__global__ void __launch_bounds__(256)
k_radix_bases(uint32_t *__restrict__ ghist)
{
    uint32_t total;
    const uint32_t v = ghist[blockIdx.x * 256 + threadIdx.x];
    const uint32_t e = block_exclusive_scan(v, &total);
    ghist[blockIdx.x * 256 + threadIdx.x] = e;
}

__global__ void __launch_bounds__(OS_THREADS, 4)
k_radix_onesweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                 uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n, int shift,
                 const uint32_t *__restrict__ gbase, uint32_t *state, uint32_t *ticket, int vals_iota)
{
    constexpr int R = 256, WARPS = OS_THREADS / 32;
    __shared__ uint32_t s_key[OS_TILE], s_val[OS_TILE];
    __shared__ uint32_t wcnt[WARPS][R];
    __shared__ uint32_t s_dstart[R], s_gpos[R];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * R; i += OS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tbase = tile * OS_TILE;
    const uint32_t tile_n = min((uint32_t)OS_TILE, n - tbase);

    const uint32_t wbase = tbase + warp * (32 * OS_ITEMS);
    uint32_t k[OS_ITEMS];
    uint32_t rk2[OS_ITEMS / 2];                 // ranks inside (warp, digit), two 16-bit values per register
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        k[r] = idx < n ? keys_in[idx] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const bool valid = wbase + r * 32 + lane < n;
        const uint32_t d = (k[r] >> shift) & (R - 1);
        // out-of-range lanes get a private pseudo-digit so they never join a real group
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (R + lane));
        const uint32_t rank = __popc(m & ((1u << lane) - 1u));
        const uint32_t prev = valid ? wcnt[warp][d] : 0u;
        __syncwarp();
        if (valid && rank == 0) wcnt[warp][d] = prev + __popc(m);
        __syncwarp();
        if (r & 1) rk2[r >> 1] |= (prev + rank) << 16; else rk2[r >> 1] = prev + rank;
    }
    __syncthreads();
    // thread d: digit d's count in this tile, exclusive offsets of the warps inside the digit
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const uint32_t c = wcnt[w][tid];
        wcnt[w][tid] = count;
        count += c;
    }
    // publish the tile's digit count, then look back for the digit's count in all preceding tiles
    uint32_t *mine = state + (size_t)tile * R + tid;
    *(volatile uint32_t *)mine = count | (tile == 0 ? OS_FLAG_PREFIX : OS_FLAG_AGG);
    uint32_t total;
    const uint32_t dstart = block_exclusive_scan(count, &total);
    s_dstart[tid] = dstart;
    uint32_t prefix = 0;
    if (tile > 0) {
        // four predecessors per round trip; stop at the first one that already knows its prefix
        int t = (int)tile - 1;
        bool done = false;
        while (!done) {
            uint32_t st[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) st[q] = t - q >= 0 ? *(const volatile uint32_t *)(state + (size_t)(t - q) * R + tid) : OS_FLAG_PREFIX;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (done) break;
                if ((st[q] >> 30) == 0u) break;            // not published yet: poll again from this tile
                prefix += st[q] & OS_VALUE_MASK;
                --t;
                if ((st[q] >> 30) == 2u) done = true;
            }
        }
        *(volatile uint32_t *)mine = (prefix + count) | OS_FLAG_PREFIX;
    }
    s_gpos[tid] = gbase[tid] + prefix;
    __syncthreads();
    // stage in sorted order (values are loaded only now: fewer live registers while ranking)
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const uint32_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = (k[r] >> shift) & (R - 1);
            const uint32_t rank = (r & 1) ? (rk2[r >> 1] >> 16) : (rk2[r >> 1] & 0xffffu);
            const uint32_t lp = s_dstart[d] + wcnt[warp][d] + rank;
            s_key[lp] = k[r];
            s_val[lp] = vals_iota ? idx : vals_in[idx];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < tile_n; i += OS_THREADS) {
        const uint32_t key = s_key[i];
        const uint32_t d = (key >> shift) & (R - 1);
        const uint32_t pos = s_gpos[d] + (i - s_dstart[d]);
        keys_out[pos] = key;
        vals_out[pos] = s_val[i];
    }
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

size_t mb_sort_workspace_bytes(uint32_t n)
{
    const size_t ntiles = ((size_t)n + OS_TILE - 1) / OS_TILE;
    // [4 passes][256] digit histograms, 4 tickets, [4 passes][ntiles][256] look-back states
    return mb_align_up((4 * 256 + 64) * sizeof(uint32_t)) + mb_align_up(4 * ntiles * 256 * sizeof(uint32_t)) + 256;
}

int mb_sort_pairs(cudaStream_t stream, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b,
                  uint32_t *vals_b, uint32_t n, const uint32_t *n_dev, int key_bits, bool vals_a_is_iota,
                  void *workspace, size_t workspace_bytes, uint32_t **keys_out, uint32_t **vals_out)
{
    *keys_out = keys_a;
    *vals_out = vals_a;
    if (n == 0) return MB_OK;
    MB_REQUIRE(n_dev == nullptr, "mb_sort_pairs: device-side counts are not supported");
    MB_REQUIRE(n < (1u << 30), "mb_sort_pairs: too many elements");
    MB_REQUIRE(workspace_bytes >= mb_sort_workspace_bytes(n), "sort workspace too small");
    const int passes = (key_bits + 7) / 8;
    MB_REQUIRE(passes >= 1 && passes <= 4, "mb_sort_pairs: bad key width");
    const size_t ntiles = ((size_t)n + OS_TILE - 1) / OS_TILE;
    MbArena arena(workspace, workspace_bytes);
    uint32_t *head = arena.take<uint32_t>(4 * 256 + 64);      // histograms, then the tickets
    uint32_t *ghist = head, *tickets = head + 4 * 256;
    uint32_t *state = arena.take<uint32_t>(4 * ntiles * 256);
    MB_CHECK_CUDA(cudaMemsetAsync(head, 0, (4 * 256 + 64) * sizeof(uint32_t), stream));
    MB_CHECK_CUDA(cudaMemsetAsync(state, 0, (size_t)passes * ntiles * 256 * sizeof(uint32_t), stream));
    size_t hblocks = ((size_t)n + 256 * 16 - 1) / (256 * 16);
    if (hblocks > (size_t)MB_NUM_SMS * 8) hblocks = (size_t)MB_NUM_SMS * 8;
    k_radix_hist_all<<<(unsigned)hblocks, 256, 0, stream>>>(keys_a, n, passes, ghist);
    MB_LAUNCHED();
    k_radix_bases<<<passes, 256, 0, stream>>>(ghist);
    MB_LAUNCHED();

    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    bool iota = vals_a_is_iota;
    for (int p = 0; p < passes; ++p) {
        k_radix_onesweep<<<(unsigned)ntiles, OS_THREADS, 0, stream>>>(kin, vin, kout, vout, n, 8 * p, ghist + p * 256,
                                                                    state + (size_t)p * ntiles * 256, tickets + p,
                                                                    iota ? 1 : 0);
        MB_LAUNCHED();
        iota = false;
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    *keys_out = kin;
    *vals_out = vin;
    return MB_OK;
}
