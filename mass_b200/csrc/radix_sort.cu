// Stable LSD radix sort of (key, value) u32 pairs and a device-wide exclusive scan.
//
// Used to order splat contributions by voxel key.  Stability is what makes the voxel reduce
// reproduce the reference's summation order (slot-major, then pixel order:
// /root/reference/mass/utils/projection.py:294-298, 319-323, 349-351) without float atomics.
//
// 8-bit digits, one contiguous block of the input per resident CTA (see below).  Inside a tile each warp
// owns 256 consecutive elements and ranks them in 8 rounds of 32 (lanes with equal digits found by ballots),
// so the order (tile, warp, round, lane) is the input order.
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

// Small inputs (digit tables, per-tile item counts: <= 2^20 values) in ONE launch: tiles of 2048 values, the
// prefix of a tile by a decoupled look-back over the tiles before it (one zeroed word per tile: value | flag in the
// top two bits).  At most 512 CTAs, all resident at once, so a tile never waits for a CTA that has not started.
constexpr uint32_t SCAN1_MAX = 1u << 20;
constexpr uint32_t LB_AGG = 1u << 30, LB_PREFIX = 2u << 30, LB_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *state)
{
    __shared__ uint32_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile * SCAN_TILE + tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= n && ((uintptr_t)in % 16 == 0)) {
        const uint4 a = *(const uint4 *)(in + base), b = *(const uint4 *)(in + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = base + i < n ? in[base + i] : 0u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(s, &total);
    if (tid < 32) {
        // warp 0: publish the tile total, then sum the totals of the tiles before (32 per step)
        uint32_t *mine = state + tile;
        if (lane == 0) *(volatile uint32_t *)mine = total | (tile == 0 ? LB_PREFIX : LB_AGG);
        uint32_t prefix = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            for (;;) {
                const int idx = t - lane;
                const uint32_t st = idx >= 0 ? *(const volatile uint32_t *)(state + idx) : LB_PREFIX;
                const uint32_t flag = st >> 30;
                const uint32_t pm = __ballot_sync(0xffffffffu, flag == 2u), zm = __ballot_sync(0xffffffffu, flag == 0u);
                const int firstp = pm ? __ffs(pm) - 1 : 32, firstz = zm ? __ffs(zm) - 1 : 32;
                const int take = firstp < firstz ? firstp + 1 : firstz;
                uint32_t x = lane < take ? (st & LB_MASK) : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
                prefix += x;
                if (firstp < firstz) break;
                t -= take;
            }
            if (lane == 0) *(volatile uint32_t *)mine = (prefix + total) | LB_PREFIX;
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t run = s_prefix + excl;
    if (base + SCAN_ITEMS <= n && ((uintptr_t)out % 16 == 0)) {
        uint4 a, b;
        a.x = run; a.y = a.x + v[0]; a.z = a.y + v[1]; a.w = a.z + v[2];
        b.x = a.w + v[3]; b.y = b.x + v[4]; b.z = b.y + v[5]; b.w = b.z + v[6];
        *(uint4 *)(out + base) = a;
        *(uint4 *)(out + base + 4) = b;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            if (base + i < n) out[base + i] = run;
            run += v[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Radix sort, 8 bits per pass.  The input is cut into one contiguous block per resident CTA (148 SMs x 2);
// per pass:
//   k_radix_block_hist     each CTA counts the digits of its block (runs of equal digits are merged before
//                          they touch the shared-memory histogram) -> table [digit][block]
//   exclusive scan of the table (digit-major): the global position of every (digit, block)
//   k_radix_block_scatter  each CTA walks its block tile by tile (4096 pairs): ranks the tile (equal digits of a
//                          warp round by ballots, per-warp digit counters), stages the pairs through shared
//                          memory in sorted order and writes them out coalesced; thread d carries the running
//                          position of digit d.
// No CTA ever waits for another one.  Stable: the global order inside a digit is (block, tile, warp, round,
// lane) = the input order.
constexpr int OS_ITEMS = 8;
constexpr int OS_THREADS = 512;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
constexpr int OS_BLOCKS = MB_NUM_SMS * 2;
constexpr int HIST_THREADS = 1024;

// elements per CTA: the tiles split evenly over the CTAs, a multiple of the tile
__host__ __device__ inline uint32_t block_share(uint32_t n, uint32_t blocks)
{
    const uint32_t ntiles = (n + OS_TILE - 1) / OS_TILE;
    return ((ntiles + blocks - 1) / blocks) * OS_TILE;
}

__global__ void __launch_bounds__(HIST_THREADS)
k_radix_block_hist(const uint32_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ n_dev,
                   uint32_t per_block, int shift, uint32_t *__restrict__ table)
{
    __shared__ uint32_t h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    if (n_dev) {                                  // the real count lives on the device: same split, computed here
        n = min(n, *n_dev);
        per_block = block_share(n, gridDim.x);
    }
    __syncthreads();
    const uint32_t beg = min(n, blockIdx.x * per_block), end = min(n, beg + per_block);
    // each thread takes 16 consecutive keys (per_block and beg are multiples of 4096)
    for (uint32_t base = beg + threadIdx.x * 16u; base < end; base += HIST_THREADS * 16u) {
        uint32_t k[16];
        if (base + 16u <= end) {
            const uint4 *p = (const uint4 *)(keys + base);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = __ldg(p + q);
                k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) k[q] = base + q < end ? keys[base + q] : 0u;
        }
        const int cnt = (int)min(16u, end - base);
        uint32_t cur = (k[0] >> shift) & 255u, run = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if (q < cnt) {
                const uint32_t d = (k[q] >> shift) & 255u;
                if (d != cur) { atomicAdd(&h[cur], run); cur = d; run = 0; }
                ++run;
            }
        }
        atomicAdd(&h[cur], run);
    }
    __syncthreads();
    if (threadIdx.x < 256) table[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];      // digit-major
}

// Bulk asynchronous copies (the TMA engine's linear mode, SASS UBLKCP): one thread hands a whole contiguous tile to the
// copy engine and the CTA waits on an mbarrier that counts the bytes; no per-thread load instructions at all.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}

// Lanes of the warp that hold the same 8-bit digit (out-of-range lanes never join a group).  match.any takes time
// in proportion to the number of DISTINCT values in the warp: with the low digits of cell keys (close to 32 distinct
// values per round) the ranking loop spent half its time waiting for it (ncu, profiles/r02z: 51 % of the scatter
// kernel's samples on the instruction after the match; the top digit, a handful of values per round, ran 2.3x faster).
// Eight ballots cost the same whatever the digits are, so they rank the low digits; the LAST pass of a sort -- the
// top digit of a spatial key, which neighbouring elements share -- keeps match.any (34 us against 48 us with ballots).
template <bool MATCH_ANY>
__device__ __forceinline__ uint32_t match_digit(uint32_t d, bool valid, int lane)
{
    if (MATCH_ANY) return __match_any_sync(0xffffffffu, valid ? d : (256u + (uint32_t)lane));
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t has = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? has : ~has;
    }
    return valid ? peers : (1u << lane);
}

// Shared memory of k_radix_block_scatter (dynamic): the input buffer of ONE tile (keys, values), which the TMA
// engine refills with the next tile while the current one -- already in registers -- is ranked, and the sorted
// staging area.
struct ScatterSmem {
    uint32_t in_key[OS_TILE], in_val[OS_TILE];
    uint32_t s_key[OS_TILE], s_val[OS_TILE];
    uint32_t wcnt[OS_THREADS / 32][256];
    uint32_t s_dstart[256], s_off[256], s_wsum[8];
    uint64_t bar;                                // counts the bytes of the input buffer's bulk copies
};

// 16 warps of 8 rounds each per tile (two CTAs per SM: 32 warps): the ranking of a warp is a chain of dependent
// shared-memory read-modify-writes, so it is the number of independent chains per SM that sets the pace.
template <bool MATCH_ANY>
__global__ void __launch_bounds__(OS_THREADS, 2)
k_radix_block_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                      uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                      const uint32_t *__restrict__ n_dev, uint32_t per_block, int shift,
                      const uint32_t *__restrict__ table, int vals_iota)
{
    constexpr int R = 256, WARPS = OS_THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScatterSmem &S = *reinterpret_cast<ScatterSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (n_dev) {
        n = min(n, *n_dev);
        per_block = block_share(n, gridDim.x);
    }
    const uint32_t beg = min(n, blockIdx.x * per_block), end = min(n, beg + per_block);
    uint32_t gpos = tid < R ? table[tid * gridDim.x + blockIdx.x] : 0u;     // running position of digit `tid` in this block
    for (int i = tid; i < WARPS * R; i += OS_THREADS) (&S.wcnt[0][0])[i] = 0;
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // a tile (16 KB of keys + 16 KB of values, contiguous) is fetched by ONE thread as two bulk copies, a multiple of
    // 16 bytes each (beg is a multiple of the tile; the array tails are padded by the allocation: a piece that starts
    // before `end` is always in bounds of the 16-byte aligned buffer)
    auto fetch = [&](uint32_t tbase) {
        if (tid == 0) {
            const uint32_t cnt = min((uint32_t)OS_TILE, end - tbase);
            const uint32_t bytes = (cnt * 4u + 15u) & ~15u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier reads of the buffer are done
            mbar_expect_tx(&S.bar, vals_iota ? bytes : 2u * bytes);
            bulk_g2s(&S.in_key[0], keys_in + tbase, bytes, &S.bar);
            if (!vals_iota) bulk_g2s(&S.in_val[0], vals_in + tbase, bytes, &S.bar);
        }
    };
    if (beg < end) fetch(beg);
    uint32_t phase = 0u;
    for (uint32_t tbase = beg; tbase < end; tbase += OS_TILE) {
        mbar_wait(&S.bar, phase);                 // this tile's input has landed in shared memory
        phase ^= 1u;
        const uint32_t tile_n = min((uint32_t)OS_TILE, end - tbase);
        const uint32_t wloc = warp * (32 * OS_ITEMS);
        uint32_t k[OS_ITEMS], v[OS_ITEMS];
        uint32_t rk2[OS_ITEMS / 2];             // ranks inside (warp, digit), two 16-bit values per register
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const uint32_t loc = wloc + r * 32 + lane;
            k[r] = loc < tile_n ? S.in_key[loc] : 0xffffffffu;
            v[r] = vals_iota ? tbase + loc : S.in_val[loc];
        }
        __syncthreads();                          // the input buffer is in registers; the staging area and wcnt are free
        if (tbase + OS_TILE < end) fetch(tbase + OS_TILE);      // next tile, while this one is ranked
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const bool valid = wloc + r * 32 + lane < tile_n;
            const uint32_t d = (k[r] >> shift) & (R - 1);
            const uint32_t m = match_digit<MATCH_ANY>(d, valid, lane);
            const uint32_t rank = __popc(m & ((1u << lane) - 1u));
            // the first lane of every group adds the group to the warp's digit counter and hands the old count to
            // its peers.  Only atomics touch the counters in this loop and a warp's shared-memory operations
            // execute in order, so the rounds need no barrier between them and their counter updates pipeline
            // instead of forming a chain of read - barrier - write - barrier.
            uint32_t prev = 0;
            if (valid && rank == 0) prev = atomicAdd(&S.wcnt[warp][d], (uint32_t)__popc(m));
            prev = __shfl_sync(0xffffffffu, prev, __ffs(m) - 1);
            if (r & 1) rk2[r >> 1] |= (prev + rank) << 16; else rk2[r >> 1] = prev + rank;
        }
        __syncthreads();
        // thread d < 256: digit d's count in this tile, exclusive offsets of the warps inside the digit
        uint32_t count = 0, inc = 0;
        if (tid < R) {
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = S.wcnt[w][tid];
                S.wcnt[w][tid] = count;
                count += c;
            }
            inc = warp_inclusive_scan(count, lane);
            if (lane == 31) S.s_wsum[warp] = inc;
        }
        __syncthreads();
        if (tid < R) {
            uint32_t dstart = inc - count;
            for (int w = 0; w < warp; ++w) dstart += S.s_wsum[w];
            S.s_dstart[tid] = dstart;
            S.s_off[tid] = gpos - dstart;         // global position of staged element i of digit d: s_off[d] + i
            gpos += count;
        }
        __syncthreads();
        // stage in sorted order
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const uint32_t loc = wloc + r * 32 + lane;
            if (loc < tile_n) {
                const uint32_t d = (k[r] >> shift) & (R - 1);
                const uint32_t rank = (r & 1) ? (rk2[r >> 1] >> 16) : (rk2[r >> 1] & 0xffffu);
                const uint32_t lp = S.s_dstart[d] + S.wcnt[warp][d] + rank;
                S.s_key[lp] = k[r];
                S.s_val[lp] = v[r];
            }
        }
        __syncthreads();
        for (int i = tid; i < WARPS * R; i += OS_THREADS) (&S.wcnt[0][0])[i] = 0;      // (for the next tile)
        for (uint32_t i = tid; i < tile_n; i += OS_THREADS) {
            const uint32_t key = S.s_key[i];
            const uint32_t d = (key >> shift) & (R - 1);
            const uint32_t pos = S.s_off[d] + i;
            keys_out[pos] = key;
            vals_out[pos] = S.s_val[i];
        }
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

// values of [0, 2^30) only (tile totals carry two flag bits): n <= 2^20 values whose sum stays below 2^30.  `state`:
// mb_scan_state_words(n) words the caller has zeroed (they are left non-zero).
size_t mb_scan_state_words(uint32_t n) { return ((size_t)n + SCAN_TILE - 1) / SCAN_TILE + 4; }

int mb_exclusive_scan_small(cudaStream_t stream, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *zeroed_state)
{
    if (n == 0) return MB_OK;
    MB_REQUIRE(n <= SCAN1_MAX, "mb_exclusive_scan_small: too many values");
    k_scan_lookback<<<(n + SCAN_TILE - 1) / SCAN_TILE, SCAN_THREADS, 0, stream>>>(in, out, n, zeroed_state);
    MB_LAUNCHED();
    return MB_OK;
}

size_t mb_sort_workspace_bytes(uint32_t n)
{
    (void)n;
    const size_t table = (size_t)256 * OS_BLOCKS;
    return mb_align_up(table * sizeof(uint32_t)) + mb_align_up(4 * mb_scan_state_words((uint32_t)table) * sizeof(uint32_t)) + 256;
}

int mb_sort_pairs(cudaStream_t stream, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b,
                  uint32_t *vals_b, uint32_t n, const uint32_t *n_dev, int key_bits, bool vals_a_is_iota,
                  void *workspace, size_t workspace_bytes, uint32_t **keys_out, uint32_t **vals_out)
{
    *keys_out = keys_a;
    *vals_out = vals_a;
    if (n == 0) return MB_OK;
    MB_REQUIRE(workspace_bytes >= mb_sort_workspace_bytes(n), "sort workspace too small");
    const int passes = (key_bits + 7) / 8;
    MB_REQUIRE(passes >= 1 && passes <= 4, "mb_sort_pairs: bad key width");
    const size_t ntiles = ((size_t)n + OS_TILE - 1) / OS_TILE;
    // n is an upper bound when n_dev is given: the CTAs then split min(n, *n_dev) among themselves
    const int blocks = ntiles < (size_t)OS_BLOCKS ? (int)ntiles : OS_BLOCKS;
    const uint32_t per_block = block_share(n, (uint32_t)blocks);
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_radix_block_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ScatterSmem)));
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_radix_block_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(ScatterSmem)));
    MbArena arena(workspace, workspace_bytes);
    const uint32_t table_n = 256u * (uint32_t)blocks;
    uint32_t *table = arena.take<uint32_t>((size_t)256 * OS_BLOCKS);
    // look-back words of the table scans, one set per pass, zeroed by ONE memset (the sum of a table is n < 2^30)
    const size_t state_words = mb_scan_state_words(256u * OS_BLOCKS);
    uint32_t *state = arena.take<uint32_t>(4 * state_words);
    MB_REQUIRE(n < (1u << 30), "mb_sort_pairs: too many pairs");
    MB_CHECK_CUDA(cudaMemsetAsync(state, 0, 4 * state_words * sizeof(uint32_t), stream));

    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    bool iota = vals_a_is_iota;
    for (int p = 0; p < passes; ++p) {
        k_radix_block_hist<<<blocks, HIST_THREADS, 0, stream>>>(kin, n, n_dev, per_block, 8 * p, table);
        MB_LAUNCHED();
        int rc = mb_exclusive_scan_small(stream, table, table, table_n, state + (size_t)p * state_words);
        if (rc) return rc;
        if (p == passes - 1 && passes > 1)
            k_radix_block_scatter<true><<<blocks, OS_THREADS, sizeof(ScatterSmem), stream>>>(kin, vin, kout, vout, n, n_dev,
                                                                                             per_block, 8 * p, table, iota ? 1 : 0);
        else
            k_radix_block_scatter<false><<<blocks, OS_THREADS, sizeof(ScatterSmem), stream>>>(kin, vin, kout, vout, n, n_dev,
                                                                                              per_block, 8 * p, table, iota ? 1 : 0);
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
