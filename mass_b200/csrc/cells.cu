// Batched mapping pipeline (sm_100a): T frames fused per call, in frame order, without float atomics.
//
// The per-frame update of a voxel is affine in the old row:
//     new = a_t * old + (alpha / W_t) * sum_i w_i^2 f_i,   a_t = 1 - alpha * S2_t / W_t,
//     W_t = sum_i w_i, S2_t = sum_i w_i^2 over the contributions i of frame t to that voxel.
// Composing the T frames of a call gives, per voxel,
//     new = (prod_t a_t) * old + sum_i [ w_i^2 * (alpha / W_t(i)) * prod_{s > t(i)} a_s ] * f_i
// i.e. a scalar pass that needs the frame order (W, S2, a, suffix products: no feature data) and a
// feature pass that is a plain weighted sum of feature rows -- order-free, so it can be arranged
// for the memory system instead of for the frame order.
//
// The 8 contributions of a pixel go to the 2x2x2 voxels above its lower-corner "cell"
// (/root/reference/mass/utils/projection.py:280-323), so the unit of sorting is the cell, not the voxel: one
// feature-row load feeds 8 accumulators.  Neighbouring pixels of a frame fall into the same cell, so the pixels
// are first grouped inside 32 x 8 image tiles; what gets sorted is the list of ITEMS (cell, tile, <= 16
// contiguous grouped pixels), ~3.7 times shorter than the pixel list.
//
//   K1a k_cell_voxelise   pixel -> {cell key, 3 in-voxel ratios} in image order
//   K1b k_tile_group      one warp per tile: group the pixels by cell in shared memory -> grouped pixel records,
//                         per-tile item lists;  K1c k_tile_compact: dense item list in (frame, tile) order
//   --  stable radix sort of (cell key, item): inside a cell the items stay in (frame, tile) order
//   K2  k_cell_index      one sweep over the sorted items: heads of cells, (cell, frame) segments and accumulate
//                         runs, their ranks (decoupled look-back), unique cell list, segment list, dense cell
//                         table, touched-voxel bitmap
//   K4  k_vox_count/emit  ordered list of touched voxels
//   K5  k_seg_sums        per segment and slot: W, S2                                  (scalars)
//   K6a k_voxel_sources   per touched voxel: segment and run ranges of its <= 8 source cells
//   K6  k_voxel_scalars   per touched voxel: per-frame W, S2 over its sources, backward product scan ->
//                         coefficient g = (alpha / W_t) * prod_{s>t} a_s per (segment, slot), A = prod a
//   K7  k_cell_accumulate per run of same-cell pixels: 8 rows P_k = sum_i w_ik^2 g_k f_i   (the hot loop)
//   K8  k_voxel_apply     per touched voxel: map = A * map + sum over its cells' runs of P
//
// Every sum runs in an order fixed by the tile grouping and the stable sort, so results are bit-reproducible
// run to run.
// Weights follow the reference's fp32 operation order; occupancy is bit-exact and values differ from
// the reference CPU path by fp32 re-association only (<= 1e-5 relative).
#include <cstdlib>
#include <mutex>
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.cuh"
#include "geometry.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ACC_THREADS = 256;

// Cells live on the map grid extended by one at the low end of each axis: a pixel in voxel 0 with
// ratio < 0.5 has its lower corner at -1 (the reference clamps that neighbour onto voxel 0).
struct CellGrid {
    int S0, S1, S2;          // map dims (y flipped, x, z)
    int E0, E1, E2;          // S + 1 cell coordinates per axis: e = corner + 1 in [0, S]
    uint32_t invalid;        // key of invalid pixels = E0*E1*E2 (sorts last)
};

__host__ __device__ inline CellGrid make_cells(int S0, int S1, int S2)
{
    CellGrid g;
    g.S0 = S0; g.S1 = S1; g.S2 = S2;
    g.E0 = S0 + 1; g.E1 = S1 + 1; g.E2 = S2 + 1;
    g.invalid = (uint32_t)g.E0 * (uint32_t)g.E1 * (uint32_t)g.E2;
    return g;
}

__device__ __forceinline__ uint32_t cell_key(const CellGrid &g, int e0, int e1, int e2)
{
    return ((uint32_t)e0 * (uint32_t)g.E1 + (uint32_t)e1) * (uint32_t)g.E2 + (uint32_t)e2;
}

// the 8 splat weights of a pixel, slot k = (d0 << 2) | (d1 << 1) | d2 (projection.py:300-323)
__device__ __forceinline__ void splat_weights(const uint4 &rec, float (&w)[8])
{
    const float q[3] = { __uint_as_float(rec.x), __uint_as_float(rec.y), __uint_as_float(rec.z) };
    float wl[3], wu[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const bool low = q[a] < 0.5f;
        wl[a] = low ? __fsub_rn(0.5f, q[a]) : __fsub_rn(1.5f, q[a]);
        wu[a] = low ? __fadd_rn(q[a], 0.5f) : __fsub_rn(q[a], 0.5f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        w[k] = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn((k & 4) ? wu[0] : wl[0], (k & 2) ? wu[1] : wl[1]),
                                          (k & 1) ? wu[2] : wl[2]));
}

// which slots of the cell at extended coordinate e = v + a (a in {0,1} per axis) land on voxel v:
// per axis, slot bit d lands on clamp(e - 1 + d, 0, S - 1)
__device__ __forceinline__ uint32_t axis_slots(int v, int a, int S)
{
    // a == 0: corner v-1: d=1 -> v always; d=0 -> clamp(v-1) == v only for v == 0
    // a == 1: corner v:   d=0 -> v always; d=1 -> clamp(v+1) == v only for v == S-1
    return a == 0 ? (2u | (v == 0 ? 1u : 0u)) : (1u | (v == S - 1 ? 2u : 0u));
}

__device__ __forceinline__ uint32_t slot_mask(int v0, int v1, int v2, int src, const CellGrid &g)
{
    const uint32_t m0 = axis_slots(v0, (src >> 2) & 1, g.S0), m1 = axis_slots(v1, (src >> 1) & 1, g.S1),
                   m2 = axis_slots(v2, src & 1, g.S2);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (((m0 >> ((k >> 2) & 1)) & 1u) && ((m1 >> ((k >> 1) & 1)) & 1u) && ((m2 >> (k & 1)) & 1u)) m |= 1u << k;
    return m;
}

// index of `key` in the ascending unique-cell list, or -1
__device__ __forceinline__ int find_cell(const uint32_t *__restrict__ ucell, uint32_t ncells, uint32_t key)
{
    uint32_t lo = 0, hi = ncells;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(ucell + mid) < key) lo = mid + 1; else hi = mid;
    }
    return (lo < ncells && __ldg(ucell + lo) == key) ? (int)lo : -1;
}

// ---------------------------------------------------------------------------------------------
// K1: after the per-pixel voxelisation, one WARP per 32 x 8 pixel tile of one frame groups the tile's 256 pixels by
// cell inside shared memory (one image row of the tile per round).  Neighbouring pixels fall into the same cell (a
// (cell, frame) segment averages ~4 pixels), so what leaves the tile is an ITEM list -- (cell key, tile, first
// slot, length <= 16) -- a few times shorter than the pixel list; the pixel records are written in grouped
// order, so an item's pixels are contiguous.  Grouping: per round, match.any finds the lanes that share a
// key; the first of them looks the key up in (or adds it to) the warp's 256-slot hash table; a pixel's rank
// inside its group is the group's count so far plus its rank among the round's matching lanes, i.e. the
// pixels of a group are in (row, column) order.  Which slot a key gets depends on the probing races of one
// round's leaders, so the ORDER OF GROUPS inside a tile may differ run to run; nothing downstream depends on
// it (items are sorted by cell key, and a cell has one group per tile).
//   item value = tile << 12 | first slot << 4 | (length - 1);  pixel record = {ratio0, ratio1, ratio2, pixel in tile}
constexpr int TILE_W = 32, TILE_H = 8, TILE_PIX = TILE_W * TILE_H;
#ifndef MB_WHASH
#define MB_WHASH 256
#endif
#ifndef MB_TG_MINB
#define MB_TG_MINB 4
#endif
#ifndef MB_VOX_MINB
#define MB_VOX_MINB 5
#endif
#ifndef MB_TASK_ITEMS
#define MB_TASK_ITEMS 128
#endif
#ifndef MB_ACC_U21
#define MB_ACC_U21 8
#endif
#ifndef MB_SEG_LONG_ITEMS
#define MB_SEG_LONG_ITEMS 4
#endif
#ifndef MB_SEG_MINB
#define MB_SEG_MINB 5
#endif
#ifndef MB_VS_MINB
#define MB_VS_MINB 4
#endif
constexpr int WHASH = MB_WHASH;                  // hash slots per warp (>= pixels of a tile)
constexpr int WHASH_BITS = MB_WHASH == 256 ? 8 : MB_WHASH == 512 ? 9 : MB_WHASH == 1024 ? 10 : -1;
static_assert(WHASH_BITS > 0, "MB_WHASH must be 256, 512 or 1024");
constexpr int ITEM_MAX = 16;                     // pixels per item
constexpr int SEG_LONG_ITEMS = MB_SEG_LONG_ITEMS; // segments of more items than this are summed by a warp, not a thread
constexpr uint32_t HASH_EMPTY = 0xffffffffu;
constexpr int TASK_ITEMS = MB_TASK_ITEMS;        // most items of an accumulate run (one warp task)
constexpr int TASK_WORDS = TASK_ITEMS / 32;      // items per lane when a warp loads a task
constexpr uint32_t MAX_TILES = 1u << 20;         // tile ids are 20 bits of the item value

__device__ __forceinline__ uint32_t item_tile(uint32_t v) { return v >> 12; }
__device__ __forceinline__ uint32_t item_pos(uint32_t v) { return (v >> 4) & (TILE_PIX - 1); }
__device__ __forceinline__ uint32_t item_len(uint32_t v) { return (v & 15u) + 1u; }

struct TileGeom {
    int H, W;                // camera
    int tiles_x, tpf;        // tiles per image row, tiles per frame
    uint32_t tpf_magic, tx_magic;     // ceil(2^32 / d) for the two divisors (0: divide)
};

// x / d for x < 2^20 (tile ids) as one multiply-high: with m = ceil(2^32 / d) and e = m * d - 2^32 < d the product
// x * m / 2^32 exceeds x / d by x * e / (d * 2^32) < 1 / d as long as x * e < 2^32, which d <= 4096 guarantees.
__host__ __device__ inline uint32_t div_magic(uint32_t d)
{
    return d >= 2u && d <= 4096u ? (uint32_t)((((uint64_t)1 << 32) + d - 1) / d) : 0u;
}

__device__ __forceinline__ uint32_t fast_div(uint32_t x, uint32_t d, uint32_t magic)
{
    return magic ? __umulhi(x, magic) : x / d;
}

__host__ __device__ inline TileGeom make_tiles(int H, int W)
{
    TileGeom t;
    t.H = H; t.W = W;
    t.tiles_x = (W + TILE_W - 1) / TILE_W;
    t.tpf = t.tiles_x * ((H + TILE_H - 1) / TILE_H);
    t.tpf_magic = div_magic((uint32_t)t.tpf);
    t.tx_magic = div_magic((uint32_t)t.tiles_x);
    return t;
}

constexpr uint32_t IDX_AGG = 1u << 30, IDX_PREFIX = 2u << 30, IDX_MASK = (1u << 30) - 1u;

// Decoupled look-back over one counter per tile: publishes `total` for `tile` and returns the sum of the totals of
// all tiles before it.  Tiles take their index from a ticket, so a tile only ever waits for tiles that started
// before it.  state: one zeroed word per tile (value | flag in the top two bits).  Called by ONE full warp with
// the same arguments in every lane; the warp inspects 32 predecessors per step.
__device__ __forceinline__ uint32_t lookback_prefix(uint32_t *state, uint32_t tile, uint32_t total, int lane)
{
    uint32_t *mine = state + tile;
    if (lane == 0) *(volatile uint32_t *)mine = total | (tile == 0 ? IDX_PREFIX : IDX_AGG);
    uint32_t prefix = 0;
    if (tile > 0) {
        int t = (int)tile - 1;                      // nearest predecessor not yet summed
        for (;;) {
            const int idx = t - lane;
            const uint32_t st = idx >= 0 ? *(const volatile uint32_t *)(state + idx) : IDX_PREFIX;
            const uint32_t flag = st >> 30;
            const uint32_t pm = __ballot_sync(FULL, flag == 2u), zm = __ballot_sync(FULL, flag == 0u);
            const int firstp = pm ? __ffs(pm) - 1 : 32, firstz = zm ? __ffs(zm) - 1 : 32;
            const int take = firstp < firstz ? firstp + 1 : firstz;      // lanes [0, take) are summed now
            uint32_t v = lane < take ? (st & IDX_MASK) : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
            prefix += v;
            if (firstp < firstz) break;
            t -= take;
        }
        if (lane == 0) *(volatile uint32_t *)mine = (prefix + total) | IDX_PREFIX;
    }
    return prefix;
}

constexpr int VOX_PPT = 4;                       // pixels per thread of k_cell_voxelise

// K1a: grid = (pixel blocks, frames): pixel -> {cell key (or >= 0xffffffe0: invalid), 3 in-voxel ratios}, in
// image order.  Pure per-pixel math at full occupancy; the grouping kernel below re-reads it.
__global__ void __launch_bounds__(256, MB_VOX_MINB)
k_cell_voxelise(const float *__restrict__ rays, const float *__restrict__ depth, const float *__restrict__ pose,
                uint32_t npix, const float *__restrict__ bins_x, int nx, const float *__restrict__ bins_y, int ny,
                const float *__restrict__ bins_z, int nz, CellGrid g, float min_d, float max_d,
                uint4 *__restrict__ pix, uint32_t *__restrict__ counters, bool keep_error_bits)
{
    __shared__ float P[12], spacing[6];
    const uint32_t t = blockIdx.y;
    if (threadIdx.x < 12) P[threadIdx.x] = pose[(size_t)t * 12 + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        const int a = threadIdx.x - 32;
        const float *bb = a == 0 ? bins_x : a == 1 ? bins_y : bins_z;
        const int nb = a == 0 ? nx : a == 1 ? ny : nz;
        spacing[2 * a] = __ldg(bb);
        spacing[2 * a + 1] = bins_scale(bb, nb);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < MB_NUM_COUNTERS && !(keep_error_bits && threadIdx.x == MB_CNT_ERROR))
        counters[threadIdx.x] = 0;
    __syncthreads();
    // VOX_PPT pixels per thread (a CTA covers 256 * VOX_PPT consecutive pixels of one frame): the loads of all of them
    // are issued before the first is used, and the per-CTA prologue above is paid once for all of them
    float ray[VOX_PPT][3], dep[VOX_PPT];
    const uint32_t p0 = blockIdx.x * (256u * VOX_PPT) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < VOX_PPT; ++i) {
        const uint32_t p = p0 + 256u * i;
        if (p < npix) {
            ray[i][0] = __ldg(rays + 3 * (size_t)p);
            ray[i][1] = __ldg(rays + 3 * (size_t)p + 1);
            ray[i][2] = __ldg(rays + 3 * (size_t)p + 2);
            dep[i] = __ldg(depth + (size_t)t * npix + p);
        }
    }
#pragma unroll
    for (int i = 0; i < VOX_PPT; ++i) {
        const uint32_t p = p0 + 256u * i;
        if (p >= npix) break;
        float r0, r1, r2;
        orient(P, ray[i][0], ray[i][1], ray[i][2], r0, r1, r2);
        const BinResult b = bin_point_fast(bins_x, nx, bins_y, ny, bins_z, nz, spacing, P[9], P[10], P[11], r0, r1, r2,
                                           dep[i], min_d, max_d);
        uint4 out = make_uint4(0xffffffffu, 0u, 0u, 0u);
        if (b.ok) {
            // map axes are (y flipped, x, z) = input axes (1, 0, 2): base_projection_layer.py:339
            const float q0 = b.q1, q1 = b.q0, q2 = b.q2;
            const int e0 = q0 < 0.5f ? b.i1 : b.i1 + 1;       // lower corner + 1
            const int e1 = q1 < 0.5f ? b.i0 : b.i0 + 1;
            const int e2 = q2 < 0.5f ? b.i2 : b.i2 + 1;
            out = make_uint4(cell_key(g, e0, e1, e2), __float_as_uint(q0), __float_as_uint(q1), __float_as_uint(q2));
        }
        pix[(size_t)t * npix + p] = out;
    }
}

struct __align__(16) WarpTile {
    uint4 px[TILE_PIX];                          // the tile's pixels {key, 3 ratios}, image order (bulk-copied in)
    uint32_t hkey[WHASH];
    uint32_t cnt[WHASH];
    uint16_t start[WHASH];
    uint64_t bar;                                // counts the bytes of the tile's bulk copies
};

__device__ __forceinline__ uint32_t tile_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K1b: one warp per 32 x 8 tile groups the tile's pixels by cell (see above) and writes the grouped records
// and the tile's items.  The tile (8 image rows of 512 contiguous bytes) comes in through the TMA engine's linear mode:
// lane 0 issues one bulk copy per row and the warp waits on its own mbarrier, the hash table being cleared meanwhile;
// keys and records are then read from shared memory.
__global__ void __launch_bounds__(256, MB_TG_MINB)
k_tile_group(const uint4 *__restrict__ pix, TileGeom tg, uint32_t ntiles, uint4 *__restrict__ rec,
             uint32_t *__restrict__ tkey, uint32_t *__restrict__ tval, uint32_t *__restrict__ tcount)
{
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    WarpTile *s_w = reinterpret_cast<WarpTile *>(tile_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x * 8u + warp;
    if (tile >= ntiles) return;
    WarpTile &S = s_w[warp];
    const uint32_t frame = fast_div(tile, (uint32_t)tg.tpf, tg.tpf_magic), tif = tile - frame * (uint32_t)tg.tpf;
    const uint32_t tyo = fast_div(tif, (uint32_t)tg.tiles_x, tg.tx_magic);
    const int y0 = (int)tyo * TILE_H, x0 = (int)(tif - tyo * (uint32_t)tg.tiles_x) * TILE_W;
    const size_t fbase = (size_t)frame * tg.H * tg.W;
    const int cols = min(TILE_W, tg.W - x0), rows = min(TILE_H, tg.H - y0);
    const uint32_t bar = tile_smem_addr(&S.bar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t row_bytes = (uint32_t)cols * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * (uint32_t)rows) : "memory");
    }
    __syncwarp();
    if (lane < rows)                                   // one row per lane: the eight copies are issued in one go
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tile_smem_addr(&S.px[lane * TILE_W])), "l"(pix + fbase + (size_t)(y0 + lane) * tg.W + x0),
                       "r"((uint32_t)cols * 16u), "r"(bar) : "memory");
    {
        // clear the table with 16-byte stores (WHASH keys and WHASH counters = WHASH / 4 pieces each)
        uint4 *hk = reinterpret_cast<uint4 *>(S.hkey), *ct = reinterpret_cast<uint4 *>(S.cnt);
        const uint4 e4 = make_uint4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY), z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = lane; i < WHASH / 4; i += 32) hk[i] = e4;
#pragma unroll
        for (int i = lane; i < WHASH / 4; i += 32) ct[i] = z4;
    }
    const uint32_t ltmask = (1u << lane) - 1u;
    __syncwarp();
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    uint32_t pkey[TILE_H];                            // keys of the lane's 8 pixels: column x0 + lane, rows y0 .. y0 + 7
#pragma unroll
    for (int r = 0; r < TILE_H; ++r) {
        pkey[r] = 0xffffffffu;
        if (r < rows && lane < cols) pkey[r] = S.px[r * TILE_W + lane].x;
    }
    uint32_t sr[TILE_H];                              // slot | rank << 16 (slot WHASH: invalid pixel)
    // ---- one grouping round per image row ------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < TILE_H; ++r) {
        const bool valid = pkey[r] < 0xffffffe0u;
        const uint32_t key = valid ? pkey[r] : 0xffffffe0u + lane;   // invalid lanes never match another lane
        const uint32_t m = __match_any_sync(FULL, key);
        const int leader = __ffs(m) - 1;
        uint32_t sl = 0;
        if (valid && lane == leader) {
            uint32_t h = (key * 2654435761u) >> (32 - WHASH_BITS);    // the top bits of the product
            for (;;) {
                const uint32_t old = atomicCAS(&S.hkey[h], HASH_EMPTY, key);
                if (old == HASH_EMPTY || old == key) break;
                h = (h + 1) & (WHASH - 1);
            }
            sl = h;
        }
        // the leader adds its group to the slot's count and hands the old count (and the slot) to its peers: only
        // atomics touch the counts in this loop, and a warp's shared-memory operations execute in order
        uint32_t prev = 0;
        if (valid && lane == leader) prev = atomicAdd(&S.cnt[sl], (uint32_t)__popc(m));
        sl = __shfl_sync(FULL, sl, leader);
        prev = __shfl_sync(FULL, prev, leader);
        sr[r] = valid ? (sl | ((prev + __popc(m & ltmask)) << 16)) : (uint32_t)WHASH;
    }
    __syncwarp();                                     // the counts are read with plain loads from here on
    // ---- lay the groups out: lane owns slots lane, lane + 32, ... ---------------------------------------------
    uint32_t mine = 0;                                // pixels | items << 16 of this lane's slots
#pragma unroll
    for (int j = 0; j < WHASH / 32; ++j) {
        const uint32_t c = S.cnt[lane + 32 * j];
        mine += c | (((c + ITEM_MAX - 1) / ITEM_MAX) << 16);
    }
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    const uint32_t nitems = __shfl_sync(FULL, inc, 31) >> 16;
    uint32_t base = inc - mine;
    const size_t tbase = (size_t)tile * TILE_PIX;
#pragma unroll
    for (int j = 0; j < WHASH / 32; ++j) {
        const int sl = lane + 32 * j;
        const uint32_t c = S.cnt[sl];
        S.start[sl] = (uint16_t)(base & 0xffffu);
        if (c) {
            const uint32_t k = S.hkey[sl], gs = base & 0xffffu, is = base >> 16;
            // nearly every group is one item (<= ITEM_MAX pixels): no loop on that path
            tkey[tbase + is] = k;
            tval[tbase + is] = (tile << 12) | (gs << 4) | (min((uint32_t)ITEM_MAX, c) - 1u);
            if (c > ITEM_MAX) {
                for (uint32_t o = ITEM_MAX, n = 1; o < c; o += ITEM_MAX, ++n) {
                    tkey[tbase + is + n] = k;
                    tval[tbase + is + n] = (tile << 12) | ((gs + o) << 4) | (min((uint32_t)ITEM_MAX, c - o) - 1u);
                }
            }
        }
        base += c | (((c + ITEM_MAX - 1) / ITEM_MAX) << 16);
    }
    __syncwarp();
    // ---- pixel records in grouped order ------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < TILE_H; ++r) {
        const uint32_t sl = sr[r] & 0xffffu;
        if (sl < WHASH) {
            const uint32_t pos = (uint32_t)S.start[sl] + (sr[r] >> 16);
            const uint4 px = S.px[r * TILE_W + lane];
            rec[tbase + pos] = make_uint4(px.y, px.z, px.w, (uint32_t)(r * TILE_W + lane));
        }
    }
    if (lane == 0) tcount[tile] = nitems;
}

// K1c: dense item list in tile order (toff = exclusive scan of the tile counts); one warp per tile
__global__ void __launch_bounds__(256)
k_tile_compact(const uint32_t *__restrict__ tkey, const uint32_t *__restrict__ tval, const uint32_t *__restrict__ tcount,
               const uint32_t *__restrict__ toff, uint32_t ntiles, uint32_t *__restrict__ ikey,
               uint32_t *__restrict__ ival, uint32_t *__restrict__ counters)
{
    const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    const int lane = threadIdx.x & 31;
    const uint32_t cnt = tcount[tile], off = toff[tile];
    const size_t tbase = (size_t)tile * TILE_PIX;
    for (uint32_t j = lane; j < cnt; j += 32) {
        ikey[off + j] = tkey[tbase + j];
        ival[off + j] = tval[tbase + j];
    }
    if (tile == ntiles - 1 && lane == 0) counters[MB_CNT_NVALID] = off + cnt;     // number of items
}

// ---------------------------------------------------------------------------------------------
// K2/K3 fused: one sweep over the sorted list finds the heads of cells, (cell, frame) segments and accumulate
// runs, ranks them and emits the unique cell list, the segment list, the dense cell table and the
// touched-voxel bitmap.  A tile is 2048 sorted items (8 warps x 8 words of 32).  The three ranks of a
// tile's first position come from a decoupled look-back over the preceding tiles' counts (tiles take
// their index from a ticket, so a tile only ever waits for tiles that started before it).
//   The sorted list is the ITEM list (K1): cell head: first item of a cell; segment head: first item of a
//   (cell, frame); run head: cell head or start of an accumulate task (a multiple of TASK_ITEMS items).
constexpr int IDX_WORDS = 8;                      // words per warp
constexpr int IDX_TILE = 256 * IDX_WORDS;         // positions per tile

struct IndexOut {
    uint32_t *smask, *soff;                       // per word of 32 positions: segment-head mask, segment rank before the word
    uint32_t *ucell, *cstart, *cseg;              // per unique cell (+ sentinel)
    uint32_t *seg_start, *seg_frame;              // per segment (+ sentinel)
    uint32_t *bitmap;                             // touched voxels
    int *ctab;                                    // dense cell key -> unique index, or null
    uint32_t *counters;
    uint32_t *state;                              // [tiles][2] look-back words (cells, segments), zeroed
    uint32_t *ticket;                             // zeroed
};

__global__ void __launch_bounds__(256)
k_cell_index(const uint32_t *__restrict__ skey, const uint32_t *__restrict__ sval, const uint32_t *__restrict__ n_dev,
             TileGeom tg, CellGrid g, const IndexOut O)
{
    __shared__ uint32_t s_tile, s_wsum[8][2], s_base[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = *n_dev;                    // number of items
    if (tid == 0) s_tile = atomicAdd(O.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if ((uint64_t)tile * IDX_TILE >= n) return;   // launched for the worst case; later tiles are empty too
    const uint32_t wbase = tile * IDX_TILE + warp * (32 * IDX_WORDS);
    const uint32_t lt = (1u << lane) - 1u;

    uint32_t key[IDX_WORDS], val[IDX_WORDS], cm[IDX_WORDS], sm[IDX_WORDS];
    uint32_t pk = 0xffffffffu, pv = 0;            // element before the warp's first one
    if (lane == 0 && wbase > 0 && wbase <= n) { pk = skey[wbase - 1]; pv = sval[wbase - 1]; }
#pragma unroll
    for (int r = 0; r < IDX_WORDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        key[r] = i < n ? skey[i] : 0xffffffffu;
        val[r] = i < n ? sval[i] : 0u;
    }
    uint32_t cw = 0, sw = 0;                      // heads in this warp's 8 words
#pragma unroll
    for (int r = 0; r < IDX_WORDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        uint32_t kprev = __shfl_up_sync(FULL, key[r], 1), vprev = __shfl_up_sync(FULL, val[r], 1);
        if (lane == 0) { kprev = pk; vprev = pv; }
        pk = __shfl_sync(FULL, key[r], 31);       // (only lane 0's copy is used)
        pv = __shfl_sync(FULL, val[r], 31);
        const bool valid = i < n;
        const bool chead = valid && (i == 0 || kprev != key[r]);
        const uint32_t tpf = (uint32_t)tg.tpf;
        const bool shead = valid && (chead || fast_div(item_tile(vprev), tpf, tg.tpf_magic) != fast_div(item_tile(val[r]), tpf, tg.tpf_magic));
        cm[r] = __ballot_sync(FULL, chead);
        sm[r] = __ballot_sync(FULL, shead);
        cw += __popc(cm[r]); sw += __popc(sm[r]);
    }
    if (lane == 0) { s_wsum[warp][0] = cw; s_wsum[warp][1] = sw; }
    __syncthreads();
    if (tid < 2) {
        // tile total of counter `tid`, published; then the look-back for the tile's base rank (one thread per counter,
        // four predecessors per step: a warp-wide look-back, 32 per step, measured 10-20 us SLOWER here -- the tiles
        // start almost together, so the wider window mostly re-reads words that are not published yet)
        uint32_t total = 0;
        for (int w = 0; w < 8; ++w) total += s_wsum[w][tid];
        uint32_t *mine = O.state + (size_t)tile * 2 + tid;
        *(volatile uint32_t *)mine = total | (tile == 0 ? IDX_PREFIX : IDX_AGG);
        uint32_t prefix = 0;
        if (tile > 0) {
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t st[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) st[q] = t - q >= 0 ? *(const volatile uint32_t *)(O.state + (size_t)(t - q) * 2 + tid) : IDX_PREFIX;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (done) break;
                    if ((st[q] >> 30) == 0u) break;
                    prefix += st[q] & IDX_MASK;
                    --t;
                    if ((st[q] >> 30) == 2u) done = true;
                }
            }
            *(volatile uint32_t *)mine = (prefix + total) | IDX_PREFIX;
        }
        s_base[tid] = prefix;
    }
    __syncthreads();
    uint32_t cb = s_base[0], sb = s_base[1];                      // ranks before this warp's first word
    for (int w = 0; w < warp; ++w) { cb += s_wsum[w][0]; sb += s_wsum[w][1]; }

#pragma unroll
    for (int r = 0; r < IDX_WORDS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const uint32_t w = i >> 5;
        if (lane == 0 && w <= (n >> 5)) { O.smask[w] = sm[r]; O.soff[w] = sb; }
        const bool valid = i < n;
        if (valid) {
            const uint32_t crank = cb + __popc(cm[r] & lt), srank = sb + __popc(sm[r] & lt);
            if ((sm[r] >> lane) & 1u) {
                O.seg_start[srank] = i;
                O.seg_frame[srank] = fast_div(item_tile(val[r]), (uint32_t)tg.tpf, tg.tpf_magic);
            }
            if ((cm[r] >> lane) & 1u) {
                const uint32_t k = key[r];
                O.ucell[crank] = k;
                if (O.ctab != nullptr) O.ctab[k] = (int)crank;
                O.cstart[crank] = i;
                O.cseg[crank] = srank;
                const int e2 = (int)(k % (uint32_t)g.E2);
                const uint32_t t = k / (uint32_t)g.E2;
                const int e1 = (int)(t % (uint32_t)g.E1), e0 = (int)(t / (uint32_t)g.E1);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int v0 = min(max(e0 - 1 + ((q >> 2) & 1), 0), g.S0 - 1);
                    const int v1 = min(max(e1 - 1 + ((q >> 1) & 1), 0), g.S1 - 1);
                    const int v2 = min(max(e2 - 1 + (q & 1), 0), g.S2 - 1);
                    const uint32_t v = ((uint32_t)v0 * (uint32_t)g.S1 + (uint32_t)v1) * (uint32_t)g.S2 + (uint32_t)v2;
                    atomicOr(&O.bitmap[v >> 5], 1u << (v & 31));
                }
            }
        }
        {
            const bool is_last = valid && i + 1 == n;
            if (is_last) {
                const uint32_t le = lane == 31 ? 0xffffffffu : ((2u << lane) - 1u);
                const uint32_t nc = cb + __popc(cm[r] & le), ns = sb + __popc(sm[r] & le);
                O.counters[MB_CNT_NVALID] = i + 1;
                O.counters[MB_CNT_CELLS] = nc;
                O.counters[MB_CNT_SEGS] = ns;
                O.cstart[nc] = i + 1;
                O.cseg[nc] = ns;
                O.seg_start[ns] = i + 1;
            }
        }
        cb += __popc(cm[r]); sb += __popc(sm[r]);
    }
}

// K3: accumulate runs.  A run is a piece of ONE cell's item list, at most TASK_ITEMS items long, cut at the cell's own
// multiples of TASK_ITEMS: it is the accumulate kernel's unit of work (one warp, one set of 8 partial rows).  Runs
// start on cell heads, so the warps that work on neighbouring cells at the same moment are also at the same place
// in those cells' frame-ordered lists -- the feature rows of one frame that share a DRAM atom across a cell border
// then meet in L2 instead of being fetched twice.  One thread per cell; run ranks by a decoupled look-back over
// CTAs of 256 cells.
__global__ void __launch_bounds__(256)
k_cell_runs(const uint32_t *__restrict__ cstart, uint32_t *__restrict__ counters, uint32_t *state, uint32_t *ticket,
            uint32_t *__restrict__ crun, uint32_t *__restrict__ rstart, uint32_t run_limit)
{
    __shared__ uint32_t s_tile, s_wsum[8], s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ncells = counters[MB_CNT_CELLS];
    // persistent CTAs take tiles of 256 cells from a ticket counter (a tile only ever waits for lower tickets)
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile * 256u >= ncells) {
            if (tile == 0 && tid == 0) { crun[0] = 0; rstart[0] = 0; counters[MB_CNT_RUNS] = 0; }  // nothing valid at all
            return;
        }
        const uint32_t u = tile * 256u + tid;
        uint32_t beg = 0, nr = 0;
        if (u < ncells) {
            beg = cstart[u];
            nr = (cstart[u + 1] - beg + TASK_ITEMS - 1) / TASK_ITEMS;
        }
        uint32_t inc = nr;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) total += s_wsum[w];
            const uint32_t prefix = lookback_prefix(state, tile, total, lane);
            if (lane == 0) s_base = prefix;
        }
        __syncthreads();
        uint32_t r0 = s_base + inc - nr;
        for (int w = 0; w < warp; ++w) r0 += s_wsum[w];
        if (u < ncells) {
            crun[u] = r0;
            for (uint32_t j = 0; j < nr; ++j) rstart[r0 + j] = beg + j * TASK_ITEMS;
            if (u == ncells - 1) {
                const uint32_t total = r0 + nr;
                crun[ncells] = total;
                rstart[total] = cstart[ncells];
                counters[MB_CNT_RUNS] = total;
                if (total > run_limit) atomicOr(&counters[MB_CNT_ERROR], 1u);      // never expected: worst_runs() bounds it
            }
        }
    }
}

// K4: ordered list of touched voxels from the bitmap, one sweep: a tile is 2048 bitmap words (8 per thread); the
// position of a tile's first voxel comes from a decoupled look-back over the tiles before it.
constexpr int VOX_WORDS = 8;
constexpr int VOX_TILE = 256 * VOX_WORDS;

__global__ void __launch_bounds__(256)
k_vox_list(const uint32_t *__restrict__ bitmap, uint32_t nwords, uint32_t *state, uint32_t *ticket,
           uint32_t *__restrict__ vlist, uint32_t *__restrict__ counters)
{
    __shared__ uint32_t s_tile, s_wsum[8], s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * VOX_TILE + tid * VOX_WORDS;
    uint32_t m[VOX_WORDS];
    if (base + VOX_WORDS <= nwords) {
        const uint4 a = __ldg((const uint4 *)(bitmap + base)), b = __ldg((const uint4 *)(bitmap + base) + 1);
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < VOX_WORDS; ++k) m[k] = base + k < nwords ? bitmap[base + k] : 0u;
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < VOX_WORDS; ++k) cnt += __popc(m[k]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += s_wsum[w];
        const uint32_t prefix = lookback_prefix(state, tile, total, lane);
        if (lane == 0) {
            s_base = prefix;
            if (tile == gridDim.x - 1) counters[MB_CNT_VOX] = prefix + total;
        }
    }
    __syncthreads();
    uint32_t o = s_base + inc - cnt;
    for (int w = 0; w < warp; ++w) o += s_wsum[w];
#pragma unroll
    for (int k = 0; k < VOX_WORDS; ++k) {
        uint32_t mm = m[k];
        while (mm) {
            const int bit = __ffs(mm) - 1;
            vlist[o++] = ((base + k) << 5) + (uint32_t)bit;
            mm &= mm - 1;
        }
    }
}

// K5: per (cell, frame) segment and slot: W = sum w, S2 = sum w^2 over the pixels of the segment's items (item
// order, pixel order inside the item).  One thread per segment; an item's records are contiguous.
__global__ void __launch_bounds__(256, MB_SEG_MINB)
k_seg_sums(const uint32_t *__restrict__ ival, const uint4 *__restrict__ rec, const uint32_t *__restrict__ seg_start,
           float2 *__restrict__ segws, uint32_t cap, uint32_t *__restrict__ counters, uint32_t *__restrict__ long_segs)
{
    const uint32_t nsegs = counters[MB_CNT_SEGS];
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsegs) return;
    // The loads of a segment form a chain (segment bounds -> item -> records); the kernel is bound by that latency, not
    // by its arithmetic.  So the bounds and the first item of the NEXT segment of this thread are requested before the
    // current one is processed: only the record loads remain on the critical path.
    uint32_t beg = __ldg(seg_start + s), end = __ldg(seg_start + s + 1);
    uint32_t v0 = __ldg(ival + beg);
    for (;;) {
        const uint32_t sn = s + stride;
        const bool more = sn < nsegs;
        uint32_t nbeg = 0, nend = 0;
        if (more) { nbeg = __ldg(seg_start + sn); nend = __ldg(seg_start + sn + 1); }
        // A segment of many items (a cell that fills a large part of a frame: a surface a few centimetres from the
        // camera puts thousands of pixels into one cell) would keep this one thread busy for milliseconds while the rest
        // of the grid has long finished: such segments go on a list and are summed by whole warps (k_seg_sums_long).
        // (this thread then writes zeros for it, which the warp kernel, next in the stream, overwrites)
        if (end - beg > (uint32_t)SEG_LONG_ITEMS) {
            long_segs[atomicAdd(&counters[MB_CNT_LONGSEGS], 1u)] = s;
            end = beg;
        }
        float W[8], S2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { W[k] = 0.f; S2[k] = 0.f; }
        for (uint32_t it = beg; it < end; ++it) {
            const uint32_t v = it == beg ? v0 : __ldg(ival + it);
            const uint4 *pr = rec + (item_tile(v) * (uint32_t)TILE_PIX + item_pos(v));      // (< 2^28: 32-bit index)
            const uint32_t len = item_len(v);
            for (uint32_t i = 0; i < len; i += 4) {
                uint4 r[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u < len) r[u] = __ldg(pr + i + u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u < len) {
                        float w[8];
                        splat_weights(r[u], w);
#pragma unroll
                        for (int k = 0; k < 8; ++k) { W[k] += w[k]; S2[k] = fmaf(w[k], w[k], S2[k]); }
                    }
            }
        }
        uint32_t nv0 = 0;
        if (more) nv0 = __ldg(ival + nbeg);             // (its address arrived while the records were being summed)
#pragma unroll
        for (int k = 0; k < 8; ++k) segws[(uint32_t)k * cap + s] = make_float2(W[k], S2[k]);    // slot-major (8 * cap < 2^31)
        if (!more) break;
        s = sn; beg = nbeg; end = nend; v0 = nv0;
    }
}

// K5b: the segments k_seg_sums put aside, one WARP each: lanes take the segment's items in turn (<= 16 pixels per
// item), then the 16 sums are reduced over the warp by a fixed butterfly.  The order of the additions differs from the
// one-thread order (it is still fixed, so results stay reproducible).
__global__ void __launch_bounds__(256)
k_seg_sums_long(const uint32_t *__restrict__ ival, const uint4 *__restrict__ rec, const uint32_t *__restrict__ seg_start,
                float2 *__restrict__ segws, uint32_t cap, const uint32_t *__restrict__ counters,
                const uint32_t *__restrict__ long_segs)
{
    const uint32_t nlong = counters[MB_CNT_LONGSEGS];
    const int lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = wid; j < nlong; j += nw) {
        const uint32_t s = long_segs[j];
        const uint32_t beg = __ldg(seg_start + s), end = __ldg(seg_start + s + 1);
        float W[8], S2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { W[k] = 0.f; S2[k] = 0.f; }
        for (uint32_t it = beg + lane; it < end; it += 32) {
            const uint32_t v = __ldg(ival + it);
            const uint4 *pr = rec + (item_tile(v) * (uint32_t)TILE_PIX + item_pos(v));
            const uint32_t len = item_len(v);
            for (uint32_t i = 0; i < len; ++i) {
                const uint4 r = __ldg(pr + i);
                float w[8];
                splat_weights(r, w);
#pragma unroll
                for (int k = 0; k < 8; ++k) { W[k] += w[k]; S2[k] = fmaf(w[k], w[k], S2[k]); }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                W[k] += __shfl_xor_sync(FULL, W[k], d);
                S2[k] += __shfl_xor_sync(FULL, S2[k], d);
            }
        // every lane holds the totals; lane k writes slot k
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (lane == k) segws[(uint32_t)k * cap + s] = make_float2(W[k], S2[k]);
    }
}

// K6a: one thread per (touched voxel, source cell): the <= 8 cells at extended coordinates v + {0,1}^3
// contribute to voxel v.  Looks each one up and records its segment range (for K6) and run range (for K8).
__global__ void __launch_bounds__(256)
k_voxel_sources(const uint32_t *__restrict__ vlist, const uint32_t *__restrict__ ucell, const int *__restrict__ ctab,
                const uint32_t *__restrict__ cseg, const uint32_t *__restrict__ crun, CellGrid g,
                uint2 *__restrict__ vseg, uint2 *__restrict__ vrun, const uint32_t *__restrict__ counters)
{
    const uint32_t nvox = counters[MB_CNT_VOX], ncells = counters[MB_CNT_CELLS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvox * 8u; i += gridDim.x * blockDim.x) {
        const uint32_t v = vlist[i >> 3];
        const int s = (int)(i & 7u);
        const int v2 = (int)(v % (uint32_t)g.S2);
        const uint32_t t01 = v / (uint32_t)g.S2;
        const int v1 = (int)(t01 % (uint32_t)g.S1), v0 = (int)(t01 / (uint32_t)g.S1);
        const uint32_t key = cell_key(g, v0 + ((s >> 2) & 1), v1 + ((s >> 1) & 1), v2 + (s & 1));
        const int u = ctab != nullptr ? ctab[key] : find_cell(ucell, ncells, key);
        uint2 sr = make_uint2(0u, 0u), rr = make_uint2(0u, 0u);
        if (u >= 0) {
            sr = make_uint2(cseg[u], cseg[u + 1]);
            rr = make_uint2(crun[u], crun[u + 1]);
        }
        vseg[i] = sr;
        vrun[i] = rr;
    }
}

// K6: one warp per touched voxel.  Each source cell's segments are sorted by frame, one segment per
// frame.  The warp adds the W / S2 sums of all sources into a per-frame table in shared memory (sources
// one after the other, so the order of the adds is fixed), turns every touched frame into
// a = 1 - alpha*S2/W and r = alpha/W, and runs a backward product scan over the frames:
// g(t) = r(t) * prod_{s>t} a(s) is the coefficient of every contribution of frame t to this voxel,
// A = prod a multiplies the old row.  All list loads of a 64-segment block are issued before the first
// is used: the kernel lives on memory-level parallelism.
__global__ void __launch_bounds__(256, MB_VS_MINB)
k_voxel_scalars(const uint32_t *__restrict__ vlist, const uint2 *__restrict__ vseg, const uint32_t *__restrict__ seg_frame,
                const float2 *__restrict__ segws, uint32_t cap, CellGrid g, float alpha, int T,
                float *__restrict__ gcoef, float *__restrict__ vA, uint32_t *__restrict__ counters)
{
    extern __shared__ float s_tab[];                  // [warps][2][Tp]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Tp = (T + 31) & ~31;
    float *tW = s_tab + (size_t)warp * 2 * Tp, *tS = tW + Tp;
    const uint32_t nvox = counters[MB_CNT_VOX];
    for (int f = lane; f < Tp; f += 32) { tW[f] = 0.f; tS[f] = 0.f; }
    __syncwarp();
    // voxels are handed out by a work queue: a voxel's cost goes with the number of frames that saw it
    uint32_t *queue = counters + MB_CNT_TASKQ + 2;
    auto next_voxel = [&]() {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(queue, 1u);
        return __shfl_sync(FULL, u, 0);
    };
    uint32_t nvt = 0;                                 // (voxel, frame) pairs this warp has seen: sum over frames of U_f
    for (uint32_t j = next_voxel(); j < nvox; j = next_voxel()) {
        const uint32_t v = vlist[j];
        uint2 mine = make_uint2(0u, 0u);
        if (lane < 8) mine = vseg[(size_t)j * 8 + lane];
        const int v2 = (int)(v % (uint32_t)g.S2);
        const uint32_t t01 = v / (uint32_t)g.S2;
        const int v1 = (int)(t01 % (uint32_t)g.S1), v0 = (int)(t01 / (uint32_t)g.S1);
        // away from the map border source s contributes exactly its slot 7 - s (the opposite corner)
        const bool interior = v0 > 0 && v0 < g.S0 - 1 && v1 > 0 && v1 < g.S1 - 1 && v2 > 0 && v2 < g.S2 - 1;
        uint32_t slo[8], shi[8];
        uint32_t maxlen = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            slo[s] = __shfl_sync(FULL, mine.x, s);
            shi[s] = __shfl_sync(FULL, mine.y, s);
            maxlen = max(maxlen, shi[s] - slo[s]);
        }
        // the tables are all zero here (zeroed once per warp, and again row by row after use)
        uint32_t rows = 0;                             // 32-frame rows with at least one touched frame
        if (interior) {
            for (uint32_t off = 0; off < maxlen; off += 32) {
                uint32_t f[8];
                float2 x[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    f[s] = 0xffffffffu;
                    if (slo[s] + off >= shi[s]) continue;          // warp-uniform: nothing left in this source
                    const uint32_t q = slo[s] + off + lane;
                    if (q < shi[s]) {
                        f[s] = __ldg(seg_frame + q);
                        x[s] = __ldg(segws + ((uint32_t)(7 - s) * cap + q));
                    }
                }
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (slo[s] + off >= shi[s]) continue;          // warp-uniform: nothing left in this source
                    if (f[s] != 0xffffffffu) {
                        tW[f[s]] += x[s].x;                        // one segment per frame and source: no two lanes share f
                        tS[f[s]] += x[s].y;
                        rows |= 1u << (f[s] >> 5);
                    }
                    __syncwarp();
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (slo[s] >= shi[s]) continue;
                const uint32_t m = slot_mask(v0, v1, v2, s, g);
                for (uint32_t q = slo[s] + lane; q < shi[s]; q += 32) {
                    const uint32_t f = seg_frame[q];
                    float W = 0.f, S2 = 0.f;
                    for (uint32_t mm = m; mm; mm &= mm - 1) {
                        const float2 x = segws[(uint32_t)(__ffs(mm) - 1) * cap + q];
                        W += x.x;
                        S2 += x.y;
                    }
                    tW[f] += W;
                    tS[f] += S2;
                    rows |= 1u << (f >> 5);
                }
                __syncwarp();
            }
        }
        rows = __reduce_or_sync(FULL, rows);
        // backward product scan over the touched rows (an untouched row has a = 1 everywhere)
        float carry = 1.0f;
        for (uint32_t rm = rows; rm;) {
            const int row = 31 - __clz(rm);
            rm &= ~(1u << row);
            const int base = row << 5;
            const float W = tW[base + lane], S2 = tS[base + lane];
            float r = 0.f, a = 1.0f;
            if (W > 0.f) { r = __fdividef(alpha, W); a = 1.0f - r * S2; }      // (2 ulp: far inside the 1e-5 budget)
            nvt += __popc(__ballot_sync(FULL, W > 0.f));
            float inc = a;                             // inclusive product over lanes >= lane
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float o = __shfl_down_sync(FULL, inc, d);
                if (lane + d < 32) inc *= o;
            }
            float exc = __shfl_down_sync(FULL, inc, 1);  // product over lanes > lane
            if (lane == 31) exc = 1.0f;
            tW[base + lane] = r * exc * carry;
            tS[base + lane] = 0.f;
            carry *= __shfl_sync(FULL, inc, 0);
        }
        __syncwarp();
        if (interior) {
            for (uint32_t off = 0; off < maxlen; off += 32) {
                uint32_t f[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    f[s] = 0xffffffffu;
                    if (slo[s] + off >= shi[s]) continue;          // warp-uniform
                    const uint32_t q = slo[s] + off + lane;
                    if (q < shi[s]) f[s] = __ldg(seg_frame + q);
                }
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (slo[s] + off >= shi[s]) continue;
                    if (f[s] != 0xffffffffu) gcoef[(uint32_t)(7 - s) * cap + (slo[s] + off + lane)] = tW[f[s]];
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (slo[s] >= shi[s]) continue;
                const uint32_t m = slot_mask(v0, v1, v2, s, g);
                for (uint32_t q = slo[s] + lane; q < shi[s]; q += 32) {
                    const float gv = tW[seg_frame[q]];
                    for (uint32_t mm = m; mm; mm &= mm - 1) gcoef[(uint32_t)(__ffs(mm) - 1) * cap + q] = gv;
                }
            }
        }
        __syncwarp();
        for (uint32_t rm = rows; rm; rm &= rm - 1) tW[((__ffs(rm) - 1) << 5) + lane] = 0.f;
        if (lane == 0) vA[j] = carry;
        __syncwarp();
    }
    if (lane == 0 && nvt) atomicAdd(&counters[MB_CNT_VOXFRAMES], nvt);
}

// ---------------------------------------------------------------------------------------------
// K7: the hot loop.  One warp per task of CH sorted pixels; lanes = channels (VEC floats per lane and
// iteration, IT iterations).  Per batch of 32 pixels the lanes first act as pixels: record + segment
// coefficient -> 8 splat coefficients per pixel, staged in shared memory; then the warp walks the 32
// pixels: one feature row load + 8 coefficient broadcasts + 8 FMAs per lane-vector.  A run of
// same-cell pixels accumulates in registers and is flushed as 8 rows of P.
struct AccArgs {
    const uint32_t *ikey, *ival;   // sorted item list
    TileGeom tg;
    const uint4 *rec;              // pixel records in tile-grouped order
    const uint32_t *smask, *soff;
    const uint32_t *rstart;        // first item of every run (+ sentinel)
    const float *gcoef;         // [8 slots][cap] coefficient per (slot, segment)
    uint32_t cap;               // padded pixel count of the chunk (< 2^28, so 8 * cap indices stay 32-bit)
    uint32_t *counters;            // read; the two work-queue heads are written
    MbFeatIndex fi;             // np = pixels per frame
    uint32_t fhw;               // feature rows per frame
    const float *features;      // [T][fhw][F] or null
    const int64_t *class_ids;   // [T][np] or null (one-hot features)
    int F;
    float *P;                   // [run][8][F]
    uint32_t round;             // launch index inside the chunk (which work-queue head to use)
    uint32_t run_base, run_cap; // runs of this round: [run_base, run_base + run_cap)
};

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// acc += c * f on VEC floats; pairs go through the packed fp32x2 FMA of sm_100
__device__ __forceinline__ void ffma2(float &a0, float &a1, float c, float f0, float f1)
{
    unsigned long long ra, rc, rf;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rc) : "f"(c));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rf) : "f"(f0), "f"(f1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(ra) : "l"(rc), "l"(rf));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ra));
}

template <int VEC>
__device__ __forceinline__ void vec_fma(float (&acc)[VEC], float c, const float (&f)[VEC])
{
    if (VEC == 1) acc[0] = fmaf(c, f[0], acc[0]);
    if (VEC == 2) ffma2(acc[0], acc[VEC > 1 ? 1 : 0], c, f[0], f[VEC > 1 ? 1 : 0]);
    if (VEC == 4) {
        ffma2(acc[0], acc[VEC > 1 ? 1 : 0], c, f[0], f[VEC > 1 ? 1 : 0]);
        ffma2(acc[VEC > 2 ? 2 : 0], acc[VEC > 2 ? 3 : 0], c, f[VEC > 2 ? 2 : 0], f[VEC > 2 ? 3 : 0]);
    }
}

// feature-row load; -DMB_ROW_L2_HINT=128 / 256 adds an L2 prefetch-size hint (measured on C2: 1.497 / 1.509 ms
// against 1.49 ms without: no gain, left off)
template <int VEC>
__device__ __forceinline__ void feat_load(float (&dst)[VEC], const float *p)
{
#ifdef MB_ROW_L2_HINT
#define MB_STR2(x) #x
#define MB_STR(x) MB_STR2(x)
    if (VEC == 1) asm volatile("ld.global.nc.L2::" MB_STR(MB_ROW_L2_HINT) "B.f32 %0, [%1];" : "=f"(dst[0]) : "l"(p));
    if (VEC == 2) asm volatile("ld.global.nc.L2::" MB_STR(MB_ROW_L2_HINT) "B.v2.f32 {%0, %1}, [%2];" : "=f"(dst[0]), "=f"(dst[VEC > 1 ? 1 : 0]) : "l"(p));
    if (VEC == 4) asm volatile("ld.global.nc.L2::" MB_STR(MB_ROW_L2_HINT) "B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[0]), "=f"(dst[VEC > 1 ? 1 : 0]), "=f"(dst[VEC > 2 ? 2 : 0]), "=f"(dst[VEC > 2 ? 3 : 0]) : "l"(p));
#else
    if (VEC == 1) dst[0] = __ldg(p);
    if (VEC == 2) { const float2 v = __ldg((const float2 *)p); dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; }
    if (VEC == 4) {
        const float4 v = __ldg((const float4 *)p);
        dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w;
    }
#endif
}

template <int VEC>
__device__ __forceinline__ void row_load(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = __ldg(p);
    if (VEC == 2) { const float2 v = __ldg((const float2 *)p); dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; }
    if (VEC == 4) {
        const float4 v = __ldg((const float4 *)p);
        dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w;
    }
}

template <int VEC>
__device__ __forceinline__ void plain_load(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = *p;
    if (VEC == 2) { const float2 v = *(const float2 *)p; dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; }
    if (VEC == 4) {
        const float4 v = *(const float4 *)p;
        dst[0] = v.x; dst[VEC > 1 ? 1 : 0] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w;
    }
}

template <int VEC>
__device__ __forceinline__ void row_store(float *p, const float (&src)[VEC])
{
    if (VEC == 1) *p = src[0];
    if (VEC == 2) *(float2 *)p = make_float2(src[0], src[VEC > 1 ? 1 : 0]);
    if (VEC == 4) *(float4 *)p = make_float4(src[0], src[VEC > 1 ? 1 : 0], src[VEC > 2 ? 2 : 0], src[VEC > 2 ? 3 : 0]);
}

// shared memory of k_cell_accumulate (dynamic: ~37 KB per CTA of 8 warps)
struct AccSmem {
    static constexpr int NW = ACC_THREADS / 32;
    float coef[NW][32][8];                          // 8 splat coefficients of every pixel of the batch being walked
    uint32_t src[NW][32];                           // feature row (or class id) of every pixel of the batch
    uint32_t pre[NW][TASK_ITEMS + 1], ival[NW][TASK_ITEMS], seg[NW][TASK_ITEMS];
    uint8_t p2i[NW][TASK_ITEMS * ITEM_MAX];         // item of every pixel of the run
};

template <int VEC, int IT, bool ONEHOT, int U>
__device__ __forceinline__ void accumulate_round(const AccArgs &A, AccSmem &SM)
{
    auto &s_coef = SM.coef; auto &s_src = SM.src; auto &s_pre = SM.pre; auto &s_ival = SM.ival; auto &s_seg = SM.seg;
    auto &s_p2i = SM.p2i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nruns = A.counters[MB_CNT_RUNS];
    // Work queue: a unit is (run, channel block); warps take the next unit from a counter when they finish one, so
    // the runs' very different pixel counts (1 .. TASK_ITEMS items of 1 .. 16 pixels each) do not leave SMs idle at the
    // end, and the runs in flight at any moment are neighbours in the sorted order.  The launches of a chunk use the
    // two queue heads in turn; each launch clears the one the next launch will use.
    if (blockIdx.x == 0 && threadIdx.x == 0) A.counters[MB_CNT_TASKQ + ((A.round + 1) & 1)] = 0;
    if (A.run_base >= nruns) return;
    uint32_t *queue = A.counters + MB_CNT_TASKQ + (A.round & 1);
    const int F = A.F;
    const uint32_t ny = (uint32_t)(F + 32 * VEC * IT - 1) / (uint32_t)(32 * VEC * IT);
    const uint32_t nunits = min(nruns - A.run_base, A.run_cap) * ny;
    const uint32_t np = A.fi.np;
    auto next_unit = [&]() {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(queue, 1u);
        return __shfl_sync(FULL, u, 0);
    };

    for (uint32_t unit = next_unit(); unit < nunits; unit = next_unit()) {
        const uint32_t rr = unit / ny;                                   // run of this round
        const int ch0 = (int)(unit - rr * ny) * (32 * VEC * IT) + lane * VEC;       // first channel of this lane
        const uint32_t base = __ldg(A.rstart + A.run_base + rr), end = __ldg(A.rstart + A.run_base + rr + 1);
        // ---- the run's items: pixel prefix, segment ranks (32 items per round of lanes) -----------------------
        uint32_t npixels = 0;
        __syncwarp();
        for (uint32_t h0 = 0; base + h0 < end; h0 += 32) {
            const uint32_t i = base + h0 + lane;
            uint32_t len = 0;
            if (i < end) {
                const uint32_t v = A.ival[i];
                len = item_len(v);
                const uint32_t w = i >> 5, bit = i & 31u;
                const uint32_t le = bit == 31u ? 0xffffffffu : ((2u << bit) - 1u);
                s_ival[warp][h0 + lane] = v;
                s_seg[warp][h0 + lane] = A.soff[w] + __popc(A.smask[w] & le) - 1u;
            }
            uint32_t inc = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t pre = npixels + inc - len;
            if (i < end) s_pre[warp][h0 + lane] = pre;
            for (uint32_t o = 0; o < len; ++o) s_p2i[warp][pre + o] = (uint8_t)(h0 + lane);
            npixels += __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
        float acc[8][IT][VEC];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int it = 0; it < IT; ++it)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[k][it][j] = 0.f;
        const char *fbase = (const char *)(A.features + ch0);          // this lane's first channel of row 0
        const uint32_t row_bytes = (uint32_t)F * (uint32_t)sizeof(float);

        for (uint32_t b0 = 0; b0 < npixels; b0 += 32) {
            // ---- lanes = pixels: coefficients of the batch ------------------------------------------------
            const uint32_t qp = b0 + lane;                                // pixel of the run
            {
                float c[8];
                uint32_t src = 0;
                if (qp < npixels) {
                    const int lo = s_p2i[warp][qp];
                    const uint32_t v = s_ival[warp][lo], off = qp - s_pre[warp][lo];
                    const uint32_t tile = item_tile(v);
                    const uint4 r = __ldg(A.rec + (tile * (uint32_t)TILE_PIX + item_pos(v) + off));
                    const uint32_t sg = s_seg[warp][lo];
                    float gk[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) gk[k] = __ldg(A.gcoef + ((uint32_t)k * A.cap + sg));
                    splat_weights(r, c);
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = c[k] * c[k] * gk[k];
                    const uint32_t frame = fast_div(tile, (uint32_t)A.tg.tpf, A.tg.tpf_magic), tif = tile - frame * (uint32_t)A.tg.tpf;
                    const uint32_t tyo = fast_div(tif, (uint32_t)A.tg.tiles_x, A.tg.tx_magic), txo = tif - tyo * (uint32_t)A.tg.tiles_x;
                    const uint32_t y = tyo * TILE_H + (r.w >> 5), x = txo * TILE_W + (r.w & 31u);
                    if (ONEHOT) {
                        const int64_t id = A.class_ids[(size_t)frame * np + y * A.fi.W + x];
                        src = (uint32_t)id;
                        if ((uint64_t)id >= (uint64_t)F) {      // functional.one_hot raises here: flag it, add nothing
                            src = 0xffffffffu;
                            atomicOr(&A.counters[MB_CNT_ERROR], 2u);
                        }
                    } else {
                        src = frame * A.fhw + (y / A.fi.ky) * A.fi.fw + x / A.fi.kx;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = 0.f;
                }
                {
                    const uint32_t src0 = __shfl_sync(FULL, src, 0);      // lane 0's pixel always exists
                    if (qp >= npixels) src = ONEHOT ? 0xffffffffu : src0;
                }
                __syncwarp();
                *(float4 *)&s_coef[warp][lane][0] = make_float4(c[0], c[1], c[2], c[3]);
                *(float4 *)&s_coef[warp][lane][4] = make_float4(c[4], c[5], c[6], c[7]);
                s_src[warp][lane] = src;
                __syncwarp();
            }
            // ---- lanes = channels: walk the batch, U rows in flight ---------------------------------------
            // (always whole groups of U: the pixels past the end of the run carry zero coefficients and the row of
            // the batch's first pixel, so a short last batch costs a few spare loads instead of a one-row-at-a-time tail)
            const int nb = (int)min(32u, npixels - b0);
            for (int jj = 0; jj < nb; jj += U) {
                float f[U][IT][VEC];
                uint32_t srcs[U];                                         // (jj is a multiple of U: 16-byte pieces)
                if (U % 4 == 0) {
#pragma unroll
                    for (int u = 0; u < U; u += 4) {
                        const uint4 q = *(const uint4 *)&s_src[warp][jj + u];
                        srcs[u] = q.x; srcs[u + 1 < U ? u + 1 : u] = q.y; srcs[u + 2 < U ? u + 2 : u] = q.z; srcs[u + 3 < U ? u + 3 : u] = q.w;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) srcs[u] = s_src[warp][jj + u];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t src = srcs[u];
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int ch = ch0 + it * 32 * VEC;
                        if (ONEHOT) {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) f[u][it][j] = (uint32_t)(ch + j) == src ? 1.0f : 0.0f;
                        } else if (ch < F) {
                            // one 32 x 32 -> 64-bit multiply-add per row address (row index x row bytes + lane base)
                            feat_load<VEC>(f[u][it], (const float *)(fbase + (uint64_t)src * row_bytes + (uint32_t)(it * 32 * VEC * 4)));
                        } else {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) f[u][it][j] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 c0 = *(const float4 *)&s_coef[warp][jj + u][0];
                    const float4 c1 = *(const float4 *)&s_coef[warp][jj + u][4];
                    const float c[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int it = 0; it < IT; ++it) vec_fma<VEC>(acc[k][it], c[k], f[u][it]);
                }
            }
        }
        // ---- the run's 8 partial rows ----------------------------------------------------------------------
        float *prow = A.P + (size_t)rr * 8 * F;
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int it = 0; it < IT; ++it) {
                const int ch = ch0 + it * 32 * VEC;
                if (ch < F) row_store<VEC>(prow + (size_t)k * F + ch, acc[k][it]);
            }
    }
}

template <int VEC, int IT, bool ONEHOT, int U>
__global__ void __launch_bounds__(ACC_THREADS, (VEC * IT <= 2 ? 4 : 1))
k_cell_accumulate(const AccArgs A)
{
    extern __shared__ __align__(16) unsigned char acc_smem_raw[];
    accumulate_round<VEC, IT, ONEHOT, U>(A, *reinterpret_cast<AccSmem *>(acc_smem_raw));
}

// K7 for single-channel maps (occupancy: F = 1).  With lanes = channels the warp would spend 8 FMA instructions per
// pixel on 32 channels of which one exists; here the lanes stay pixels: the 8 splat coefficients (times the pixel's
// value) of a batch of 32 pixels are summed over the warp by a fixed exchange tree -- 9 shuffles: at every step a lane
// keeps half of its values and receives the partner's other half -- and added to the run's 8 partial values, which
// stay in registers.  The sums run in a fixed order, so results stay bit-reproducible.  (Measured on 500 C2 frames: accumulate
// stage 0.72 -> 0.38 ms.  The same scheme with one tree per distinct class of a batch was tried for class ids: slower
// on the synthetic scene, 0.86 against 0.71 ms, whose classes change from frame to frame and put ~8 distinct ids into
// a batch; class ids stay on the lanes = channels kernel.)
struct AccSmemSingle {
    static constexpr int NW = ACC_THREADS / 32;
    uint32_t pre[NW][TASK_ITEMS + 1], ival[NW][TASK_ITEMS], seg[NW][TASK_ITEMS];
    uint8_t p2i[NW][TASK_ITEMS * ITEM_MAX];         // item of every pixel of the run
};

__device__ __forceinline__ void accumulate_round_single(const AccArgs &A, AccSmemSingle &SM)
{
    auto &s_pre = SM.pre; auto &s_ival = SM.ival; auto &s_seg = SM.seg; auto &s_p2i = SM.p2i;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nruns = A.counters[MB_CNT_RUNS];
    if (blockIdx.x == 0 && threadIdx.x == 0) A.counters[MB_CNT_TASKQ + ((A.round + 1) & 1)] = 0;
    if (A.run_base >= nruns) return;
    uint32_t *queue = A.counters + MB_CNT_TASKQ + (A.round & 1);
    const uint32_t nunits = min(nruns - A.run_base, A.run_cap);      // one unit per run: there are no channel blocks
    auto next_unit = [&]() {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(queue, 1u);
        return __shfl_sync(FULL, u, 0);
    };
    // which of its 8 values a lane ends up with after the exchange tree
    const int mine_k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);

    for (uint32_t rr = next_unit(); rr < nunits; rr = next_unit()) {
        const uint32_t base = __ldg(A.rstart + A.run_base + rr), end = __ldg(A.rstart + A.run_base + rr + 1);
        // ---- the run's items: pixel prefix, segment ranks (32 items per round of lanes) -----------------------
        uint32_t npixels = 0;
        __syncwarp();
        for (uint32_t h0 = 0; base + h0 < end; h0 += 32) {
            const uint32_t i = base + h0 + lane;
            uint32_t len = 0;
            if (i < end) {
                const uint32_t v = A.ival[i];
                len = item_len(v);
                const uint32_t w = i >> 5, bit = i & 31u;
                const uint32_t le = bit == 31u ? 0xffffffffu : ((2u << bit) - 1u);
                s_ival[warp][h0 + lane] = v;
                s_seg[warp][h0 + lane] = A.soff[w] + __popc(A.smask[w] & le) - 1u;
            }
            uint32_t inc = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t pre = npixels + inc - len;
            if (i < end) s_pre[warp][h0 + lane] = pre;
            for (uint32_t o = 0; o < len; ++o) s_p2i[warp][pre + o] = (uint8_t)(h0 + lane);
            npixels += __shfl_sync(FULL, inc, 31);
        }
        __syncwarp();
        float total = 0.f;                                                // the run's sum of value mine_k (all lanes of a quad alike)

        for (uint32_t b0 = 0; b0 < npixels; b0 += 32) {
            const uint32_t qp = b0 + lane;                                // pixel of the run
            float cv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cv[k] = 0.f;
            if (qp < npixels) {
                const int lo = s_p2i[warp][qp];
                const uint32_t v = s_ival[warp][lo], off = qp - s_pre[warp][lo];
                const uint32_t tile = item_tile(v);
                const uint4 r = __ldg(A.rec + (tile * (uint32_t)TILE_PIX + item_pos(v) + off));
                const uint32_t sg = s_seg[warp][lo];
                float gk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) gk[k] = __ldg(A.gcoef + ((uint32_t)k * A.cap + sg));
                const uint32_t frame = fast_div(tile, (uint32_t)A.tg.tpf, A.tg.tpf_magic), tif = tile - frame * (uint32_t)A.tg.tpf;
                const uint32_t tyo = fast_div(tif, (uint32_t)A.tg.tiles_x, A.tg.tx_magic), txo = tif - tyo * (uint32_t)A.tg.tiles_x;
                const uint32_t y = tyo * TILE_H + (r.w >> 5), x = txo * TILE_W + (r.w & 31u);
                // (F == 1: the pixel's only channel)
                const float value = __ldg(A.features + (size_t)(frame * A.fhw + (y / A.fi.ky) * A.fi.fw + x / A.fi.kx));
                splat_weights(r, cv);
#pragma unroll
                for (int k = 0; k < 8; ++k) cv[k] = cv[k] * cv[k] * gk[k] * value;
            }
            // ---- the exchange tree: 8 values x 32 lanes -> lane l holds the warp's sum of value mine_k(l) -------------
            float a[4], b[2], t;
            {
                const bool up = (lane & 16) != 0;                        // the upper half keeps values 4..7
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float send = up ? cv[j] : cv[j + 4], keep = up ? cv[j + 4] : cv[j];
                    a[j] = keep + __shfl_xor_sync(FULL, send, 16);
                }
            }
            {
                const bool up = (lane & 8) != 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float send = up ? a[j] : a[j + 2], keep = up ? a[j + 2] : a[j];
                    b[j] = keep + __shfl_xor_sync(FULL, send, 8);
                }
            }
            {
                const bool up = (lane & 4) != 0;
                const float send = up ? b[0] : b[1], keep = up ? b[1] : b[0];
                t = keep + __shfl_xor_sync(FULL, send, 4);
            }
            t += __shfl_xor_sync(FULL, t, 2);
            t += __shfl_xor_sync(FULL, t, 1);
            total += t;
        }
        // ---- the run's 8 partial rows (one float each) -------------------------------------------------------
        if ((lane & 3) == 0) A.P[(size_t)rr * 8 + mine_k] = total;
    }
}

__global__ void __launch_bounds__(ACC_THREADS, 4)
k_cell_accumulate_single(const AccArgs A)
{
    extern __shared__ __align__(16) unsigned char acc_smem_raw[];
    accumulate_round_single(A, *reinterpret_cast<AccSmemSingle *>(acc_smem_raw));
}

// K8: one warp per touched voxel: map = A * map + sum of the P rows of its cells' runs (round 0), or
// map += sum (later rounds, when the runs did not fit one P buffer).  The old row and the first two P
// rows of every source are requested before anything is added.
struct ApplyArgs {
    const uint32_t *vlist;
    const uint2 *vrun;          // [voxel][8] run range of each source cell (from K6a)
    const float *vA, *P;
    const uint32_t *counters;
    CellGrid g;
    int F;
    float *map, *affine_a;
    uint32_t run_base, run_cap;
    // fold into a sparse partial instead of the map (frame-sharded scenes): row slot of every touched voxel
    // (top bit: the row is new in this chunk; 0xffffffff: no room, skipped), rows and their coefficients
    const uint32_t *vslot;
    float *part_a, *part_b;
};

template <int VEC, int IT>
__device__ __forceinline__ void apply_round(const ApplyArgs &A, int chblock)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nvox = A.counters[MB_CNT_VOX], nruns = A.counters[MB_CNT_RUNS];
    if (A.run_base >= nruns && A.run_base > 0) return;
    const uint32_t run_end = A.run_base + A.run_cap;
    const int F = A.F;
    const int ch0 = chblock * (32 * VEC * IT) + lane * VEC;
    // P row of (run, slot) for this lane: one 32 x 32 -> 64-bit multiply-add ((run - run_base) * 8 + slot < 2^32: run_cap <= 2^28)
    const char *pbase = (const char *)(A.P + ch0);
    const uint32_t row_bytes = (uint32_t)F * (uint32_t)sizeof(float);
    auto prow_of = [&](uint32_t run, int slot, int it) {
        return (const float *)(pbase + (uint64_t)((run - A.run_base) * 8u + (uint32_t)slot) * row_bytes + (uint32_t)(it * 32 * VEC * 4));
    };
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = wid; j < nvox; j += nw) {
        const uint32_t v = A.vlist[j];
        uint32_t lo = 0, hi = 0;
        if (lane < 8) {
            const uint2 rr = A.vrun[(size_t)j * 8 + lane];
            lo = max(rr.x, A.run_base);
            hi = min(rr.y, run_end);
        }
        const float a = A.run_base == 0 ? A.vA[j] : 1.0f;
        float *grow = A.map + (size_t)v * F;
        bool fresh = false;                            // sparse partial: the row starts from the identity (A = 1, B = 0)
        uint32_t slot = 0;
        if (A.vslot != nullptr) {
            slot = A.vslot[j];
            if (slot == 0xffffffffu) continue;
            fresh = (slot >> 31) != 0u && A.run_base == 0;
            slot &= 0x7fffffffu;
            grow = A.part_b + (size_t)slot * F;
        }
        float old[IT][VEC];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int ch = ch0 + it * 32 * VEC;
#pragma unroll
            for (int q = 0; q < VEC; ++q) old[it][q] = 0.f;
            if (ch < F && !fresh) {
                if (VEC == 1) old[it][0] = grow[ch];
                if (VEC == 2) { const float2 o = *(const float2 *)(grow + ch); old[it][0] = o.x; old[it][VEC > 1 ? 1 : 0] = o.y; }
                if (VEC == 4) { const float4 o = *(const float4 *)(grow + ch); old[it][0] = o.x; old[it][VEC > 1 ? 1 : 0] = o.y; old[it][VEC > 2 ? 2 : 0] = o.z; old[it][VEC > 2 ? 3 : 0] = o.w; }
            }
        }
        const int v2 = (int)(v % (uint32_t)A.g.S2);
        const uint32_t t01 = v / (uint32_t)A.g.S2;
        const int v1 = (int)(t01 % (uint32_t)A.g.S1), v0 = (int)(t01 / (uint32_t)A.g.S1);
        // away from the map border source s contributes exactly its slot 7 - s (the opposite corner)
        const bool interior = v0 > 0 && v0 < A.g.S0 - 1 && v1 > 0 && v1 < A.g.S1 - 1 && v2 > 0 && v2 < A.g.S2 - 1;
        float acc[IT][VEC];
#pragma unroll
        for (int it = 0; it < IT; ++it)
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[it][q] = 0.f;
        uint32_t slo[8], shi[8];
        bool any = false;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            slo[s] = __shfl_sync(FULL, lo, s);
            shi[s] = __shfl_sync(FULL, hi, s);
            any = any || slo[s] < shi[s];
        }
        if (A.run_base > 0 && !any) continue;
        if (interior) {
            // first two runs of every source: 16 independent row loads
            float x[8][2][IT][VEC];
#pragma unroll
            for (int s = 0; s < 8; ++s)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int ch = ch0 + it * 32 * VEC;
#pragma unroll
                        for (int q = 0; q < VEC; ++q) x[s][r][it][q] = 0.f;
                        if (slo[s] + r < shi[s] && ch < F) row_load<VEC>(x[s][r][it], prow_of(slo[s] + r, 7 - s, it));
                    }
#pragma unroll
            for (int s = 0; s < 8; ++s) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int q = 0; q < VEC; ++q) acc[it][q] += x[s][r][it][q];
                for (uint32_t e = slo[s] + 2; e < shi[s]; ++e) {
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int ch = ch0 + it * 32 * VEC;
                        if (ch < F) {
                            float y[VEC];
                            row_load<VEC>(y, prow_of(e, 7 - s, it));
#pragma unroll
                            for (int q = 0; q < VEC; ++q) acc[it][q] += y[q];
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (slo[s] >= shi[s]) continue;
                const uint32_t m = slot_mask(v0, v1, v2, s, A.g);
                for (uint32_t e = slo[s]; e < shi[s]; ++e) {
                    for (uint32_t mm = m; mm; mm &= mm - 1) {
                        const int k = __ffs(mm) - 1;
#pragma unroll
                        for (int it = 0; it < IT; ++it) {
                            const int ch = ch0 + it * 32 * VEC;
                            if (ch < F) {
                                float y[VEC];
                                row_load<VEC>(y, prow_of(e, k, it));
#pragma unroll
                                for (int q = 0; q < VEC; ++q) acc[it][q] += y[q];
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int ch = ch0 + it * 32 * VEC;
            if (ch < F) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) old[it][q] = fmaf(a, old[it][q], acc[it][q]);
                row_store<VEC>(grow + ch, old[it]);
            }
        }
        if (A.affine_a != nullptr && A.run_base == 0 && chblock == 0 && lane == 0) {
            // fold output: 2.0 marks a voxel no chunk has touched yet (a product of a's never exceeds 1)
            const float prev = A.affine_a[v];
            A.affine_a[v] = (prev == 2.0f ? 1.0f : prev) * a;
        }
        if (A.vslot != nullptr && A.run_base == 0 && chblock == 0 && lane == 0)
            A.part_a[slot] = fresh ? a : A.part_a[slot] * a;
    }
}

template <int VEC, int IT>
__global__ void __launch_bounds__(256)
k_voxel_apply(const ApplyArgs A)
{
    apply_round<VEC, IT>(A, (int)blockIdx.y);
}

// Rounds 1 .. rounds-1 of the feature pass in ONE cooperative launch.  The run buffer is sized far below the worst
// case (every pixel its own cell), so the host plans several accumulate + apply rounds, but a real call fits the
// first: instead of a pair of (empty) launches per planned round this kernel checks the run count on the device and,
// in the rare case that runs are left, does the remaining rounds itself with grid-wide barriers between the phases.
template <int VEC, int IT, bool ONEHOT, int U>
__global__ void __launch_bounds__(ACC_THREADS, (VEC * IT <= 2 ? 4 : 1))
k_overflow_rounds(AccArgs A, ApplyArgs Y, int rounds)
{
    extern __shared__ __align__(16) unsigned char acc_smem_raw[];
    AccSmem &SM = *reinterpret_cast<AccSmem *>(acc_smem_raw);
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t nruns = A.counters[MB_CNT_RUNS];
    const int cblocks = (A.F + 32 * VEC * IT - 1) / (32 * VEC * IT);
    for (int r = 1; r < rounds; ++r) {
        const uint64_t base = (uint64_t)r * A.run_cap;
        if (base >= nruns) break;                        // (the same decision in every thread of the grid)
        A.run_base = Y.run_base = (uint32_t)base;
        A.round = (uint32_t)r;
        accumulate_round<VEC, IT, ONEHOT, U>(A, SM);
        grid.sync();
        for (int cb = 0; cb < cblocks; ++cb) apply_round<VEC, IT>(Y, cb);
        grid.sync();
    }
}

__global__ void __launch_bounds__(ACC_THREADS, 2)
k_overflow_rounds_single(AccArgs A, ApplyArgs Y, int rounds)
{
    extern __shared__ __align__(16) unsigned char acc_smem_raw[];
    AccSmemSingle &SM = *reinterpret_cast<AccSmemSingle *>(acc_smem_raw);
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t nruns = A.counters[MB_CNT_RUNS];
    for (int r = 1; r < rounds; ++r) {
        const uint64_t base = (uint64_t)r * A.run_cap;
        if (base >= nruns) break;                        // (the same decision in every thread of the grid)
        A.run_base = Y.run_base = (uint32_t)base;
        A.round = (uint32_t)r;
        accumulate_round_single(A, SM);
        grid.sync();
        apply_round<1, 1>(Y, 0);                         // (F == 1: one channel block of one float per lane)
        grid.sync();
    }
}

// Ordered affine combine of frame-sharded partial maps: map[idx[i]] = a[i] * map[idx[i]] + b[i].
// One warp per row, lanes stride over the channels; idx holds distinct voxels, so rows never collide.
__global__ void __launch_bounds__(256)
k_affine_apply_rows(float *__restrict__ map, int F, const int64_t *__restrict__ idx, const float *__restrict__ a,
                    const float *__restrict__ b, int64_t n)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = wid; i < n; i += nw) {
        float *row = map + (size_t)idx[i] * F;
        const float *brow = b + (size_t)i * F;
        const float ai = a[i];
        for (int ch = lane; ch < F; ch += 32) row[ch] = fmaf(ai, row[ch], brow[ch]);
    }
}

// ---- sparse partials of frame-sharded scenes ---------------------------------------------------------
// A partial is the action of a contiguous run of frames on ANY map: M[v] <- a[v] * M[v] + b[v] for the voxels the
// frames touched.  It lives in one buffer {count | index[cap] | a[cap] | b[cap][F]} (offsets: partial_layout) so that a
// peer GPU can read it through ONE mapped pointer; slot_table[V] (-1 = no row) finds a voxel's row while chunks are
// being folded in.
struct PartialView {
    uint32_t *count;
    int64_t *index;
    float *a, *b;
};

__host__ __device__ inline void partial_layout(uint32_t capacity, int F, size_t off[4], size_t *total)
{
    off[0] = 0;                                                        // count (+ padding)
    off[1] = 256;                                                      // index
    off[2] = off[1] + (((size_t)capacity * sizeof(int64_t) + 255) / 256) * 256;     // a
    off[3] = off[2] + (((size_t)capacity * sizeof(float) + 255) / 256) * 256;       // b
    if (total) *total = off[3] + (((size_t)capacity * F * sizeof(float) + 255) / 256) * 256;
}

__host__ __device__ inline PartialView partial_view(void *buffer, uint32_t capacity, int F)
{
    size_t off[4];
    partial_layout(capacity, F, off, nullptr);
    char *p = (char *)buffer;
    PartialView v;
    v.count = (uint32_t *)(p + off[0]);
    v.index = (int64_t *)(p + off[1]);
    v.a = (float *)(p + off[2]);
    v.b = (float *)(p + off[3]);
    return v;
}

// one thread per touched voxel of the chunk: find or create its row
__global__ void __launch_bounds__(256)
k_partial_slots(const uint32_t *__restrict__ vlist, const uint32_t *__restrict__ counters, int32_t *__restrict__ slot_table,
                PartialView P, uint32_t capacity, uint32_t *__restrict__ vslot, uint32_t *__restrict__ errors)
{
    const uint32_t nvox = counters[MB_CNT_VOX];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nvox; j += gridDim.x * blockDim.x) {
        const uint32_t v = vlist[j];
        int32_t s = slot_table[v];
        uint32_t out;
        if (s >= 0) {
            out = (uint32_t)s;
        } else {
            const uint32_t n = atomicAdd(P.count, 1u);
            if (n < capacity) {
                slot_table[v] = (int32_t)n;
                P.index[n] = (int64_t)v;
                out = n | 0x80000000u;
            } else {
                atomicOr(errors, 4u);                                   // partial full: the row is dropped, check() raises
                out = 0xffffffffu;
            }
        }
        vslot[j] = out;
    }
}

// slot_table[index[i]] = -1 for the rows in use (the sparse way back to an empty partial); count is zeroed by the caller
__global__ void __launch_bounds__(256)
k_partial_clear(int32_t *__restrict__ slot_table, PartialView P, uint32_t capacity)
{
    const uint32_t n = min(*P.count, capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        slot_table[P.index[i]] = -1;
}

// map[index[i]] = a[i] * map[index[i]] + b[i] for the rows of a partial; the row count is read on the device.
// `P` may point into a PEER GPU's memory (mapped over NVLink): the rows then cross the link inside this kernel,
// overlapped row by row with their application -- there is no separate exchange step.
template <int VEC>
__global__ void __launch_bounds__(256)
k_affine_apply_partial(float *__restrict__ map, int F, PartialView P, uint32_t capacity)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(*(const volatile uint32_t *)P.count, capacity);
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    // two rows per warp and iteration: both rows' loads are in flight before either is used
    for (uint32_t i = 2 * wid; i < n; i += 2 * nw) {
        const bool two = i + 1 < n;
        const int64_t v0 = P.index[i], v1 = two ? P.index[i + 1] : 0;
        const float a0 = P.a[i], a1 = two ? P.a[i + 1] : 0.f;
        float *r0 = map + (size_t)v0 * F, *r1 = map + (size_t)v1 * F;
        const float *b0 = P.b + (size_t)i * F, *b1 = b0 + F;
        for (int ch = lane * VEC; ch < F; ch += 32 * VEC) {
            float x0[VEC], x1[VEC], m0[VEC], m1[VEC];
            plain_load<VEC>(x0, b0 + ch);                   // (coherent loads: the rows may have been written by a peer)
            if (two) plain_load<VEC>(x1, b1 + ch);
            plain_load<VEC>(m0, r0 + ch);
            if (two) plain_load<VEC>(m1, r1 + ch);
#pragma unroll
            for (int q = 0; q < VEC; ++q) m0[q] = fmaf(a0, m0[q], x0[q]);
            row_store<VEC>(r0 + ch, m0);
            if (two) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) m1[q] = fmaf(a1, m1[q], x1[q]);
                row_store<VEC>(r1 + ch, m1);
            }
        }
    }
}

// All peers' partials into local staging slots in ONE kernel: the CTAs are dealt round-robin to the world - 1 peers,
// starting with the rank after `self`, so at every moment each rank reads from all its peers and each GPU's NVLink
// egress serves its 7 readers evenly (applying partial g straight out of rank g's memory on every rank at once makes
// rank g's egress the bottleneck of step g: measured 6.2 ms for 8 x 0.5 M rows against ~1 ms of link time).  A slot has
// the layout of a partial buffer; the row count is read on the device.
struct PullArgs {
    const void *peer[16];           // partial buffers of all ranks (mapped peer memory), indexed by rank
    void *slot[16];                 // local staging slot of every rank
    int world, self;
    uint32_t capacity;
    int F;
};

__global__ void __launch_bounds__(256)
k_partial_pull(const PullArgs A)
{
    const int npeers = A.world - 1;
    if (npeers <= 0) return;
    const int p = blockIdx.x % npeers, g = (A.self + 1 + p) % A.world;
    const uint32_t cta = blockIdx.x / npeers, nctas = (gridDim.x - p + npeers - 1) / npeers;
    size_t off[4];
    partial_layout(A.capacity, A.F, off, nullptr);
    const char *src = (const char *)A.peer[g];
    char *dst = (char *)A.slot[g];
    const uint32_t n = min(*(const volatile uint32_t *)src, A.capacity);
    if (cta == 0 && threadIdx.x == 0) *(uint32_t *)dst = n;
    const size_t bytes[3] = { (size_t)n * sizeof(int64_t), (size_t)n * sizeof(float), (size_t)n * A.F * sizeof(float) };
#pragma unroll
    for (int sec = 0; sec < 3; ++sec) {
        const uint4 *s4 = (const uint4 *)(src + off[sec + 1]);
        uint4 *d4 = (uint4 *)(dst + off[sec + 1]);
        const size_t n16 = (bytes[sec] + 15) / 16;                     // (sections are padded to 256 bytes)
        size_t i = (size_t)cta * 256 + threadIdx.x;
        const size_t stride = (size_t)nctas * 256;
        // four independent 16-byte loads in flight per thread: the link latency is a few microseconds
        for (; i + 3 * stride < n16; i += 4 * stride) {
            const uint4 v0 = s4[i], v1 = s4[i + stride], v2 = s4[i + 2 * stride], v3 = s4[i + 3 * stride];
            d4[i] = v0; d4[i + stride] = v1; d4[i + 2 * stride] = v2; d4[i + 3 * stride] = v3;
        }
        for (; i < n16; i += stride) d4[i] = s4[i];
    }
}

// ---------------------------------------------------------------------------------------------
struct CellBuffers {
    uint32_t *counters;
    uint4 *rec, *pix;
    uint32_t *keys_a, *keys_b, *pids_a, *pids_b;
    uint32_t *tcount, *toff;
    uint32_t *smask, *soff;
    uint32_t *rstart;                     // first item of every accumulate run (+ sentinel)
    uint32_t *run_state;                  // look-back words of K3
    uint32_t *scan_state;                 // look-back words of the tile-count scan
    uint32_t *idx_state, *vox_state;      // look-back words of K2 and K4; the bitmap follows: zeroed by one memset
    size_t state_bytes;
    uint32_t *ucell, *cstart, *cseg, *crun, *seg_start, *seg_frame;
    float2 *segws;
    float *gcoef;
    uint32_t *bitmap, *vlist;
    float *vA;
    uint2 *vseg, *vrun;
    uint32_t *vslot;            // sparse-partial row of every touched voxel (fold into a partial)
    uint32_t *long_segs;        // segments put aside by k_seg_sums (more than SEG_LONG_ITEMS items each)
    int *ctab;                  // dense cell key -> unique cell index (or null: binary search)
    char *scan_ws, *sort_ws;
    size_t scan_bytes, sort_bytes;
    float *P;
    size_t P_floats;
};

// carves everything but P; returns the bytes used.  n = padded pixel count of the chunk (tiles * 1024): the
// capacity of every per-pixel, per-item and per-segment array.
size_t carve_cells(CellBuffers &b, void *ws, size_t bytes, uint32_t n, const CellGrid &g)
{
    MbArena a(ws, bytes);
    const size_t ntiles = (size_t)n / TILE_PIX;
    const size_t words = (size_t)n / 32 + 2;
    const size_t V = (size_t)g.S0 * g.S1 * g.S2, vwords = V / 32 + 1;
    const size_t ncap = (size_t)n < (size_t)g.invalid ? (size_t)n : (size_t)g.invalid;   // unique cells
    const size_t vcap = V < 8 * ncap ? V : 8 * ncap;                                       // touched voxels
    b.counters = a.take<uint32_t>(MB_NUM_COUNTERS);       // first: mb_layer_update_status reads them at offset 0
    b.rec = a.take<uint4>(n);
    b.pix = a.take<uint4>(n);
    b.keys_a = a.take<uint32_t>(n); b.keys_b = a.take<uint32_t>(n);
    b.pids_a = a.take<uint32_t>(n); b.pids_b = a.take<uint32_t>(n);
    b.tcount = a.take<uint32_t>(ntiles + 1); b.toff = a.take<uint32_t>(ntiles + 1);
    b.smask = a.take<uint32_t>(words); b.soff = a.take<uint32_t>(words);
    {
        // (sizes are multiples of 4 words: the bitmap behind them is read in 16-byte pieces)
        const size_t iwords = ((((size_t)n + IDX_TILE - 1) / IDX_TILE * 2 + 8) + 3) & ~(size_t)3;
        const size_t rwords = (((ncap + 255) / 256 + 8) + 3) & ~(size_t)3;
        const size_t vxwords = (((vwords + VOX_TILE - 1) / VOX_TILE + 8) + 3) & ~(size_t)3;
        const size_t swords = (mb_scan_state_words((uint32_t)ntiles + 1) + 3) & ~(size_t)3;
        b.idx_state = a.take<uint32_t>(iwords + rwords + swords + vxwords + vwords + VOX_WORDS);
        b.run_state = b.idx_state + iwords;
        b.scan_state = b.run_state + rwords;
        b.vox_state = b.scan_state + swords;
        b.bitmap = b.vox_state + vxwords;             // touched-voxel bitmap
        b.state_bytes = (iwords + rwords + swords + vxwords + vwords + VOX_WORDS) * sizeof(uint32_t);
    }
    b.ucell = a.take<uint32_t>(ncap + 1); b.cstart = a.take<uint32_t>(ncap + 1);
    b.cseg = a.take<uint32_t>(ncap + 1); b.crun = a.take<uint32_t>(ncap + 1);
    b.rstart = a.take<uint32_t>(ncap + ((size_t)n + TASK_ITEMS - 1) / TASK_ITEMS + 2);     // worst_runs() + sentinel
    b.seg_start = a.take<uint32_t>((size_t)n + 1); b.seg_frame = a.take<uint32_t>((size_t)n + 1);
    b.segws = a.take<float2>((size_t)n * 8);
    b.gcoef = a.take<float>((size_t)n * 8);
    b.vlist = a.take<uint32_t>(vcap + 1);
    b.vA = a.take<float>(vcap + 1);
    b.vseg = a.take<uint2>((vcap + 1) * 8);
    b.vrun = a.take<uint2>((vcap + 1) * 8);
    b.vslot = a.take<uint32_t>(vcap + 1);
    b.long_segs = a.take<uint32_t>((size_t)n / (SEG_LONG_ITEMS + 1) + 16);      // (items <= pixels <= n)
    {
        // dense lookup table over the extended grid when it is not out of proportion to the batch
        // (it is re-initialised by every call: 4 bytes per cell of the extended grid; a one-frame call on a large
        // map is better off with the binary search over its few thousand cells)
        const size_t budget = (size_t)32 * n;
        const bool dense = (size_t)g.invalid * sizeof(int) <= budget;
        b.ctab = dense ? a.take<int>((size_t)g.invalid) : nullptr;
        if (!dense) b.ctab = nullptr;
    }
    const size_t scan_n = ntiles + 1;
    b.scan_bytes = mb_scan_workspace_bytes((uint32_t)scan_n);
    b.scan_ws = a.take<char>(b.scan_bytes);
    b.sort_bytes = mb_sort_workspace_bytes(n);
    b.sort_ws = a.take<char>(b.sort_bytes);
    b.P = a.take<float>(0);
    const size_t used = mb_align_up(a.used);
    b.P_floats = bytes > used ? (bytes - used) / sizeof(float) : 0;
    return used;
}

// most runs a call can produce: every cell starts one, every task start (TASK_ITEMS items) may split one
size_t worst_runs(uint32_t n, const CellGrid &g)
{
    const size_t ncap = (size_t)n < (size_t)g.invalid ? (size_t)n : (size_t)g.invalid;
    return ncap + ((size_t)n + TASK_ITEMS - 1) / TASK_ITEMS;
}

void pick_vec(const float *features, const float *map, const float *P, int F, int &vec, int &it)
{
    const uintptr_t al = (features ? (uintptr_t)features : 0) | (uintptr_t)map | (uintptr_t)P;
    vec = 1;
    if (F % 4 == 0 && al % 16 == 0) vec = 4;
    else if (F % 2 == 0 && al % 8 == 0) vec = 2;
    // one iteration per lane and several channel blocks (grid.y) beat two iterations: three times the warps per SM
    // for twice the (cheap) coefficient staging; two iterations only where a second block would be nearly empty
    it = (F > 32 * vec && F <= 32 * vec + 8 * vec) ? 2 : 1;
}

template <int VEC, int IT, bool ONEHOT, int U>
int launch_accumulate(cudaStream_t stream, const AccArgs &A)
{
    auto kern = k_cell_accumulate<VEC, IT, ONEHOT, U>;
    int per_sm = 1;
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem)));
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ACC_THREADS, sizeof(AccSmem)));
    if (per_sm < 1) per_sm = 1;
    static const int cap_per_sm = getenv("MASSB200_ACC_PER_SM") ? atoi(getenv("MASSB200_ACC_PER_SM")) : 0;   // measurement aid
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
    kern<<<MB_NUM_SMS * per_sm, ACC_THREADS, sizeof(AccSmem), stream>>>(A);      // persistent: the kernel walks (run, channel block) units
    MB_LAUNCHED();
    return MB_OK;
}

int dispatch_accumulate(cudaStream_t stream, const AccArgs &A, int vec, int it)
{
    const bool oh = A.class_ids != nullptr;
#define MB_ACC(V, I, UU)                                                              \
    if (vec == V && it == I)                                                          \
        return oh ? launch_accumulate<V, I, true, UU>(stream, A) : launch_accumulate<V, I, false, UU>(stream, A)
    MB_ACC(1, 1, 8); MB_ACC(1, 2, 4); MB_ACC(2, 1, MB_ACC_U21); MB_ACC(2, 2, 4); MB_ACC(4, 1, 4); MB_ACC(4, 2, 2);
#undef MB_ACC
    mb_set_error("internal: no accumulate kernel for vec %d it %d", vec, it);
    return MB_ERR_ARG;
}

template <int VEC, int IT>
int launch_apply(cudaStream_t stream, const ApplyArgs &A)
{
    const int cblocks = (A.F + 32 * VEC * IT - 1) / (32 * VEC * IT);
    dim3 grid(MB_NUM_SMS * 8, cblocks);
    k_voxel_apply<VEC, IT><<<grid, 256, 0, stream>>>(A);
    MB_LAUNCHED();
    return MB_OK;
}

template <int VEC, int IT, bool ONEHOT, int U>
int launch_overflow(cudaStream_t stream, const AccArgs &A, const ApplyArgs &Y, int rounds)
{
    auto kern = k_overflow_rounds<VEC, IT, ONEHOT, U>;
    int per_sm = 1;
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem)));
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ACC_THREADS, sizeof(AccSmem)));
    if (per_sm < 1) per_sm = 1;
    AccArgs a = A;
    ApplyArgs y = Y;
    int r = rounds;
    void *args[] = { &a, &y, &r };
    // cooperative: every CTA is resident (grid = SMs x occupancy), so the grid-wide barriers between the phases are safe
    MB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(MB_NUM_SMS * per_sm), dim3(ACC_THREADS), args,
                                              sizeof(AccSmem), stream));
    MB_LAUNCHED();
    return MB_OK;
}

int dispatch_overflow(cudaStream_t stream, const AccArgs &A, const ApplyArgs &Y, int rounds, int vec, int it)
{
    const bool oh = A.class_ids != nullptr;
#define MB_OVF(V, I, UU)                                                              \
    if (vec == V && it == I)                                                          \
        return oh ? launch_overflow<V, I, true, UU>(stream, A, Y, rounds) : launch_overflow<V, I, false, UU>(stream, A, Y, rounds)
    MB_OVF(1, 1, 8); MB_OVF(1, 2, 4); MB_OVF(2, 1, MB_ACC_U21); MB_OVF(2, 2, 4); MB_OVF(4, 1, 4); MB_OVF(4, 2, 2);
#undef MB_OVF
    mb_set_error("internal: no overflow kernel for vec %d it %d", vec, it);
    return MB_ERR_ARG;
}

// single-channel feature maps (occupancy): lanes = pixels
bool single_channel(const AccArgs &A) { return A.features != nullptr && A.F == 1; }

int launch_accumulate_single(cudaStream_t stream, const AccArgs &A)
{
    int per_sm = 1;
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_cell_accumulate_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmemSingle)));
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cell_accumulate_single, ACC_THREADS, sizeof(AccSmemSingle)));
    if (per_sm < 1) per_sm = 1;
    k_cell_accumulate_single<<<MB_NUM_SMS * per_sm, ACC_THREADS, sizeof(AccSmemSingle), stream>>>(A);
    MB_LAUNCHED();
    return MB_OK;
}

int launch_overflow_single(cudaStream_t stream, const AccArgs &A, const ApplyArgs &Y, int rounds)
{
    int per_sm = 1;
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_overflow_rounds_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmemSingle)));
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_overflow_rounds_single, ACC_THREADS, sizeof(AccSmemSingle)));
    if (per_sm < 1) per_sm = 1;
    AccArgs a = A;
    ApplyArgs y = Y;
    int r = rounds;
    void *args[] = { &a, &y, &r };
    MB_CHECK_CUDA(cudaLaunchCooperativeKernel((const void *)k_overflow_rounds_single, dim3(MB_NUM_SMS * per_sm), dim3(ACC_THREADS),
                                              args, sizeof(AccSmemSingle), stream));
    MB_LAUNCHED();
    return MB_OK;
}

int dispatch_apply(cudaStream_t stream, const ApplyArgs &A, int vec, int it)
{
#define MB_APP(V, I) if (vec == V && it == I) return launch_apply<V, I>(stream, A)
    MB_APP(1, 1); MB_APP(1, 2); MB_APP(2, 1); MB_APP(2, 2); MB_APP(4, 1); MB_APP(4, 2);
#undef MB_APP
    mb_set_error("internal: no apply kernel for vec %d it %d", vec, it);
    return MB_ERR_ARG;
}

}  // namespace

// ---- optional stage timing (mb_profile_stages / mb_profile_read): CUDA events between the stages of the
// last chunk processed while profiling was enabled.  Off by default: no events are recorded.
namespace {
constexpr int N_STAGES = 6;      // voxelise, sort, index, scalars, accumulate, apply
bool g_profile = false;
cudaEvent_t g_ev[N_STAGES + 1];
bool g_ev_ready = false, g_ev_recorded = false;

int stage_mark(cudaStream_t stream, int i)
{
    if (!g_profile) return MB_OK;
    if (!g_ev_ready) {
        for (int k = 0; k <= N_STAGES; ++k) MB_CHECK_CUDA(cudaEventCreate(&g_ev[k]));
        g_ev_ready = true;
    }
    MB_CHECK_CUDA(cudaEventRecord(g_ev[i], stream));
    if (i == N_STAGES) g_ev_recorded = true;
    return MB_OK;
}
}  // namespace

int mbk_profile_enable(int enable) { g_profile = enable != 0; return MB_OK; }

// ---- side branch of a batched update: the small, latency-bound launches between the index sweep and the voxel
// scalar pass (accumulate runs, touched-voxel list, per-voxel source ranges: ~40 us of nearly empty GPU) and the
// 4-byte-per-cell table memset do not depend on the segment sums, so they run on a library-owned stream next to
// them: fork and join with events, which also makes them parallel branches when the caller captures the update in
// a CUDA graph.  MASSB200_NO_FORK=1 keeps everything on the caller's stream (measurement aid).
namespace {
struct SideBranch {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, ctab_done = nullptr, indexed = nullptr, sources_done = nullptr;
};
constexpr int MAX_DEVICES = 64;
SideBranch g_side[MAX_DEVICES];
std::mutex g_side_mutex;

// the calling device's side branch, or null (forking disabled / device index out of range / the branch would have to
// be created while the caller's stream is being captured: stream and event creation stay out of captures, that call
// runs in line)
int side_branch(cudaStream_t caller, SideBranch **out)
{
    static const bool no_fork = getenv("MASSB200_NO_FORK") != nullptr;
    *out = nullptr;
    if (no_fork) return MB_OK;
    int dev = 0;
    MB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) return MB_OK;
    SideBranch &sb = g_side[dev];
    if (sb.stream == nullptr) {
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        MB_CHECK_CUDA(cudaStreamIsCapturing(caller, &capturing));
        if (capturing != cudaStreamCaptureStatusNone) return MB_OK;
        cudaStream_t st;
        MB_CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        cudaEvent_t *ev[4] = { &sb.fork, &sb.ctab_done, &sb.indexed, &sb.sources_done };
        for (auto e : ev) MB_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        sb.stream = st;
    }
    *out = &sb;
    return MB_OK;
}
}  // namespace

int mbk_profile_read(float *ms_host, int capacity)
{
    if (!g_ev_recorded) return 0;
    if (cudaEventSynchronize(g_ev[N_STAGES]) != cudaSuccess) return 0;
    int n = 0;
    for (int k = 0; k < N_STAGES && k < capacity; ++k, ++n)
        if (cudaEventElapsedTime(&ms_host[k], g_ev[k], g_ev[k + 1]) != cudaSuccess) return n;
    return n;
}

int mbk_affine_apply_rows(cudaStream_t stream, float *map, int F, const int64_t *idx, const float *a, const float *b,
                          int64_t n)
{
    if (n <= 0) return MB_OK;
    k_affine_apply_rows<<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, F, idx, a, b, n);
    MB_LAUNCHED();
    return MB_OK;
}

size_t mbk_partial_buffer_layout(uint32_t capacity, int F, size_t *offsets)
{
    size_t off[4], total;
    partial_layout(capacity, F, off, &total);
    if (offsets) for (int i = 0; i < 4; ++i) offsets[i] = off[i];
    return total;
}

int mbk_partial_reset(cudaStream_t stream, int32_t *slot_table, int64_t voxels, void *buffer)
{
    MB_CHECK_CUDA(cudaMemsetAsync(slot_table, 0xff, (size_t)voxels * sizeof(int32_t), stream));
    MB_CHECK_CUDA(cudaMemsetAsync(buffer, 0, 256, stream));
    return MB_OK;
}

int mbk_partial_clear(cudaStream_t stream, int32_t *slot_table, void *buffer, uint32_t capacity, int F)
{
    k_partial_clear<<<MB_NUM_SMS * 4, 256, 0, stream>>>(slot_table, partial_view(buffer, capacity, F), capacity);
    MB_LAUNCHED();
    MB_CHECK_CUDA(cudaMemsetAsync(buffer, 0, 256, stream));
    return MB_OK;
}

int mbk_partial_pull(cudaStream_t stream, const void *const *peer_buffers_host, void *const *slots_host, int world,
                     int self, uint32_t capacity, int F)
{
    MB_REQUIRE(world >= 1 && world <= 16 && self >= 0 && self < world, "mb_partial_pull: at most 16 ranks");
    if (world == 1) return MB_OK;
    PullArgs A;
    for (int g = 0; g < 16; ++g) { A.peer[g] = g < world ? peer_buffers_host[g] : nullptr; A.slot[g] = g < world ? slots_host[g] : nullptr; }
    A.world = world; A.self = self; A.capacity = capacity; A.F = F;
    k_partial_pull<<<MB_NUM_SMS * 8, 256, 0, stream>>>(A);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_affine_apply_partial(cudaStream_t stream, float *map, int F, const void *buffer, uint32_t capacity)
{
    const PartialView pv = partial_view(const_cast<void *>(buffer), capacity, F);
    const uintptr_t al = (uintptr_t)map | (uintptr_t)pv.b;
    if (F % 4 == 0 && al % 16 == 0) k_affine_apply_partial<4><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, F, pv, capacity);
    else if (F % 2 == 0 && al % 8 == 0) k_affine_apply_partial<2><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, F, pv, capacity);
    else k_affine_apply_partial<1><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, F, pv, capacity);
    MB_LAUNCHED();
    return MB_OK;
}

// bytes per P run (8 rows of F floats)
static size_t run_bytes(int F) { return (size_t)8 * F * sizeof(float); }

// P budget asked for by default: room for min(worst case, one run per 8 pixels), at least 128 MB
static size_t default_P_bytes(uint32_t n, const CellGrid &g, int F)
{
    const size_t worst = worst_runs(n, g) * run_bytes(F);
    size_t want = ((size_t)n / 8 + 1024) * run_bytes(F);
    if (want < ((size_t)128 << 20)) want = (size_t)128 << 20;
    return want < worst ? want : worst;
}

// padded pixel count of a chunk of T frames (tiles * 1024), 0 if the chunk is too large
static uint64_t padded_pixels(int H, int W, int T)
{
    const TileGeom tg = make_tiles(H, W);
    const uint64_t ntiles = (uint64_t)T * tg.tpf;
    if (ntiles >= MAX_TILES || T > MB_MAX_CHUNK_FRAMES) return 0;
    return ntiles * TILE_PIX;
}

size_t mbk_batch_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F)
{
    const CellGrid g = make_cells(ny - 1, nx - 1, nz - 1);
    CellBuffers b;
    const uint32_t n = (uint32_t)padded_pixels(H, W, T);
    return carve_cells(b, nullptr, 0, n, g) + default_P_bytes(n, g, F) + 512;
}

// smallest workspace that takes T frames in one chunk: the run buffer then holds 1/64 of the worst-case
// runs (never less than a few tasks' worth), i.e. the feature pass may take up to 64 rounds
size_t mbk_batch_min_workspace_bytes(int H, int W, int nx, int ny, int nz, int T, int F)
{
    const CellGrid g = make_cells(ny - 1, nx - 1, nz - 1);
    CellBuffers b;
    const uint32_t n = (uint32_t)padded_pixels(H, W, T);
    const size_t wr = worst_runs(n, g);
    size_t minruns = wr / 64 + 512;
    if (minruns > wr) minruns = wr;
    return carve_cells(b, nullptr, 0, n, g) + minruns * run_bytes(F) + 512;
}

// largest number of frames one chunk can hold at all (20-bit tile ids; MB_MAX_CHUNK_FRAMES frames)
int mbk_batch_max_chunk_frames(int H, int W)
{
    const TileGeom tg = make_tiles(H, W);
    int t = (int)((MAX_TILES - 1) / (uint32_t)tg.tpf);
    if (t > MB_MAX_CHUNK_FRAMES) t = MB_MAX_CHUNK_FRAMES;
    return t;
}

// frames per internal chunk for a given workspace; 0 if even one frame does not fit
int mbk_batch_frames_that_fit(int H, int W, int nx, int ny, int nz, int F, size_t workspace_bytes, int T)
{
    auto fits = [&](int t) {
        if (padded_pixels(H, W, t) == 0) return false;
        return mbk_batch_min_workspace_bytes(H, W, nx, ny, nz, t, F) <= workspace_bytes;
    };
    if (fits(T)) return T;
    int best = 0;
    for (int t = 1; t <= T; t = t < 8 ? t + 1 : t * 2) {
        if (fits(t)) best = t; else break;
    }
    return best;
}

// One chunk of T frames.
int mbk_batch_update(cudaStream_t stream, const float *rays, const float *depth, const float *features,
                     const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                     const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                     float *map, float *affine_a, float alpha, float min_d, float max_d, void *workspace,
                     size_t workspace_bytes, const MbSparseFold *sparse, bool keep_error_bits)
{
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    const TileGeom tg = make_tiles(H, W);
    const uint64_t padded = padded_pixels(H, W, T);
    MB_REQUIRE(padded != 0 && padded < 0x7fffffffull, "too many frames or pixels per chunk");
    const uint32_t n = (uint32_t)padded, ntiles = n / TILE_PIX;
    const CellGrid g = make_cells(ny - 1, nx - 1, nz - 1);
    MB_REQUIRE((uint64_t)g.E0 * g.E1 * g.E2 < 0xffffffe0ull, "map too large for 32-bit cell keys");
    MB_REQUIRE(class_ids != nullptr || (uint64_t)T * fh * fw < 0xffffffffull, "too many feature rows per chunk");
    CellBuffers b;
    MB_REQUIRE(carve_cells(b, workspace, workspace_bytes, n, g) + 512 <= workspace_bytes, "batch workspace too small");
    const size_t V = (size_t)g.S0 * g.S1 * g.S2;
    const uint32_t vwords = (uint32_t)(V / 32 + 1);
    int vec, it;
    pick_vec(features, sparse ? partial_view(sparse->buffer, sparse->capacity, F).b : map, b.P, F, vec, it);
    const size_t run_cap_sz = b.P_floats / ((size_t)8 * F);
    const size_t wruns = worst_runs(n, g);
    MB_REQUIRE(run_cap_sz >= (wruns < 512 ? wruns : 512), "batch workspace too small for the run buffer");
    const uint32_t run_cap = (uint32_t)(run_cap_sz < (1ull << 28) ? run_cap_sz : (1ull << 28));      // (run * 8 + slot stays 32-bit)
    const int rounds = (int)((wruns + run_cap - 1) / run_cap);
    MB_REQUIRE(rounds <= 4096, "batch workspace far too small for the run buffer");

    // K1: voxelise + group inside tiles, compact the items; then sort the items by cell
    int rc;
    if ((rc = stage_mark(stream, 0))) return rc;
    // the side branch's events are shared by all callers on this device: one enqueue at a time
    std::lock_guard<std::mutex> side_lock(g_side_mutex);
    SideBranch *side = nullptr;
    if ((rc = side_branch(stream, &side))) return rc;
    if (side != nullptr && b.ctab) {
        // the dense cell table is cleared next to the front end (only the index sweep needs it)
        MB_CHECK_CUDA(cudaEventRecord(side->fork, stream));
        MB_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        MB_CHECK_CUDA(cudaMemsetAsync(b.ctab, 0xff, (size_t)g.invalid * sizeof(int), side->stream));
        MB_CHECK_CUDA(cudaEventRecord(side->ctab_done, side->stream));
    }
    // every look-back word of the call and the touched-voxel bitmap: one memset up front
    MB_CHECK_CUDA(cudaMemsetAsync(b.idx_state, 0, b.state_bytes, stream));
    uint32_t *tkey = b.keys_b, *tval = b.pids_b;          // tile-local item lists live in the sort's second buffers
    dim3 vgrid((npix + 256 * VOX_PPT - 1) / (256 * VOX_PPT), (unsigned)T);
    k_cell_voxelise<<<vgrid, 256, 0, stream>>>(rays, depth, pose, npix, bins_x, nx, bins_y, ny, bins_z, nz, g, min_d, max_d,
                                               b.pix, b.counters, keep_error_bits);
    MB_LAUNCHED();
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_tile_group, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * sizeof(WarpTile))));
    k_tile_group<<<(ntiles + 7) / 8, 256, 8 * sizeof(WarpTile), stream>>>(b.pix, tg, ntiles, b.rec, tkey, tval, b.tcount);
    MB_LAUNCHED();
    if ((rc = mb_exclusive_scan_small(stream, b.tcount, b.toff, ntiles, b.scan_state))) return rc;
    k_tile_compact<<<(ntiles + 7) / 8, 256, 0, stream>>>(tkey, tval, b.tcount, b.toff, ntiles, b.keys_a, b.pids_a,
                                                         b.counters);
    MB_LAUNCHED();
    if ((rc = stage_mark(stream, 1))) return rc;
    int bits = 1;
    while (bits < 32 && (((uint64_t)1) << bits) <= (uint64_t)g.invalid) ++bits;
    uint32_t *ikey, *ival;
    const uint32_t *n_items = b.counters + MB_CNT_NVALID;
    rc = mb_sort_pairs(stream, b.keys_a, b.pids_a, b.keys_b, b.pids_b, n, n_items, bits, false, b.sort_ws, b.sort_bytes,
                       &ikey, &ival);
    if (rc) return rc;

    // K2: index sweep over the sorted items
    if ((rc = stage_mark(stream, 2))) return rc;
    if (b.ctab) {
        if (side != nullptr) MB_CHECK_CUDA(cudaStreamWaitEvent(stream, side->ctab_done, 0));
        else MB_CHECK_CUDA(cudaMemsetAsync(b.ctab, 0xff, (size_t)g.invalid * sizeof(int), stream));
    }
    {
        const uint32_t itiles = (n + IDX_TILE - 1) / IDX_TILE;
        IndexOut O;
        O.smask = b.smask; O.soff = b.soff;
        O.ucell = b.ucell; O.cstart = b.cstart; O.cseg = b.cseg;
        O.seg_start = b.seg_start; O.seg_frame = b.seg_frame; O.bitmap = b.bitmap; O.ctab = b.ctab;
        O.counters = b.counters; O.state = b.idx_state + 4; O.ticket = b.idx_state;
        k_cell_index<<<itiles, 256, 0, stream>>>(ikey, ival, n_items, tg, g, O);
        MB_LAUNCHED();
    }
    // K3, K4, K6a on the side branch (or in line): they need the index sweep only
    cudaStream_t sstream = stream;
    if (side != nullptr) {
        MB_CHECK_CUDA(cudaEventRecord(side->indexed, stream));
        MB_CHECK_CUDA(cudaStreamWaitEvent(side->stream, side->indexed, 0));
        sstream = side->stream;
    }
    // K3: accumulate runs (cell-aligned pieces of <= TASK_ITEMS items)
    {
        const uint64_t lim = (uint64_t)rounds * run_cap;
        k_cell_runs<<<MB_NUM_SMS * 4, 256, 0, sstream>>>(b.cstart, b.counters, b.run_state + 4, b.run_state, b.crun, b.rstart,
                                                         (uint32_t)(lim < 0xffffffffull ? lim : 0xffffffffull));
        MB_LAUNCHED();
    }
    // K4
    k_vox_list<<<(vwords + VOX_TILE - 1) / VOX_TILE, 256, 0, sstream>>>(b.bitmap, vwords, b.vox_state + 4, b.vox_state, b.vlist,
                                                                        b.counters);
    MB_LAUNCHED();
    // K6a
    k_voxel_sources<<<MB_NUM_SMS * 8, 256, 0, sstream>>>(b.vlist, b.ucell, b.ctab, b.cseg, b.crun, g, b.vseg, b.vrun,
                                                         b.counters);
    MB_LAUNCHED();
    if (side != nullptr) MB_CHECK_CUDA(cudaEventRecord(side->sources_done, side->stream));
    // K5, K6
    if ((rc = stage_mark(stream, 3))) return rc;
    {
        // one full wave: every CTA takes an equal share of the segments, so a partial second wave would idle SMs
        int per_sm = 1;
        MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seg_sums, 256, 0));
        if (per_sm < 1) per_sm = 1;
        static const int cap_per_sm = getenv("MASSB200_SCALAR_PER_SM") ? atoi(getenv("MASSB200_SCALAR_PER_SM")) : 0;   // measurement aid
        if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
        k_seg_sums<<<MB_NUM_SMS * per_sm, 256, 0, stream>>>(ival, b.rec, b.seg_start, b.segws, n, b.counters, b.long_segs);
        MB_LAUNCHED();
        k_seg_sums_long<<<MB_NUM_SMS * 8, 256, 0, stream>>>(ival, b.rec, b.seg_start, b.segws, n, b.counters, b.long_segs);
        MB_LAUNCHED();
    }
    {
        const size_t smem = (size_t)8 * 2 * ((T + 31) & ~31) * sizeof(float);
        MB_CHECK_CUDA(cudaFuncSetAttribute(k_voxel_scalars, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_voxel_scalars, 256, smem));
        if (per_sm < 1) per_sm = 1;
        static const int cap_per_sm = getenv("MASSB200_SCALAR_PER_SM") ? atoi(getenv("MASSB200_SCALAR_PER_SM")) : 0;   // measurement aid
        if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
        if (side != nullptr) MB_CHECK_CUDA(cudaStreamWaitEvent(stream, side->sources_done, 0));
        k_voxel_scalars<<<MB_NUM_SMS * per_sm, 256, smem, stream>>>(b.vlist, b.vseg, b.seg_frame, b.segws, n, g, alpha, T,
                                                                    b.gcoef, b.vA, b.counters);
        MB_LAUNCHED();
    }
    // K7, K8 (one round unless the runs outgrow the P buffer)
    AccArgs A;
    A.ikey = ikey; A.ival = ival; A.tg = tg; A.rec = b.rec; A.smask = b.smask; A.soff = b.soff; A.rstart = b.rstart;
    A.gcoef = b.gcoef; A.cap = n; A.counters = b.counters;
    A.fi = MbFeatIndex{ npix, (uint32_t)W, (uint32_t)(H / fh), (uint32_t)(W / fw), (uint32_t)fw };
    A.fhw = (uint32_t)fh * (uint32_t)fw;
    A.features = features; A.class_ids = class_ids; A.F = F; A.P = b.P; A.run_cap = run_cap;
    ApplyArgs Y;
    Y.vlist = b.vlist; Y.vrun = b.vrun; Y.vA = b.vA; Y.P = b.P; Y.counters = b.counters;
    Y.g = g; Y.F = F; Y.map = map; Y.affine_a = affine_a; Y.run_cap = run_cap;
    Y.vslot = nullptr; Y.part_a = Y.part_b = nullptr;
    if (sparse != nullptr) {
        // fold into a sparse partial: rows are found / created per touched voxel, then the apply kernel writes them
        const PartialView pv = partial_view(sparse->buffer, sparse->capacity, F);
        k_partial_slots<<<MB_NUM_SMS * 4, 256, 0, stream>>>(b.vlist, b.counters, sparse->slot_table, pv, sparse->capacity,
                                                            b.vslot, b.counters + MB_CNT_ERROR);
        MB_LAUNCHED();
        Y.vslot = b.vslot; Y.part_a = pv.a; Y.part_b = pv.b; Y.map = pv.b;
    }
    // round 0 holds every run of a real call; the rounds the worst case would need beyond it are one cooperative launch
    // that looks at the run count on the device
    A.run_base = Y.run_base = 0;
    A.round = 0;
    static const bool no_single = getenv("MASSB200_NO_SINGLE") != nullptr;    // measurement aid: lanes = channels for every input
    const bool single = single_channel(A) && !no_single;
    auto accumulate = [&]() {
        return single ? launch_accumulate_single(stream, A) : dispatch_accumulate(stream, A, vec, it);
    };
    if ((rc = stage_mark(stream, 4))) return rc;
    if ((rc = accumulate())) return rc;
    if ((rc = stage_mark(stream, 5))) return rc;
    if ((rc = dispatch_apply(stream, Y, vec, it))) return rc;
    static const bool no_coop = getenv("MASSB200_NO_COOP") != nullptr;        // measurement aid: one launch pair per round
    if (rounds > 1 && !no_coop) {
        rc = single ? launch_overflow_single(stream, A, Y, rounds) : dispatch_overflow(stream, A, Y, rounds, vec, it);
        if (rc) return rc;
    }
    for (int r = 1; r < rounds && no_coop; ++r) {
        A.run_base = Y.run_base = (uint32_t)r * run_cap;
        A.round = (uint32_t)r;
        if ((rc = accumulate())) return rc;
        if ((rc = dispatch_apply(stream, Y, vec, it))) return rc;
    }
    return stage_mark(stream, 6);
}
