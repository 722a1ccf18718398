// Batched mapping pipeline (sm_100a): T frames fused per call, in frame order.
//
//   K1  k_voxelise_batch   pixel -> 16-byte record {packed lower-corner voxel, 3 in-voxel ratios}
//                          + number of 4x4x4 voxel bricks its 2x2x2 splat footprint overlaps
//   --  exclusive scan of the brick counts (entry offsets, deterministic compaction)
//   K1' k_emit_entries     one (brick id, pixel id) entry per overlapped brick, in (frame, pixel) order
//   --  stable radix sort of the entries by brick id: inside a brick the entries stay in
//       (frame, pixel) order
//   K2  k_brick_reduce     one CTA per touched brick walks its entries frame by frame:
//         lanes = entries : recompute the 8 splat contributions (voxel, weight) of each entry
//         CTA             : stable counting sort of the contributions by voxel (64 bins, shared memory)
//         lanes = channels: one warp per voxel accumulates  W = sum w, S2 = sum w^2, B = sum w^2 f
//                           in registers and applies the per-voxel affine update of the frame
//                               new = (1 - alpha*S2/W) * old + (alpha/W) * B        
//                           to the map row (read-modify-write; the rows of a brick stay in L1/L2
//                           while its CTA walks the frames).
//
// No float atomics anywhere: every voxel of every frame is summed by one warp in an order fixed by
// the sorted entry list, so the result is bit-reproducible.  Weights follow the reference's fp32
// operation order (/root/reference/mass/utils/projection.py:280-323); occupancy is therefore
// bit-exact and values differ from the reference CPU path by fp32 re-association only.
#include "common.cuh"
#include "kernels.cuh"
#include "geometry.cuh"

// Optional phase timing of k_brick_reduce_qw (build with -DMB_PHASE_TIMING; read with
// mb_debug_phase_cycles).  Not compiled into the shipped library.
#ifdef MB_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define MB_T0() long long t_prev = clock64(); long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define MB_TICK(i) do { const long long t_now = clock64(); t_acc[i] += t_now - t_prev; t_prev = t_now; } while (0)
#define MB_TFLUSH() do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_phase_cycles[i_], (unsigned long long)t_acc[i_]); } while (0)
#else
#define MB_T0()
#define MB_TICK(i)
#define MB_TFLUSH()
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RED_THREADS = 256;               // threads per brick CTA = entries per chunk
constexpr int RED_WARPS = RED_THREADS / 32;
constexpr int RED_CONTRIB = RED_THREADS * 8;   // contributions per chunk
constexpr uint32_t REC_INVALID = 0xffffffffu;
constexpr uint32_t NO_FRAME = 0xffffffffu;

// brick geometry of the map (bricks are 4 x 4 x 4 voxels)
struct MbBricks {
    int S0, S1, S2;      // map dims (y flipped, x, z)
    int N0, N1, N2;      // bricks per axis
};

__host__ __device__ inline MbBricks make_bricks(int S0, int S1, int S2)
{
    MbBricks b;
    b.S0 = S0; b.S1 = S1; b.S2 = S2;
    b.N0 = (S0 + 3) >> 2; b.N1 = (S1 + 3) >> 2; b.N2 = (S2 + 3) >> 2;
    return b;
}

// record key: lower-corner coordinate + 1 per axis (the corner may be -1 at the map border, where
// the reference clamps it: projection.py:280-291), 11 + 11 + 10 bits
__device__ __forceinline__ uint32_t pack_corner(int b0, int b1, int b2)
{
    return (uint32_t)(b0 + 1) | ((uint32_t)(b1 + 1) << 11) | ((uint32_t)(b2 + 1) << 22);
}

struct Footprint {
    int lo[3], hi[3];      // clamped voxel coordinates of the two neighbours per axis
};

__device__ __forceinline__ Footprint footprint_of(uint32_t key, const MbBricks &g)
{
    Footprint f;
    const int b0 = (int)(key & 2047u) - 1, b1 = (int)((key >> 11) & 2047u) - 1, b2 = (int)(key >> 22) - 1;
    f.lo[0] = max(b0, 0); f.hi[0] = min(b0 + 1, g.S0 - 1);
    f.lo[1] = max(b1, 0); f.hi[1] = min(b1 + 1, g.S1 - 1);
    f.lo[2] = max(b2, 0); f.hi[2] = min(b2 + 1, g.S2 - 1);
    return f;
}

// ---------------------------------------------------------------------------------------------
// K1: grid = (pixel blocks, frames)
__global__ void __launch_bounds__(256)
k_voxelise_batch(const float *__restrict__ rays, const float *__restrict__ depth, const float *__restrict__ pose,
                 uint32_t npix, const float *__restrict__ bins_x, int nx, const float *__restrict__ bins_y, int ny,
                 const float *__restrict__ bins_z, int nz, MbBricks g, float min_d, float max_d,
                 uint4 *__restrict__ rec, uint32_t *__restrict__ cnt, uint32_t *__restrict__ counters)
{
    __shared__ float P[12];
    const uint32_t t = blockIdx.y;
    if (threadIdx.x < 12) P[threadIdx.x] = pose[(size_t)t * 12 + threadIdx.x];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < MB_NUM_COUNTERS) counters[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const size_t pid = (size_t)t * npix + p;
    float r0, r1, r2;
    orient(P, rays[3 * (size_t)p], rays[3 * (size_t)p + 1], rays[3 * (size_t)p + 2], r0, r1, r2);
    const BinResult b = bin_point(bins_x, nx, bins_y, ny, bins_z, nz, P[9], P[10], P[11], r0, r1, r2,
                                  depth[pid], min_d, max_d);
    uint4 out = make_uint4(REC_INVALID, 0u, 0u, 0u);
    uint32_t nb = 0;
    if (b.ok) {
        // map axes are (y flipped, x, z) = input axes (1, 0, 2): base_projection_layer.py:339
        const float q0 = b.q1, q1 = b.q0, q2 = b.q2;
        const int c0 = q0 < 0.5f ? b.i1 - 1 : b.i1;
        const int c1 = q1 < 0.5f ? b.i0 - 1 : b.i0;
        const int c2 = q2 < 0.5f ? b.i2 - 1 : b.i2;
        out = make_uint4(pack_corner(c0, c1, c2), __float_as_uint(q0), __float_as_uint(q1), __float_as_uint(q2));
        const Footprint f = footprint_of(out.x, g);
        nb = (1u + ((f.lo[0] >> 2) != (f.hi[0] >> 2))) * (1u + ((f.lo[1] >> 2) != (f.hi[1] >> 2))) *
             (1u + ((f.lo[2] >> 2) != (f.hi[2] >> 2)));
    }
    rec[pid] = out;
    cnt[pid] = nb;
}

// K1': offs = exclusive scan of cnt
__global__ void __launch_bounds__(256)
k_emit_entries(const uint4 *__restrict__ rec, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ offs,
               uint32_t ntotal, MbBricks g, uint32_t *__restrict__ keys, uint32_t *__restrict__ pids,
               uint32_t *__restrict__ counters)
{
    const uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= ntotal) return;
    const uint32_t nb = cnt[pid];
    uint32_t o = offs[pid];
    if (pid == ntotal - 1) counters[MB_CNT_ENTRIES] = o + nb;
    if (nb == 0) return;
    const Footprint f = footprint_of(rec[pid].x, g);
    const int a0 = f.lo[0] >> 2, a1 = f.hi[0] >> 2, b0 = f.lo[1] >> 2, b1 = f.hi[1] >> 2, c0 = f.lo[2] >> 2,
              c1 = f.hi[2] >> 2;
    for (int a = a0; a <= a1; ++a)
        for (int b = b0; b <= b1; ++b)
            for (int c = c0; c <= c1; ++c) {
                keys[o] = (uint32_t)((a * g.N1 + b) * g.N2 + c);
                pids[o] = pid;
                ++o;
            }
}

// One (start, entries) pair per touched brick of the sorted entry list, bucketed by log2(entries) so
// that the reduce kernel can hand out the heaviest bricks first.  The order inside a bucket is
// arbitrary: bricks are independent, the result does not depend on it.
__global__ void __launch_bounds__(256)
k_brick_heads(const uint32_t *__restrict__ keys, uint32_t nmax, uint2 *__restrict__ bricks,
              uint32_t *__restrict__ counters)
{
    const uint32_t n = min(nmax, counters[MB_CNT_ENTRIES]);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool head = i < n && (i == 0 || keys[i - 1] != keys[i]);
    const uint32_t m = __ballot_sync(FULL, head);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&counters[MB_CNT_BRICKS], (uint32_t)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (!head) return;
    const uint32_t key = keys[i];
    uint32_t lo = i + 1, hi = n;                       // first index past the brick's segment
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] == key) lo = mid + 1; else hi = mid;
    }
    const uint32_t count = lo - i;
    bricks[base + __popc(m & ((1u << lane) - 1u))] = make_uint2(i, count);
    atomicAdd(&counters[MB_CNT_BUCKET + (31 - __clz(count))], 1u);
}

__global__ void __launch_bounds__(256)
k_brick_order(const uint2 *__restrict__ bricks, uint32_t *__restrict__ order, uint32_t *__restrict__ counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[MB_CNT_BRICKS]) return;
    const int bucket = 31 - __clz(bricks[i].y);
    uint32_t base = 0;
    for (int b = 31; b > bucket; --b) base += counters[MB_CNT_BUCKET + b];
    order[base + atomicAdd(&counters[MB_CNT_FILL + bucket], 1u)] = i;
}

// ---------------------------------------------------------------------------------------------
// K2
struct ReduceArgs {
    const uint32_t *keys, *pids;
    const uint2 *bricks;        // (start, entries) per touched brick
    const uint32_t *order;      // brick list indices, heaviest bricks first
    uint32_t *counters;
    uint32_t nmax;
    const uint4 *rec;
    MbFeatIndex fi;             // np = pixels per frame
    uint32_t fhw;               // feature rows per frame
    const float *features;      // [T][fhw][F] or null
    const float *features_end;  // one past the last feature row
    const int64_t *class_ids;   // [T][np] or null
    int F;
    float *map;                 // [S0][S1][S2][F], updated in place
    float *affine_a;            // optional [S0*S1*S2]: multiplied by every frame's a (affine output mode)
    MbBricks g;
    float alpha;
};

template <int VEC>
__device__ __forceinline__ void vec_load(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = *p;
    if (VEC == 2) { const float2 v = *(const float2 *)p; dst[0] = v.x; dst[1] = v.y; }
    if (VEC == 4) { const float4 v = *(const float4 *)p; dst[0] = v.x; dst[1] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w; }
}

template <int VEC>
__device__ __forceinline__ void vec_load_nc(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = __ldg(p);
    if (VEC == 2) { const float2 v = __ldg((const float2 *)p); dst[0] = v.x; dst[1] = v.y; }
    if (VEC == 4) { const float4 v = __ldg((const float4 *)p); dst[0] = v.x; dst[1] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w; }
}

template <int VEC>
__device__ __forceinline__ void vec_store(float *p, const float (&src)[VEC])
{
    if (VEC == 1) *p = src[0];
    if (VEC == 2) *(float2 *)p = make_float2(src[0], src[1]);
    if (VEC == 4) *(float4 *)p = make_float4(src[0], src[1], src[VEC > 2 ? 2 : 0], src[VEC > 2 ? 3 : 0]);
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One frame's affine update of a voxel row held in registers: row = a*row + (alpha/W)*B, a = 1 - alpha*S2/W.
template <int VEC, int IT>
__device__ __forceinline__ void apply_frame(float alpha, float W, float S2, float (&row)[IT][VEC],
                                            float (&acc)[IT][VEC], float &aprod)
{
    const float a = 1.0f - alpha * S2 / W;
    const float sc = alpha / W;
#pragma unroll
    for (int it = 0; it < IT; ++it)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            row[it][j] = fmaf(a, row[it][j], sc * acc[it][j]);
            acc[it][j] = 0.f;
        }
    aprod *= a;
}

// Shared memory plan of k_brick_reduce (dynamic):
//   s_con   [2048] float4  sorted contributions {w, w^2, tag = entry | frame offset << 8, -}
//   s_part  [64][RS]       partial sums of a frame that continues into the next chunk
//   s_feat  [256][F]       staged feature rows of the chunk's entries (STAGE only)
//   s_cnt   [8][64], s_start[64], s_total[64], s_W[64], s_S2[64], s_esrc[256], s_eframe[256]
template <int VEC, int IT, bool ONEHOT, bool STAGE>
__global__ void __launch_bounds__(RED_THREADS)
k_brick_reduce(const ReduceArgs A)
{
    constexpr int RS = 32 * VEC * IT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_con = (float4 *)smem_raw;
    float *s_part = (float *)(s_con + RED_CONTRIB);
    uint32_t *s_cnt = (uint32_t *)(s_part + 64 * RS);
    uint32_t *s_start = s_cnt + RED_WARPS * 64;
    uint32_t *s_total = s_start + 64;
    float *s_W = (float *)(s_total + 64);
    float *s_S2 = s_W + 64;
    uint32_t *s_esrc = (uint32_t *)(s_S2 + 64);
    uint32_t *s_eframe = s_esrc + RED_THREADS;
    float *s_feat = (float *)(s_eframe + RED_THREADS);
    __shared__ uint32_t s_ticket, s_next, s_cont;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ltmask = (1u << lane) - 1u;
    const uint32_t n = min(A.nmax, A.counters[MB_CNT_ENTRIES]);
    const uint32_t nbricks = A.counters[MB_CNT_BRICKS];
    const uint32_t np = A.fi.np;
    const int F = A.F;

    if (tid < 64) { s_W[tid] = 0.f; s_S2[tid] = 0.f; }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(&A.counters[MB_CNT_TICKET], 1u);
        __syncthreads();
        const uint32_t ticket = s_ticket;
        if (ticket >= nbricks) break;
        const uint2 brick = A.bricks[A.order[ticket]];
        const uint32_t bend = brick.x + brick.y;
        const uint32_t bkey = A.keys[brick.x];
        const int bz = bkey % A.g.N2, by = (bkey / A.g.N2) % A.g.N1, bx = bkey / (A.g.N2 * A.g.N1);
        const int org0 = bx * 4, org1 = by * 4, org2 = bz * 4;
        uint32_t pending = NO_FRAME;          // frame whose partial sums sit in s_part / s_W / s_S2

        for (uint32_t pos = brick.x; pos < bend; pos += RED_THREADS) {
            // ---- entries of this chunk ---------------------------------------------------------------
            const uint32_t idx = pos + tid;
            const bool active = idx < bend;
            const int nact = min((uint32_t)RED_THREADS, bend - pos);
            const uint32_t pid = active ? A.pids[idx] : 0u;
            const uint32_t frame = pid / np;
            uint32_t src = 0;
            if (active) {
                const uint32_t p = pid - frame * np;
                if (ONEHOT) {
                    src = (uint32_t)A.class_ids[pid];
                } else if (A.fi.kx == 1 && A.fi.ky == 1) {
                    src = frame * A.fhw + p;
                } else {
                    const uint32_t y = p / A.fi.W, x = p - y * A.fi.W;
                    src = frame * A.fhw + (y / A.fi.ky) * A.fi.fw + x / A.fi.kx;
                }
            }
            s_esrc[tid] = src;
            s_eframe[tid] = frame;
            if (tid == nact - 1) {
                // does the chunk's last frame continue in the next chunk of this brick?
                const uint32_t nxt = idx + 1;
                s_cont = (nact == RED_THREADS && nxt < bend && A.pids[nxt] / np == frame) ? 1u : 0u;
            }
            for (int i = tid; i < RED_WARPS * 64; i += RED_THREADS) s_cnt[i] = 0;
            if (tid == 0) s_next = 0;
            __syncthreads();
            const uint32_t f_first = s_eframe[0], f_last = s_eframe[nact - 1];
            const bool cont = s_cont != 0;

            if (STAGE) {   // feature rows of the chunk -> shared memory, asynchronously
                const int per_row = F / VEC;
                for (int i = tid; i < nact * per_row; i += RED_THREADS) {
                    const int e = i / per_row, c = (i - e * per_row) * VEC;
                    cp_async<4 * VEC>(s_feat + e * F + c, A.features + (size_t)s_esrc[e] * F + c);
                }
            }

            // ---- lanes = entries: one contribution per voxel parity class ------------------------------
            uint32_t cv[8];      // local voxel (0..63) or 0xff
            float cw[8], cw2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { cv[k] = 0xffu; cw[k] = 0.f; cw2[k] = 0.f; }
            if (active) {
                const uint4 r = A.rec[pid];
                const Footprint f = footprint_of(r.x, A.g);
                const float q[3] = { __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w) };
                float wl[3], wu[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {               // projection.py:300-316
                    const bool low = q[a] < 0.5f;
                    wl[a] = low ? __fsub_rn(0.5f, q[a]) : __fsub_rn(1.5f, q[a]);
                    wu[a] = low ? __fadd_rn(q[a], 0.5f) : __fsub_rn(q[a], 0.5f);
                }
                const int l0lo = f.lo[0] - org0, l1lo = f.lo[1] - org1, l2lo = f.lo[2] - org2;
                const int l0hi = f.hi[0] - org0, l1hi = f.hi[1] - org1, l2hi = f.hi[2] - org2;
                const bool clamped = f.lo[0] == f.hi[0] || f.lo[1] == f.hi[1] || f.lo[2] == f.hi[2];
                if (!clamped) {
                    // the two neighbours of an axis have opposite parity: class k takes, per axis, the
                    // neighbour whose coordinate parity equals the class bit
                    const int p0 = f.lo[0] & 1, p1 = f.lo[1] & 1, p2 = f.lo[2] & 1;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool u0 = (((k >> 2) & 1) != p0), u1 = (((k >> 1) & 1) != p1), u2 = ((k & 1) != p2);
                        const int l0 = u0 ? l0hi : l0lo, l1 = u1 ? l1hi : l1lo, l2 = u2 ? l2hi : l2lo;
                        if ((unsigned)l0 < 4u && (unsigned)l1 < 4u && (unsigned)l2 < 4u) {
                            cv[k] = (uint32_t)((l0 << 4) | (l1 << 2) | l2);
                            const float w = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn(u0 ? wu[0] : wl[0], u1 ? wu[1] : wl[1]),
                                                                       u2 ? wu[2] : wl[2]));   // projection.py:319-323
                            cw[k] = w;
                            cw2[k] = w * w;
                        }
                    }
                } else {
                    // map border: clamping folds neighbours onto one voxel (projection.py:280-291); the
                    // folded slots are summed into one contribution of that voxel's class
                    for (int s = 0; s < 8; ++s) {
                        const int l0 = (s & 4) ? l0hi : l0lo, l1 = (s & 2) ? l1hi : l1lo, l2 = (s & 1) ? l2hi : l2lo;
                        if ((unsigned)l0 < 4u && (unsigned)l1 < 4u && (unsigned)l2 < 4u) {
                            const float w = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn((s & 4) ? wu[0] : wl[0], (s & 2) ? wu[1] : wl[1]),
                                                                       (s & 1) ? wu[2] : wl[2]));
                            const int cls = ((l0 & 1) << 2) | ((l1 & 1) << 1) | (l2 & 1);
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (k == cls) {
                                    cv[k] = (uint32_t)((l0 << 4) | (l1 << 2) | l2);
                                    cw[k] += w;
                                    cw2[k] += w * w;
                                }
                        }
                    }
                }
            }

            // ---- stable counting sort of the contributions by voxel (entry order inside a voxel) --------
            uint32_t rk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool ok = cv[k] != 0xffu;
                const uint32_t m = __match_any_sync(FULL, ok ? cv[k] : 64u + lane);
                rk[k] = __popc(m & ltmask);
                if (ok && rk[k] == 0) s_cnt[warp * 64 + cv[k]] = __popc(m);   // classes never share a voxel
            }
            __syncthreads();
            if (tid < 64) {
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < RED_WARPS; ++w) {
                    const uint32_t c = s_cnt[w * 64 + tid];
                    s_cnt[w * 64 + tid] = run;
                    run += c;
                }
                s_total[tid] = run;
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t t0 = s_total[2 * lane], t1 = s_total[2 * lane + 1];
                uint32_t inc = t0 + t1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += u;
                }
                const uint32_t ex = inc - (t0 + t1);
                s_start[2 * lane] = ex;
                s_start[2 * lane + 1] = ex + t0;
            }
            __syncthreads();
            {
                const uint32_t tag = (uint32_t)tid | ((frame - f_first) << 8);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (cv[k] != 0xffu)
                        s_con[s_start[cv[k]] + s_cnt[warp * 64 + cv[k]] + rk[k]] =
                            make_float4(cw[k], cw2[k], __uint_as_float(tag), 0.f);
            }
            if (STAGE) cp_async_wait_all();
            __syncthreads();

            // ---- lanes = channels: one warp per voxel walks its contributions in frame order -----------
            const uint32_t goff_last = f_last - f_first;
            for (;;) {
                int v = 0;
                if (lane == 0) v = (int)atomicAdd(&s_next, 1u);
                v = __shfl_sync(FULL, v, 0);
                if (v >= 64) break;
                const uint32_t nv = s_total[v];
                const float Wp = pending != NO_FRAME ? s_W[v] : 0.f;
                if (nv == 0 && !(Wp > 0.f)) continue;
                if (nv == 0 && cont && goff_last == 0) continue;       // still pending, nothing new

                const size_t vox = ((size_t)(org0 + (v >> 4)) * A.g.S1 + (org1 + ((v >> 2) & 3))) * A.g.S2 +
                                   (org2 + (v & 3));
                float *grow = A.map + vox * (size_t)F;
                float row[IT][VEC], acc[IT][VEC];
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const int ch = (it * 32 + lane) * VEC;
                    if (ch < F) vec_load<VEC>(row[it], grow + ch);
                    else
#pragma unroll
                        for (int j = 0; j < VEC; ++j) row[it][j] = 0.f;
                }
                float W = 0.f, S2 = 0.f, aprod = 1.0f;
                bool dirty = false;                                     // row changed: write it back
                uint32_t cur = 0;                                       // frame offset being accumulated
                if (Wp > 0.f) {                                         // carried in from the previous chunk
                    W = Wp;
                    S2 = s_S2[v];
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc[it][j] = s_part[v * RS + (it * 32 + lane) * VEC + j];
                } else {
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc[it][j] = 0.f;
                    if (nv) cur = __float_as_uint(s_con[s_start[v]].z) >> 8;
                }
                const float4 *con = s_con + s_start[v];
                for (uint32_t k = 0; k < nv; ++k) {
                    const float4 c = con[k];
                    const uint32_t tag = __float_as_uint(c.z);
                    const uint32_t e = tag & 255u, g = tag >> 8;
                    if (g != cur) {                                     // next frame: close the previous one
                        apply_frame<VEC, IT>(A.alpha, W, S2, row, acc, aprod);
                        dirty = true;
                        W = 0.f; S2 = 0.f;
                        cur = g;
                    }
                    W += c.x;
                    S2 += c.y;
                    if (ONEHOT) {
                        const int cls = (int)s_esrc[e];
#pragma unroll
                        for (int it = 0; it < IT; ++it)
#pragma unroll
                            for (int j = 0; j < VEC; ++j)
                                if ((it * 32 + lane) * VEC + j == cls) acc[it][j] += c.y;
                    } else {
                        const float *frow = STAGE ? s_feat + e * F : A.features + (size_t)s_esrc[e] * F;
#pragma unroll
                        for (int it = 0; it < IT; ++it) {
                            const int ch = (it * 32 + lane) * VEC;
                            if (ch < F) {
                                float fv[VEC];
                                if (STAGE) vec_load<VEC>(fv, frow + ch); else vec_load_nc<VEC>(fv, frow + ch);
#pragma unroll
                                for (int j = 0; j < VEC; ++j) acc[it][j] = fmaf(c.y, fv[j], acc[it][j]);
                            }
                        }
                    }
                }
                const bool keep = cont && cur == goff_last;            // the frame goes on in the next chunk
                if (keep) {
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) s_part[v * RS + (it * 32 + lane) * VEC + j] = acc[it][j];
                } else {
                    apply_frame<VEC, IT>(A.alpha, W, S2, row, acc, aprod);
                    dirty = true;
                }
                __syncwarp();
                if (lane == 0) { s_W[v] = keep ? W : 0.f; s_S2[v] = keep ? S2 : 0.f; }
                if (dirty) {
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int ch = (it * 32 + lane) * VEC;
                        if (ch < F) vec_store<VEC>(grow + ch, row[it]);
                    }
                    if (A.affine_a != nullptr && lane == 0) A.affine_a[vox] = A.affine_a[vox] * aprod;
                }
            }
            pending = cont ? f_last : NO_FRAME;
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2, group form (even feature sizes).  The work item is one (brick, frame) group of the sorted
// entry list.  Everything expensive about a group -- its per-voxel sums W, S2, B -- does not
// depend on the map, so groups are reduced fully in parallel by whichever CTA draws them; only the
// final row update  new = a*old + b  has to follow the frame order of the brick.  That order is
// enforced by a per-brick progress counter: a group applies its update when the counter has
// reached its first entry, then advances it past its last entry (release/acquire through L2).
// Groups are handed out frame-major, so a group's predecessor was always handed out earlier, to a
// CTA that never waits on a later ticket: the wait cannot deadlock (and is bounded anyway).
// Frame-major dispatch also means each frame's feature rows are consumed by all bricks at about
// the same time: they are read from HBM once and re-read from L2.
//
// Inside a group: 128 entries per chunk, two threads per entry (4 voxel parity classes each);
// LPV lanes own one voxel with 8 channels per lane (four packed f32x2 FMAs per contribution),
// 32/LPV voxels per warp pass; feature rows of the chunk are staged in shared memory by cp.async.
constexpr int GRP_CHUNK = 128;                 // entries per chunk
constexpr int GRP_CONTRIB = GRP_CHUNK * 8;
constexpr uint32_t SPIN_LIMIT = 1u << 22;      // polls (~ seconds) before a wait gives up and flags an error

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// group heads of the brick-sorted entry list: entry i starts a group if its brick or its frame differs
// from entry i-1.  One 32-bit mask + count per 32 entries; brick heads also reset the brick's
// progress counter to the brick's first entry.
__global__ void __launch_bounds__(256)
k_group_flags(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pids, uint32_t nmax, uint32_t np,
              uint32_t *__restrict__ masks, uint32_t *__restrict__ wcount, uint32_t *__restrict__ progress,
              uint32_t *__restrict__ counters)
{
    const uint32_t n = min(nmax, counters[MB_CNT_ENTRIES]);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool head = false;
    if (i < n) {
        const uint32_t k = keys[i];
        const bool bhead = i == 0 || keys[i - 1] != k;
        head = bhead || pids[i - 1] / np != pids[i] / np;
        if (bhead) progress[k] = i;
    }
    const uint32_t m = __ballot_sync(FULL, head);
    if ((threadIdx.x & 31) == 0) {
        masks[i >> 5] = m;
        wcount[i >> 5] = __popc(m);
    }
}

// ordered list of group starts (+ sentinel n at the end) and the per-frame group histogram
__global__ void __launch_bounds__(256)
k_group_emit(const uint32_t *__restrict__ masks, const uint32_t *__restrict__ woffs, const uint32_t *__restrict__ pids,
             uint32_t nmax, uint32_t np, uint32_t *__restrict__ gstart, uint32_t *__restrict__ frame_hist,
             uint32_t *__restrict__ counters)
{
    const uint32_t n = min(nmax, counters[MB_CNT_ENTRIES]);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if ((i & ~31u) >= n) return;
    const uint32_t m = masks[i >> 5];
    const uint32_t base = woffs[i >> 5];
    if (i < n && ((m >> lane) & 1u)) {
        gstart[base + __popc(m & ((1u << lane) - 1u))] = i;
        atomicAdd(&frame_hist[pids[i] / np], 1u);
    }
    if (n > 0 && i == n - 1) {
        const uint32_t total = base + __popc(m & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1u)));
        counters[MB_CNT_GROUPS] = total;
        gstart[total] = n;
    }
}

// frame-major dispatch order (any order inside a frame: its groups belong to different bricks)
// One descriptor per group, in dispatch order: {first entry, entries | frame << 20, brick id,
// brick coordinates packed 10 + 10 + 10 bits}.
__global__ void __launch_bounds__(256)
k_group_order(const uint32_t *__restrict__ gstart, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pids,
              uint32_t np, int T, MbBricks g, const uint32_t *__restrict__ frame_hist,
              uint32_t *__restrict__ frame_fill, uint4 *__restrict__ desc, const uint32_t *__restrict__ counters)
{
    extern __shared__ uint32_t s_base[];      // exclusive prefix of frame_hist
    for (int t = threadIdx.x; t < T; t += blockDim.x) s_base[t] = frame_hist[t];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < T; ++t) { const uint32_t c = s_base[t]; s_base[t] = run; run += c; }
    }
    __syncthreads();
    const uint32_t ngroups = counters[MB_CNT_GROUPS];
    for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += gridDim.x * blockDim.x) {
        const uint32_t beg = gstart[gi], end = gstart[gi + 1];
        const uint32_t f = pids[beg] / np, key = keys[beg];
        const uint32_t bz = key % g.N2, by = (key / g.N2) % g.N1, bx = key / (g.N2 * g.N1);
        desc[s_base[f] + atomicAdd(&frame_fill[f], 1u)] =
            make_uint4(beg, (end - beg) | (f << 20), key, bx | (by << 10) | (bz << 20));
    }
}

struct GroupArgs {
    const uint32_t *pids;
    const uint4 *desc;
    uint32_t *progress, *counters;
    const uint4 *rec;
    MbFeatIndex fi;
    uint32_t fhw;
    const float *features, *features_end;
    int F;
    float *map, *affine_a;
    MbBricks g;
    float alpha;
};

template <int LPV, bool STAGE>
__global__ void __launch_bounds__(RED_THREADS)
k_group_reduce(const GroupArgs A)
{
    constexpr int G = 32 / LPV;               // voxels per warp pass
    constexpr int RSP = LPV * 8;              // floats per voxel row in s_out
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_con = (float4 *)smem_raw;                         // [GRP_CONTRIB] sorted {w, w^2, feature row tag, -}
    float *s_out = (float *)(s_con + GRP_CONTRIB);              // [64][RSP] per-voxel B of the group
    uint32_t *s_cnt = (uint32_t *)(s_out + 64 * RSP);           // [RED_WARPS][64]
    uint32_t *s_start = s_cnt + RED_WARPS * 64;                 // [64]
    uint32_t *s_total = s_start + 64;                           // [64]
    uint32_t *s_vorder = s_total + 64;                          // [64]
    float *s_W = (float *)(s_vorder + 64);                      // [64]
    float *s_S2 = s_W + 64;                                     // [64]
    uint32_t *s_esrc = (uint32_t *)(s_S2 + 64);                 // [GRP_CHUNK]
    float *s_feat = (float *)(s_esrc + GRP_CHUNK);              // [GRP_CHUNK][RSF] staged feature rows
    __shared__ uint32_t s_next, s_turn;
    __shared__ uint4 s_desc[2];               // descriptor of the current and of the prefetched next group

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane % LPV, grp = lane / LPV;
    const int ent = tid & (GRP_CHUNK - 1), half = tid >> 7;     // entry of this thread, class half (0: 0-3, 1: 4-7)
    const uint32_t ltmask = (1u << lane) - 1u;
    const uint32_t ngroups = A.counters[MB_CNT_GROUPS];
    const uint32_t np = A.fi.np;
    const int F = A.F;
    const int RSF = (F + 3) & ~3;
    const int cb = sub * 8;
    MB_T0();

    if (tid == 0) {
        const uint32_t t = atomicAdd(&A.counters[MB_CNT_TICKET], 1u);
        s_desc[0] = t < ngroups ? A.desc[t] : make_uint4(0u, 0u, 0xffffffffu, 0u);
    }
    for (uint32_t it = 0;; ++it) {
        __syncthreads();
        const uint4 d = s_desc[it & 1];
        if (d.z == 0xffffffffu) break;
        const uint32_t gbeg = d.x, gend = d.x + (d.y & 0xfffffu), bkey = d.z, frame = d.y >> 20;
        const int org0 = (int)(d.w & 1023u) * 4, org1 = (int)((d.w >> 10) & 1023u) * 4, org2 = (int)(d.w >> 20) * 4;
        if (tid == 0) {
            // draw the next group now (its descriptor load overlaps this group's work) and take a first
            // look at the brick's progress counter
            const uint32_t t = atomicAdd(&A.counters[MB_CNT_TICKET], 1u);
            s_desc[(it + 1) & 1] = t < ngroups ? A.desc[t] : make_uint4(0u, 0u, 0xffffffffu, 0u);
            s_turn = ld_acquire(A.progress + bkey) == gbeg ? 1u : 0u;
        }
        if (tid < 64) { s_W[tid] = 0.f; s_S2[tid] = 0.f; }
        MB_TICK(0);

        for (uint32_t pos = gbeg; pos < gend; pos += GRP_CHUNK) {
            const int nact = min((uint32_t)GRP_CHUNK, gend - pos);
            const bool active = ent < nact;
            const uint32_t pid = active ? A.pids[pos + ent] : 0u;
            uint32_t src = 0;
            if (active) {
                const uint32_t p = pid - frame * np;
                if (A.fi.kx == 1 && A.fi.ky == 1) {
                    src = frame * A.fhw + p;
                } else {
                    const uint32_t y = p / A.fi.W, x = p - y * A.fi.W;
                    src = frame * A.fhw + (y / A.fi.ky) * A.fi.fw + x / A.fi.kx;
                }
            }
            if (half == 0) s_esrc[ent] = src;
            for (int i = tid; i < RED_WARPS * 64; i += RED_THREADS) s_cnt[i] = 0;
            if (tid == 0) s_next = 0;
            __syncthreads();
            MB_TICK(1);

            if (STAGE) {
                // feature rows -> shared memory in 16-byte pieces, from the 16-byte boundary at or before
                // the row (rows are 8-byte aligned: the data starts 0 or 2 floats into the slot)
                const int per_row = RSF >> 2;
                for (int i = tid; i < nact * per_row; i += RED_THREADS) {
                    const int e = i / per_row, c = i - e * per_row;
                    const size_t first = ((size_t)s_esrc[e] * F) & ~(size_t)3;
                    const float *gp = A.features + first + 4 * c;
                    float *d = s_feat + e * RSF + 4 * c;
                    if (gp + 4 <= A.features_end) cp_async<16>(d, gp);
                    else if (gp + 2 <= A.features_end) cp_async<8>(d, gp);         // tail of the very last row
                }
            }

            // ---- two threads per entry: one contribution per voxel parity class (4 classes each) -------
            uint32_t cv[4];
            float cw[4], cw2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { cv[k] = 0xffu; cw[k] = 0.f; cw2[k] = 0.f; }
            if (active) {
                const uint4 r = A.rec[pid];
                const Footprint f = footprint_of(r.x, A.g);
                const float q[3] = { __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w) };
                float wl[3], wu[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {               // projection.py:300-316
                    const bool low = q[a] < 0.5f;
                    wl[a] = low ? __fsub_rn(0.5f, q[a]) : __fsub_rn(1.5f, q[a]);
                    wu[a] = low ? __fadd_rn(q[a], 0.5f) : __fsub_rn(q[a], 0.5f);
                }
                const int l0lo = f.lo[0] - org0, l1lo = f.lo[1] - org1, l2lo = f.lo[2] - org2;
                const int l0hi = f.hi[0] - org0, l1hi = f.hi[1] - org1, l2hi = f.hi[2] - org2;
                const bool clamped = f.lo[0] == f.hi[0] || f.lo[1] == f.hi[1] || f.lo[2] == f.hi[2];
                if (!clamped) {
                    // the two neighbours of an axis have opposite parity: class (half, k) takes, per axis,
                    // the neighbour whose coordinate parity equals the class bit
                    const int p0 = f.lo[0] & 1, p1 = f.lo[1] & 1, p2 = f.lo[2] & 1;
                    const bool u0 = half != p0;
                    const int l0 = u0 ? l0hi : l0lo;
                    const float w0 = u0 ? wu[0] : wl[0];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool u1 = (((k >> 1) & 1) != p1), u2 = ((k & 1) != p2);
                        const int l1 = u1 ? l1hi : l1lo, l2 = u2 ? l2hi : l2lo;
                        if ((unsigned)l0 < 4u && (unsigned)l1 < 4u && (unsigned)l2 < 4u) {
                            cv[k] = (uint32_t)((l0 << 4) | (l1 << 2) | l2);
                            const float w = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn(w0, u1 ? wu[1] : wl[1]),
                                                                       u2 ? wu[2] : wl[2]));   // projection.py:319-323
                            cw[k] = w;
                            cw2[k] = w * w;
                        }
                    }
                } else {
                    // map border: clamping folds neighbours onto one voxel (projection.py:280-291); folded
                    // slots are summed into one contribution of that voxel's class
                    for (int s = 0; s < 8; ++s) {
                        const int l0 = (s & 4) ? l0hi : l0lo, l1 = (s & 2) ? l1hi : l1lo, l2 = (s & 1) ? l2hi : l2lo;
                        if ((unsigned)l0 < 4u && (unsigned)l1 < 4u && (unsigned)l2 < 4u && (l0 & 1) == half) {
                            const float w = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn((s & 4) ? wu[0] : wl[0], (s & 2) ? wu[1] : wl[1]),
                                                                       (s & 1) ? wu[2] : wl[2]));
                            const int cls = ((l1 & 1) << 1) | (l2 & 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k == cls) {
                                    cv[k] = (uint32_t)((l0 << 4) | (l1 << 2) | l2);
                                    cw[k] += w;
                                    cw2[k] += w * w;
                                }
                        }
                    }
                }
            }

            // ---- stable counting sort of the contributions by voxel (entry order inside a voxel) ---------
            uint32_t rk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ok = cv[k] != 0xffu;
                const uint32_t m = __match_any_sync(FULL, ok ? cv[k] : 64u + lane);
                rk[k] = __popc(m & ltmask);
                if (ok && rk[k] == 0) s_cnt[warp * 64 + cv[k]] = __popc(m);   // classes never share a voxel
            }
            __syncthreads();
            MB_TICK(2);
            if (tid < 64) {
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < RED_WARPS; ++w) {
                    const uint32_t c = s_cnt[w * 64 + tid];
                    s_cnt[w * 64 + tid] = run;
                    run += c;
                }
                s_total[tid] = run;
            }
            __syncthreads();
            if (warp == 0) {
                const uint32_t t0 = s_total[2 * lane], t1 = s_total[2 * lane + 1];
                uint32_t inc = t0 + t1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(FULL, inc, d);
                    if (lane >= d) inc += u;
                }
                const uint32_t ex = inc - (t0 + t1);
                s_start[2 * lane] = ex;
                s_start[2 * lane + 1] = ex + t0;
            } else if (tid >= 64 && tid < 128) {
                // voxels by decreasing contribution count: a warp pass takes G neighbours of this order, so
                // the G segments walked in lock-step have similar lengths
                const int v = tid - 64;
                const uint32_t mine = s_total[v];
                int rank = 0;
                for (int u = 0; u < 64; ++u) {
                    const uint32_t o = s_total[u];
                    rank += (o > mine || (o == mine && u < v)) ? 1 : 0;
                }
                s_vorder[rank] = (uint32_t)v;
            }
            __syncthreads();
            {
                // tag: where the entry's feature row sits (staged: float offset in s_feat; else entry slot)
                uint32_t tag = (uint32_t)ent;
                if (STAGE) tag = (uint32_t)(ent * RSF) + (uint32_t)(((size_t)src * F) & 3);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cv[k] != 0xffu)
                        s_con[s_start[cv[k]] + s_cnt[warp * 64 + cv[k]] + rk[k]] =
                            make_float4(cw[k], cw2[k], __uint_as_float(tag), 0.f);
            }
            MB_TICK(3);
            if (STAGE) cp_async_wait_all();
            __syncthreads();
            MB_TICK(4);

            // ---- LPV lanes per voxel, G voxels per warp pass -----------------------------------------------
            for (;;) {
                int t = 0;
                if (lane == 0) t = (int)atomicAdd(&s_next, (uint32_t)G);
                t = __shfl_sync(FULL, t, 0);
                if (t >= 64) break;
                const int v = (int)s_vorder[t + grp];
                const uint32_t nv = s_total[v];
                const uint32_t maxnv = __reduce_max_sync(FULL, nv);
                if (maxnv == 0) break;                                   // sorted by count: nothing left
                float2 acc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
                float W = 0.f, S2 = 0.f;
                const float Wp = s_W[v], S2p = s_S2[v];                  // sums of earlier chunks of a long group
                const float4 *con = s_con + s_start[v];
                for (uint32_t k = 0; k < maxnv; ++k) {
                    if (k < nv) {
                        const float4 c = con[k];
                        const uint32_t tag = __float_as_uint(c.z);
                        W += c.x;
                        S2 += c.y;
                        const float2 w2 = make_float2(c.y, c.y);
                        if (STAGE) {
                            const float *frow = s_feat + tag + cb;
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[j] = ffma2(w2, *(const float2 *)(frow + 2 * j), acc[j]);
                        } else {
                            const float *frow = A.features + (size_t)s_esrc[tag] * F + cb;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (cb + 2 * j < F) acc[j] = ffma2(w2, __ldg((const float2 *)(frow + 2 * j)), acc[j]);
                        }
                    }
                }
                __syncwarp();
                if (nv) {
                    float *o = s_out + v * RSP + cb;
                    if (Wp > 0.f) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 p = *(const float2 *)(o + 2 * j);
                            acc[j] = make_float2(p.x + acc[j].x, p.y + acc[j].y);
                        }
                        W = Wp + W;
                        S2 = S2p + S2;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) *(float2 *)(o + 2 * j) = acc[j];
                    if (sub == 0) { s_W[v] = W; s_S2[v] = S2; }
                }
            }
            __syncthreads();
            MB_TICK(5);
        }

        // ---- the brick's turn: all earlier frames of this brick have been applied ---------------------
        if (tid == 0 && s_turn == 0u) {
            uint32_t spins = 0;
            while (ld_acquire(A.progress + bkey) != gbeg) {
                __nanosleep(64);
                if (++spins > SPIN_LIMIT) { atomicOr(&A.counters[MB_CNT_ERROR], 1u); break; }
            }
        }
        __syncthreads();
        MB_TICK(6);
        for (int t = warp * G; t < 64; t += RED_WARPS * G) {
            const int v = t + grp;
            const float W = s_W[v];
            if (W > 0.f) {
                const float r = __frcp_rn(W);
                const float sc = A.alpha * r;
                const float a = 1.0f - sc * s_S2[v];
                const size_t vox = ((size_t)(org0 + (v >> 4)) * A.g.S1 + (org1 + ((v >> 2) & 3))) * A.g.S2 +
                                   (org2 + (v & 3));
                float *grow = A.map + vox * (size_t)F + cb;
                const float *o = s_out + v * RSP + cb;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (cb + 2 * j < F) {
                        const float2 old = __ldcg((const float2 *)(grow + 2 * j));       // L2: written by other SMs
                        const float2 b = *(const float2 *)(o + 2 * j);
                        *(float2 *)(grow + 2 * j) = ffma2(make_float2(a, a), old, make_float2(sc * b.x, sc * b.y));
                    }
                if (A.affine_a != nullptr && sub == 0) A.affine_a[vox] = __ldcg(A.affine_a + vox) * a;
            }
        }
        __syncthreads();                                  // every row store of the CTA is ordered before ...
        if (tid == 0) {
            __threadfence();                              // ... this fence, which publishes them at GPU scope
            st_release(A.progress + bkey, gend);
        }
        MB_TICK(7);
    }
    MB_TFLUSH();
}

size_t group_smem_bytes(int LPV, int F, bool stage)
{
    const int RSF = (F + 3) & ~3;
    return (size_t)GRP_CONTRIB * 16 + (size_t)64 * LPV * 8 * 4 + (size_t)(RED_WARPS * 64 + 64 * 5 + GRP_CHUNK) * 4 +
           (stage ? (size_t)(GRP_CHUNK * RSF + LPV * 8 + 4) * 4 : 0);
}

template <int LPV>
int launch_group_reduce(cudaStream_t stream, const GroupArgs &A, bool stage)
{
    const size_t smem = group_smem_bytes(LPV, A.F, stage);
    auto kern = stage ? k_group_reduce<LPV, true> : k_group_reduce<LPV, false>;
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RED_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    kern<<<MB_NUM_SMS * per_sm, RED_THREADS, smem, stream>>>(A);
    MB_LAUNCHED();
    return MB_OK;
}

bool group_form_applies(const float *features, const float *map, int F)
{
    return features != nullptr && F % 2 == 0 && F <= 256 && (uintptr_t)features % 16 == 0 && (uintptr_t)map % 8 == 0;
}

int dispatch_group_reduce(cudaStream_t stream, const GroupArgs &A)
{
    const int lanes = (A.F + 7) / 8;
    const bool stage = (size_t)GRP_CHUNK * ((A.F + 3) & ~3) * 4 <= 49152;
    if (lanes <= 1) return launch_group_reduce<1>(stream, A, stage);
    if (lanes <= 2) return launch_group_reduce<2>(stream, A, stage);
    if (lanes <= 4) return launch_group_reduce<4>(stream, A, stage);
    if (lanes <= 8) return launch_group_reduce<8>(stream, A, stage);
    if (lanes <= 16) return launch_group_reduce<16>(stream, A, stage);
    return launch_group_reduce<32>(stream, A, stage);
}

size_t reduce_smem_bytes(int VEC, int IT, int F, bool stage)
{
    return (size_t)RED_CONTRIB * 16 + (size_t)64 * 32 * VEC * IT * 4 + (size_t)(RED_WARPS * 64 + 64 * 4 + 2 * RED_THREADS) * 4 +
           (stage ? (size_t)RED_THREADS * F * 4 : 0);
}

template <int VEC, int IT, bool STAGE>
int launch_brick_reduce(cudaStream_t stream, const ReduceArgs &A)
{
    const size_t smem = reduce_smem_bytes(VEC, IT, A.F, STAGE);
    const bool onehot = A.class_ids != nullptr;
    auto kern = onehot ? k_brick_reduce<VEC, IT, true, false> : k_brick_reduce<VEC, IT, false, STAGE>;
    const size_t bytes = onehot ? reduce_smem_bytes(VEC, IT, A.F, false) : smem;
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int per_sm = 1;
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RED_THREADS, bytes));
    if (per_sm < 1) per_sm = 1;
    kern<<<MB_NUM_SMS * per_sm, RED_THREADS, bytes, stream>>>(A);
    MB_LAUNCHED();
    return MB_OK;
}

int dispatch_brick_reduce(cudaStream_t stream, const ReduceArgs &A)
{
    const int F = A.F;
    const bool a16 = A.features == nullptr || ((uintptr_t)A.features % 16 == 0);
    const bool m16 = (uintptr_t)A.map % 16 == 0;
    int vec = 1;
    if (F % 4 == 0 && a16 && m16) vec = 4;
    else if (F % 2 == 0 && (uintptr_t)A.features % 8 == 0 && (uintptr_t)A.map % 8 == 0) vec = 2;
    const int need = (F + 32 * vec - 1) / (32 * vec);
    // feature rows are staged through shared memory while a chunk's rows fit in 64 KB
    const bool stage = A.features != nullptr && (size_t)RED_THREADS * F * 4 <= 65536;
#define MB_GO(V, I) return launch_brick_reduce<V, I, false>(stream, A)
#define MB_GO_STAGED(V, I)                                                                            \
    do {                                                                                              \
        if (stage) return launch_brick_reduce<V, I, true>(stream, A);                                 \
        return launch_brick_reduce<V, I, false>(stream, A);                                           \
    } while (0)
    if (vec == 4) { if (need <= 1) MB_GO_STAGED(4, 1); if (need <= 2) MB_GO(4, 2); if (need <= 4) MB_GO(4, 4); }
    if (vec == 2) { if (need <= 1) MB_GO_STAGED(2, 1); if (need <= 2) MB_GO(2, 2); if (need <= 4) MB_GO(2, 4); }
    if (vec == 1) { if (need <= 1) MB_GO_STAGED(1, 1); if (need <= 2) MB_GO_STAGED(1, 2); if (need <= 4) MB_GO(1, 4); if (need <= 8) MB_GO(1, 8); }
#undef MB_GO
#undef MB_GO_STAGED
    mb_set_error("feature_size %d not supported by the batched path (<= 512 if a multiple of 4, <= 256 otherwise)", F);
    return MB_ERR_ARG;
}

struct BatchBuffers {
    uint4 *rec;
    uint32_t *cnt, *offs, *keys_a, *keys_b, *pids_a, *pids_b, *order, *counters;
    uint32_t *masks, *wcount, *woffs, *gstart, *progress, *frame_hist;   // group form
    uint4 *desc;
    uint2 *bricks;
    char *scan_ws, *sort_ws;
    size_t scan_bytes, sort_bytes;
};

size_t carve_batch(BatchBuffers &b, void *ws, size_t bytes, uint32_t ntotal, size_t nbricks_total, int T)
{
    MbArena a(ws, bytes);
    const size_t nent = (size_t)ntotal * 8;           // upper bound: 8 bricks per pixel
    size_t nbrick_cap = nbricks_total;
    if (nbrick_cap > nent) nbrick_cap = nent;
    b.counters = a.take<uint32_t>(MB_NUM_COUNTERS);   // first: mb_layer_update_status finds them at offset 0
    b.rec = a.take<uint4>(ntotal);
    b.cnt = a.take<uint32_t>(ntotal);
    b.offs = a.take<uint32_t>(ntotal);
    b.keys_a = a.take<uint32_t>(nent);
    b.keys_b = a.take<uint32_t>(nent);
    b.pids_a = a.take<uint32_t>(nent);
    b.pids_b = a.take<uint32_t>(nent);
    b.bricks = a.take<uint2>(nbrick_cap);     // one per touched brick
    b.order = a.take<uint32_t>(nent);         // brick form: one per brick; group form: one per group
    b.masks = a.take<uint32_t>(nent / 32 + 1);
    b.wcount = a.take<uint32_t>(nent / 32 + 1);
    b.woffs = a.take<uint32_t>(nent / 32 + 1);
    b.gstart = a.take<uint32_t>(nent + 1);
    b.desc = a.take<uint4>(nent);
    b.progress = a.take<uint32_t>(nbricks_total);
    b.frame_hist = a.take<uint32_t>(2 * (size_t)T);
    b.scan_bytes = mb_scan_workspace_bytes(ntotal);          // >= the group-word scan (nent / 32 words)
    b.scan_ws = a.take<char>(b.scan_bytes);
    b.sort_bytes = mb_sort_workspace_bytes((uint32_t)nent);
    b.sort_ws = a.take<char>(b.sort_bytes);
    return a.used + 256;
}

}  // namespace

static size_t total_bricks(int nx, int ny, int nz)
{
    const MbBricks g = make_bricks(ny - 1, nx - 1, nz - 1);
    return (size_t)g.N0 * g.N1 * g.N2;
}

// frames per internal chunk for a given workspace; 0 if even one frame does not fit
int mbk_batch_frames_that_fit(uint32_t npix, int nx, int ny, int nz, size_t workspace_bytes, int T)
{
    const size_t cap = total_bricks(nx, ny, nz);
    BatchBuffers b;
    if ((uint64_t)T * npix * 8 < 0xffffffffull && carve_batch(b, nullptr, 0, (uint32_t)T * npix, cap, T) <= workspace_bytes)
        return T;
    int best = 0;
    for (int t = 1; t <= T; t = t < 8 ? t + 1 : t * 2) {
        if ((uint64_t)t * npix * 8 >= 0xffffffffull) break;
        if (carve_batch(b, nullptr, 0, (uint32_t)t * npix, cap, t) <= workspace_bytes) best = t; else break;
    }
    return best;
}

size_t mbk_batch_workspace_bytes(uint32_t npix, int nx, int ny, int nz, int T)
{
    BatchBuffers b;
    return carve_batch(b, nullptr, 0, (uint32_t)T * npix, total_bricks(nx, ny, nz), T);
}

// One chunk of T frames (T * npix * 8 < 2^32).
int mbk_batch_update(cudaStream_t stream, const float *rays, const float *depth, const float *features,
                     const int64_t *class_ids, const float *pose, int T, int H, int W, int fh, int fw, int F,
                     const float *bins_x, int nx, const float *bins_y, int ny, const float *bins_z, int nz,
                     float *map, float *affine_a, float alpha, float min_d, float max_d, void *workspace,
                     size_t workspace_bytes)
{
    const uint32_t npix = (uint32_t)H * (uint32_t)W;
    const uint32_t ntotal = (uint32_t)T * npix;
    const MbBricks g = make_bricks(ny - 1, nx - 1, nz - 1);
    MB_REQUIRE(g.S0 <= 2046 && g.S1 <= 2046 && g.S2 <= 1022, "map too large for the packed voxel record");
    MB_REQUIRE(class_ids != nullptr || (uint64_t)T * fh * fw < 0xffffffffull, "too many feature rows per chunk");
    MB_REQUIRE(T <= 4096, "too many frames per chunk");
    MB_REQUIRE(npix < (1u << 20), "frames of 2^20 pixels or more are not supported by the batched path");
    MB_REQUIRE(g.N0 <= 1024 && g.N1 <= 1024 && g.N2 <= 1024, "map too large for the packed brick coordinates");
    BatchBuffers b;
    MB_REQUIRE(carve_batch(b, workspace, workspace_bytes, ntotal, (size_t)g.N0 * g.N1 * g.N2, T) <= workspace_bytes,
               "batch workspace too small");
    const uint32_t nent = ntotal * 8u;

    dim3 grid((npix + 255) / 256, (unsigned)T);
    k_voxelise_batch<<<grid, 256, 0, stream>>>(rays, depth, pose, npix, bins_x, nx, bins_y, ny, bins_z, nz, g,
                                               min_d, max_d, b.rec, b.cnt, b.counters);
    MB_LAUNCHED();
    int rc = mb_exclusive_scan_u32(stream, b.cnt, b.offs, ntotal, b.scan_ws, b.scan_bytes);
    if (rc) return rc;
    k_emit_entries<<<(ntotal + 255) / 256, 256, 0, stream>>>(b.rec, b.cnt, b.offs, ntotal, g, b.keys_a, b.pids_a,
                                                             b.counters);
    MB_LAUNCHED();
    int bits = 1;
    while ((1u << bits) < (uint32_t)(g.N0 * g.N1 * g.N2)) ++bits;
    uint32_t *keys, *pids;
    rc = mb_sort_pairs(stream, b.keys_a, b.pids_a, b.keys_b, b.pids_b, nent, b.counters + MB_CNT_ENTRIES, bits, false,
                       b.sort_ws, b.sort_bytes, &keys, &pids);
    if (rc) return rc;
    if (group_form_applies(features, map, F)) {
        const uint32_t nwords = nent / 32 + 1;
        MB_CHECK_CUDA(cudaMemsetAsync(b.frame_hist, 0, 2 * (size_t)T * sizeof(uint32_t), stream));
        k_group_flags<<<(nent + 255) / 256, 256, 0, stream>>>(keys, pids, nent, npix, b.masks, b.wcount, b.progress,
                                                              b.counters);
        MB_LAUNCHED();
        rc = mb_exclusive_scan_u32(stream, b.wcount, b.woffs, nwords, b.scan_ws, b.scan_bytes);
        if (rc) return rc;
        k_group_emit<<<(nent + 255) / 256, 256, 0, stream>>>(b.masks, b.woffs, pids, nent, npix, b.gstart, b.frame_hist,
                                                             b.counters);
        MB_LAUNCHED();
        k_group_order<<<MB_NUM_SMS * 2, 256, (size_t)T * sizeof(uint32_t), stream>>>(
            b.gstart, keys, pids, npix, T, g, b.frame_hist, b.frame_hist + T, b.desc, b.counters);
        MB_LAUNCHED();
        GroupArgs G;
        G.pids = pids; G.desc = b.desc; G.progress = b.progress;
        G.counters = b.counters; G.rec = b.rec;
        G.fi = MbFeatIndex{ npix, (uint32_t)W, (uint32_t)(H / fh), (uint32_t)(W / fw), (uint32_t)fw };
        G.fhw = (uint32_t)fh * (uint32_t)fw;
        G.features = features; G.features_end = features + (size_t)T * fh * fw * F;
        G.F = F; G.map = map; G.affine_a = affine_a; G.g = g; G.alpha = alpha;
        return dispatch_group_reduce(stream, G);
    }
    k_brick_heads<<<(nent + 255) / 256, 256, 0, stream>>>(keys, nent, b.bricks, b.counters);
    MB_LAUNCHED();
    {
        size_t cap = (size_t)g.N0 * g.N1 * g.N2;
        if (cap > nent) cap = nent;
        k_brick_order<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(b.bricks, b.order, b.counters);
        MB_LAUNCHED();
    }

    ReduceArgs A;
    A.keys = keys; A.pids = pids; A.bricks = b.bricks; A.order = b.order; A.counters = b.counters; A.nmax = nent; A.rec = b.rec;
    A.fi = MbFeatIndex{ npix, (uint32_t)W, (uint32_t)(H / fh), (uint32_t)(W / fw), (uint32_t)fw };
    A.fhw = (uint32_t)fh * (uint32_t)fw;
    A.features = features; A.features_end = features ? features + (size_t)T * fh * fw * F : nullptr;
    A.class_ids = class_ids; A.F = F; A.map = map; A.affine_a = affine_a; A.g = g;
    A.alpha = alpha;
    return dispatch_brick_reduce(stream, A);
}

#ifdef MB_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int mb_debug_phase_cycles(unsigned long long *out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
    return 0;
}
#endif
