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

// brick segment starts of the sorted entry list (order of the list is irrelevant: bricks are independent)
__global__ void __launch_bounds__(256)
k_brick_heads(const uint32_t *__restrict__ keys, uint32_t nmax, uint32_t *__restrict__ starts,
              uint32_t *__restrict__ counters)
{
    const uint32_t n = min(nmax, counters[MB_CNT_ENTRIES]);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool head = i < n && (i == 0 || keys[i - 1] != keys[i]);
    const uint32_t m = __ballot_sync(FULL, head);
    if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counters[MB_CNT_BRICKS], (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (head) starts[base + __popc(m & ((1u << lane) - 1u))] = i;
    }
}

// ---------------------------------------------------------------------------------------------
// K2
struct ReduceArgs {
    const uint32_t *keys, *pids, *starts;
    uint32_t *counters;
    uint32_t nmax;
    const uint4 *rec;
    MbFeatIndex fi;             // np = pixels per frame
    uint32_t fhw;               // feature rows per frame
    const float *features;      // [T][fhw][F] or null
    const int64_t *class_ids;   // [T][np] or null
    int F;
    float *map;                 // [S0][S1][S2][F], updated in place
    float *affine_a;            // optional [S0*S1*S2]: multiplied by the frame's a (affine output mode)
    MbBricks g;
    float alpha;
};

template <int VEC> struct VecT;
template <> struct VecT<1> { typedef float type; };
template <> struct VecT<2> { typedef float2 type; };
template <> struct VecT<4> { typedef float4 type; };

template <int VEC>
__device__ __forceinline__ void vec_load(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = __ldg(p);
    if (VEC == 2) { const float2 v = __ldg((const float2 *)p); dst[0] = v.x; dst[1] = v.y; }
    if (VEC == 4) { const float4 v = __ldg((const float4 *)p); dst[0] = v.x; dst[1] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w; }
}

template <int VEC>
__device__ __forceinline__ void vec_load_rw(float (&dst)[VEC], const float *p)
{
    if (VEC == 1) dst[0] = *p;
    if (VEC == 2) { const float2 v = *(const float2 *)p; dst[0] = v.x; dst[1] = v.y; }
    if (VEC == 4) { const float4 v = *(const float4 *)p; dst[0] = v.x; dst[1] = v.y; dst[VEC > 2 ? 2 : 0] = v.z; dst[VEC > 2 ? 3 : 0] = v.w; }
}

template <int VEC>
__device__ __forceinline__ void vec_store(float *p, const float (&src)[VEC])
{
    if (VEC == 1) *p = src[0];
    if (VEC == 2) *(float2 *)p = make_float2(src[0], src[1]);
    if (VEC == 4) *(float4 *)p = make_float4(src[0], src[1], src[VEC > 2 ? 2 : 0], src[VEC > 2 ? 3 : 0]);
}

// Applies one frame's affine update to the map row of voxel `vox`.
template <int VEC, int IT>
__device__ __forceinline__ void apply_row(const ReduceArgs &A, size_t vox, int lane, float W, float S2,
                                          const float (&acc)[IT][VEC])
{
    const float a = 1.0f - A.alpha * S2 / W;
    const float sc = A.alpha / W;
    float *row = A.map + vox * (size_t)A.F;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int c = (it * 32 + lane) * VEC;
        if (c < A.F) {
            float old[VEC], out[VEC];
            vec_load_rw<VEC>(old, row + c);
#pragma unroll
            for (int j = 0; j < VEC; ++j) out[j] = fmaf(a, old[j], sc * acc[it][j]);
            vec_store<VEC>(row + c, out);
        }
    }
    if (A.affine_a != nullptr && lane == 0) A.affine_a[vox] = A.affine_a[vox] * a;
}

template <int VEC, int IT, bool ONEHOT>
__global__ void __launch_bounds__(RED_THREADS)
k_brick_reduce(const ReduceArgs A)
{
    constexpr int RS = 32 * VEC * IT;                       // floats per partial row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_part = (float *)smem_raw;                      // [64][RS] partial sums of a multi-chunk frame
    float2 *s_con = (float2 *)(s_part + 64 * RS);           // [RED_CONTRIB] sorted (w, source row)
    uint32_t *s_cnt = (uint32_t *)(s_con + RED_CONTRIB);    // [RED_WARPS][64]
    uint32_t *s_start = s_cnt + RED_WARPS * 64;             // [64]
    uint32_t *s_total = s_start + 64;                       // [64]
    float *s_W = (float *)(s_total + 64);                   // [64]
    float *s_S2 = s_W + 64;                                 // [64]
    __shared__ uint32_t s_ticket, s_frame0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = min(A.nmax, A.counters[MB_CNT_ENTRIES]);
    const uint32_t nbricks = A.counters[MB_CNT_BRICKS];
    const uint32_t np = A.fi.np;

    for (int i = tid; i < 64 * RS; i += RED_THREADS) s_part[i] = 0.f;
    if (tid < 64) { s_W[tid] = 0.f; s_S2[tid] = 0.f; }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(&A.counters[MB_CNT_TICKET], 1u);
        __syncthreads();
        const uint32_t ticket = s_ticket;
        if (ticket >= nbricks) break;
        uint32_t pos = A.starts[ticket];
        const uint32_t bkey = A.keys[pos];
        const int bz = bkey % A.g.N2, by = (bkey / A.g.N2) % A.g.N1, bx = bkey / (A.g.N2 * A.g.N1);
        uint32_t pending = NO_FRAME;          // frame whose partial sums sit in s_part / s_W / s_S2

        for (;;) {
            // ---- next chunk: the leading entries of [pos, pos + 256) that share one frame ----------
            const uint32_t idx = pos + tid;
            const bool mine = idx < n && A.keys[idx] == bkey;
            const uint32_t pid = mine ? A.pids[idx] : 0u;
            const uint32_t frame = pid / np;
            if (tid == 0) s_frame0 = mine ? frame : NO_FRAME;
            __syncthreads();
            const uint32_t f0 = s_frame0;
            const bool active = mine && frame == f0;
            const int nact = f0 == NO_FRAME ? 0 : __syncthreads_count(active);
            const bool last = nact < RED_THREADS;           // the frame's entries end inside this chunk

            // ---- a finished multi-chunk frame whose sums are still pending ------------------------
            if (pending != NO_FRAME && pending != f0) {
                for (int v = warp; v < 64; v += RED_WARPS) {
                    const float W = s_W[v];
                    if (W > 0.f) {
                        float acc[IT][VEC];
#pragma unroll
                        for (int it = 0; it < IT; ++it)
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                acc[it][j] = s_part[v * RS + (it * 32 + lane) * VEC + j];
                                s_part[v * RS + (it * 32 + lane) * VEC + j] = 0.f;
                            }
                        const size_t vox = ((size_t)(bx * 4 + (v >> 4)) * A.g.S1 + (by * 4 + ((v >> 2) & 3))) * A.g.S2 +
                                           (bz * 4 + (v & 3));
                        apply_row<VEC, IT>(A, vox, lane, W, s_S2[v], acc);
                        __syncwarp();
                        if (lane == 0) { s_W[v] = 0.f; s_S2[v] = 0.f; }
                    }
                }
                pending = NO_FRAME;
            }
            if (f0 == NO_FRAME) break;                       // brick finished

            // ---- lanes = entries: the 8 contributions of this thread's entry -----------------------
            uint32_t cv[8];      // local voxel (0..63) or 0xff
            float cw[8];
            uint32_t src = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) { cv[s] = 0xffu; cw[s] = 0.f; }
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
                const int org[3] = { bx * 4, by * 4, bz * 4 };
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int c0 = (s & 4) ? f.hi[0] : f.lo[0], c1 = (s & 2) ? f.hi[1] : f.lo[1],
                              c2 = (s & 1) ? f.hi[2] : f.lo[2];
                    const int l0 = c0 - org[0], l1 = c1 - org[1], l2 = c2 - org[2];
                    if ((unsigned)l0 < 4u && (unsigned)l1 < 4u && (unsigned)l2 < 4u) {
                        cv[s] = (uint32_t)((l0 << 4) | (l1 << 2) | l2);
                        const float w0 = (s & 4) ? wu[0] : wl[0], w1 = (s & 2) ? wu[1] : wl[1],
                                    w2 = (s & 1) ? wu[2] : wl[2];
                        cw[s] = __fadd_rn(1e-9f, __fmul_rn(__fmul_rn(w0, w1), w2));   // projection.py:319-323
                    }
                }
                const uint32_t p = pid - f0 * np;
                if (ONEHOT) {
                    src = (uint32_t)A.class_ids[pid];
                } else if (A.fi.kx == 1 && A.fi.ky == 1) {
                    src = f0 * A.fhw + p;
                } else {
                    const uint32_t y = p / A.fi.W, x = p - y * A.fi.W;
                    src = f0 * A.fhw + (y / A.fi.ky) * A.fi.fw + x / A.fi.kx;
                }
            }

            // ---- stable counting sort of the chunk's contributions by voxel ------------------------
            for (int i = tid; i < RED_WARPS * 64; i += RED_THREADS) s_cnt[i] = 0;
            __syncthreads();
            uint32_t rk[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const bool ok = cv[s] != 0xffu;
                const uint32_t m = __match_any_sync(FULL, ok ? cv[s] : 64u + lane);
                const uint32_t rank = __popc(m & ((1u << lane) - 1u));
                const uint32_t prev = ok ? s_cnt[warp * 64 + cv[s]] : 0u;
                __syncwarp();
                if (ok && rank == 0) s_cnt[warp * 64 + cv[s]] = prev + __popc(m);
                __syncwarp();
                rk[s] = prev + rank;
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
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (cv[s] != 0xffu)
                    s_con[s_start[cv[s]] + s_cnt[warp * 64 + cv[s]] + rk[s]] =
                        make_float2(cw[s], __uint_as_float(src));
            __syncthreads();

            // ---- lanes = channels: one warp per voxel --------------------------------------------------
            const bool merge = pending == f0;
            for (int v = warp; v < 64; v += RED_WARPS) {
                const uint32_t nv = s_total[v];
                const float Wp = merge ? s_W[v] : 0.f;
                if (nv == 0 && !(last && Wp > 0.f)) continue;
                float acc[IT][VEC];
#pragma unroll
                for (int it = 0; it < IT; ++it)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[it][j] = 0.f;
                float W = 0.f, S2 = 0.f;
                const float2 *con = s_con + s_start[v];
                for (uint32_t k = 0; k < nv; ++k) {
                    const float2 c = con[k];
                    const float w = c.x, w2 = w * w;
                    W += w;
                    S2 += w2;
                    if (ONEHOT) {
                        const int cls = (int)__float_as_uint(c.y);
#pragma unroll
                        for (int it = 0; it < IT; ++it)
#pragma unroll
                            for (int j = 0; j < VEC; ++j)
                                if ((it * 32 + lane) * VEC + j == cls) acc[it][j] += w2;
                    } else {
                        const float *frow = A.features + (size_t)__float_as_uint(c.y) * A.F;
#pragma unroll
                        for (int it = 0; it < IT; ++it) {
                            const int ch = (it * 32 + lane) * VEC;
                            if (ch < A.F) {
                                float f[VEC];
                                vec_load<VEC>(f, frow + ch);
#pragma unroll
                                for (int j = 0; j < VEC; ++j) acc[it][j] = fmaf(w2, f[j], acc[it][j]);
                            }
                        }
                    }
                }
                if (merge && Wp > 0.f) {                      // earlier chunks of the same frame
                    W = Wp + W;
                    S2 = s_S2[v] + S2;
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            acc[it][j] = s_part[v * RS + (it * 32 + lane) * VEC + j] + acc[it][j];
                            if (last) s_part[v * RS + (it * 32 + lane) * VEC + j] = 0.f;
                        }
                }
                if (last) {
                    const size_t vox = ((size_t)(bx * 4 + (v >> 4)) * A.g.S1 + (by * 4 + ((v >> 2) & 3))) * A.g.S2 +
                                       (bz * 4 + (v & 3));
                    apply_row<VEC, IT>(A, vox, lane, W, S2, acc);
                    __syncwarp();
                    if (merge && lane == 0) { s_W[v] = 0.f; s_S2[v] = 0.f; }
                } else {
#pragma unroll
                    for (int it = 0; it < IT; ++it)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) s_part[v * RS + (it * 32 + lane) * VEC + j] = acc[it][j];
                    __syncwarp();
                    if (lane == 0) { s_W[v] = W; s_S2[v] = S2; }
                }
            }
            pending = last ? NO_FRAME : f0;
            pos += (uint32_t)nact;
            __syncthreads();
        }
    }
}

size_t reduce_smem_bytes(int VEC, int IT)
{
    return (size_t)64 * 32 * VEC * IT * 4 + (size_t)RED_CONTRIB * 8 + (size_t)(RED_WARPS * 64 + 64 + 64) * 4 + 2 * 64 * 4;
}

template <int VEC, int IT>
int launch_brick_reduce(cudaStream_t stream, const ReduceArgs &A)
{
    const size_t smem = reduce_smem_bytes(VEC, IT);
    const bool onehot = A.class_ids != nullptr;
    auto kern = onehot ? k_brick_reduce<VEC, IT, true> : k_brick_reduce<VEC, IT, false>;
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RED_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    kern<<<MB_NUM_SMS * per_sm, RED_THREADS, smem, stream>>>(A);
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
#define MB_GO(V, I) return launch_brick_reduce<V, I>(stream, A)
    if (vec == 4) { if (need <= 1) MB_GO(4, 1); if (need <= 2) MB_GO(4, 2); if (need <= 4) MB_GO(4, 4); }
    if (vec == 2) { if (need <= 1) MB_GO(2, 1); if (need <= 2) MB_GO(2, 2); if (need <= 4) MB_GO(2, 4); }
    if (vec == 1) { if (need <= 1) MB_GO(1, 1); if (need <= 2) MB_GO(1, 2); if (need <= 4) MB_GO(1, 4); if (need <= 8) MB_GO(1, 8); }
#undef MB_GO
    mb_set_error("feature_size %d not supported by the batched path (<= 512 if a multiple of 4, <= 256 otherwise)", F);
    return MB_ERR_ARG;
}

struct BatchBuffers {
    uint4 *rec;
    uint32_t *cnt, *offs, *keys_a, *keys_b, *pids_a, *pids_b, *starts, *counters;
    char *scan_ws, *sort_ws;
    size_t scan_bytes, sort_bytes;
};

size_t carve_batch(BatchBuffers &b, void *ws, size_t bytes, uint32_t ntotal)
{
    MbArena a(ws, bytes);
    const size_t nent = (size_t)ntotal * 8;           // upper bound: 8 bricks per pixel
    b.rec = a.take<uint4>(ntotal);
    b.cnt = a.take<uint32_t>(ntotal);
    b.offs = a.take<uint32_t>(ntotal);
    b.keys_a = a.take<uint32_t>(nent);
    b.keys_b = a.take<uint32_t>(nent);
    b.pids_a = a.take<uint32_t>(nent);
    b.pids_b = a.take<uint32_t>(nent);
    b.starts = a.take<uint32_t>(nent);        // one per touched brick (<= entries)
    b.counters = a.take<uint32_t>(64);
    b.scan_bytes = mb_scan_workspace_bytes(ntotal);
    b.scan_ws = a.take<char>(b.scan_bytes);
    b.sort_bytes = mb_sort_workspace_bytes((uint32_t)nent);
    b.sort_ws = a.take<char>(b.sort_bytes);
    return a.used + 256;
}

}  // namespace

// frames per internal chunk for a given workspace; 0 if even one frame does not fit
int mbk_batch_frames_that_fit(uint32_t npix, size_t workspace_bytes, int T)
{
    {
        BatchBuffers all;
        if ((uint64_t)T * npix * 8 < 0xffffffffull && carve_batch(all, nullptr, 0, (uint32_t)T * npix) <= workspace_bytes)
            return T;
    }
    int best = 0;
    for (int t = 1; t <= T; t = t < 8 ? t + 1 : t * 2) {
        if ((uint64_t)t * npix * 8 >= 0xffffffffull) break;
        BatchBuffers b;
        if (carve_batch(b, nullptr, 0, (uint32_t)t * npix) <= workspace_bytes) best = t; else break;
    }
    return best;
}

size_t mbk_batch_workspace_bytes(uint32_t npix, int T)
{
    BatchBuffers b;
    return carve_batch(b, nullptr, 0, (uint32_t)T * npix);
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
    BatchBuffers b;
    MB_REQUIRE(carve_batch(b, workspace, workspace_bytes, ntotal) <= workspace_bytes, "batch workspace too small");
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
    k_brick_heads<<<(nent + 255) / 256, 256, 0, stream>>>(keys, nent, b.starts, b.counters);
    MB_LAUNCHED();

    ReduceArgs A;
    A.keys = keys; A.pids = pids; A.starts = b.starts; A.counters = b.counters; A.nmax = nent; A.rec = b.rec;
    A.fi = MbFeatIndex{ npix, (uint32_t)W, (uint32_t)(H / fh), (uint32_t)(W / fw), (uint32_t)fw };
    A.fhw = (uint32_t)fh * (uint32_t)fw;
    A.features = features; A.class_ids = class_ids; A.F = F; A.map = map; A.affine_a = affine_a; A.g = g;
    A.alpha = alpha;
    return dispatch_brick_reduce(stream, A);
}
