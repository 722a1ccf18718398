// K3 / K4: instance extraction and matching kernels (sm_100a).
//
//   mbk_class_presence   /root/reference/mass/nn/applications/semantic_projection_layer.py:309-317
//                        (avg_pool3d box mean of one class channel, threshold, any over z)
//   mbk_instance_pool    semantic_projection_layer.py:329-357 (per bounding box: confidence, expected
//                        position, size, pooled instance feature)
//   mbk_pairwise_l2      /root/reference/mass/utils/experimentation.py:261-265, 277-280
//   mbk_lsap             experimentation.py:284-287 -> scipy.optimize.linear_sum_assignment (third party;
//                        Crouse 2016 shortest augmenting path), one CTA, float64 duals
//
// All reductions run in a fixed order (per-thread strided partial sums, then a fixed shared-memory
// tree), so results are reproducible run to run.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// class presence.  Pass 1 (pad == 0): image[y][x] = any_z(map[y][x][z][c] > thr).
__global__ void __launch_bounds__(256)
k_presence_nopad(const float *__restrict__ map, int S0, int S1, int S2, int F, int c, float thr,
                 uint8_t *__restrict__ image)
{
    // one warp per (y, x) column, lanes stride over z
    const int lane = threadIdx.x & 31;
    const size_t col = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= (size_t)S0 * S1) return;
    const float *p = map + col * (size_t)S2 * F + c;
    bool any = false;
    for (int z = lane; z < S2; z += 32) any |= p[(size_t)z * F] > thr;
    any = __any_sync(FULL, any);
    if (lane == 0) image[col] = any ? 1 : 0;
}

// Box sums for pad > 0 are separable: along z (from the strided class channel), along x, along y.
// out[y][x][z] = sum_{dz=-pad..pad} map[y][x][z+dz][c]   (zero outside the map)
__global__ void __launch_bounds__(128)
k_boxsum_z(const float *__restrict__ map, int S0, int S1, int S2, int F, int c, int pad, float *__restrict__ out)
{
    extern __shared__ float s_col[];
    const size_t col = blockIdx.x;
    const float *p = map + col * (size_t)S2 * F + c;
    for (int z = threadIdx.x; z < S2; z += blockDim.x) s_col[z] = p[(size_t)z * F];
    __syncthreads();
    for (int z = threadIdx.x; z < S2; z += blockDim.x) {
        float s = 0.f;
        for (int d = -pad; d <= pad; ++d) {
            const int zz = z + d;
            if (zz >= 0 && zz < S2) s += s_col[zz];
        }
        out[col * S2 + z] = s;
    }
}

// sums along one of the two planar axes: `stride` elements between neighbours, `len` positions
__global__ void __launch_bounds__(256)
k_boxsum_axis(const float *__restrict__ in, size_t total, int len, size_t stride, int pad, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int pos = (int)((i / stride) % (size_t)len);
    float s = 0.f;
    for (int d = -pad; d <= pad; ++d) {
        const int q = pos + d;
        if (q >= 0 && q < len) s += in[i + (ptrdiff_t)d * (ptrdiff_t)stride];
    }
    out[i] = s;
}

__global__ void __launch_bounds__(256)
k_presence_from_sums(const float *__restrict__ sums, int S0, int S1, int S2, float count, float thr,
                     uint8_t *__restrict__ image)
{
    const int lane = threadIdx.x & 31;
    const size_t col = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= (size_t)S0 * S1) return;
    bool any = false;
    for (int z = lane; z < S2; z += 32) any |= __fdiv_rn(sums[col * S2 + z], count) > thr;
    any = __any_sync(FULL, any);
    if (lane == 0) image[col] = any ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// column summaries (the readers of the whole map that sit next to the path):
//   amax[y][x][f]  = max over z of map[y][x][z][f]                       (/root/reference/agent.py:330-331, 391-392)
//   blocked[y][x]  = any z in [z_lo, z_hi): sum_f |map[y][x][z][f]| > thr  (mass/navigation_policy.py:207-216)
// One warp per (y, x) column: the column is one contiguous block of S2 * F floats, read once.
template <int VEC>
__global__ void __launch_bounds__(256)
k_column_summary(const float *__restrict__ map, size_t ncols, int S2, int F, int z_lo, int z_hi, float thr,
                 float *__restrict__ amax, uint8_t *__restrict__ blocked)
{
    constexpr int UZ = 8;                            // z rows requested before the first is used
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t col = wid; col < ncols; col += nw) {
        const float *p = map + col * (size_t)S2 * F;
        float rowsum_hi = 0.f;                       // unused; keeps the structure symmetric
        (void)rowsum_hi;
        bool any = false;
        const int nblk = (F + 32 * VEC - 1) / (32 * VEC);
        for (int cb = 0; cb < nblk; ++cb) {          // channel blocks of 32 * VEC (one pass over z per block)
            const int f = (cb * 32 + lane) * VEC;
            const bool on = f < F;
            float m[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) m[j] = -INFINITY;
            for (int z0 = 0; z0 < S2; z0 += UZ) {
                float v[UZ][VEC];
#pragma unroll
                for (int u = 0; u < UZ; ++u) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] = 0.f;
                    if (on && z0 + u < S2) {
                        const float *q = p + (size_t)(z0 + u) * F + f;
                        if (VEC == 1) v[u][0] = __ldg(q);
                        if (VEC == 2) { const float2 t = __ldg((const float2 *)q); v[u][0] = t.x; v[u][VEC > 1 ? 1 : 0] = t.y; }
                        if (VEC == 4) {
                            const float4 t = __ldg((const float4 *)q);
                            v[u][0] = t.x; v[u][VEC > 1 ? 1 : 0] = t.y; v[u][VEC > 2 ? 2 : 0] = t.z; v[u][VEC > 2 ? 3 : 0] = t.w;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UZ; ++u) {
                    if (on && z0 + u < S2) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) m[j] = fmaxf(m[j], v[u][j]);
                    }
                    if (blocked != nullptr && nblk == 1 && z0 + u >= z_lo && z0 + u < z_hi) {
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) s += fabsf(v[u][j]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                        any |= s > thr;
                    }
                }
            }
            if (amax != nullptr && on) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) amax[col * F + f + j] = m[j];
            }
        }
        if (blocked != nullptr && nblk > 1) {
            // wide rows: a second sweep over the slice (L1/L2 hits), lanes stride over the channels of one z row
            for (int z = z_lo; z < z_hi; ++z) {
                float s = 0.f;
                for (int f = lane; f < F; f += 32) s += fabsf(__ldg(p + (size_t)z * F + f));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                any |= s > thr;
            }
        }
        if (blocked != nullptr && lane == 0) blocked[col] = any ? 1 : 0;
    }
}

// perception hand-off (mass/thor/segmentation_config.py:314-334): the detector's instances
// (mask, class, score) -> per-class mask counts -> arg-max class id per pixel (first maximum: class 0 where
// nothing was detected), without the [H, W, 54] float buffer and without leaving the device.
constexpr int IDS_MAX_CLASSES = 128;
__global__ void __launch_bounds__(128)
k_masks_to_ids(const uint8_t *__restrict__ masks, const int64_t *__restrict__ classes, const float *__restrict__ scores,
               int n, size_t npix, int num_classes, float threshold, int64_t *__restrict__ ids)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    unsigned short cnt[IDS_MAX_CLASSES];
    for (int c = 0; c < num_classes; ++c) cnt[c] = 0;
    for (int i = 0; i < n; ++i) {
        if (__ldg(scores + i) < threshold) continue;              // "skip if the model is not confident enough"
        const int64_t c = __ldg(classes + i);
        if (c >= 0 && c < num_classes && __ldg(masks + (size_t)i * npix + p)) ++cnt[c];
    }
    int best = 0;
    for (int c = 1; c < num_classes; ++c)
        if (cnt[c] > cnt[best]) best = c;
    ids[p] = best;
}

// top-down rendering (mass/nn/base_projection_layer.py:345-379): per (y, x) column the feature row
// of the top-most voxel of [z_lo, z_hi) that has any non-zero channel, zeros if there is none (the reference's
// cumsum * mask arg-max picks the last filled voxel, and row z_lo -- all zero -- when none is filled).
// One warp per column, walking down from the top of the slice and stopping at the first filled voxel.
__global__ void __launch_bounds__(256)
k_top_down(const float *__restrict__ map, size_t ncols, int S2, int F, int z_lo, int z_hi, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t col = wid; col < ncols; col += nw) {
        const float *p = map + col * (size_t)S2 * F;
        int top = -1;
        for (int z = z_hi - 1; z >= z_lo && top < 0; --z) {
            bool nz = false;
            for (int f = lane; f < F; f += 32) nz |= __ldg(p + (size_t)z * F + f) != 0.f;
            if (__any_sync(FULL, nz)) top = z;
        }
        float *o = out + col * F;
        for (int f = lane; f < F; f += 32) o[f] = top >= 0 ? __ldg(p + (size_t)top * F + f) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// instance pooling: one CTA per bounding box (x, y, w, h) over the full depth of the map.
//   out row = [confidence, coord_x, coord_y, coord_z, size, feature[FF]]
constexpr int POOL_THREADS = 256;
constexpr int POOL_LIST = 2048;      // non-zero voxels of the box processed per round

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
    const int tid = threadIdx.x;
    s_red[tid] = v;
    __syncthreads();
    for (int step = POOL_THREADS / 2; step > 0; step >>= 1) {
        if (tid < step) s_red[tid] += s_red[tid + step];
        __syncthreads();
    }
    const double r = s_red[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(POOL_THREADS)
k_instance_pool(const int *__restrict__ boxes, const float *__restrict__ sem, int S0, int S1, int S2, int F, int c,
                const float *__restrict__ feat, int FF, const float *__restrict__ mx, const float *__restrict__ my,
                const float *__restrict__ mz, float *__restrict__ out)
{
    __shared__ double s_red[POOL_THREADS];
    __shared__ uint32_t s_list[POOL_LIST];     // box-local voxel ids with a non-zero mask
    __shared__ float s_lw[POOL_LIST];          // their mask values
    __shared__ uint32_t s_count;
    extern __shared__ float s_facc[];          // [POOL_THREADS / 32][FF] per-warp feature sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // c >= 0: boxes are (x, y, w, h), all of class c; c < 0: boxes are (x, y, w, h, class)
    const int bstride = c >= 0 ? 4 : 5;
    const int bx = boxes[bstride * blockIdx.x], by = boxes[bstride * blockIdx.x + 1], bw = boxes[bstride * blockIdx.x + 2],
              bh = boxes[bstride * blockIdx.x + 3];
    if (c < 0) c = boxes[bstride * blockIdx.x + 4];
    const uint32_t nvox = (uint32_t)bw * bh * S2;
    float *orow = out + (size_t)blockIdx.x * (5 + FF);

    // pass 1: total, sum of squares, first moments (double accumulators, fixed order)
    double t = 0, t2 = 0, tx = 0, ty = 0, tz = 0;
    for (uint32_t i = tid; i < nvox; i += POOL_THREADS) {
        const int z = i % S2;
        const int x = (i / S2) % bw, y = i / (S2 * bw);
        const float m = sem[(((size_t)(by + y) * S1 + (bx + x)) * S2 + z) * F + c];
        t += m;
        t2 += (double)m * m;
        tx += (double)m * mx[bx + x];
        ty += (double)m * my[by + y];
        tz += (double)m * mz[z];
    }
    t = block_sum(t, s_red);
    t2 = block_sum(t2, s_red);
    tx = block_sum(tx, s_red);
    ty = block_sum(ty, s_red);
    tz = block_sum(tz, s_red);
    const double denom = (double)((float)t + 1e-9f);       // mask_roi.sum() + 1e-9 in fp32
    if (tid == 0) {
        orow[0] = (float)(t2 / denom);
        orow[1] = (float)(tx / denom);
        orow[2] = (float)(ty / denom);
        orow[3] = (float)(tz / denom);
        orow[4] = (float)t;
    }
    if (feat == nullptr || FF == 0) return;

    // pass 2: pooled feature = sum_vox mask * feat[vox] / denom over the non-zero voxels only
    const int nwarps = POOL_THREADS / 32;
    for (int k = tid; k < nwarps * FF; k += POOL_THREADS) s_facc[k] = 0.f;
    for (uint32_t base = 0; base < nvox; base += POOL_LIST) {
        if (tid == 0) s_count = 0;
        __syncthreads();
        // ordered compaction of the non-zero voxels of this slice (deterministic list order)
        for (uint32_t off = 0; off < POOL_LIST; off += POOL_THREADS) {
            const uint32_t i = base + off + tid;
            float m = 0.f;
            if (i < nvox) {
                const int z = i % S2;
                const int x = (i / S2) % bw, y = i / (S2 * bw);
                m = sem[(((size_t)(by + y) * S1 + (bx + x)) * S2 + z) * F + c];
            }
            const bool nz = m != 0.f;
            const uint32_t bal = __ballot_sync(FULL, nz);
            __shared__ uint32_t s_wbase[POOL_THREADS / 32];
            if (lane == 0) s_wbase[warp] = __popc(bal);
            __syncthreads();
            uint32_t before = s_count;
            for (int w = 0; w < warp; ++w) before += s_wbase[w];
            if (nz) {
                const uint32_t slot = before + __popc(bal & ((1u << lane) - 1u));
                s_list[slot] = i;
                s_lw[slot] = m;
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t add = 0;
                for (int w = 0; w < nwarps; ++w) add += s_wbase[w];
                s_count += add;
            }
            __syncthreads();
        }
        const uint32_t cnt = s_count;
        // warp w takes list items w, w + nwarps, ...; lanes are channels
        for (uint32_t k = warp; k < cnt; k += nwarps) {
            const uint32_t i = s_list[k];
            const float m = s_lw[k];
            const int z = i % S2;
            const int x = (i / S2) % bw, y = i / (S2 * bw);
            const float *frow = feat + (((size_t)(by + y) * S1 + (bx + x)) * S2 + z) * FF;
            for (int ch = lane; ch < FF; ch += 32) s_facc[warp * FF + ch] = fmaf(m, __ldg(frow + ch), s_facc[warp * FF + ch]);
        }
        __syncthreads();
    }
    for (int ch = tid; ch < FF; ch += POOL_THREADS) {
        double s = 0;
        for (int w = 0; w < nwarps; ++w) s += s_facc[w * FF + ch];
        orow[5 + ch] = (float)(s / denom);
    }
}

// ---------------------------------------------------------------------------------------------
// pairwise L2: one warp per (i, j), direct differences (the GEMM form drifts: SURVEY.md section 7)
__global__ void __launch_bounds__(256)
k_pairwise_l2(const float *__restrict__ a, int n, const float *__restrict__ b, int m, int d, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const size_t pair = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pair >= (size_t)n * m) return;
    const int i = (int)(pair / m), j = (int)(pair % m);
    const float *pa = a + (size_t)i * d, *pb = b + (size_t)j * d;
    double s = 0;
    for (int k = lane; k < d; k += 32) {
        const float diff = __fsub_rn(pa[k], pb[k]);
        s += (double)diff * (double)diff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) out[pair] = (float)sqrt(s);
}

// ---------------------------------------------------------------------------------------------
// cosine similarity + best match (NOT on the reference's path, which matches by L2 + assignment: SURVEY.md F3;
// offered because the task statement names it).  One warp per query row i: sim(i, j) = a_i.b_j / (|a_i| |b_j|)
// with float64 accumulation, best j = the first maximum (torch.argmax's tie rule).  For the instance counts of
// this path (<= a few hundred rows of 256) this is a latency-sized job; a tcgen05 GEMM would only pay off
// beyond a few thousand instances.
__global__ void __launch_bounds__(256)
k_cosine_best_match(const float *__restrict__ a, int n, const float *__restrict__ b, int m, int d,
                    int64_t *__restrict__ best, float *__restrict__ best_sim)
{
    const int lane = threadIdx.x & 31;
    const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const float *pa = a + (size_t)i * d;
    double na = 0;
    for (int k = lane; k < d; k += 32) na += (double)pa[k] * pa[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) na += __shfl_xor_sync(FULL, na, o);
    double top = -INFINITY;
    int arg = 0;
    for (int j = 0; j < m; ++j) {
        const float *pb = b + (size_t)j * d;
        double dot = 0, nb = 0;
        for (int k = lane; k < d; k += 32) {
            const double x = pb[k];
            dot += (double)pa[k] * x;
            nb += x * x;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(FULL, dot, o);
            nb += __shfl_xor_sync(FULL, nb, o);
        }
        const double den = sqrt(na) * sqrt(nb);
        const double sim = den > 0 ? dot / den : 0.0;
        if (sim > top) { top = sim; arg = j; }
    }
    if (lane == 0) {
        best[i] = m > 0 ? arg : -1;
        best_sim[i] = m > 0 ? (float)top : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// rectangular linear sum assignment, nr <= nc (the host wrapper transposes otherwise).
// Follows the scan order and tie rule of SURVEY.md Appendix B: columns are scanned in the order of
// the `remaining` list (initialised in reverse, swap-removed); among equal shortest path costs the
// first scanned column wins unless a later scanned one is unassigned, in which case the last
// unassigned one wins.
constexpr int LSAP_THREADS = 256;

struct Best {
    double val;
    int first;      // first scan position attaining val
    int lastfree;   // last scan position attaining val whose column is unassigned (-1: none)
};

__device__ __forceinline__ Best best_merge(const Best &x, const Best &y)   // x covers earlier positions than y
{
    if (x.val < y.val) return x;
    if (y.val < x.val) return y;
    Best r;
    r.val = x.val;
    r.first = x.first < 0 ? y.first : x.first;
    r.lastfree = y.lastfree >= 0 ? y.lastfree : x.lastfree;
    return r;
}

__global__ void __launch_bounds__(LSAP_THREADS)
k_lsap(const float *__restrict__ cost32, const double *__restrict__ cost64, int nr, int nc, int transposed,
       double *__restrict__ work, int *__restrict__ iwork, int64_t *__restrict__ col4row_out, int *__restrict__ status)
{
    // cost(i, j): element of the nr x nc problem; if transposed the stored matrix is nc x nr
    double *u = work, *v = u + nr, *spc = v + nc;
    int *path = iwork, *row4col = path + nc, *col4row = row4col + nc, *remaining = col4row + nr;
    int *SR = remaining + nc, *SC = SR + nr;
    __shared__ Best s_best[LSAP_THREADS];
    __shared__ int s_i, s_sink, s_nrem, s_fail;
    __shared__ double s_minval;
    const int tid = threadIdx.x;

    for (int k = tid; k < nr; k += LSAP_THREADS) { u[k] = 0.0; col4row[k] = -1; }
    for (int k = tid; k < nc; k += LSAP_THREADS) { v[k] = 0.0; row4col[k] = -1; }
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (int cur = 0; cur < nr; ++cur) {
        for (int k = tid; k < nc; k += LSAP_THREADS) { remaining[k] = nc - k - 1; spc[k] = INFINITY; path[k] = -1; SC[k] = 0; }
        for (int k = tid; k < nr; k += LSAP_THREADS) SR[k] = 0;
        if (tid == 0) { s_i = cur; s_sink = -1; s_nrem = nc; s_minval = 0.0; }
        __syncthreads();
        while (s_sink == -1) {
            const int i = s_i, nrem = s_nrem;
            const double minval = s_minval, ui = u[i];
            if (tid == 0) SR[i] = 1;
            Best mine;
            mine.val = INFINITY; mine.first = -1; mine.lastfree = -1;
            // each thread scans a contiguous slice of the scan order
            const int per = (nrem + LSAP_THREADS - 1) / LSAP_THREADS;
            const int lo = tid * per, hi = min(nrem, lo + per);
            for (int it = lo; it < hi; ++it) {
                const int j = remaining[it];
                const double c = cost64 ? (transposed ? cost64[(size_t)j * nr + i] : cost64[(size_t)i * nc + j])
                                        : (double)(transposed ? cost32[(size_t)j * nr + i] : cost32[(size_t)i * nc + j]);
                const double r = minval + c - ui - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                Best b;
                b.val = spc[j]; b.first = it; b.lastfree = row4col[j] == -1 ? it : -1;
                mine = best_merge(mine, b);
            }
            s_best[tid] = mine;
            __syncthreads();
            for (int step = 1; step < LSAP_THREADS; step <<= 1) {      // ordered tree: left covers earlier positions
                if ((tid & (2 * step - 1)) == 0 && tid + step < LSAP_THREADS) s_best[tid] = best_merge(s_best[tid], s_best[tid + step]);
                __syncthreads();
            }
            if (tid == 0) {
                const Best b = s_best[0];
                if (b.first < 0 || b.val == INFINITY) {
                    s_fail = 1;
                    s_sink = -2;
                } else {
                    const int index = b.lastfree > b.first ? b.lastfree : b.first;
                    const int j = remaining[index];
                    s_minval = b.val;
                    if (row4col[j] == -1) s_sink = j; else s_i = row4col[j];
                    SC[j] = 1;
                    remaining[index] = remaining[nrem - 1];
                    s_nrem = nrem - 1;
                }
            }
            __syncthreads();
        }
        if (s_fail) break;
        const double minval = s_minval;
        const int sink = s_sink;
        __syncthreads();
        for (int k = tid; k < nr; k += LSAP_THREADS)
            if (k == cur) u[k] += minval;
            else if (SR[k]) u[k] += minval - spc[col4row[k]];
        for (int k = tid; k < nc; k += LSAP_THREADS)
            if (SC[k]) v[k] -= minval - spc[k];
        __syncthreads();
        if (tid == 0) {
            int j = sink;
            for (;;) {
                const int i = path[j];
                row4col[j] = i;
                const int t = col4row[i];
                col4row[i] = j;
                j = t;
                if (i == cur) break;
            }
        }
        __syncthreads();
    }
    for (int k = tid; k < nr; k += LSAP_THREADS) col4row_out[k] = col4row[k];
    if (tid == 0) *status = s_fail;
}

}  // namespace

size_t mbk_class_presence_workspace_bytes(int S0, int S1, int S2, int pad)
{
    if (pad <= 0) return 256;
    return 2 * mb_align_up((size_t)S0 * S1 * S2 * sizeof(float)) + 256;
}

int mbk_class_presence(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int c, int pad, float thr,
                       uint8_t *image, void *workspace, size_t workspace_bytes)
{
    const size_t cols = (size_t)S0 * S1;
    const unsigned blocks = (unsigned)((cols * 32 + 255) / 256);
    if (pad <= 0) {
        k_presence_nopad<<<blocks, 256, 0, stream>>>(map, S0, S1, S2, F, c, thr, image);
        MB_LAUNCHED();
        return MB_OK;
    }
    MB_REQUIRE(workspace && workspace_bytes >= mbk_class_presence_workspace_bytes(S0, S1, S2, pad),
               "class presence workspace too small");
    MbArena arena(workspace, workspace_bytes);
    const size_t total = cols * S2;
    float *t0 = arena.take<float>(total), *t1 = arena.take<float>(total);
    k_boxsum_z<<<(unsigned)cols, 128, (size_t)S2 * sizeof(float), stream>>>(map, S0, S1, S2, F, c, pad, t0);
    MB_LAUNCHED();
    const unsigned eb = (unsigned)((total + 255) / 256);
    k_boxsum_axis<<<eb, 256, 0, stream>>>(t0, total, S1, (size_t)S2, pad, t1);              // along x
    MB_LAUNCHED();
    k_boxsum_axis<<<eb, 256, 0, stream>>>(t1, total, S0, (size_t)S1 * S2, pad, t0);         // along y
    MB_LAUNCHED();
    const int k = 2 * pad + 1;
    k_presence_from_sums<<<blocks, 256, 0, stream>>>(t0, S0, S1, S2, (float)(k * k * k), thr, image);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_column_summary(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float thr,
                       float *amax, uint8_t *blocked)
{
    const size_t ncols = (size_t)S0 * S1;
    const uintptr_t al = (uintptr_t)map | (uintptr_t)amax;
    if (F % 4 == 0 && al % 16 == 0)
        k_column_summary<4><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, ncols, S2, F, z_lo, z_hi, thr, amax, blocked);
    else if (F % 2 == 0 && al % 8 == 0)
        k_column_summary<2><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, ncols, S2, F, z_lo, z_hi, thr, amax, blocked);
    else
        k_column_summary<1><<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, ncols, S2, F, z_lo, z_hi, thr, amax, blocked);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_masks_to_ids(cudaStream_t stream, const uint8_t *masks, const int64_t *classes, const float *scores, int n,
                     size_t npix, int num_classes, float threshold, int64_t *ids)
{
    MB_REQUIRE(num_classes >= 1 && num_classes <= IDS_MAX_CLASSES, "mb_masks_to_ids: between 1 and %d classes", IDS_MAX_CLASSES);
    MB_REQUIRE(n < 65536, "mb_masks_to_ids: too many instances");
    k_masks_to_ids<<<(unsigned)((npix + 127) / 128), 128, 0, stream>>>(masks, classes, scores, n, npix, num_classes,
                                                                      threshold, ids);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_top_down(cudaStream_t stream, const float *map, int S0, int S1, int S2, int F, int z_lo, int z_hi, float *out)
{
    k_top_down<<<MB_NUM_SMS * 8, 256, 0, stream>>>(map, (size_t)S0 * S1, S2, F, z_lo, z_hi, out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_instance_pool(cudaStream_t stream, const int *boxes, int nboxes, const float *sem, int S0, int S1, int S2, int F,
                      int c, const float *feat, int FF, const float *mx, const float *my, const float *mz, float *out)
{
    if (nboxes <= 0) return MB_OK;
    const size_t smem = (size_t)(POOL_THREADS / 32) * (feat ? FF : 0) * sizeof(float);
    MB_REQUIRE(smem <= 40000, "instance feature size %d too large", FF);
    k_instance_pool<<<nboxes, POOL_THREADS, smem, stream>>>(boxes, sem, S0, S1, S2, F, c, feat, feat ? FF : 0, mx, my, mz, out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_pairwise_l2(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, float *out)
{
    if (n <= 0 || m <= 0) return MB_OK;
    const size_t threads = (size_t)n * m * 32;
    k_pairwise_l2<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(a, n, b, m, d, out);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_cosine_best_match(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                          float *best_sim)
{
    if (n <= 0) return MB_OK;
    k_cosine_best_match<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, stream>>>(a, n, b, m, d, best, best_sim);
    MB_LAUNCHED();
    return MB_OK;
}

size_t mbk_lsap_workspace_bytes(int n, int m)
{
    const size_t nr = n < m ? n : m, nc = n < m ? m : n;
    return mb_align_up((nr + 2 * nc) * sizeof(double)) + mb_align_up((4 * nc + 3 * nr) * sizeof(int)) +
           mb_align_up(nr * sizeof(int64_t)) + 1024;
}

// cost is n x m (float32 or float64, exactly one non-null); writes min(n, m) pairs sorted by row
int mbk_lsap(cudaStream_t stream, const float *cost32, const double *cost64, int n, int m, int64_t *rows, int64_t *cols,
             int *status, void *workspace, size_t workspace_bytes);

namespace {
__global__ void k_lsap_finish(const int64_t *__restrict__ col4row, int nr, int nc, int transposed, int64_t *rows,
                              int64_t *cols, int *scratch)
{
    // not transposed: pairs (i, col4row[i]).  transposed: the problem's rows are the matrix's columns:
    // pairs (row = col4row[j], col = j), to be listed by increasing row.
    if (!transposed) {
        for (int i = threadIdx.x; i < nr; i += blockDim.x) { rows[i] = i; cols[i] = col4row[i]; }
        return;
    }
    for (int i = threadIdx.x; i < nc; i += blockDim.x) scratch[i] = -1;
    __syncthreads();
    for (int j = threadIdx.x; j < nr; j += blockDim.x) scratch[col4row[j]] = j;
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = 0;
        for (int i = 0; i < nc; ++i)
            if (scratch[i] >= 0) { rows[k] = i; cols[k] = scratch[i]; ++k; }
    }
}
}  // namespace

int mbk_lsap(cudaStream_t stream, const float *cost32, const double *cost64, int n, int m, int64_t *rows, int64_t *cols,
             int *status, void *workspace, size_t workspace_bytes)
{
    if (n <= 0 || m <= 0) return MB_OK;
    MB_REQUIRE(workspace && workspace_bytes >= mbk_lsap_workspace_bytes(n, m), "lsap workspace too small");
    const int transposed = m < n ? 1 : 0;
    const int nr = transposed ? m : n, nc = transposed ? n : m;
    MbArena arena(workspace, workspace_bytes);
    double *work = arena.take<double>((size_t)nr + 2 * nc);
    int *iwork = arena.take<int>((size_t)4 * nc + 3 * nr);
    int64_t *col4row = arena.take<int64_t>((size_t)nr);
    k_lsap<<<1, LSAP_THREADS, 0, stream>>>(cost32, cost64, nr, nc, transposed, work, iwork, col4row, status);
    MB_LAUNCHED();
    k_lsap_finish<<<1, 256, 0, stream>>>(col4row, nr, nc, transposed, rows, cols, iwork);
    MB_LAUNCHED();
    return MB_OK;
}
