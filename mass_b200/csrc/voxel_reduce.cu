// K2: deterministic voxel reduce + map update (sm_100a).
//
// Input: splat contributions sorted by voxel key with a STABLE sort, so inside one voxel segment
// they appear in the reference's order (contribution id = slot * N + point,
// /root/reference/mass/utils/projection.py:294-298, 319-323).  One warp owns one voxel segment:
// lanes are feature channels, the segment is walked sequentially, nothing is combined with
// atomics, so the result is independent of scheduling.
//
// MB_MODE_EXACT restates projection.py:335-351 operation by operation (no FMA):
//     W = sum w;  new = sum_i (((1 - a*w_i)*old + (a*w_i)*f_i) * w_i) / W
// and is bitwise equal to the reference CPU path.  MB_MODE_FAST uses the algebraically equal
// per-voxel affine form new = (1 - a*S2/W)*old + (a/W)*sum w_i^2 f_i with one
// pass, one division per voxel and FMAs; it differs from EXACT by fp32 re-association only.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// weight of slot s for a point with in-voxel ratios q (projection.py:305-323)
__device__ __forceinline__ float slot_weight(const float4 q, int s)
{
    const float r0 = q.x, r1 = q.y, r2 = q.z;
    const float w0 = (s & 4) ? (r0 < 0.5f ? __fadd_rn(r0, 0.5f) : __fsub_rn(r0, 0.5f))
                             : (r0 < 0.5f ? __fsub_rn(0.5f, r0) : __fsub_rn(1.5f, r0));
    const float w1 = (s & 2) ? (r1 < 0.5f ? __fadd_rn(r1, 0.5f) : __fsub_rn(r1, 0.5f))
                             : (r1 < 0.5f ? __fsub_rn(0.5f, r1) : __fsub_rn(1.5f, r1));
    const float w2 = (s & 1) ? (r2 < 0.5f ? __fadd_rn(r2, 0.5f) : __fsub_rn(r2, 0.5f))
                             : (r2 < 0.5f ? __fsub_rn(0.5f, r2) : __fsub_rn(1.5f, r2));
    return __fadd_rn(1e-9f, __fmul_rn(__fmul_rn(w0, w1), w2));
}

__global__ void __launch_bounds__(256)
k_segment_heads(const uint32_t *__restrict__ keys, uint32_t n, uint32_t invalid, uint32_t *__restrict__ heads,
                uint32_t *__restrict__ counters)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool head = false;
    if (i < n) {
        const uint32_t k = keys[i];
        head = k < invalid && (i == 0 || keys[i - 1] != k);
    }
    const uint32_t m = __ballot_sync(FULL, head);
    if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counters[MB_CNT_HEADS], (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (head) heads[base + __popc(m & ((1u << lane) - 1u))] = i;
    }
}

template <int CPL, bool EXACT, bool ONEHOT>
__global__ void __launch_bounds__(256)
k_voxel_reduce(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n,
               const uint32_t *__restrict__ heads, const uint32_t *__restrict__ counters,
               const float4 *__restrict__ pt_ratio, MbFeatIndex fi, const float *__restrict__ features,
               const int64_t *__restrict__ class_ids, int F, float *__restrict__ map, MbGrid g, float alpha)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nheads = counters[MB_CNT_HEADS];

    for (uint32_t seg = warp; seg < nheads; seg += nwarps) {
        const uint32_t start = heads[seg];
        const uint32_t key = keys[start];
        float *row = map + mb_key_to_voxel(g, key) * (size_t)F;
        float old[CPL], acc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int c = lane + 32 * k;
            old[k] = c < F ? row[c] : 0.f;
            acc[k] = 0.f;
        }
        float W = 0.f, S2 = 0.f;

        if (EXACT) {   // pass A: W in contribution order
            for (uint32_t base = start;; base += 32) {
                const uint32_t idx = base + lane;
                const bool mine = idx < n && keys[idx] == key;
                const uint32_t m = __ballot_sync(FULL, mine);
                const int cnt = m == FULL ? 32 : __ffs(~m) - 1;
                float w = 0.f;
                if (lane < cnt) {
                    const uint32_t ci = vals[idx];
                    const uint32_t s = ci / fi.np;
                    w = slot_weight(pt_ratio[ci - s * fi.np], (int)s);
                }
                for (int j = 0; j < cnt; ++j) W = __fadd_rn(W, __shfl_sync(FULL, w, j));
                if (cnt < 32) break;
            }
        }

        for (uint32_t base = start;; base += 32) {
            const uint32_t idx = base + lane;
            const bool mine = idx < n && keys[idx] == key;
            const uint32_t m = __ballot_sync(FULL, mine);
            const int cnt = m == FULL ? 32 : __ffs(~m) - 1;
            float w = 0.f;
            uint32_t src = 0;   // feature row index, or class id in one-hot mode
            if (lane < cnt) {
                const uint32_t ci = vals[idx];
                const uint32_t s = ci / fi.np;
                const uint32_t p = ci - s * fi.np;
                w = slot_weight(pt_ratio[p], (int)s);
                if (ONEHOT) {
                    src = (uint32_t)class_ids[p];
                } else if (fi.kx == 1 && fi.ky == 1) {
                    src = p;
                } else {
                    const uint32_t y = p / fi.W, x = p - y * fi.W;
                    src = (y / fi.ky) * fi.fw + x / fi.kx;
                }
            }
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float wj = __shfl_sync(FULL, w, j);
                const uint32_t sj = __shfl_sync(FULL, src, j);
                const float *frow = features + (size_t)sj * F;
                if (EXACT) {
                    const float aw = __fmul_rn(alpha, wj);
                    const float keep = __fsub_rn(1.0f, aw);
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const int c = lane + 32 * k;
                        if (c < F) {
                            const float f = ONEHOT ? (sj == (uint32_t)c ? 1.0f : 0.0f) : __ldg(frow + c);
                            float t = __fadd_rn(__fmul_rn(keep, old[k]), __fmul_rn(aw, f));
                            t = __fdiv_rn(__fmul_rn(t, wj), W);
                            acc[k] = __fadd_rn(acc[k], t);
                        }
                    }
                } else {
                    const float w2 = wj * wj;
                    W += wj;
                    S2 += w2;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const int c = lane + 32 * k;
                        if (c < F) {
                            const float f = ONEHOT ? (sj == (uint32_t)c ? 1.0f : 0.0f) : __ldg(frow + c);
                            acc[k] = fmaf(w2, f, acc[k]);
                        }
                    }
                }
            }
            if (cnt < 32) break;
        }

#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int c = lane + 32 * k;
            if (c < F) {
                float out = acc[k];
                if (!EXACT) {
                    const float a = 1.0f - alpha * S2 / W;
                    out = fmaf(a, old[k], (alpha / W) * acc[k]);
                }
                row[c] = out;
            }
        }
    }
}

template <int CPL>
int launch_reduce(cudaStream_t stream, const uint32_t *keys, const uint32_t *vals, uint32_t n,
                  const uint32_t *heads, const uint32_t *counters, const float4 *pt_ratio,
                  const MbFeatIndex &fi, const float *features, const int64_t *class_ids, int F, float *map,
                  const MbGrid &g, float alpha, int mode)
{
    const int blocks = MB_NUM_SMS * 8;
    const bool onehot = class_ids != nullptr;
#define MB_RED(EX, OH)                                                                                   \
    k_voxel_reduce<CPL, EX, OH><<<blocks, 256, 0, stream>>>(keys, vals, n, heads, counters, pt_ratio, fi, \
                                                            features, class_ids, F, map, g, alpha)
    if (mode == MB_MODE_EXACT) {
        if (onehot) MB_RED(true, true); else MB_RED(true, false);
    } else {
        if (onehot) MB_RED(false, true); else MB_RED(false, false);
    }
#undef MB_RED
    MB_LAUNCHED();
    return MB_OK;
}

}  // namespace

int mbk_segment_heads(cudaStream_t stream, const uint32_t *keys, uint32_t n, const MbGrid &g,
                      uint32_t *heads, uint32_t *counters)
{
    if (n == 0) return MB_OK;
    k_segment_heads<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, g.invalid, heads, counters);
    MB_LAUNCHED();
    return MB_OK;
}

int mbk_voxel_reduce(cudaStream_t stream, const uint32_t *keys, const uint32_t *vals, uint32_t n,
                     const uint32_t *heads, const uint32_t *counters, const float4 *pt_ratio,
                     const MbFeatIndex &fi, const float *features, const int64_t *class_ids, int F,
                     float *map, const MbGrid &g, float alpha, int mode)
{
    if (n == 0) return MB_OK;
    MB_REQUIRE(F >= 1 && F <= 512, "feature_size %d not supported (1..512)", F);
#define MB_GO(C) return launch_reduce<C>(stream, keys, vals, n, heads, counters, pt_ratio, fi, features, \
                                         class_ids, F, map, g, alpha, mode)
    if (F <= 32) MB_GO(1);
    if (F <= 64) MB_GO(2);
    if (F <= 128) MB_GO(4);
    if (F <= 256) MB_GO(8);
    MB_GO(16);
#undef MB_GO
}
