// Device-side geometry shared by the per-frame and the batched kernels: the reference's fp32
// operation sequence for ray orientation, unprojection and binning, written with
// explicit round-to-nearest intrinsics so no FMA contraction can change a bit.
//   transform_rays   /root/reference/mass/utils/projection.py:104-110
//   bin_rays         /root/reference/mass/utils/projection.py:182-230
//   neighbour pairs  /root/reference/mass/utils/projection.py:280-291
#pragma once
#include "common.cuh"

namespace {

// torch.bucketize(x, bins, right=True) - 1 == (number of edges <= x) - 1, in [-1, n-1].
// Guess from the (near-uniform) table spacing, then walk to the exact answer on the table
// itself: the table is ATen's arange output and is the only authority on edge positions.
__device__ __forceinline__ int bucket_right(const float *__restrict__ bins, int n, float x)
{
    if (x != x) return n - 1;                     // NaN: ATen's upper bound runs off the end
    const float b0 = __ldg(bins), b1 = __ldg(bins + n - 1);
    if (!(x >= b0)) return -1;
    int i;
    if (x >= b1) {
        i = n - 1;
    } else {
        const float g = (x - b0) * ((float)(n - 1) / (b1 - b0));
        i = g >= (float)(n - 1) ? n - 1 : (int)g;
        if (i < 0) i = 0;
    }
    while (i + 1 < n && __ldg(bins + i + 1) <= x) ++i;
    while (i >= 0 && __ldg(bins + i) > x) --i;
    return i;
}

struct BinResult {
    int i0, i1, i2;      // bucket per input axis (x, y, z); valid only if ok
    float q0, q1, q2;    // ratio per input axis (axis 1 already flipped together with i1)
    bool ok;
};

// world point + binning of one pixel.  ray = oriented ray, o = origin.
__device__ __forceinline__ BinResult bin_point(const float *__restrict__ bins0, int n0,
                                               const float *__restrict__ bins1, int n1,
                                               const float *__restrict__ bins2, int n2, float o0, float o1,
                                               float o2, float r0, float r1, float r2, float d,
                                               float min_d, float max_d)
{
    BinResult b;
    const float x0 = __fadd_rn(o0, __fmul_rn(r0, d));
    const float x1 = __fadd_rn(o1, __fmul_rn(r1, d));
    const float x2 = __fadd_rn(o2, __fmul_rn(r2, d));
    const int i0 = bucket_right(bins0, n0, x0);
    const int i1 = bucket_right(bins1, n1, x1);
    const int i2 = bucket_right(bins2, n2, x2);
    b.ok = (d >= min_d) && (d <= max_d) && i0 >= 0 && i0 < n0 - 1 && i1 >= 0 && i1 < n1 - 1 &&
           i2 >= 0 && i2 < n2 - 1;
    b.i0 = i0; b.i1 = i1; b.i2 = i2;
    b.q0 = b.q1 = b.q2 = 0.f;
    if (b.ok) {
        const float l0 = __ldg(bins0 + i0), h0 = __ldg(bins0 + i0 + 1);
        const float l1 = __ldg(bins1 + i1), h1 = __ldg(bins1 + i1 + 1);
        const float l2 = __ldg(bins2 + i2), h2 = __ldg(bins2 + i2 + 1);
        b.q0 = __fdiv_rn(__fsub_rn(x0, l0), __fsub_rn(h0, l0));
        b.q1 = __fsub_rn(1.0f, __fdiv_rn(__fsub_rn(x1, l1), __fsub_rn(h1, l1)));
        b.q2 = __fdiv_rn(__fsub_rn(x2, l2), __fsub_rn(h2, l2));
        b.i1 = n1 - 2 - i1;
    }
    return b;
}

// Same search with the table spacing precomputed by the caller (b0 = bins[0], scale = (n-1) / (bins[n-1] - b0))
// and the two edges around the answer returned, so the ratio needs no further loads.
struct Bucket { int i; float lo, hi; };

__device__ __forceinline__ Bucket bucket_right_edges(const float *__restrict__ bins, int n, float x, float b0, float scale)
{
    Bucket r;
    r.lo = r.hi = 0.f;
    if (x != x) { r.i = n - 1; return r; }          // NaN: ATen's upper bound runs off the end
    if (!(x >= b0)) { r.i = -1; return r; }
    const float g = (x - b0) * scale;
    int i = g >= (float)(n - 2) ? n - 2 : (int)g;
    if (i < 0) i = 0;
    float lo = __ldg(bins + i), hi = __ldg(bins + i + 1);
    while (x >= hi) {
        if (++i >= n - 1) { r.i = n - 1; return r; }
        lo = hi;
        hi = __ldg(bins + i + 1);
    }
    while (x < lo) {                                 // x >= bins[0], so i stays >= 0
        --i;
        hi = lo;
        lo = __ldg(bins + i);
    }
    r.i = i; r.lo = lo; r.hi = hi;
    return r;
}

// table spacing helpers: call once per axis (per CTA), keep in shared memory
__device__ __forceinline__ float bins_scale(const float *__restrict__ bins, int n)
{
    return (float)(n - 1) / (__ldg(bins + n - 1) - __ldg(bins));
}

// Bucket of x for callers that only need VALID buckets: returns true iff bins[0] <= x < bins[n-1] (and x is
// not NaN), i.e. iff torch.bucketize(x, bins, right=True) - 1 lies in [0, n-2]; then i is that value and
// lo = bins[i], hi = bins[i+1].  The guess from the table spacing is checked against the table itself (the only
// authority on edge positions); on a near-uniform table it is right except within an ulp or two of an edge, so
// everything else -- one step to a neighbour, the walk general tables need, points outside the table -- lives in
// an out-of-line function that a warp only enters when one of its lanes needs it.
struct BucketFix { int k; float lo, hi; int ok; };

__device__ __noinline__ BucketFix bucket_walk(const float *__restrict__ bins, int n, float x, int k, float lo, float hi)
{
    BucketFix r;
    r.ok = 0;
    if (x >= __ldg(bins) && x < __ldg(bins + n - 1)) {          // (false for NaN)
        while (x >= hi && k < n - 2) { ++k; lo = hi; hi = __ldg(bins + k + 1); }
        while (x < lo && k > 0) { --k; hi = lo; lo = __ldg(bins + k); }
        r.ok = (x >= lo && x < hi) ? 1 : 0;
    }
    r.k = k; r.lo = lo; r.hi = hi;
    return r;
}

__device__ __forceinline__ bool bucket_valid(const float *__restrict__ bins, int n, float x, float b0, float scale,
                                             int &i, float &lo, float &hi)
{
    const float gidx = (x - b0) * scale;
    int k = (int)fminf(fmaxf(gidx, 0.f), (float)(n - 2));        // NaN -> 0
    lo = __ldg(bins + k);
    hi = __ldg(bins + k + 1);
    bool ok = x >= lo && x < hi;
    if (!ok) {
        const BucketFix f = bucket_walk(bins, n, x, k, lo, hi);
        k = f.k; lo = f.lo; hi = f.hi; ok = f.ok != 0;
    }
    i = k;
    return ok;
}

// bin_point with precomputed spacing: spacing = {b0_x, scale_x, b0_y, scale_y, b0_z, scale_z}.  Indices and
// ratios are only meaningful when ok.
__device__ __forceinline__ BinResult bin_point_fast(const float *__restrict__ bins0, int n0,
                                                    const float *__restrict__ bins1, int n1,
                                                    const float *__restrict__ bins2, int n2,
                                                    const float *__restrict__ spacing, float o0, float o1, float o2,
                                                    float r0, float r1, float r2, float d, float min_d, float max_d)
{
    BinResult b;
    const float x0 = __fadd_rn(o0, __fmul_rn(r0, d));
    const float x1 = __fadd_rn(o1, __fmul_rn(r1, d));
    const float x2 = __fadd_rn(o2, __fmul_rn(r2, d));
    float l0, h0, l1, h1, l2, h2;
    int i1;
    const bool v0 = bucket_valid(bins0, n0, x0, spacing[0], spacing[1], b.i0, l0, h0);
    const bool v1 = bucket_valid(bins1, n1, x1, spacing[2], spacing[3], i1, l1, h1);
    const bool v2 = bucket_valid(bins2, n2, x2, spacing[4], spacing[5], b.i2, l2, h2);
    b.ok = (d >= min_d) && (d <= max_d) && v0 && v1 && v2;
    b.q0 = __fdiv_rn(__fsub_rn(x0, l0), __fsub_rn(h0, l0));
    b.q1 = __fsub_rn(1.0f, __fdiv_rn(__fsub_rn(x1, l1), __fsub_rn(h1, l1)));
    b.q2 = __fdiv_rn(__fsub_rn(x2, l2), __fsub_rn(h2, l2));
    b.i1 = n1 - 2 - i1;
    return b;
}

__device__ __forceinline__ void orient(const float *__restrict__ R, float a, float b, float c, float &o0,
                                       float &o1, float &o2)
{
    o0 = __fadd_rn(__fadd_rn(__fmul_rn(a, R[0]), __fmul_rn(b, R[1])), __fmul_rn(c, R[2]));
    o1 = __fadd_rn(__fadd_rn(__fmul_rn(a, R[3]), __fmul_rn(b, R[4])), __fmul_rn(c, R[5]));
    o2 = __fadd_rn(__fadd_rn(__fmul_rn(a, R[6]), __fmul_rn(b, R[7])), __fmul_rn(c, R[8]));
}

// per-axis neighbour pair of update_feature_map (projection.py:280-291)
__device__ __forceinline__ void axis_pair(int ind, float ratio, int size, int &lo, int &hi)
{
    if (ratio < 0.5f) {
        lo = ind - 1 < 0 ? 0 : ind - 1;
        hi = ind;
    } else {
        lo = ind;
        hi = ind + 1 > size - 1 ? size - 1 : ind + 1;
    }
}

}  // namespace
