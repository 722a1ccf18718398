// Cosine similarity + best match on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), for
// instance-by-feature matrices that are a real dense contraction (thousands of rows; the SIMT kernel of instances.cu
// serves the reference's sizes).  The decision is EXACT: the tensor cores only rank candidates.
//
//   pass 1  S~ = A B^T in 3 x TF32 (a = a_hi + a_lo; hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) -> per row the
//           largest approximate cosine
//   pass 2  the same tiles again: every column whose approximate cosine is within EPS of the row's largest is
//           re-evaluated in float64 exactly as the SIMT kernel does; the first maximum wins (torch.argmax's tie rule)
//   finish  the column splits of a row are merged in column order
//
// A pre-pass splits both matrices ONCE into their TF32 high and low parts and lays them out tile by tile in the
// canonical K-major, no-swizzle UMMA layout -- 128 rows x 32 floats per tile = 16 KB contiguous: 16-byte chunks of 4
// floats, chunk c of row r at ((c * 128 + r) * 16) bytes, i.e. core matrices of 8 rows x 16 bytes, 128 bytes between
// 8-row groups (SBO), 2048 bytes between K chunks (LBO).  The contraction kernel is warp-specialised: a CTA of 6 warps
// owns a 128-row block of A and walks 128-column tiles of B;
//   warp 0 (one lane)  producer: four 16 KB bulk copies per K block (TMA engine, cp.async.bulk, SASS UBLKCP) into a
//                      3-stage shared-memory ring, completion counted by an mbarrier per stage;
//   warp 1 (one lane)  issues the MMAs (tcgen05.mma.cta_group::1.kind::tf32, 12 per K block); tcgen05.commit frees the
//                      stage and, after the last K block, hands the accumulator (one of two TMEM stages) to the epilogue;
//   warps 2-5          epilogue: thread <-> accumulator row, tcgen05.ld 4 x 32 columns, running maximum (pass 1) or
//                      exact re-evaluation of the candidates (pass 2), then the TMEM stage goes back to the MMA warp.
// There is no reference counterpart (the reference matches by L2 distance + assignment).
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int TC_M = 128, TC_N = 128, TC_BK = 32, TC_THREADS = 128;
constexpr uint32_t TC_TILE_BYTES = TC_M * TC_BK * 4;             // 16 KB per operand tile
constexpr float TC_EPS = 2e-3f;                                  // >> the 3xTF32 error of a cosine (~1e-5)
constexpr unsigned FULLM = 0xffffffffu;

constexpr int TC_STAGES = 3, TC_ACC = 2;
constexpr int TC_THREADS2 = 192;                                 // producer warp, MMA warp, 4 epilogue warps

struct __align__(1024) TcSmem {
    unsigned char tile[TC_STAGES][4][TC_TILE_BYTES];             // per stage: A hi, A lo, B hi, B lo
    uint64_t full[TC_STAGES], empty[TC_STAGES], tfull[TC_ACC], tempty[TC_ACC];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor: K-major, SWIZZLE_NONE, start address / LBO / SBO in units of 16 bytes, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((TC_M * 16u) >> 4) << 16;                    // leading byte offset: next 16-byte chunk along K
    d |= (uint64_t)(128u >> 4) << 32;                            // stride byte offset: next group of 8 rows
    d |= (uint64_t)1 << 46;                                      // descriptor version (sm_100)
    return d;
}

// instruction descriptor: D f32, A/B tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// one warp per row: float64 sum of squares (the SIMT kernel's summation order) and 1 / norm as float (0 for a zero row)
__global__ void __launch_bounds__(256)
k_row_norms(const float *__restrict__ x, int rows, int d, double *__restrict__ norm2, float *__restrict__ inv)
{
    const int lane = threadIdx.x & 31;
    const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= rows) return;
    const float *p = x + (size_t)i * d;
    double s = 0;
    for (int k = lane; k < d; k += 32) s += (double)p[k] * p[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULLM, s, o);
    if (lane == 0) {
        norm2[i] = s;
        inv[i] = s > 0 ? (float)(1.0 / sqrt(s)) : 0.f;
    }
}

// order-preserving map float -> uint32 (for atomicMax on floats of either sign)
__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct TcArgs {
    const float *a, *b;
    int n, m, d;
    const double *na2, *nb2;        // sums of squares
    const float *inva, *invb;       // 1 / norm (0 for zero rows)
    uint32_t *rowmax;               // [n] ordered-uint of the largest approximate cosine (pass 1 writes, pass 2 reads)
    double *part_sim;               // [n][splits] best exact cosine of the split (pass 2)
    int *part_arg;                  // [n][splits]
    int tiles_per_split;
    const float *atiles, *btiles;   // pre-split operand tiles: [row block][K block][hi, lo][16 KB]
};

// pre-pass: rows [rb*128, rb*128+128) x K block kb of x -> one 16 KB tile of TF32 high parts and one of low parts
// (x = hi + lo exactly; the tensor core ignores the low 13 mantissa bits of lo), zero padded past rows / d
__global__ void __launch_bounds__(128)
k_tc_split(const float *__restrict__ x, int rows, int d, int nkb, float *__restrict__ tiles)
{
    const int rb = blockIdx.x, kb = blockIdx.y, t = threadIdx.x;
    const int row = rb * TC_M + t, k0 = kb * TC_BK;
    float4 *hi_t = (float4 *)(tiles + ((size_t)(rb * nkb + kb) * 2 + 0) * (TC_TILE_BYTES / 4));
    float4 *lo_t = (float4 *)(tiles + ((size_t)(rb * nkb + kb) * 2 + 1) * (TC_TILE_BYTES / 4));
#pragma unroll
    for (int c = 0; c < TC_BK / 4; ++c) {
        float v[4] = { 0.f, 0.f, 0.f, 0.f };
        const int k = k0 + 4 * c;
        if (row < rows) {
            if (k + 4 <= d && (d & 3) == 0) {
                const float4 q = __ldg((const float4 *)(x + (size_t)row * d + k));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                for (int e = 0; e < 4; ++e)
                    if (k + e < d) v[e] = x[(size_t)row * d + k + e];
            }
        }
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
            l[e] = v[e] - h[e];
        }
        hi_t[c * TC_M + t] = make_float4(h[0], h[1], h[2], h[3]);
        lo_t[c * TC_M + t] = make_float4(l[0], l[1], l[2], l[3]);
    }
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

template <int PASS>
__global__ void __launch_bounds__(TC_THREADS2, 1)
k_cosine_tc(const TcArgs A)
{
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    TcSmem &S = *reinterpret_cast<TcSmem *>(tc_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     :: "r"(smem_u32(&S.tmem_base)), "r"((uint32_t)(TC_ACC * TC_N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
        for (int i = 0; i < TC_ACC; ++i) { mbar_init(&S.tfull[i], 1); mbar_init(&S.tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = S.tmem_base;

    const int rb = blockIdx.x;
    const int ntiles = (A.m + TC_N - 1) / TC_N;
    const int jt0 = blockIdx.y * A.tiles_per_split, jt1 = min(ntiles, jt0 + A.tiles_per_split);
    const int nkb = (A.d + TC_BK - 1) / TC_BK;
    const size_t tile_floats = TC_TILE_BYTES / 4;

    if (warp == 0) {
        // ===== producer: bulk copies of the pre-split tiles =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int jt = jt0; jt < jt1; ++jt)
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(&S.empty[stage]), phase ^ 1u);          // the MMAs have left this stage
                    mbar_expect_tx(&S.full[stage], 4u * TC_TILE_BYTES);
                    const float *ta = A.atiles + (size_t)(rb * nkb + kb) * 2 * tile_floats;
                    const float *tb = A.btiles + (size_t)(jt * nkb + kb) * 2 * tile_floats;
                    bulk_g2s(S.tile[stage][0], ta, TC_TILE_BYTES, &S.full[stage]);
                    bulk_g2s(S.tile[stage][1], ta + tile_floats, TC_TILE_BYTES, &S.full[stage]);
                    bulk_g2s(S.tile[stage][2], tb, TC_TILE_BYTES, &S.full[stage]);
                    bulk_g2s(S.tile[stage][3], tb + tile_floats, TC_TILE_BYTES, &S.full[stage]);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, aphase = 0;
            for (int jt = jt0; jt < jt1; ++jt) {
                mbar_wait(smem_u32(&S.tempty[acc]), aphase ^ 1u);              // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)(acc * TC_N);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(smem_u32(&S.full[stage]), phase);                // the tiles have landed
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t ah = smem_u32(S.tile[stage][0]), al = smem_u32(S.tile[stage][1]);
                    const uint32_t bh = smem_u32(S.tile[stage][2]), bl = smem_u32(S.tile[stage][3]);
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {                    // one MMA covers K = 8 (two chunks)
                        const uint32_t o = (uint32_t)ks * 2u * TC_M * 16u;
                        umma_tf32(d_tmem, umma_desc(ah + o), umma_desc(bh + o), (kb | ks) ? 1u : 0u);
                        umma_tf32(d_tmem, umma_desc(ah + o), umma_desc(bl + o), 1u);
                        umma_tf32(d_tmem, umma_desc(al + o), umma_desc(bh + o), 1u);
                    }
                    umma_commit(&S.empty[stage]);                               // frees the stage when the MMAs are done
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&S.tfull[acc]);                                     // the accumulator is complete
                if (++acc == TC_ACC) { acc = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: thread <-> accumulator row =====
        const int quarter = warp & 3;                                           // the TMEM lanes this warp may read
        const int row = rb * TC_M + quarter * 32 + lane;
        const bool row_ok = row < A.n;
        const float inv_na = row_ok ? A.inva[row] : 0.f;
        float best_approx = -INFINITY;
        double top = -INFINITY;
        int arg = -1;
        const float thr = (PASS == 1 && row_ok) ? ord2f(A.rowmax[row]) - TC_EPS : 0.f;
        int acc = 0;
        uint32_t aphase = 0;
        for (int jt = jt0; jt < jt1; ++jt) {
            const int col0 = jt * TC_N;
            mbar_wait(smem_u32(&S.tfull[acc]), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < TC_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TC_N + c * 32), v);
                if (row_ok) {
#pragma unroll 1
                    for (int q = 0; q < 32; ++q) {
                        const int j = col0 + c * 32 + q;
                        if (j >= A.m) break;
                        const float approx = __uint_as_float(v[q]) * inv_na * __ldg(A.invb + j);
                        if (PASS == 0) {
                            best_approx = fmaxf(best_approx, approx);
                        } else if (approx >= thr) {
                            // exact, as the SIMT kernel: float64 dot over the row pair, first maximum wins
                            const float *pa = A.a + (size_t)row * A.d, *pb = A.b + (size_t)j * A.d;
                            double dot = 0;
                            for (int k = 0; k < A.d; ++k) dot += (double)pa[k] * (double)pb[k];
                            const double den = sqrt(A.na2[row]) * sqrt(A.nb2[j]);
                            const double sim = den > 0 ? dot / den : 0.0;
                            if (sim > top) { top = sim; arg = j; }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.tempty[acc]);                         // this warp's quarter has left the accumulator
            if (++acc == TC_ACC) { acc = 0; aphase ^= 1u; }
        }
        if (row_ok) {
            if (PASS == 0) {
                if (jt1 > jt0) atomicMax(A.rowmax + row, f2ord(best_approx));
            } else {
                A.part_sim[(size_t)row * gridDim.y + blockIdx.y] = top;
                A.part_arg[(size_t)row * gridDim.y + blockIdx.y] = arg;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"((uint32_t)(TC_ACC * TC_N)) : "memory");
}

// merge the column splits of every row in column order: strictly larger wins, so the first maximum stays
__global__ void __launch_bounds__(256)
k_cosine_finish(const double *__restrict__ part_sim, const int *__restrict__ part_arg, int n, int splits, int m,
                int64_t *__restrict__ best, float *__restrict__ best_sim)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double top = -INFINITY;
    int arg = -1;
    for (int s = 0; s < splits; ++s) {
        const double v = part_sim[(size_t)i * splits + s];
        const int j = part_arg[(size_t)i * splits + s];
        if (j >= 0 && v > top) { top = v; arg = j; }
    }
    best[i] = m > 0 ? (arg >= 0 ? arg : 0) : -1;
    best_sim[i] = m > 0 ? (float)top : 0.f;
}

struct TcPlan {
    int splits, tiles_per_split;
    size_t off_na2, off_nb2, off_inva, off_invb, off_rowmax, off_psim, off_parg, off_atiles, off_btiles, total;
};

TcPlan tc_plan(int n, int m, int d)
{
    TcPlan p;
    const int row_blocks = (n + TC_M - 1) / TC_M, ntiles = (m + TC_N - 1) / TC_N;
    int splits = (MB_NUM_SMS + row_blocks - 1) / row_blocks;                 // one CTA per SM (192 KB of tile stages each)
    if (splits > ntiles) splits = ntiles;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (ntiles + splits - 1) / splits;
    p.splits = (ntiles + p.tiles_per_split - 1) / p.tiles_per_split;
    if (p.splits < 1) p.splits = 1;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = mb_align_up(o + bytes); return at; };
    p.off_na2 = take((size_t)n * sizeof(double));
    p.off_nb2 = take((size_t)m * sizeof(double));
    p.off_inva = take((size_t)n * sizeof(float));
    p.off_invb = take((size_t)m * sizeof(float));
    p.off_rowmax = take((size_t)n * sizeof(uint32_t));
    p.off_psim = take((size_t)n * p.splits * sizeof(double));
    p.off_parg = take((size_t)n * p.splits * sizeof(int));
    const size_t nkb = (size_t)(d + TC_BK - 1) / TC_BK;
    p.off_atiles = take((size_t)row_blocks * nkb * 2 * TC_TILE_BYTES);
    p.off_btiles = take((size_t)ntiles * nkb * 2 * TC_TILE_BYTES);
    p.total = o + 256;
    return p;
}

}  // namespace

size_t mbk_cosine_tc_workspace_bytes(int n, int m, int d) { return tc_plan(n, m, d).total; }

int mbk_cosine_best_match_tc(cudaStream_t stream, const float *a, int n, const float *b, int m, int d, int64_t *best,
                             float *best_sim, void *workspace, size_t workspace_bytes)
{
    if (n <= 0) return MB_OK;
    const TcPlan p = tc_plan(n, m > 0 ? m : 1, d);
    MB_REQUIRE(workspace && workspace_bytes >= p.total, "cosine best match: workspace too small");
    char *w = (char *)workspace;
    TcArgs A;
    A.a = a; A.b = b; A.n = n; A.m = m; A.d = d;
    A.na2 = (const double *)(w + p.off_na2); A.nb2 = (const double *)(w + p.off_nb2);
    A.inva = (const float *)(w + p.off_inva); A.invb = (const float *)(w + p.off_invb);
    A.rowmax = (uint32_t *)(w + p.off_rowmax);
    A.part_sim = (double *)(w + p.off_psim); A.part_arg = (int *)(w + p.off_parg);
    A.tiles_per_split = p.tiles_per_split;
    A.atiles = (const float *)(w + p.off_atiles); A.btiles = (const float *)(w + p.off_btiles);
    if (m <= 0) {
        k_cosine_finish<<<(n + 255) / 256, 256, 0, stream>>>(A.part_sim, A.part_arg, n, 0, 0, best, best_sim);
        MB_LAUNCHED();
        return MB_OK;
    }
    k_row_norms<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, stream>>>(a, n, d, (double *)(w + p.off_na2), (float *)(w + p.off_inva));
    MB_LAUNCHED();
    k_row_norms<<<(unsigned)(((size_t)m * 32 + 255) / 256), 256, 0, stream>>>(b, m, d, (double *)(w + p.off_nb2), (float *)(w + p.off_invb));
    MB_LAUNCHED();
    MB_CHECK_CUDA(cudaMemsetAsync(A.rowmax, 0, (size_t)n * sizeof(uint32_t), stream));      // ordered-uint 0 = below every float
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_cosine_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem)));
    MB_CHECK_CUDA(cudaFuncSetAttribute(k_cosine_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem)));
    const int nkb = (d + TC_BK - 1) / TC_BK;
    k_tc_split<<<dim3((n + TC_M - 1) / TC_M, nkb), 128, 0, stream>>>(a, n, d, nkb, (float *)(w + p.off_atiles));
    MB_LAUNCHED();
    k_tc_split<<<dim3((m + TC_N - 1) / TC_N, nkb), 128, 0, stream>>>(b, m, d, nkb, (float *)(w + p.off_btiles));
    MB_LAUNCHED();
    dim3 grid((n + TC_M - 1) / TC_M, p.splits);
    k_cosine_tc<0><<<grid, TC_THREADS2, sizeof(TcSmem), stream>>>(A);
    MB_LAUNCHED();
    k_cosine_tc<1><<<grid, TC_THREADS2, sizeof(TcSmem), stream>>>(A);
    MB_LAUNCHED();
    k_cosine_finish<<<(n + 255) / 256, 256, 0, stream>>>(A.part_sim, A.part_arg, n, p.splits, m, best, best_sim);
    MB_LAUNCHED();
    return MB_OK;
}
