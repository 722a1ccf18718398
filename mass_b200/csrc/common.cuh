// Shared host/device helpers of libmassb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/massb200.h"

#define MB_API extern "C" __attribute__((visibility("default")))

void mb_set_error(const char *fmt, ...);
void mb_count_launch(int n = 1);

#define MB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t mb_e_ = (expr);                                                           \
        if (mb_e_ != cudaSuccess) {                                                           \
            mb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                   \
                         cudaGetErrorString(mb_e_));                                          \
            return MB_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// after every kernel launch: picks up launch-configuration errors and counts the launch
#define MB_LAUNCHED()                                                                         \
    do {                                                                                      \
        mb_count_launch();                                                                    \
        MB_CHECK_CUDA(cudaGetLastError());                                                    \
    } while (0)

#define MB_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            mb_set_error(__VA_ARGS__);                                                        \
            return MB_ERR_ARG;                                                                \
        }                                                                                     \
    } while (0)

static inline size_t mb_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct MbArena {
    char *base;
    size_t size, used;
    MbArena(void *p, size_t n) : base((char *)p), size(n), used(0) {}
    template <typename T> T *take(size_t count)
    {
        size_t off = mb_align_up(used);
        used = off + count * sizeof(T);
        return (T *)(base ? base + off : nullptr);
    }
    bool ok() const { return used <= size; }
};

constexpr int MB_NUM_SMS = 148;          // B200: 2 dies x 74 SMs
constexpr int MB_BRICK = 8;              // voxel bricks are 8 x 8 x 8
constexpr int MB_BRICK_VOX = 512;

// Geometry of the voxel map and of its brick-major key space.
//   key = brick_id * 512 + ((i0 & 7) << 6 | (i1 & 7) << 3 | (i2 & 7)),
//   brick_id = ((i0 >> 3) * B1 + (i1 >> 3)) * B2 + (i2 >> 3)
// Brick-major keys keep the voxels a camera frustum hits together close in the sorted
// order, so consecutive voxel segments re-use the same feature rows out of L1/L2.
struct MbGrid {
    int S0, S1, S2;   // map dims (y flipped, x, z)
    int B0, B1, B2;   // bricks per axis
    uint32_t invalid; // first key past the last brick: marks contributions of invalid pixels
};

static inline MbGrid mb_make_grid(int S0, int S1, int S2)
{
    MbGrid g;
    g.S0 = S0; g.S1 = S1; g.S2 = S2;
    g.B0 = (S0 + 7) / 8; g.B1 = (S1 + 7) / 8; g.B2 = (S2 + 7) / 8;
    g.invalid = (uint32_t)((uint64_t)g.B0 * g.B1 * g.B2 * MB_BRICK_VOX);
    return g;
}

static inline int mb_key_bits(const MbGrid &g)
{
    int bits = 1;
    while (bits < 32 && (((uint64_t)1) << bits) <= (uint64_t)g.invalid) ++bits;
    return bits;
}

__device__ __forceinline__ uint32_t mb_voxel_key(const MbGrid &g, int i0, int i1, int i2)
{
    uint32_t brick = (uint32_t)(((i0 >> 3) * g.B1 + (i1 >> 3)) * g.B2 + (i2 >> 3));
    return brick * MB_BRICK_VOX + (uint32_t)(((i0 & 7) << 6) | ((i1 & 7) << 3) | (i2 & 7));
}

__device__ __forceinline__ size_t mb_key_to_voxel(const MbGrid &g, uint32_t key)
{
    uint32_t brick = key >> 9, local = key & 511;
    int b2 = brick % g.B2;
    uint32_t t = brick / g.B2;
    int b1 = t % g.B1, b0 = t / g.B1;
    int i0 = b0 * 8 + (local >> 6), i1 = b1 * 8 + ((local >> 3) & 7), i2 = b2 * 8 + (local & 7);
    return ((size_t)i0 * g.S1 + i1) * g.S2 + i2;
}

// ---- internal stage interfaces (defined in the .cu files) ----------------------------------

// radix_sort.cu: stable LSD radix sort of (key, value) u32 pairs on bits [0, key_bits).
// keys_a/vals_a hold the input; the sorted result lands in *keys_out / *vals_out (one of a/b).
// If vals_a_is_iota the first pass synthesises vals = 0..n-1 instead of reading them.
size_t mb_sort_workspace_bytes(uint32_t n);
// n is the host-side element count (or an upper bound of it); if n_dev is not null the device
// value min(n, *n_dev) is the real count.
int mb_sort_pairs(cudaStream_t stream, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b,
                  uint32_t *vals_b, uint32_t n, const uint32_t *n_dev, int key_bits, bool vals_a_is_iota,
                  void *workspace, size_t workspace_bytes, uint32_t **keys_out, uint32_t **vals_out);

// one-launch exclusive scan of n <= 2^20 values whose sum stays below 2^30 (in place allowed); zeroed_state:
// mb_scan_state_words(n) words the caller has zeroed
size_t mb_scan_state_words(uint32_t n);
int mb_exclusive_scan_small(cudaStream_t stream, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *zeroed_state);
// exclusive prefix sum of n u32 values (in place allowed)
size_t mb_scan_workspace_bytes(uint32_t n);
int mb_exclusive_scan_u32(cudaStream_t stream, const uint32_t *in, uint32_t *out, uint32_t n,
                          void *workspace, size_t workspace_bytes);
