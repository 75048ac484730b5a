// Shared by profile_scan.cu (float32 / float64 rows) and filter_scan.cu (8-byte quantised rows):
// the parameter block of the averaged-profile scans and the tile geometry of the filter kernels.
#pragma once
#include "common.cuh"

#define RS_FAST_W   24
#define FT_THREADS  128
#define FT_P        9                       // windows per thread; odd => 7*P-word stride is bank-conflict free
#define FT_TILE     (FT_THREADS * FT_P)     // 1152 positions, 1152*28 B is a multiple of 16
#define FT_STAGES   3

#define EX_THREADS  128
#define EX_TILE     1024
#define EX_STAGES   2

struct ProfileParams {
    const uint8_t *codes;        // may be NULL in the exact kernel (no separators)
    const void    *profile;
    double        *dense_out;    // exact kernel, dense mode
    int64_t        n;            // rows == symbols
    int64_t        padded;       // rs_padded_count(n)
    int64_t        n_tiles;
    double         threshold;
    float          filt_thr;     // threshold - guard band (fp32, rounded down)
    int            mode;         // RS_MODE_*
    int            W;
    int            dense;
    HitStage       st;
    int            defer;        // 1: emit filter candidates only (positions); the exact rows live on the host
    int            count_on;     // quantised rows: add the letters of rows [0, count_rows) to counts8
    int64_t        pos_base;     // added to every emitted position (chunked scans)
    int64_t        count_rows;
    unsigned long long *counts8;
    float          sf[RS_MAX_W * RS_CHANNELS];   // filter table: fp32, rounded up
    double         sd[RS_MAX_W * RS_CHANNELS];   // exact structure table
    double         qd[RS_MAX_W * 4];             // exact sequence table (A,C,G,U)
};

__host__ __device__ constexpr uint32_t ru16(uint32_t x) { return (x + 15u) & ~15u; }


static inline float f32_round_up(double v)
{
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}
static inline float f32_round_down(double v)
{
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
}

// Workspace of the candidate-only scans (rs_filter_profile): staged positions, per-tile segments, ordering scratch.
struct FilterWork { int64_t off_pos, off_seg, off_scan, total; };
static inline FilterWork rs_filter_layout(int64_t n, int64_t cap)
{
    FilterWork w;
    const int64_t tiles = (n > 0 ? n : 0) / RS_MIN_TILE + 16;
    int64_t off = 0;
    w.off_pos = off;  off += rs_roundup((cap > 0 ? cap : 0) * 8, 256);
    w.off_seg = off;  off += rs_roundup(tiles * 16, 256);
    w.off_scan = off; off += rs_roundup(rs_order_tmp_bytes(tiles), 256);
    w.total = off;
    return w;
}
int rs_filter_q8_launch(const ProfileParams &prm, int W, cudaStream_t stream);   // filter_scan.cu
