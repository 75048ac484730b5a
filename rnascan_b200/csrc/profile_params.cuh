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

// Block-cooperative handling of the windows of one tile that passed the fp32 filter.
//
// The filter leaves every thread with a bit mask of its own P consecutive windows.  Evaluating those windows
// where they are found serialises a warp behind each candidate (fp64, ~50 dependent operations); at low
// thresholds, where most warps hold one, the kernel then runs at 1/32 of its width.  Instead the candidates
// of the whole tile are compacted, in position order, into a shared queue; the CTA's threads take one
// candidate each (`eval(window) -> is it a hit`), the hits are compacted in place (still in order), ONE atomic
// claims their slice of the staging area and `emit(window, slot)` writes them -- all THREADS wide.
template <int THREADS, int P, typename EVAL, typename EMIT>
__device__ __forceinline__ void resolve_tile_candidates(const HitStage &st, int64_t tile, unsigned candmask,
                                                        EVAL eval, EMIT emit)
{
    __shared__ uint16_t s_queue[THREADS * P];
    __shared__ unsigned s_warp[THREADS / 32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- candidates -> queue, position order (thread-major, windows of a thread are consecutive)
    const unsigned cnt = __popc(candmask);
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = 0, n_cand = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        const unsigned v = s_warp[w];
        if (w < warp) before += v;
        n_cand += v;
    }
    unsigned slot = before + incl - cnt;
#pragma unroll
    for (int i = 0; i < P; i++)
        if (candmask & (1u << i)) s_queue[slot++] = (uint16_t)(tid * P + i);
    if (tid == 0) atomicAdd(st.counters + 1, (unsigned long long)n_cand);
    __syncthreads();
    // ---- evaluate THREADS candidates per round, compact the hits in place
    unsigned n_hits = 0;
    for (unsigned r0 = 0; r0 < n_cand; r0 += THREADS) {
        const unsigned q = r0 + tid;
        const int w = q < n_cand ? (int)s_queue[q] : 0;
        const bool hit = q < n_cand && eval(w);
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();                         // every queue entry of this round has been read
        unsigned rank = __popc(ballot & ((1u << lane) - 1u)), round_hits = 0;
#pragma unroll
        for (int k = 0; k < THREADS / 32; k++) {
            const unsigned v = s_warp[k];
            if (k < warp) rank += v;
            round_hits += v;
        }
        if (hit) s_queue[n_hits + rank] = (uint16_t)w;
        n_hits += round_hits;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long base = n_hits ? atomicAdd(st.counters, (unsigned long long)n_hits) : 0ull;
        s_base = base;
        st.tile_seg[tile] = make_ulonglong2(base, (unsigned long long)n_hits);
    }
    __syncthreads();
    const unsigned long long base = s_base;
    for (unsigned k = tid; k < n_hits; k += THREADS)
        if ((int64_t)(base + k) < st.capacity) emit((int)s_queue[k], (int64_t)(base + k));
    __syncthreads();                             // the staged tile stays in use until every hit is written
}

// Workspace of the candidate-only scans (rs_filter_profile): staged positions, per-tile segments, ordering scratch.
struct FilterWork { int64_t off_pos, off_sym, off_seg, off_scan, total; };
static inline FilterWork rs_filter_layout(int64_t n, int64_t cap)
{
    FilterWork w;
    const int64_t tiles = (n > 0 ? n : 0) / RS_MIN_TILE + 16;
    int64_t off = 0;
    w.off_pos = off;  off += rs_roundup((cap > 0 ? cap : 0) * 8, 256);
    w.off_sym = off;  off += rs_roundup((cap > 0 ? cap : 0) * 8, 256);      // candidates' packed symbols (4-bit rows)
    w.off_seg = off;  off += rs_roundup(tiles * 16, 256);
    w.off_scan = off; off += rs_roundup(rs_order_tmp_bytes(tiles), 256);
    w.total = off;
    return w;
}
int rs_filter_q8_launch(const ProfileParams &prm, int W, cudaStream_t stream);   // filter_scan.cu
int rs_filter_q4_launch(const ProfileParams &prm, int W, cudaStream_t stream);   // filter_scan.cu
