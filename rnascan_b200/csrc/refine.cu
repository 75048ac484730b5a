// Second half of the combined-mode decision, applied to an ordered candidate list.
//
// combine() (/root/reference/rnascan/rnascan.py:416-434) is an inner join of two result sets that
// were each thresholded with the same -m: a combined hit exists iff the structure score AND the
// sequence score exceed it.  The structure side does not depend on the data's background (the
// averaged-profile mode cannot compute one, rnascan.py:533-540), the sequence side does
// (rnascan.py:507-511).  So the structure-only candidate scan (rs_scan_fused, RS_MODE_STRUCT) can run
// while the background histogram, its all-reduce and the host log-odds are still in flight on another
// stream; this kernel then scores the few candidates with the sequence PSSM exactly as _pwm.c:34-68
// (fp64 adds in j order, one cast to float, any non-ACGU symbol => NaN => no hit) and keeps the
// survivors in position order.
#include "common.cuh"

#define RF_THREADS 256
#define RF_PER     4
#define RF_TILE    (RF_THREADS * RF_PER)      // candidates per CTA

struct RefineParams {
    const uint8_t *codes;
    int64_t        n;
    const unsigned long long *n_cand;   // device: candidates found by the first pass (may exceed capacity)
    const int64_t *in_pos;
    const double  *in_str;              // may be NULL
    double         threshold;
    int            W;
    HitStage       st;
    double         qd[RS_MAX_W * 4];    // exact sequence table (A,C,G,U)
};

__global__ void __launch_bounds__(RF_THREADS) refine_seq_kernel(const __grid_constant__ RefineParams prm)
{
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    unsigned long long ncand = *prm.n_cand;
    if ((int64_t)ncand > prm.st.capacity) ncand = (unsigned long long)prm.st.capacity;
    const int64_t k0 = tile * RF_TILE + (int64_t)tid * RF_PER;
    unsigned mask = 0;
    float sq[RF_PER];
#pragma unroll
    for (int i = 0; i < RF_PER; i++) {
        sq[i] = 0.f;
        if ((unsigned long long)(k0 + i) < ncand) {
            const int64_t pos = prm.in_pos[k0 + i];
            double q;
            if (pos >= 0 && pos + prm.W <= prm.n && rs_exact_onehot_window<4, 4>(prm.codes + pos, prm.qd, prm.W, q)) {
                const float qf = (float)q;                          // _pwm.c:65
                sq[i] = qf;
                if ((double)qf > prm.threshold) mask |= 1u << i;    // SURVEY.md note N1
            }
        }
    }
    const int any = __syncthreads_or(mask != 0);
    if (any) {
        emit_tile_hits<RF_THREADS>(prm.st, tile, mask, RF_PER, [&](int i, int64_t k) {
            prm.st.pos[k] = prm.in_pos[k0 + i];
            if (prm.st.str) prm.st.str[k] = prm.in_str[k0 + i];
            prm.st.seq[k] = sq[i];
        });
    } else if (tid == 0) {
        prm.st.tile_seg[tile] = make_ulonglong2(0ull, 0ull);
    }
}

extern "C" int rs_refine_hits_seq(const uint8_t *d_codes, int64_t n, const double *seq_table, int W, double threshold,
                                  const uint64_t *d_n_candidates, int64_t hit_capacity, int64_t *d_hit_pos,
                                  float *d_hit_seq, double *d_hit_struct, uint64_t *d_counters2, void *d_work,
                                  int64_t work_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_codes || !seq_table || !d_n_candidates || !d_counters2 || n < 0) {
        rs_set_error("rs_refine_hits_seq: null argument"); return RS_ERR_INVALID;
    }
    if (W < 1 || W > RS_MAX_W) { rs_set_error("motif width %d outside [1, %d]", W, RS_MAX_W); return RS_ERR_INVALID; }
    if (threshold != threshold) { rs_set_error("threshold is NaN"); return RS_ERR_INVALID; }
    if (hit_capacity < 0 || (hit_capacity > 0 && (!d_hit_pos || !d_hit_seq))) {
        rs_set_error("bad hit buffers"); return RS_ERR_INVALID;
    }
    if ((const void *)d_n_candidates == (const void *)d_counters2) {
        rs_set_error("d_n_candidates and d_counters2 must be different buffers"); return RS_ERR_INVALID;
    }
    RS_CUDA(cudaMemsetAsync(d_counters2, 0, 2 * sizeof(uint64_t), st));
    if (hit_capacity == 0) return RS_OK;
    const int64_t n_tiles = (hit_capacity + RF_TILE - 1) / RF_TILE;
    if (n_tiles > n / RS_MIN_TILE + 16) {
        rs_set_error("hit_capacity %lld is out of proportion to the stream length %lld", (long long)hit_capacity, (long long)n);
        return RS_ERR_INVALID;
    }
    WorkLayout wl = rs_work_layout(n, hit_capacity);
    if (!d_work || work_bytes < wl.total) { rs_set_error("workspace too small: need %lld bytes", (long long)wl.total); return RS_ERR_WORKSPACE; }

    RefineParams prm = {};
    prm.codes = d_codes; prm.n = n; prm.n_cand = (const unsigned long long *)d_n_candidates;
    prm.in_pos = d_hit_pos; prm.in_str = d_hit_struct; prm.threshold = threshold; prm.W = W;
    uint8_t *wk = (uint8_t *)d_work;
    prm.st.pos = (int64_t *)(wk + wl.off_pos);
    prm.st.seq = (float *)(wk + wl.off_seq);
    prm.st.str = d_hit_struct ? (double *)(wk + wl.off_str) : nullptr;
    prm.st.tile_seg = (ulonglong2 *)(wk + wl.off_seg);
    prm.st.counters = (unsigned long long *)d_counters2;
    prm.st.capacity = hit_capacity;
    for (int k = 0; k < W * 4; k++) prm.qd[k] = seq_table[k];
    refine_seq_kernel<<<(unsigned)n_tiles, RF_THREADS, 0, st>>>(prm);
    RS_CUDA(cudaGetLastError());
    OrderDest od = {d_hit_pos, d_hit_seq, d_hit_struct, nullptr, nullptr, 0};
    return rs_order_hits(prm.st, n_tiles, od, wk + wl.off_scan, st);
}

// ------------------------------------------------------------------------------------------------
// The same decision for candidates that carry their window's symbols with them (rs_filter_profile on 4-bit rows:
// symbol j in bits 2j, 2j+1 of the 64-bit payload, bit 63 = a symbol that is not A,C,G,U): the symbol stream is
// not on the device at all in that pipeline.  Survivors stay in order; their positions come back.
struct RefinePackedParams {
    const int64_t *in_pos;
    const unsigned long long *in_sym;
    int64_t        n_cand;
    double         threshold;
    int            W;
    HitStage       st;
    double         qd[RS_MAX_W * 4];
};

__global__ void __launch_bounds__(RF_THREADS) refine_packed_kernel(const __grid_constant__ RefinePackedParams prm)
{
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t k0 = tile * RF_TILE + (int64_t)tid * RF_PER;
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < RF_PER; i++) {
        if (k0 + i < prm.n_cand) {
            const unsigned long long sym = prm.in_sym[k0 + i];
            if (!(sym >> 63)) {
                double q = 0.0;
                for (int j = 0; j < prm.W; j++) q = __dadd_rn(q, prm.qd[j * 4 + (int)((sym >> (2 * j)) & 3ull)]);
                if ((double)(float)q > prm.threshold) mask |= 1u << i;      // _pwm.c:65 + SURVEY.md note N1
            }
        }
    }
    const int any = __syncthreads_or(mask != 0);
    if (any) {
        emit_tile_hits<RF_THREADS>(prm.st, tile, mask, RF_PER, [&](int i, int64_t k) { prm.st.pos[k] = prm.in_pos[k0 + i]; });
    } else if (tid == 0) {
        prm.st.tile_seg[tile] = make_ulonglong2(0ull, 0ull);
    }
}

extern "C" int64_t rs_refine_packed_workspace_bytes(int64_t n_cand)
{
    const int64_t cap = n_cand > 0 ? n_cand : 0, tiles = cap / RF_TILE + 2;
    return rs_roundup(cap * 8, 256) + rs_roundup(tiles * 16, 256) + rs_roundup(rs_order_tmp_bytes(tiles), 256);
}

extern "C" int rs_refine_candidates_packed(const int64_t *d_cand_pos, const uint64_t *d_cand_sym, int64_t n_cand,
                                           const double *seq_table, int W, double threshold, int64_t *d_out_pos,
                                           uint64_t *d_counters2, void *d_work, int64_t work_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cand < 0 || !seq_table || !d_counters2) { rs_set_error("rs_refine_candidates_packed: bad argument"); return RS_ERR_INVALID; }
    if (W < 1 || W > 31) { rs_set_error("packed symbols hold motif widths 1..31"); return RS_ERR_INVALID; }
    if (threshold != threshold) { rs_set_error("threshold is NaN"); return RS_ERR_INVALID; }
    RS_CUDA(cudaMemsetAsync(d_counters2, 0, 2 * sizeof(uint64_t), st));
    if (n_cand == 0) return RS_OK;
    if (!d_cand_pos || !d_cand_sym || !d_out_pos) { rs_set_error("rs_refine_candidates_packed: null buffer"); return RS_ERR_INVALID; }
    if (!d_work || work_bytes < rs_refine_packed_workspace_bytes(n_cand)) {
        rs_set_error("workspace too small: need %lld bytes", (long long)rs_refine_packed_workspace_bytes(n_cand));
        return RS_ERR_WORKSPACE;
    }
    const int64_t n_tiles = (n_cand + RF_TILE - 1) / RF_TILE, tiles = n_cand / RF_TILE + 2;
    RefinePackedParams prm = {};
    prm.in_pos = d_cand_pos; prm.in_sym = (const unsigned long long *)d_cand_sym; prm.n_cand = n_cand;
    prm.threshold = threshold; prm.W = W;
    uint8_t *wk = (uint8_t *)d_work;
    prm.st.pos = (int64_t *)wk;
    prm.st.tile_seg = (ulonglong2 *)(wk + rs_roundup(n_cand * 8, 256));
    prm.st.counters = (unsigned long long *)d_counters2;
    prm.st.capacity = n_cand;
    for (int k = 0; k < W * 4; k++) prm.qd[k] = seq_table[k];
    refine_packed_kernel<<<(unsigned)n_tiles, RF_THREADS, 0, st>>>(prm);
    RS_CUDA(cudaGetLastError());
    OrderDest od = {d_out_pos, nullptr, nullptr, nullptr, nullptr, 0};
    return rs_order_hits(prm.st, n_tiles, od, wk + rs_roundup(n_cand * 8, 256) + rs_roundup(tiles * 16, 256), st);
}
