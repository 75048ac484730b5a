// Exact decision for an ordered list of candidate windows whose rows were gathered on the host.
//
// Second half of the filter + gather + resolve scans (rs_filter_profile): the float64 rows the reference
// reads with pd.read_table (/root/reference/rnascan/rnascan.py:296-297) never travel to the device in
// bulk; the filter runs on a float32 or 8-byte quantised shadow, the host gathers the W rows (and W
// symbols) of every candidate window, and this kernel scores them exactly as the reference does:
//   structure  rnascan.py:302-307   sum_j nan_to_num(dot(profile[i+j], pssm[j]))  (float64, ddot order)
//   sequence   _pwm.c:34-68         float64 adds in j order, one cast to float32
//   decision   strict `>` on both (Biopython search / rnascan.py:308; combine() = AND, rnascan.py:416-434)
// Survivors keep their order (tile staging + rs_order_hits, as every other scan here).
#include "common.cuh"

#define RV_THREADS 256
#define RV_PER     4
#define RV_TILE    (RV_THREADS * RV_PER)

struct ResolveParams {
    const int64_t *cand_pos;
    const void    *rows;         // [n_cand][W][7] float32 | float64
    const uint8_t *codes;        // [n_cand][W]
    int64_t        n_cand;
    double         threshold;
    int            W;
    int            mode;
    int            rows_f64;
    HitStage       st;
    double         sd[RS_MAX_W * RS_CHANNELS];
    double         qd[RS_MAX_W * 4];
};

template <typename PT>
__device__ __forceinline__ bool resolve_one(const ResolveParams &prm, int64_t k, float &sq, double &sc)
{
    const int W = prm.W;
    const PT *rows = reinterpret_cast<const PT *>(prm.rows) + (size_t)k * W * RS_CHANNELS;
    const uint8_t *codes = prm.codes + (size_t)k * W;
    sc = rs_exact_profile_window<PT>(rows, prm.sd, W);
    sq = 0.f;
    if (!(sc > prm.threshold)) return false;
    if (prm.mode == RS_MODE_AND) {
        double q;
        if (!rs_exact_onehot_window<4, 4>(codes, prm.qd, W, q)) return false;
        sq = (float)q;                                   // _pwm.c:65
        return (double)sq > prm.threshold;               // SURVEY.md note N1
    }
    return rs_no_separator(codes, W);
}

__global__ void __launch_bounds__(RV_THREADS) resolve_kernel(const __grid_constant__ ResolveParams prm)
{
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t k0 = tile * RV_TILE + (int64_t)tid * RV_PER;
    unsigned mask = 0;
    float sq[RV_PER];
    double sc[RV_PER];
#pragma unroll
    for (int i = 0; i < RV_PER; i++) {
        sq[i] = 0.f; sc[i] = 0.0;
        if (k0 + i < prm.n_cand) {
            const bool hit = prm.rows_f64 ? resolve_one<double>(prm, k0 + i, sq[i], sc[i])
                                          : resolve_one<float>(prm, k0 + i, sq[i], sc[i]);
            if (hit) mask |= 1u << i;
        }
    }
    const int any = __syncthreads_or(mask != 0);
    if (any) {
        emit_tile_hits<RV_THREADS>(prm.st, tile, mask, RV_PER, [&](int i, int64_t k) {
            prm.st.pos[k] = prm.cand_pos[k0 + i];
            prm.st.str[k] = sc[i];
            if (prm.st.seq) prm.st.seq[k] = sq[i];
        });
    } else if (tid == 0) {
        prm.st.tile_seg[tile] = make_ulonglong2(0ull, 0ull);
    }
}

struct ResolveWork { int64_t off_pos, off_str, off_seq, off_seg, off_scan, total; };
static ResolveWork resolve_layout(int64_t n_cand)
{
    ResolveWork w;
    const int64_t cap = n_cand > 0 ? n_cand : 0;
    const int64_t tiles = cap / RV_TILE + 2;
    int64_t off = 0;
    w.off_pos = off;  off += rs_roundup(cap * 8, 256);
    w.off_str = off;  off += rs_roundup(cap * 8, 256);
    w.off_seq = off;  off += rs_roundup(cap * 4, 256);
    w.off_seg = off;  off += rs_roundup(tiles * 16, 256);
    w.off_scan = off; off += rs_roundup(rs_order_tmp_bytes(tiles), 256);
    w.total = off;
    return w;
}

extern "C" int64_t rs_resolve_workspace_bytes(int64_t n_cand) { return resolve_layout(n_cand).total; }

extern "C" int rs_resolve_candidates(const int64_t *d_cand_pos, int64_t n_cand, const void *d_win_rows, int rows_dtype,
                                     const uint8_t *d_win_codes, const double *seq_table, const double *struct_table,
                                     int W, double threshold, int64_t *d_hit_pos, float *d_hit_seq,
                                     double *d_hit_struct, uint64_t *d_counters2, void *d_work, int64_t work_bytes,
                                     void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cand < 0 || !struct_table || !d_counters2) { rs_set_error("rs_resolve_candidates: bad argument"); return RS_ERR_INVALID; }
    if (rows_dtype != RS_F32 && rows_dtype != RS_F64) { rs_set_error("rows_dtype must be RS_F32 or RS_F64"); return RS_ERR_INVALID; }
    if (W < 1 || W > RS_MAX_W) { rs_set_error("motif width %d outside [1, %d]", W, RS_MAX_W); return RS_ERR_INVALID; }
    if (threshold != threshold) { rs_set_error("threshold is NaN"); return RS_ERR_INVALID; }
    RS_CUDA(cudaMemsetAsync(d_counters2, 0, 2 * sizeof(uint64_t), st));
    if (n_cand == 0) return RS_OK;
    if (!d_cand_pos || !d_win_rows || !d_win_codes || !d_hit_pos || !d_hit_struct || (seq_table && !d_hit_seq)) {
        rs_set_error("rs_resolve_candidates: null buffer"); return RS_ERR_INVALID;
    }
    const ResolveWork wl = resolve_layout(n_cand);
    if (!d_work || work_bytes < wl.total) { rs_set_error("workspace too small: need %lld bytes", (long long)wl.total); return RS_ERR_WORKSPACE; }

    ResolveParams prm = {};
    prm.cand_pos = d_cand_pos; prm.rows = d_win_rows; prm.codes = d_win_codes; prm.n_cand = n_cand;
    prm.threshold = threshold; prm.W = W; prm.mode = seq_table ? RS_MODE_AND : RS_MODE_STRUCT;
    prm.rows_f64 = rows_dtype == RS_F64;
    uint8_t *wk = (uint8_t *)d_work;
    prm.st.pos = (int64_t *)(wk + wl.off_pos);
    prm.st.str = (double *)(wk + wl.off_str);
    prm.st.seq = seq_table ? (float *)(wk + wl.off_seq) : nullptr;
    prm.st.tile_seg = (ulonglong2 *)(wk + wl.off_seg);
    prm.st.counters = (unsigned long long *)d_counters2;
    prm.st.capacity = n_cand;
    for (int k = 0; k < W * RS_CHANNELS; k++) prm.sd[k] = struct_table[k];
    if (seq_table) for (int k = 0; k < W * 4; k++) prm.qd[k] = seq_table[k];
    const int64_t n_tiles = (n_cand + RV_TILE - 1) / RV_TILE;
    resolve_kernel<<<(unsigned)n_tiles, RV_THREADS, 0, st>>>(prm);
    RS_CUDA(cudaGetLastError());
    OrderDest od = {d_hit_pos, seq_table ? d_hit_seq : nullptr, d_hit_struct, nullptr, nullptr, 0};
    return rs_order_hits(prm.st, n_tiles, od, wk + wl.off_scan, st);
}
