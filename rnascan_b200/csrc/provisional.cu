// Provisional log-odds tables computed ON THE DEVICE from device-resident background counts.
//
// The reference turns counts into a background (rnascan.py:445-457: p = (count + 1) / (sum + |A|)) and
// the background into log-odds (Biopython log_odds called at rnascan.py:248: background renormalised,
// math.log(p / b, 2)).  The exact table needs the host's libm (bit-identical to Python's math.log), which
// puts a device -> host -> device round trip between the histogram and the scan.  A scan can start
// without it from THIS table: the same formulas in IEEE double arithmetic with the device's log2
// (<= 1 ulp), so every entry is within a few ulps of the exact one, plus a margin that bounds the
// difference of any W-term window score.  Scans that use it (rs_scan_onehot_begin) only select
// CANDIDATES (provisional score + margin > threshold, a superset of the exact hits); the exact host
// table decides and scores afterwards (rs_scan_onehot_finish).
#include "common.cuh"

struct ProvParams {
    int W, A;
    double prob[RS_MAX_W * RS_CHANNELS];
};

// out[0 .. W*A) = table (row-major [W][A], device column order), out[W*A] = margin
__global__ void provisional_table_kernel(const unsigned long long *__restrict__ counts8,
                                         const __grid_constant__ ProvParams prm, double *__restrict__ out)
{
    __shared__ double s_bg[8];
    __shared__ double s_rowmax[RS_MAX_W];
    const int W = prm.W, A = prm.A;
    if (threadIdx.x == 0) {
        long long total = A;
        for (int c = 0; c < A; c++) total += (long long)counts8[c];
        double bg[8], norm = 0.0;
        for (int c = 0; c < A; c++) {
            bg[c] = ((double)(long long)counts8[c] + 1.0) / (double)total;
            norm += bg[c];
        }
        for (int c = 0; c < A; c++) s_bg[c] = bg[c] / norm;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        double rowmax = 0.0;
        for (int c = 0; c < A; c++) {
            const double p = prm.prob[j * A + c], b = s_bg[c];
            double v;
            if (b > 0) v = p > 0 ? log2(p / b) : -INFINITY;          // p <= 0 / NaN: as motifs.log_odds
            else       v = p > 0 ? INFINITY : nan("");
            out[j * A + c] = v;
            if (isfinite(v)) rowmax = fmax(rowmax, fabs(v));
        }
        s_rowmax[j] = rowmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // per entry |exact - provisional| <= 2^-36 (1 + |t|): the true gap (a few ulps of t from the two
        // log implementations, ~1 ulp of p/b from the summation order of the renormalisation) is four
        // orders of magnitude smaller.  Summed over the W rows of a window, doubled for the roundings of
        // the two W-term sums themselves.
        double m = 0.0;
        for (int j = 0; j < W; j++) m += 1.0 + s_rowmax[j];
        out[W * A] = 2.0 * ldexp(m, -36);
    }
}

extern "C" int rs_provisional_table(const uint64_t *d_counts8, const double *prob, int W, int alphabet,
                                    double *d_table_margin, void *stream)
{
    if (!d_counts8 || !prob || !d_table_margin) { rs_set_error("rs_provisional_table: null argument"); return RS_ERR_INVALID; }
    if (alphabet != 4 && alphabet != 7) { rs_set_error("alphabet must be 4 (A,C,G,U) or 7 (B,E,H,L,M,R,T)"); return RS_ERR_INVALID; }
    if (W < 1 || W > RS_MAX_W) { rs_set_error("motif width %d outside [1, %d]", W, RS_MAX_W); return RS_ERR_INVALID; }
    ProvParams prm = {};
    prm.W = W; prm.A = alphabet;
    for (int k = 0; k < W * alphabet; k++) prm.prob[k] = prob[k];
    provisional_table_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_counts8, prm,
                                                                 d_table_margin);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
