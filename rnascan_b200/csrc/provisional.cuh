// Provisional log-odds on the device (see provisional.cu): shared by rs_provisional_table and the
// kernels of the two-call one-hot scans, which rebuild the small table per CTA instead of paying a launch.
#pragma once
#include "common.cuh"

#define RS_PROV_MAX_W 16          // the two-call scans cover W <= 16

struct ProvProb {                 // motif probabilities, row-major [W][A], device column order
    int W, A;
    double p[RS_PROV_MAX_W * 8];
};

// All threads of the CTA call this (contains __syncthreads).  tab: [W][TS] doubles in shared memory
// (columns >= A are zeroed), returns the margin bounding |exact - provisional| of any window score.
template <int TS>
__device__ __forceinline__ double rs_prov_table_cta(const unsigned long long *__restrict__ counts8,
                                                    const ProvProb &prm, double *tab)
{
    __shared__ double s_bg[8];
    __shared__ double s_rowmax[RS_PROV_MAX_W];
    __shared__ double s_margin;
    const int W = prm.W, A = prm.A;
    if (threadIdx.x == 0) {
        long long total = A;
        for (int c = 0; c < A; c++) total += (long long)counts8[c];
        double bg[8], norm = 0.0;
        for (int c = 0; c < A; c++) {
            bg[c] = ((double)(long long)counts8[c] + 1.0) / (double)total;      // rnascan.py:445-457
            norm += bg[c];
        }
        for (int c = 0; c < A; c++) s_bg[c] = bg[c] / norm;                      // Biopython renormalises
    }
    __syncthreads();
    // one (row, column) entry per thread: the double-precision log2 is a long dependent chain, W * A of them
    // side by side instead of A in a row per thread
    for (int e = threadIdx.x; e < W * TS; e += blockDim.x) {
        const int j = e / TS, c = e - j * TS;
        double v = 0.0;
        if (c < A) {
            const double p = prm.p[j * A + c], b = s_bg[c];
            if (b > 0) v = p > 0 ? log2(p / b) : -INFINITY;                      // p <= 0 / NaN: as motifs.log_odds
            else       v = p > 0 ? INFINITY : nan("");
        }
        tab[e] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        double rowmax = 0.0;
        for (int c = 0; c < A; c++) {
            const double v = tab[j * TS + c];
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
        s_margin = 2.0 * ldexp(m, -36);
    }
    __syncthreads();
    return s_margin;
}
