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
#include "provisional.cuh"

// out[0 .. W*A) = table (row-major [W][A], device column order), out[W*A] = margin
__global__ void provisional_table_kernel(const unsigned long long *__restrict__ counts8,
                                         const __grid_constant__ ProvProb prm, double *__restrict__ out)
{
    __shared__ double s_tab[RS_PROV_MAX_W * 8];
    const double margin = rs_prov_table_cta<8>(counts8, prm, s_tab);
    for (int k = threadIdx.x; k < prm.W * prm.A; k += blockDim.x) out[k] = s_tab[(k / prm.A) * 8 + k % prm.A];
    if (threadIdx.x == 0) out[prm.W * prm.A] = margin;
}

extern "C" int rs_provisional_table(const uint64_t *d_counts8, const double *prob, int W, int alphabet,
                                    double *d_table_margin, void *stream)
{
    if (!d_counts8 || !prob || !d_table_margin) { rs_set_error("rs_provisional_table: null argument"); return RS_ERR_INVALID; }
    if (alphabet != 4 && alphabet != 7) { rs_set_error("alphabet must be 4 (A,C,G,U) or 7 (B,E,H,L,M,R,T)"); return RS_ERR_INVALID; }
    if (W < 1 || W > RS_PROV_MAX_W) { rs_set_error("motif width %d outside [1, %d]", W, RS_PROV_MAX_W); return RS_ERR_INVALID; }
    ProvProb prm = {};
    prm.W = W; prm.A = alphabet;
    for (int k = 0; k < W * alphabet; k++) prm.p[k] = prob[k];
    provisional_table_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_counts8, prm,
                                                                 d_table_margin);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
