// Put staged hits into position order in ONE launch: scan kernels stage each tile's hits (already in
// position order) at an atomically claimed offset and record (offset, count) per tile; here one thread
// per tile takes part in an exclusive prefix sum of the counts (tile order == position order) and
// copies its segment to its final place.
//
// Single pass with look-back: a CTA takes a ticket (its virtual index, so a CTA only ever waits for
// CTAs that already run), scans its 1024 tile counts, publishes its aggregate and sums the aggregates
// of all CTAs before it (read in parallel by its threads; n_ctas is ~100 per 10^8 symbols).
#include "common.cuh"

#define OR_THREADS 1024                       // tiles per CTA (one per thread)

struct OrderTmp {
    unsigned long long ticket;
    unsigned long long pad;
    unsigned long long agg[1];                // [n_ctas]: 0 = not ready, else aggregate + 1
};

// Optional sequence check folded into the pass (rs_scan_fused_resolve): the staged entries of a tile
// are structure-only candidates; those whose sequence score (_pwm.c:34-68 arithmetic) also exceeds the
// threshold are compacted to the front of the tile's segment before the counts are summed.
struct SeqRefine {
    const uint8_t *codes;
    int64_t n;
    double threshold;
    int W;
    unsigned long long *total_out;            // receives the number of survivors
    double qd[RS_MAX_W * 4];
};

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long &total)
{
    __shared__ unsigned long long s_w[OR_THREADS / 32];
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();                          // s_w / s_total may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long x = s_w[lane], xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        s_w[lane] = xi - x;                   // exclusive warp offsets
        if (lane == 31) s_total = xi;
    }
    __syncthreads();
    total = s_total;
    return s_w[warp] + incl - v;
}

template <bool REFINE>
__global__ void __launch_bounds__(OR_THREADS) order_kernel(HitStage st, int64_t n_tiles, OrderTmp *tmp, int n_ctas,
                                                           OrderDest od, const __grid_constant__ SeqRefine rf)
{
    __shared__ unsigned s_vb;
    if (threadIdx.x == 0) s_vb = (unsigned)atomicAdd(&tmp->ticket, 1ull);
    __syncthreads();
    const unsigned vb = s_vb;
    const int64_t tile = (int64_t)vb * OR_THREADS + threadIdx.x;
    ulonglong2 sg = tile < n_tiles ? st.tile_seg[tile] : make_ulonglong2(0ull, 0ull);
    if (REFINE && sg.y) {
        unsigned long long kept = 0;
        for (unsigned long long h = 0; h < sg.y; h++) {
            const unsigned long long from = sg.x + h;
            if ((int64_t)from >= st.capacity) break;                   // dropped by the first pass (overflow)
            const int64_t pos = st.pos[from];
            double q;
            if (pos + rf.W <= rf.n && rs_exact_onehot_window<4, 4>(rf.codes + pos, rf.qd, rf.W, q)) {
                const float qf = (float)q;                              // _pwm.c:65
                if ((double)qf > rf.threshold) {                        // SURVEY.md note N1
                    const unsigned long long to = sg.x + kept;
                    st.pos[to] = pos;
                    st.str[to] = st.str[from];
                    st.seq[to] = qf;
                    kept++;
                }
            }
        }
        sg.y = kept;
    }
    unsigned long long total;
    const unsigned long long excl = block_exclusive_scan(sg.y, total);
    if (threadIdx.x == 0) {
        __threadfence();
        ((volatile unsigned long long *)tmp->agg)[vb] = total + 1ull;
    }
    // look back: sum of the aggregates of CTAs [0, vb)
    unsigned long long part = 0;
    for (unsigned p = threadIdx.x; p < vb; p += OR_THREADS) {
        unsigned long long v;
        unsigned spins = 0;
        unsigned long long t_start = 0ull;
        while ((v = ((volatile unsigned long long *)tmp->agg)[p]) == 0ull) {
            __nanosleep(20);
            if ((++spins & 0xFFFFu) == 0u && rs_spin_expired(t_start)) __trap();   // a minute without the predecessor's aggregate
        }
        part += v - 1ull;
    }
    unsigned long long before;
    block_exclusive_scan(part, before);       // `before` = block-wide sum
    if (vb == (unsigned)(n_ctas - 1) && threadIdx.x == 0) {
        tmp->ticket = 0ull;                   // every CTA holds its ticket by now
        if (REFINE) *rf.total_out = before + total;
    }
    // copy this tile's segment to its final place
    const unsigned long long out_base = od.out_base ? *od.out_base : 0ull;
    const unsigned long long dst = out_base + before + excl;
    for (unsigned long long h = 0; h < sg.y; h++) {
        const unsigned long long from = sg.x + h, to = dst + h;
        if ((int64_t)from < st.capacity && (int64_t)to < st.capacity) {
            od.pos[to] = st.pos[from];
            if (od.seq && st.seq) od.seq[to] = st.seq[from];
            if (od.str && st.str) od.str[to] = st.str[from];
            if (od.out_motif) od.out_motif[to] = od.motif_id;
        }
    }
}

int64_t rs_order_tmp_bytes(int64_t max_tiles)
{
    return 16 + 8 * (max_tiles / OR_THREADS + 2);
}

static int order_impl(const HitStage &st, int64_t n_tiles, const OrderDest &dst, void *d_scan_tmp, cudaStream_t stream,
                      const SeqRefine *rf)
{
    if (n_tiles <= 0) return RS_OK;
    const int n_ctas = (int)((n_tiles + OR_THREADS - 1) / OR_THREADS);
    OrderTmp *tmp = (OrderTmp *)d_scan_tmp;
    RS_CUDA(cudaMemsetAsync(tmp, 0, (size_t)(2 + n_ctas) * sizeof(unsigned long long), stream));
    if (rf) {
        order_kernel<true><<<n_ctas, OR_THREADS, 0, stream>>>(st, n_tiles, tmp, n_ctas, dst, *rf);
    } else {
        SeqRefine none = {};
        order_kernel<false><<<n_ctas, OR_THREADS, 0, stream>>>(st, n_tiles, tmp, n_ctas, dst, none);
    }
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int rs_order_hits(const HitStage &st, int64_t n_tiles, const OrderDest &dst, void *d_scan_tmp, cudaStream_t stream)
{
    return order_impl(st, n_tiles, dst, d_scan_tmp, stream, nullptr);
}

int rs_order_hits_seq_refined(const HitStage &st, int64_t n_tiles, const OrderDest &dst, void *d_scan_tmp,
                              cudaStream_t stream, const uint8_t *d_codes, int64_t n, const double *seq_table, int W,
                              double threshold, unsigned long long *d_total_out)
{
    SeqRefine rf = {};
    rf.codes = d_codes; rf.n = n; rf.threshold = threshold; rf.W = W; rf.total_out = d_total_out;
    for (int k = 0; k < W * 4; k++) rf.qd[k] = seq_table[k];
    if (n_tiles <= 0) return cudaMemsetAsync(d_total_out, 0, 8, stream) == cudaSuccess ? RS_OK : RS_ERR_CUDA;
    return order_impl(st, n_tiles, dst, d_scan_tmp, stream, &rf);
}
