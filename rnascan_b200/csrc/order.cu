// Put staged hits into position order: exclusive prefix sum of the per-tile hit counts
// (tile order == position order) followed by a segmented copy.  Two small launches.
#include "common.cuh"

#define OR_THREADS 1024
#define OR_PER     8
#define OR_CHUNK   (OR_THREADS * OR_PER)      // tiles per CTA

struct OrderTmp {
    unsigned long long ticket;
    unsigned long long pad;
    unsigned long long block_sum[1];          // [n_blocks], becomes exclusive offsets
};

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long &total)
{
    __shared__ unsigned long long s_w[OR_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long x = lane < OR_THREADS / 32 ? s_w[lane] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        s_w[lane] = xi - x;                   // exclusive warp offsets
    }
    __syncthreads();
    unsigned long long excl = s_w[warp] + incl - v;
    // total = offset of last warp + its inclusive sum
    __shared__ unsigned long long s_total;
    if (threadIdx.x == OR_THREADS - 1) s_total = excl + v;
    __syncthreads();
    total = s_total;
    return excl;
}

// pass 1: per-CTA sums of tile counts; the last CTA to finish turns them into exclusive offsets
__global__ void __launch_bounds__(OR_THREADS) order_sum_kernel(const ulonglong2 *__restrict__ seg, int64_t n_tiles,
                                                               OrderTmp *tmp, int n_blocks)
{
    const int64_t base = (int64_t)blockIdx.x * OR_CHUNK + (int64_t)threadIdx.x * OR_PER;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < OR_PER; k++)
        if (base + k < n_tiles) s += seg[base + k].y;
    unsigned long long total;
    block_exclusive_scan(s, total);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        tmp->block_sum[blockIdx.x] = total;
        __threadfence();
        unsigned long long t = atomicAdd(&tmp->ticket, 1ull);
        s_last = (t == (unsigned long long)(n_blocks - 1));
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // n_blocks is small (<= a few thousand): serial chunks of OR_THREADS
        unsigned long long carry = 0;
        for (int b0 = 0; b0 < n_blocks; b0 += OR_THREADS) {
            const int b = b0 + threadIdx.x;
            unsigned long long v = b < n_blocks ? ((volatile unsigned long long *)tmp->block_sum)[b] : 0;
            unsigned long long tot;
            unsigned long long ex = block_exclusive_scan(v, tot);
            if (b < n_blocks) tmp->block_sum[b] = carry + ex;
            carry += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) tmp->ticket = 0;     // re-arm for the next call
    }
}

// pass 2: exclusive offsets per tile, then copy each non-empty segment to its final place
__global__ void __launch_bounds__(OR_THREADS) order_copy_kernel(HitStage st, int64_t n_tiles, const OrderTmp *tmp,
                                                                OrderDest od)
{
    int64_t *__restrict__ out_pos = od.pos;
    float *__restrict__ out_seq = od.seq;
    double *__restrict__ out_str = od.str;
    const unsigned long long out_base = od.out_base ? *od.out_base : 0ull;
    const int64_t base = (int64_t)blockIdx.x * OR_CHUNK + (int64_t)threadIdx.x * OR_PER;
    ulonglong2 sg[OR_PER];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < OR_PER; k++) {
        sg[k] = base + k < n_tiles ? st.tile_seg[base + k] : make_ulonglong2(0ull, 0ull);
        s += sg[k].y;
    }
    unsigned long long total;
    unsigned long long dst = tmp->block_sum[blockIdx.x] + block_exclusive_scan(s, total);
#pragma unroll
    for (int k = 0; k < OR_PER; k++) {
        for (unsigned long long h = 0; h < sg[k].y; h++) {
            const unsigned long long from = sg[k].x + h, to = out_base + dst + h;
            if ((int64_t)from < st.capacity && (int64_t)to < st.capacity) {
                out_pos[to] = st.pos[from];
                if (out_seq && st.seq) out_seq[to] = st.seq[from];
                if (out_str && st.str) out_str[to] = st.str[from];
                if (od.out_motif) od.out_motif[to] = od.motif_id;
            }
        }
        dst += sg[k].y;
    }
}

int rs_order_hits(const HitStage &st, int64_t n_tiles, const OrderDest &dst, void *d_scan_tmp, cudaStream_t stream)
{
    if (n_tiles <= 0) return RS_OK;
    const int n_blocks = (int)((n_tiles + OR_CHUNK - 1) / OR_CHUNK);
    OrderTmp *tmp = (OrderTmp *)d_scan_tmp;
    RS_CUDA(cudaMemsetAsync(tmp, 0, 2 * sizeof(unsigned long long), stream));
    order_sum_kernel<<<n_blocks, OR_THREADS, 0, stream>>>(st.tile_seg, n_tiles, tmp, n_blocks);
    RS_CUDA(cudaGetLastError());
    if (st.capacity > 0) {
        order_copy_kernel<<<n_blocks, OR_THREADS, 0, stream>>>(st, n_tiles, tmp, dst);
        RS_CUDA(cudaGetLastError());
    }
    return RS_OK;
}
