// Background letter counts over the symbol stream (exact integers).
//
// Replaces the per-letter Seq.count() passes of /root/reference/rnascan/rnascan.py:450-453.
// A symbol is counted in bin (code & 7) iff bit 3 of its code is clear (separators 0xFF,
// ambiguous bases 0x0C/0x0F and lower-case structure letters are not counted -- SURVEY.md H11).
//
// HBM-bound at 1 B/symbol: each thread streams 16-byte vectors; per 32-bit word the three
// index bit-planes are masked with the validity plane and the 8 subset sums
// (v, b0, b1, b2, b0b1, b0b2, b1b2, b0b1b2) are accumulated as packed byte counters in
// registers (no shared-memory atomics in the loop).  Bin counts follow by inclusion-exclusion
// once per thread, then warp shuffles -> shared memory -> ONE global atomic per bin per CTA.
#include "common.cuh"

#define HI_THREADS 256

__device__ __forceinline__ unsigned bytesum(unsigned x)       // sum of the four byte lanes
{
    return __dp4a(x, 0x01010101u, 0u);
}

__global__ void __launch_bounds__(HI_THREADS) hist_kernel(const uint8_t *__restrict__ codes, int64_t n,
                                                          unsigned long long *__restrict__ counts)
{
    // subset sums: index bit s set <=> plane s is in the product; element 0 = valid count
    unsigned long long tot[8];
#pragma unroll
    for (int k = 0; k < 8; k++) tot[k] = 0;

    const int64_t nvec = n / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(codes);
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;

    while (i < nvec) {
        unsigned acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0;
        // up to 15 vectors (60 words, each adds <= 1 per byte lane ... 4 words/vector => <= 60 < 256)
        for (int rep = 0; rep < 15 && i < nvec; rep++, i += gstride) {
            const uint4 q = __ldg(v + i);
            const unsigned w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const unsigned w = w4[t];
                const unsigned nv = ~(w >> 3) & 0x01010101u;      // valid: bit 3 clear
                const unsigned b0 = w & nv;
                const unsigned b1 = (w >> 1) & nv;
                const unsigned b2 = (w >> 2) & nv;
                acc[0] += nv;
                acc[1] += b0;
                acc[2] += b1;
                acc[4] += b2;
                acc[3] += b0 & b1;
                acc[5] += b0 & b2;
                acc[6] += b1 & b2;
                acc[7] += b0 & b1 & b2;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) tot[k] += bytesum(acc[k]);
    }
    // tail symbols (n % 16) by the first threads of block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 15)) {
        const unsigned c = codes[nvec * 16 + threadIdx.x];
        if (!(c & 8)) {
            const unsigned b0 = c & 1, b1 = (c >> 1) & 1, b2 = (c >> 2) & 1;
            tot[0] += 1; tot[1] += b0; tot[2] += b1; tot[4] += b2;
            tot[3] += b0 & b1; tot[5] += b0 & b2; tot[6] += b1 & b2; tot[7] += b0 & b1 & b2;
        }
    }
    // inclusion-exclusion: count of index k = sum over supersets T of bits(k) of (-1)^{|T|-|k|} S[T]
    long long cnt[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        long long c = 0;
#pragma unroll
        for (int T = 0; T < 8; T++)
            if ((T & k) == k) c += (__popc(T ^ k) & 1) ? -(long long)tot[T] : (long long)tot[T];
        cnt[k] = c;
    }
    __shared__ unsigned long long s_bins[8];
    if (threadIdx.x < 8) s_bins[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        long long c = cnt[k];
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_bins[k], (unsigned long long)c);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_bins[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_bins[threadIdx.x]);
}

extern "C" int rs_hist(const uint8_t *d_codes, int64_t n, uint64_t *d_counts8, void *stream)
{
    if (!d_codes || !d_counts8 || n < 0) { rs_set_error("rs_hist: bad argument"); return RS_ERR_INVALID; }
    if ((uintptr_t)d_codes & 15) { rs_set_error("codes pointer must be 16-byte aligned"); return RS_ERR_INVALID; }
    if (n == 0) return RS_OK;
    int64_t blocks = (n / 16 + HI_THREADS - 1) / HI_THREADS;
    int64_t cap = (int64_t)rs_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    hist_kernel<<<(unsigned)blocks, HI_THREADS, 0, (cudaStream_t)stream>>>(d_codes, n,
                                                                         (unsigned long long *)d_counts8);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}
