// Microbenchmark (scratch, not part of the library): TMEM -> register read throughput of tcgen05.ld.32x32b on sm_100a,
// without any MMA running.  One CTA per SM allocates all 512 TMEM columns; NW warps (NW/4 per lane quadrant) read them
// ITER times with DEPTH loads in flight per warp.  Prints bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_bw tools/micro/tmem_bw.cu && build/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ldwait(uint32_t (&r)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 : : "memory");
}

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) tmem_read(int n_warps, int iters, unsigned long long *out_clk, uint32_t *sink)
{
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < n_warps) {
        const int per_q = n_warps / 4;                       // warps sharing this lane quadrant split the 16 chunks
        const int first = (warp >> 2) * (16 / per_q), count = 16 / per_q;
        uint32_t a[32], b[32];
        for (int it = 0; it < iters; it++) {
            if (DEPTH == 1) {
                for (int c = 0; c < count; c++) { ld32(base + 32u * (first + c), a); ldwait(a); acc ^= a[lane & 31] ^ a[0]; }
            } else {
                ld32(base + 32u * first, a);
                for (int c = 0; c < count; c += 2) {
                    ldwait(a);
                    ld32(base + 32u * (first + c + 1), b);
                    acc ^= a[0] ^ a[31];
                    ldwait(b);
                    if (c + 2 < count) ld32(base + 32u * (first + c + 2), a);
                    acc ^= b[0] ^ b[31];
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *d_clk; uint32_t *d_sink;
    cudaMalloc(&d_clk, sms * 8); cudaMalloc(&d_sink, 4);
    const int iters = 2000;
    for (int depth = 1; depth <= 2; depth++)
        for (int nw = 4; nw <= 16; nw *= 2) {
            if (depth == 1) tmem_read<1><<<sms, 512>>>(nw, iters, d_clk, d_sink);
            else            tmem_read<2><<<sms, 512>>>(nw, iters, d_clk, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long clk0;
            cudaMemcpy(&clk0, d_clk, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * 128 * 512 * 4;     // all 128 lanes x 512 columns x 4 B per iteration
            printf("warps %2d depth %d: %8.0f clk per full TMEM read (256 KB), %.1f B/clk/SM\n", nw, depth,
                   (double)clk0 / iters, bytes / (double)clk0);
        }
    return 0;
}
