// Sequence-PSSM threshold scan for W <= 8 through an EXACT k-mer decision table.
//
// Replaces, like onehot_scan.cu, _pwm.c:34-68 driven per window by Biopython search()
// (rnascan.py:263) -- but the per-window work is one table lookup instead of W dependent
// fp64 adds.  The nucleotide alphabet has 4 letters, so a window is a 2W-bit number: for
// W <= 8 every possible window fits a 16-bit index and the hit decision
//        (double)(float)(sum_j table[j][code_j])  >  threshold           (_pwm.c:36-65, N1)
// can be tabulated for all 65 536 8-mers up front, IN THE REFERENCE'S OWN ARITHMETIC
// (kmer_lut_kernel: sequential fp64 adds in j order, one cast to float).  The table is
// therefore not a filter with a guard band but the exact decision; scores are recomputed
// only for the windows that are reported.  An 8-mer holds Q = 9 - W consecutive windows,
// so one byte lookup decides Q positions at once.
//
// Per thread: 28 consecutive positions (7-word stride => conflict-free LDS.32), 9 words of
// symbol bytes are squeezed to 2 bits/symbol with one AND + one IMAD per word
// ((w & 0x03030303) * 0x01041040 gathers the four 2-bit fields into the top byte) and
// merged with PRMT; each lookup index is a funnel shift of that 72-bit string.  Invalid
// symbols (N, separators, padding) are ignored on the fast path -- they can only create
// false candidates, never lose a hit -- and candidates of threads whose bytes hold any
// invalid symbol are re-checked.
//
// The hot kernel is a pure stream: a producer warp feeds a 3-stage shared-memory ring with
// bulk async copies (full/empty mbarriers, no CTA-wide barrier), 8 consumer warps turn
// symbols into one 28-bit hit mask per thread (one coalesced 4-byte store) plus one count
// per warp segment (896 positions).  No atomics, no divergent emission.  Ordering is then a
// prefix sum over the segment counts (kmer_segscan_kernel) and an expansion pass
// (kmer_expand_kernel, one warp per segment: ballot/popc prefix inside the segment) that
// writes positions and exact scores straight to their final, position-sorted place.
#include "common.cuh"

#define KM_CONSUMERS   256                         // consumer threads (8 warps)
#define KM_THREADS     (KM_CONSUMERS + 32)         // + one producer warp
#define KM_P           28                          // positions per thread
#define KM_TILE        (KM_CONSUMERS * KM_P)       // 7168 = 28 * 256
#define KM_STAGES      3
#define KM_STAGE_BYTES (KM_TILE + 16)              // +8 symbols of halo, rounded to 16
#define KM_LUT_BYTES   65536
#define KM_WARPS       (KM_CONSUMERS / 32)
#define KM_SEG         (32 * KM_P)                 // 896 positions per warp segment
#define SS_THREADS     1024
#define SS_PER         8
#define SS_CHUNK       (SS_THREADS * SS_PER)       // segments per scan block

struct KmerWork {                 // carved out of the caller's workspace
    uint8_t  *lut;                // [65536]
    uint32_t *mask;               // [n_tiles * 256] hit mask per consumer thread
    uint32_t *segcnt;             // [n_segs] hits per warp segment
    uint32_t *seglocal;           // [n_segs] exclusive offset inside its scan block
    unsigned long long *blockbase;// [n_scan_blocks] exclusive offsets of scan blocks
    unsigned long long *ticket;   // [1]
};

struct KmerParams {
    const uint8_t *codes;
    int64_t        n, padded, n_tiles;
    KmerWork       wk;
};

struct KmerTable { double t[8 * 4]; };

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bit r of lut[idx] = window of W symbols starting at symbol r of the 8-mer `idx` is a hit
__global__ void kmer_lut_kernel(uint8_t *__restrict__ lut, const KmerTable tab, int W, double threshold)
{
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int Q = 9 - W;
    unsigned bits = 0;
    for (int r = 0; r < Q; r++) {
        double s = 0.0;
        for (int j = 0; j < W; j++) s = __dadd_rn(s, tab.t[j * 4 + ((idx >> (2 * (r + j))) & 3u)]);
        const float f = (float)s;                                  // _pwm.c:65
        if ((double)f > threshold) bits |= 1u << r;                // SURVEY.md note N1
    }
    lut[idx] = (uint8_t)bits;
}

template <int W>
__global__ void __launch_bounds__(KM_THREADS, 2) kmer_scan_kernel(const __grid_constant__ KmerParams prm)
{
    constexpr int Q = 9 - W;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);          // [STAGES] tile landed
    uint64_t *empty = full + KM_STAGES;                           // [STAGES] tile consumed by all warps
    uint64_t *lutbar = empty + KM_STAGES;
    uint8_t *s_lut = smem + 128;
    uint8_t *stages = s_lut + KM_LUT_BYTES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < KM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], KM_WARPS); }
        mbar_init(lutbar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t stride = gridDim.x, first = blockIdx.x;
    const int64_t my_tiles = first < prm.n_tiles ? (prm.n_tiles - first + stride - 1) / stride : 0;

    if (warp == KM_WARPS) {
        // ---------------- producer warp: one lane drives the TMA engine
        if (lane == 0) {
            mbar_expect_tx(lutbar, KM_LUT_BYTES);
            bulk_g2s(s_lut, prm.wk.lut, KM_LUT_BYTES, lutbar);
            for (int64_t it = 0; it < my_tiles; it++) {
                const int s = (int)(it % KM_STAGES);
                if (it >= KM_STAGES) mbar_wait(&empty[s], (uint32_t)(((it / KM_STAGES) - 1) & 1));
                const int64_t t0 = (first + it * stride) * KM_TILE;
                const uint32_t bytes = (uint32_t)min((int64_t)KM_STAGE_BYTES, prm.padded - t0);
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(stages + (size_t)s * KM_STAGE_BYTES, prm.codes + t0, bytes, &full[s]);
            }
        }
        return;
    }

    // -------------------- consumer warps
    mbar_wait(lutbar, 0);
    for (int64_t it = 0; it < my_tiles; it++) {
        const int s = (int)(it % KM_STAGES);
        mbar_wait(&full[s], (uint32_t)((it / KM_STAGES) & 1));

        const int64_t tile = first + it * stride;
        const uint8_t *sym = stages + (size_t)s * KM_STAGE_BYTES + tid * KM_P;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(sym);
        const int64_t g0 = tile * KM_TILE + (int64_t)tid * KM_P;

        // ---- 36 symbol bytes -> 72 bits, 2 per symbol, symbol k at bits 2k
        uint32_t x[9], inv = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const uint32_t w = wp[k];
            inv |= w & 0x0C0C0C0Cu;                                // any symbol that is not A/C/G/U
            x[k] = (w & 0x03030303u) * 0x01041040u;                // four 2-bit fields -> top byte
        }
        const uint32_t r0 = __byte_perm(__byte_perm(x[0], x[1], 0x0073), __byte_perm(x[2], x[3], 0x0073), 0x5410);
        const uint32_t r1 = __byte_perm(__byte_perm(x[4], x[5], 0x0073), __byte_perm(x[6], x[7], 0x0073), 0x5410);
        const uint32_t r2 = x[8] >> 24;

        // ---- one decision byte per Q positions
        uint32_t hit = 0;
#pragma unroll
        for (int p = 0; p < KM_P; p += Q) {
            const int sh = 2 * p;
            uint32_t idx;
            if (sh + 16 <= 32)      idx = (r0 >> sh) & 0xFFFFu;
            else if (sh < 32)       idx = __funnelshift_r(r0, r1, sh) & 0xFFFFu;
            else if (sh + 16 <= 64) idx = (r1 >> (sh - 32)) & 0xFFFFu;
            else                    idx = __funnelshift_r(r1, r2, sh - 32) & 0xFFFFu;
            hit |= (uint32_t)s_lut[idx] << p;
        }
        hit &= (1u << KM_P) - 1u;

        // ---- candidates next to an invalid symbol or the end of the stream: check the bytes
        if (hit != 0 && (inv != 0 || g0 + KM_P + W > prm.n)) {
            uint32_t keep = 0;
            for (uint32_t h = hit; h; h &= h - 1) {
                const int p = __ffs(h) - 1;
                bool ok = g0 + p + W <= prm.n;
                for (int j = 0; j < W; j++) ok = ok && sym[p + j] < 4;
                if (ok) keep |= 1u << p;
            }
            hit = keep;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);                     // this warp is done with the stage

        prm.wk.mask[tile * KM_CONSUMERS + tid] = hit;
        const unsigned total = __reduce_add_sync(0xffffffffu, (unsigned)__popc(hit));
        if (lane == 0) prm.wk.segcnt[tile * KM_WARPS + warp] = total;
    }
}

// ---- exclusive prefix sum over the segment counts ------------------------------------------
__device__ __forceinline__ unsigned long long ss_block_scan(unsigned long long v, unsigned long long &total)
{
    __shared__ unsigned long long s_w[SS_THREADS / 32];
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads();                          // protect s_w / s_total across repeated calls
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long x = s_w[lane], xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        s_w[lane] = xi - x;
        if (lane == 31) s_total = xi;
    }
    __syncthreads();
    total = s_total;
    return s_w[warp] + incl - v;
}

__global__ void __launch_bounds__(SS_THREADS) kmer_segscan_kernel(KmerWork wk, int64_t n_segs, int n_blocks,
                                                                  unsigned long long *counters)
{
    const int64_t base = (int64_t)blockIdx.x * SS_CHUNK + (int64_t)threadIdx.x * SS_PER;
    unsigned c[SS_PER];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SS_PER; k++) {
        c[k] = base + k < n_segs ? wk.segcnt[base + k] : 0u;
        s += c[k];
    }
    unsigned long long total;
    unsigned long long ex = ss_block_scan(s, total);
#pragma unroll
    for (int k = 0; k < SS_PER; k++) {
        if (base + k < n_segs) wk.seglocal[base + k] = (uint32_t)ex;
        ex += c[k];
    }
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        wk.blockbase[blockIdx.x] = total;
        __threadfence();
        s_last = atomicAdd(wk.ticket, 1ull) == (unsigned long long)(n_blocks - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        unsigned long long carry = 0;
        for (int b0 = 0; b0 < n_blocks; b0 += SS_THREADS) {
            const int b = b0 + threadIdx.x;
            unsigned long long v = b < n_blocks ? ((volatile unsigned long long *)wk.blockbase)[b] : 0ull;
            unsigned long long tot;
            unsigned long long e = ss_block_scan(v, tot);
            if (b < n_blocks) wk.blockbase[b] = carry + e;
            carry += tot;
        }
        if (threadIdx.x == 0) { counters[0] = carry; *wk.ticket = 0ull; }
    }
}

// ---- expansion: one warp per segment writes its hits, in order, with exact scores ----------
struct ExpandParams {
    const uint8_t *codes;
    KmerWork wk;
    int64_t n_segs, capacity;
    OrderDest od;
    int W;
    double ta[8 * 4];
};

__global__ void __launch_bounds__(256) kmer_expand_kernel(const __grid_constant__ ExpandParams prm)
{
    const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (seg >= prm.n_segs) return;
    // four independent loads issued together (one memory round trip instead of three)
    const uint32_t segcnt = prm.wk.segcnt[seg];
    const uint32_t hit = prm.wk.mask[seg * 32 + lane];
    const unsigned long long bbase = prm.wk.blockbase[seg / SS_CHUNK];
    const uint32_t slocal = prm.wk.seglocal[seg];
    const unsigned long long obase = prm.od.out_base ? *prm.od.out_base : 0ull;
    if (segcnt == 0) return;
    const unsigned cnt = __popc(hit);
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    unsigned long long k = bbase + slocal + (incl - cnt) + obase;
    const int64_t g0 = seg * KM_SEG + (int64_t)lane * KM_P;       // segments tile the stream contiguously
    for (uint32_t h = hit; h; h &= h - 1, k++) {
        if ((int64_t)k >= prm.capacity) break;
        const int p = __ffs(h) - 1;
        const uint8_t *c = prm.codes + g0 + p;
        double s = 0.0;
        for (int j = 0; j < prm.W; j++) s = __dadd_rn(s, prm.ta[j * 4 + (c[j] & 3)]);
        prm.od.pos[k] = g0 + p;
        prm.od.seq[k] = (float)s;                                  // _pwm.c:65
        if (prm.od.out_motif) prm.od.out_motif[k] = prm.od.motif_id;
    }
}

template <int W>
static int launch_kmer(const KmerParams &prm, cudaStream_t stream)
{
    const size_t smem = 128 + KM_LUT_BYTES + (size_t)KM_STAGES * KM_STAGE_BYTES;
    static bool configured[RS_MAX_DEVICES] = {};          // the attribute is per device
    const int dev = rs_current_device();
    if (!configured[dev]) {
        RS_CUDA(cudaFuncSetAttribute(kmer_scan_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    int64_t grid = (int64_t)rs_sm_count() * 2;
    if (grid > prm.n_tiles) grid = prm.n_tiles;
    rs_prof_start(stream);
    kmer_scan_kernel<W><<<(unsigned)grid, KM_THREADS, smem, stream>>>(prm);
    rs_prof_stop(stream);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

int64_t rs_kmer_work_bytes(int64_t n)
{
    const int64_t n_tiles = (n > 0 ? n : 0) / KM_TILE + 1;
    const int64_t n_segs = n_tiles * KM_WARPS;
    const int64_t n_blocks = n_segs / SS_CHUNK + 1;
    return KM_LUT_BYTES + rs_roundup(n_tiles * KM_CONSUMERS * 4, 256) + 2 * rs_roundup(n_segs * 4, 256) +
           rs_roundup(n_blocks * 8, 256) + 256;
}

// Called by rs_scan_seq (onehot_scan.cu) when W <= 8.
int rs_scan_seq_kmer(const uint8_t *d_codes, int64_t n, const double *table, int W, double threshold, int64_t cap,
                     int64_t *d_hit_pos, float *d_hit_score, uint64_t *d_counters2, void *d_work, cudaStream_t st)
{
    WorkLayout wl = rs_work_layout(n, cap);
    uint8_t *wk = (uint8_t *)d_work + wl.off_lut;
    KmerParams prm = {};
    prm.codes = d_codes; prm.n = n; prm.padded = rs_padded_count(n);
    prm.n_tiles = (n + KM_TILE - 1) / KM_TILE;
    const int64_t n_segs = prm.n_tiles * KM_WARPS;
    const int n_blocks = (int)((n_segs + SS_CHUNK - 1) / SS_CHUNK);
    int64_t off = 0;
    prm.wk.lut = wk + off;                          off += KM_LUT_BYTES;
    prm.wk.mask = (uint32_t *)(wk + off);           off += rs_roundup(prm.n_tiles * KM_CONSUMERS * 4, 256);
    prm.wk.segcnt = (uint32_t *)(wk + off);         off += rs_roundup(n_segs * 4, 256);
    prm.wk.seglocal = (uint32_t *)(wk + off);       off += rs_roundup(n_segs * 4, 256);
    prm.wk.blockbase = (unsigned long long *)(wk + off); off += rs_roundup((int64_t)n_blocks * 8, 256);
    prm.wk.ticket = (unsigned long long *)(wk + off);
    RS_CUDA(cudaMemsetAsync(prm.wk.ticket, 0, 8, st));

    KmerTable kt = {};
    for (int k = 0; k < W * 4; k++) kt.t[k] = table[k];
    kmer_lut_kernel<<<KM_LUT_BYTES / 256, 256, 0, st>>>(prm.wk.lut, kt, W, threshold);
    RS_CUDA(cudaGetLastError());
    int rc;
    switch (W) {
    case 1: rc = launch_kmer<1>(prm, st); break;
    case 2: rc = launch_kmer<2>(prm, st); break;
    case 3: rc = launch_kmer<3>(prm, st); break;
    case 4: rc = launch_kmer<4>(prm, st); break;
    case 5: rc = launch_kmer<5>(prm, st); break;
    case 6: rc = launch_kmer<6>(prm, st); break;
    case 7: rc = launch_kmer<7>(prm, st); break;
    case 8: rc = launch_kmer<8>(prm, st); break;
    default: rs_set_error("internal: k-mer scan needs W <= 8"); return RS_ERR_INVALID;
    }
    if (rc) return rc;
    kmer_segscan_kernel<<<n_blocks, SS_THREADS, 0, st>>>(prm.wk, n_segs, n_blocks, (unsigned long long *)d_counters2);
    RS_CUDA(cudaGetLastError());
    if (cap > 0) {
        ExpandParams ep = {};
        ep.codes = d_codes; ep.wk = prm.wk; ep.n_segs = n_segs; ep.capacity = cap; ep.W = W;
        ep.od = OrderDest{d_hit_pos, d_hit_score, nullptr, nullptr, nullptr, 0};
        for (int k = 0; k < W * 4; k++) ep.ta[k] = table[k];
        const int64_t blocks = (n_segs * 32 + 255) / 256;
        kmer_expand_kernel<<<(unsigned)blocks, 256, 0, st>>>(ep);
        RS_CUDA(cudaGetLastError());
    }
    return RS_OK;
}
