// Averaged-structure filter scan over QUANTISED rows: 8 bytes per position.
//
// Replaces, like profile_scan.cu, the pandas double loop of /root/reference/rnascan/rnascan.py:293-315
// (and, with a sequence table, the sequence scan + inner join of rnascan.py:258-275,416-434) -- for
// inputs whose exact float64 rows (what pd.read_table gives the reference, rnascan.py:296) stay in HOST
// memory.  What travels to the device is the filter form only:
//
//     row = { q_B, q_E, q_H, q_L, q_M, q_R, q_T, code }      q_c = rint(p_c * 255 / scale), uint8
//
// i.e. 8 B per position instead of 1 + 28 (float32) or 1 + 56 (float64).  The kernel evaluates
// sum_j sum_c q[i+j][c] * sf[j][c] in fp32 (sf = table * scale / 255, rounded up) and keeps every window
// whose value is not provably below the threshold: the guard band covers the quantisation step
// (scale / 510 per entry, times sum |table|) and the fp32 evaluation.  Candidates come back as ordered
// positions; the host gathers their exact rows and rs_resolve_candidates (resolve.cu) decides and scores
// them in the reference's arithmetic, so hit sets and scores never depend on the quantisation.
//
// The sequence symbol rides in byte 7 of the row, so the background counts of the sequence
// (rnascan.py:450-453) can be taken in the same pass: the filter does not need the sequence log-odds
// (combine() is an AND of two separately thresholded sets, rnascan.py:416-434).
//
// Layout of a CTA as in fused_filter_kernel: persistent, 128 threads x 9 consecutive windows, tiles of
// 1152 + W - 1 rows fetched with 1-D bulk copies (TMA engine) into a 4-stage ring.  One LDS.64 per row
// (18-word thread stride: conflict-free per half-warp), bytes widened with PRMT + FADD (no I2F).
//
// RS_ROWS_Q4 is the same idea at 4 bytes per position: seven 4-bit channels q_c = floor(p_c * 15 / scale) in the low
// nibbles of a 32-bit word, the symbol (A,C,G,U = 0..3, 0xC other, 0xF separator) in the top nibble.  Flooring
// makes the quantisation error one-sided (0 <= p - q * scale / 15 < scale / 15), so the guard band only needs
// the table's POSITIVE entries: scale / 15 * sum max(table, 0) -- about 2 score units for a typical 7 x 7 motif,
// against 0.2 for the 8-bit form.  Worth it when the threshold is high and the link is the bottleneck; each
// candidate carries its window's symbols (2 bits each) so that a sequence table that arrives later
// (rs_refine_candidates_packed) can thin the list on the device before anything is gathered on the host.
#include "profile_params.cuh"

#define Q8_STAGES 4

__device__ __forceinline__ float q8_to_float(uint32_t w, int k)
{
    // 0x4B0000xx as a float is 2^23 + xx
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)) - 8388608.f;
}

// symbols of window i (tile-relative) sit in byte 7 of rows i .. i+W-1
__device__ __forceinline__ bool q8_deferred_window(const ProfileParams &prm, const uint8_t *rows8, int i, int64_t gpos)
{
    const int W = prm.W;
    if (gpos + W > prm.n) return false;
    const uint8_t *c = rows8 + (size_t)i * 8 + 7;
    if (prm.mode == RS_MODE_AND) {
        double s = 0.0;
        for (int j = 0; j < W; j++) {
            const int idx = c[j * 8] & 7;
            if (idx >= 4) return false;                       // _pwm.c:61-66: any other symbol => NaN
            s = __dadd_rn(s, prm.qd[j * 4 + idx]);
        }
        return (double)(float)s > prm.threshold;              // _pwm.c:65 + SURVEY.md note N1
    }
    for (int j = 0; j < W; j++)
        if (c[j * 8] == RS_SEP) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------ 4-bit rows
__device__ __forceinline__ float q4_to_float(uint32_t w, int k)
{
    return __uint_as_float(((w >> (4 * k)) & 0xFu) | 0x4B000000u) - 8388608.f;
}

// The W symbols of the window starting at row i, 2 bits each (symbol j in bits 2j, 2j+1); bit 63 set when one of
// them is not A,C,G,U, bit 62 when one of them is a separator.
__device__ __forceinline__ unsigned long long q4_window_symbols(const uint32_t *rows, int i, int W)
{
    unsigned long long sym = 0ull;
    for (int j = 0; j < W; j++) {
        const uint32_t c = rows[i + j] >> 28;
        if (c >= 4u) sym |= (c == 0xFu) ? (3ull << 62) : (1ull << 63);
        sym |= (unsigned long long)(c & 3u) << (2 * j);
    }
    return sym;
}

__device__ __forceinline__ bool q4_deferred_window(const ProfileParams &prm, const uint32_t *rows, int i, int64_t gpos)
{
    const int W = prm.W;
    if (gpos + W > prm.n) return false;
    const unsigned long long sym = q4_window_symbols(rows, i, W);
    if (prm.mode == RS_MODE_AND) {
        if (sym >> 63) return false;                          // _pwm.c:61-66: any other symbol => NaN
        double s = 0.0;
        for (int j = 0; j < W; j++) s = __dadd_rn(s, prm.qd[j * 4 + (int)((sym >> (2 * j)) & 3ull)]);
        return (double)(float)s > prm.threshold;              // _pwm.c:65 + SURVEY.md note N1
    }
    return ((sym >> 62) & 3ull) != 3ull;                      // structure only: no separator inside the window
}

template <int W>
__global__ void __launch_bounds__(FT_THREADS, 4) filter_q4_kernel(const __grid_constant__ ProfileParams prm)
{
    constexpr int ROWS = FT_TILE + W - 1;
    constexpr uint32_t STAGE_BYTES = ru16(ROWS * 4);
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *stages = smem + 128;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < Q8_STAGES; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t stride = gridDim.x;
    const int64_t first = blockIdx.x;
    const int64_t my_tiles = first < prm.n_tiles ? (prm.n_tiles - first + stride - 1) / stride : 0;
    const int64_t rows_end = prm.padded * 4;

    auto issue = [&](int64_t it) {
        const int s = (int)(it % Q8_STAGES);
        const int64_t start = (first + it * stride) * FT_TILE * 4;
        const uint32_t bytes = (uint32_t)min((int64_t)STAGE_BYTES, rows_end - start);
        mbar_expect_tx(&bars[s], bytes);
        bulk_g2s(stages + (size_t)s * STAGE_BYTES, reinterpret_cast<const uint8_t *>(prm.profile) + start, bytes, &bars[s]);
    };
    if (tid == 0)
        for (int64_t it = 0; it < Q8_STAGES - 1 && it < my_tiles; it++) issue(it);

    unsigned cnt[4] = {0u, 0u, 0u, 0u};

    for (int64_t it = 0; it < my_tiles; it++) {
        const int s = (int)(it % Q8_STAGES);
        if (tid == 0 && it + Q8_STAGES - 1 < my_tiles) issue(it + Q8_STAGES - 1);
        mbar_wait(&bars[s], (uint32_t)((it / Q8_STAGES) & 1));

        const int64_t tile = first + it * stride;
        const int64_t t0 = tile * FT_TILE;
        const uint32_t *tile_rows = reinterpret_cast<const uint32_t *>(stages + (size_t)s * STAGE_BYTES);
        const uint32_t *rows = tile_rows + tid * FT_P;       // 9-word thread stride: conflict-free

        float acc[FT_P];
#pragma unroll
        for (int i = 0; i < FT_P; i++) acc[i] = 0.f;
        unsigned packed = 0;
#pragma unroll
        for (int r = 0; r < FT_P + W - 1; r++) {
            const uint32_t v = rows[r];
            float x[RS_CHANNELS];
#pragma unroll
            for (int c = 0; c < RS_CHANNELS; c++) x[c] = q4_to_float(v, c);
            if (r < FT_P) {
                const unsigned code = v >> 28;
                const bool counted = code < 4u && t0 + tid * FT_P + r < prm.count_rows;
                packed += counted ? (1u << (8u * code)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < W; j++) {
                const int i = r - j;
                if (i >= 0 && i < FT_P) {
#pragma unroll
                    for (int c = 0; c < RS_CHANNELS; c++)
                        acc[i] = fmaf(x[c], prm.sf[j * RS_CHANNELS + c], acc[i]);
                }
            }
        }
        cnt[0] += packed & 0xffu; cnt[1] += (packed >> 8) & 0xffu;
        cnt[2] += (packed >> 16) & 0xffu; cnt[3] += packed >> 24;

        unsigned candmask = 0;
#pragma unroll
        for (int i = 0; i < FT_P; i++)
            if (!(acc[i] <= prm.filt_thr)) candmask |= 1u << i;

        const int any = __syncthreads_or(candmask != 0);
        if (any) {
            resolve_tile_candidates<FT_THREADS, FT_P>(
                prm.st, tile, candmask,
                [&](int w) { return q4_deferred_window(prm, tile_rows, w, t0 + w); },
                [&](int w, int64_t k) {
                    prm.st.pos[k] = prm.pos_base + t0 + w;
                    if (prm.st.str)                          // the window's symbols ride with the candidate
                        prm.st.str[k] = __longlong_as_double((long long)q4_window_symbols(tile_rows, w, prm.W));
                });
        } else if (tid == 0) {
            prm.st.tile_seg[tile] = make_ulonglong2(0ull, 0ull);
        }
    }

    if (prm.count_on) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned v = cnt[k];
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((tid & 31) == 0 && v) atomicAdd(prm.counts8 + k, (unsigned long long)v);
        }
    }
}

template <int W>
static int launch_q4(const ProfileParams &prm, cudaStream_t stream)
{
    constexpr int ROWS = FT_TILE + W - 1;
    constexpr uint32_t STAGE_BYTES = ru16(ROWS * 4);
    const size_t smem = 128 + (size_t)Q8_STAGES * STAGE_BYTES;
    static bool configured[RS_MAX_DEVICES] = {};
    const int dev = rs_current_device();
    if (!configured[dev]) {
        RS_CUDA(cudaFuncSetAttribute(filter_q4_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    int64_t grid = (int64_t)rs_grid_sms() * 4;
    if (grid > prm.n_tiles) grid = prm.n_tiles;
    rs_prof_start(stream);
    filter_q4_kernel<W><<<(unsigned)grid, FT_THREADS, smem, stream>>>(prm);
    rs_prof_stop(stream);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

template <int W>
struct Q4Dispatch {
    static int run(int w, const ProfileParams &prm, cudaStream_t stream)
    {
        if (w == W) return launch_q4<W>(prm, stream);
        return Q4Dispatch<W - 1>::run(w, prm, stream);
    }
};
template <>
struct Q4Dispatch<0> {
    static int run(int, const ProfileParams &, cudaStream_t)
    {
        rs_set_error("internal: no 4-bit filter kernel for this W");
        return RS_ERR_INVALID;
    }
};

int rs_filter_q4_launch(const ProfileParams &prm, int W, cudaStream_t stream)
{
    return Q4Dispatch<RS_FAST_W>::run(W, prm, stream);
}

// ------------------------------------------------------------------------------------------------ 8-bit rows
template <int W>
__global__ void __launch_bounds__(FT_THREADS, 4) filter_q8_kernel(const __grid_constant__ ProfileParams prm)
{
    constexpr int ROWS = FT_TILE + W - 1;
    constexpr uint32_t STAGE_BYTES = ru16(ROWS * 8);
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *stages = smem + 128;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < Q8_STAGES; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int64_t stride = gridDim.x;
    const int64_t first = blockIdx.x;
    const int64_t my_tiles = first < prm.n_tiles ? (prm.n_tiles - first + stride - 1) / stride : 0;
    const int64_t rows_end = prm.padded * 8;

    auto issue = [&](int64_t it) {
        const int s = (int)(it % Q8_STAGES);
        const int64_t start = (first + it * stride) * FT_TILE * 8;
        const uint32_t bytes = (uint32_t)min((int64_t)STAGE_BYTES, rows_end - start);
        mbar_expect_tx(&bars[s], bytes);
        bulk_g2s(stages + (size_t)s * STAGE_BYTES, reinterpret_cast<const uint8_t *>(prm.profile) + start, bytes, &bars[s]);
    };
    if (tid == 0)
        for (int64_t it = 0; it < Q8_STAGES - 1 && it < my_tiles; it++) issue(it);

    unsigned cnt[4] = {0u, 0u, 0u, 0u};                       // background counts of this thread's rows

    for (int64_t it = 0; it < my_tiles; it++) {
        const int s = (int)(it % Q8_STAGES);
        if (tid == 0 && it + Q8_STAGES - 1 < my_tiles) issue(it + Q8_STAGES - 1);
        mbar_wait(&bars[s], (uint32_t)((it / Q8_STAGES) & 1));

        const int64_t tile = first + it * stride;
        const int64_t t0 = tile * FT_TILE;
        const uint8_t *rows8 = stages + (size_t)s * STAGE_BYTES;
        const uint2 *rows = reinterpret_cast<const uint2 *>(rows8) + tid * FT_P;

        float acc[FT_P];
#pragma unroll
        for (int i = 0; i < FT_P; i++) acc[i] = 0.f;
        unsigned packed = 0;                                   // four byte counters, <= 9 each
#pragma unroll
        for (int r = 0; r < FT_P + W - 1; r++) {
            const uint2 v = rows[r];
            float x[RS_CHANNELS];
            x[0] = q8_to_float(v.x, 0); x[1] = q8_to_float(v.x, 1); x[2] = q8_to_float(v.x, 2);
            x[3] = q8_to_float(v.x, 3); x[4] = q8_to_float(v.y, 0); x[5] = q8_to_float(v.y, 1);
            x[6] = q8_to_float(v.y, 2);
            if (r < FT_P) {                                    // rows this thread owns: symbols A,C,G,U = 0..3
                const unsigned code = v.y >> 24;
                const bool counted = code < 4u && t0 + tid * FT_P + r < prm.count_rows;
                packed += counted ? (1u << (8u * code)) : 0u;
            }
#pragma unroll
            for (int j = 0; j < W; j++) {
                const int i = r - j;
                if (i >= 0 && i < FT_P) {
#pragma unroll
                    for (int c = 0; c < RS_CHANNELS; c++)
                        acc[i] = fmaf(x[c], prm.sf[j * RS_CHANNELS + c], acc[i]);
                }
            }
        }
        cnt[0] += packed & 0xffu; cnt[1] += (packed >> 8) & 0xffu;
        cnt[2] += (packed >> 16) & 0xffu; cnt[3] += packed >> 24;

        unsigned candmask = 0;
#pragma unroll
        for (int i = 0; i < FT_P; i++)
            if (!(acc[i] <= prm.filt_thr)) candmask |= 1u << i;

        const int any = __syncthreads_or(candmask != 0);
        if (any) {
            resolve_tile_candidates<FT_THREADS, FT_P>(
                prm.st, tile, candmask,
                [&](int w) { return q8_deferred_window(prm, rows8, w, t0 + w); },
                [&](int w, int64_t k) { prm.st.pos[k] = prm.pos_base + t0 + w; });
        } else if (tid == 0) {
            prm.st.tile_seg[tile] = make_ulonglong2(0ull, 0ull);
        }
        // every thread has passed the barrier above after its last read of this stage, and the next copy
        // into it is issued by thread 0 one iteration later
    }

    if (prm.count_on) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned v = cnt[k];
            for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((tid & 31) == 0 && v) atomicAdd(prm.counts8 + k, (unsigned long long)v);
        }
    }
}

template <int W>
static int launch_q8(const ProfileParams &prm, cudaStream_t stream)
{
    constexpr int ROWS = FT_TILE + W - 1;
    constexpr uint32_t STAGE_BYTES = ru16(ROWS * 8);
    const size_t smem = 128 + (size_t)Q8_STAGES * STAGE_BYTES;
    static bool configured[RS_MAX_DEVICES] = {};
    const int dev = rs_current_device();
    if (!configured[dev]) {
        RS_CUDA(cudaFuncSetAttribute(filter_q8_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    int64_t grid = (int64_t)rs_grid_sms() * 4;
    if (grid > prm.n_tiles) grid = prm.n_tiles;
    rs_prof_start(stream);
    filter_q8_kernel<W><<<(unsigned)grid, FT_THREADS, smem, stream>>>(prm);
    rs_prof_stop(stream);
    RS_CUDA(cudaGetLastError());
    return RS_OK;
}

template <int W>
struct Q8Dispatch {
    static int run(int w, const ProfileParams &prm, cudaStream_t stream)
    {
        if (w == W) return launch_q8<W>(prm, stream);
        return Q8Dispatch<W - 1>::run(w, prm, stream);
    }
};
template <>
struct Q8Dispatch<0> {
    static int run(int, const ProfileParams &, cudaStream_t)
    {
        rs_set_error("internal: no quantised filter kernel for this W");
        return RS_ERR_INVALID;
    }
};

int rs_filter_q8_launch(const ProfileParams &prm, int W, cudaStream_t stream)
{
    return Q8Dispatch<RS_FAST_W>::run(W, prm, stream);
}
