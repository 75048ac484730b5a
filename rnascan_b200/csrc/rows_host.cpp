// Host side of the filter + gather + resolve scans: the exact profile rows (float64 as pd.read_table
// parses them, /root/reference/rnascan/rnascan.py:296-297, or float32) stay in host memory; these
// routines derive the filter forms that travel to the device and gather the rows of candidate windows.
// Plain C++ on host threads; no CUDA.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/rnascan_b200.h"

void rs_set_error(const char *fmt, ...);

namespace {

template <typename F>
void parallel_blocks(int64_t n, int threads, int64_t block, F fn)     // fn(begin, end) over [0, n)
{
    if (n <= 0) return;
    const int64_t n_blocks = (n + block - 1) / block;
    if (threads < 1) threads = 1;
    if ((int64_t)threads > n_blocks) threads = (int)n_blocks;
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= n_blocks) break;
            const int64_t lo = b * block, hi = lo + block < n ? lo + block : n;
            fn(lo, hi);
        }
    };
    if (threads == 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
}

template <typename T>
void stats_impl(const T *rows, int64_t n, int threads, double *out4)
{
    double rowmax = 0.0, valmax = 0.0;
    struct Acc { double rowmax = 0, valmax = 0; int64_t bad = 0, neg = 0; };
    std::vector<Acc> accs;
    const int64_t block = 1 << 16;
    const int64_t n_blocks = (n + block - 1) / block;
    accs.resize((size_t)(n_blocks > 0 ? n_blocks : 1));
    parallel_blocks(n, threads, block, [&](int64_t lo, int64_t hi) {
        Acc a;
        for (int64_t r = lo; r < hi; r++) {
            double s = 0.0;
            for (int c = 0; c < 7; c++) {
                const double v = (double)rows[r * 7 + c];
                if (!std::isfinite(v)) { a.bad++; continue; }
                const double av = fabs(v);
                s += av;
                if (v < 0) a.neg++;
                if (av > a.valmax) a.valmax = av;
            }
            if (s > a.rowmax) a.rowmax = s;
        }
        accs[(size_t)(lo / block)] = a;
    });
    int64_t nb = 0, nn = 0;
    for (const Acc &a : accs) {
        if (a.rowmax > rowmax) rowmax = a.rowmax;
        if (a.valmax > valmax) valmax = a.valmax;
        nb += a.bad; nn += a.neg;
    }
    out4[0] = rowmax; out4[1] = (double)nb; out4[2] = (double)nn; out4[3] = valmax;
}

template <typename T>
int64_t quantize_impl(const T *rows, int64_t n, const uint8_t *codes, double scale, uint8_t *out, int threads)
{
    const double k = 255.0 / scale;
    std::atomic<int64_t> bad(0);
    parallel_blocks(n, threads, 1 << 16, [&](int64_t lo, int64_t hi) {
        int64_t b = 0;
        for (int64_t r = lo; r < hi; r++) {
            uint8_t *o = out + r * 8;
            for (int c = 0; c < 7; c++) {
                const double v = (double)rows[r * 7 + c];
                if (!(v >= 0.0 && v <= scale)) { b++; o[c] = 255; continue; }      // NaN lands here too
                const double q = nearbyint(v * k);
                o[c] = (uint8_t)(q > 255.0 ? 255.0 : q);
            }
            o[7] = codes ? codes[r] : 0;
        }
        if (b) bad.fetch_add(b);
    });
    return bad.load();
}

template <typename T>
int64_t quantize4_impl(const T *rows, int64_t n, const uint8_t *codes, double scale, uint32_t *out, int threads)
{
    const double k = 15.0 / scale;
    std::atomic<int64_t> bad(0);
    parallel_blocks(n, threads, 1 << 16, [&](int64_t lo, int64_t hi) {
        int64_t b = 0;
        for (int64_t r = lo; r < hi; r++) {
            uint32_t w = 0;
            for (int c = 0; c < 7; c++) {
                const double v = (double)rows[r * 7 + c];
                if (!(v >= 0.0 && v <= scale)) { b++; continue; }                  // NaN lands here too
                double q = floor(v * k);
                if (q > 15.0) q = 15.0;
                while (q > 0.0 && q * (scale / 15.0) > v) q -= 1.0;             // the floor must never overshoot p
                w |= (uint32_t)q << (4 * c);
            }
            w |= (uint32_t)((codes ? codes[r] : 0) & 0xF) << 28;
            out[r] = w;
        }
        if (b) bad.fetch_add(b);
    });
    return bad.load();
}

template <typename T>
void gather_impl(const T *rows, int64_t n_rows, const uint8_t *codes, int64_t code_stride, const int64_t *pos,
                 int64_t n_cand, int W, T *out_rows, uint8_t *out_codes, int threads)
{
    parallel_blocks(n_cand, threads, 1 << 12, [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; k++) {
            const int64_t p = pos[k];
            T *o = out_rows + (size_t)k * W * 7;
            uint8_t *oc = out_codes + (size_t)k * W;
            if (p < 0 || p + W > n_rows) {                    // not a window of the stream: resolved as "no hit"
                memset(o, 0, sizeof(T) * (size_t)W * 7);
                memset(oc, RS_SEP, (size_t)W);
                continue;
            }
            memcpy(o, rows + (size_t)p * 7, sizeof(T) * (size_t)W * 7);
            if (codes && code_stride == -4) {                 // top nibble of 4-byte quantised rows
                const uint32_t *q4 = reinterpret_cast<const uint32_t *>(codes);
                for (int j = 0; j < W; j++) {
                    const uint32_t c = q4[p + j] >> 28;
                    oc[j] = c < 4u ? (uint8_t)c : (c == 0xFu ? (uint8_t)RS_SEP : (uint8_t)RS_RNA_OTHER);
                }
            } else if (codes) {
                for (int j = 0; j < W; j++) oc[j] = codes[(size_t)(p + j) * code_stride];
            } else {
                memset(oc, 0, (size_t)W);
            }
        }
    });
}

}  // namespace

extern "C" int rs_host_rows_stats(const void *rows, int rows_dtype, int64_t n_rows, int threads, double *out4)
{
    if ((!rows && n_rows > 0) || !out4 || n_rows < 0) { rs_set_error("rs_host_rows_stats: bad argument"); return RS_ERR_INVALID; }
    if (rows_dtype == RS_F32) stats_impl((const float *)rows, n_rows, threads, out4);
    else if (rows_dtype == RS_F64) stats_impl((const double *)rows, n_rows, threads, out4);
    else { rs_set_error("rows_dtype must be RS_F32 or RS_F64"); return RS_ERR_INVALID; }
    return RS_OK;
}

extern "C" int rs_host_rows_to_f32(const double *rows, int64_t n_values, float *out, int threads)
{
    if ((!rows || !out) && n_values > 0) { rs_set_error("rs_host_rows_to_f32: null buffer"); return RS_ERR_INVALID; }
    parallel_blocks(n_values, threads, 1 << 18, [&](int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; k++) out[k] = (float)rows[k];      // round to nearest even
    });
    return RS_OK;
}

extern "C" int rs_host_copy(void *dst, const void *src, int64_t n_bytes, int threads)
{
    if ((!dst || !src) && n_bytes > 0) { rs_set_error("rs_host_copy: null buffer"); return RS_ERR_INVALID; }
    parallel_blocks(n_bytes, threads, (int64_t)1 << 22, [&](int64_t lo, int64_t hi) {
        memcpy((char *)dst + lo, (const char *)src + lo, (size_t)(hi - lo));
    });
    return RS_OK;
}

extern "C" int rs_host_quantize_q8(const void *rows, int rows_dtype, int64_t n_rows, const uint8_t *codes, double scale,
                                   uint8_t *out_rows8, int threads, int64_t *n_out_of_range)
{
    if ((!rows || !out_rows8) && n_rows > 0) { rs_set_error("rs_host_quantize_q8: null buffer"); return RS_ERR_INVALID; }
    if (!(scale > 0.0) || !std::isfinite(scale)) { rs_set_error("rs_host_quantize_q8: scale must be positive and finite"); return RS_ERR_INVALID; }
    int64_t bad;
    if (rows_dtype == RS_F32) bad = quantize_impl((const float *)rows, n_rows, codes, scale, out_rows8, threads);
    else if (rows_dtype == RS_F64) bad = quantize_impl((const double *)rows, n_rows, codes, scale, out_rows8, threads);
    else { rs_set_error("rows_dtype must be RS_F32 or RS_F64"); return RS_ERR_INVALID; }
    if (n_out_of_range) *n_out_of_range = bad;
    return RS_OK;
}

extern "C" int rs_host_quantize_q4(const void *rows, int rows_dtype, int64_t n_rows, const uint8_t *codes, double scale,
                                   uint32_t *out_rows4, int threads, int64_t *n_out_of_range)
{
    if ((!rows || !out_rows4) && n_rows > 0) { rs_set_error("rs_host_quantize_q4: null buffer"); return RS_ERR_INVALID; }
    if (!(scale > 0.0) || !std::isfinite(scale)) { rs_set_error("rs_host_quantize_q4: scale must be positive and finite"); return RS_ERR_INVALID; }
    int64_t bad;
    if (rows_dtype == RS_F32) bad = quantize4_impl((const float *)rows, n_rows, codes, scale, out_rows4, threads);
    else if (rows_dtype == RS_F64) bad = quantize4_impl((const double *)rows, n_rows, codes, scale, out_rows4, threads);
    else { rs_set_error("rows_dtype must be RS_F32 or RS_F64"); return RS_ERR_INVALID; }
    if (n_out_of_range) *n_out_of_range = bad;
    return RS_OK;
}

extern "C" int rs_host_gather_windows(const void *rows, int rows_dtype, int64_t n_rows, const uint8_t *codes,
                                      int64_t code_stride, const int64_t *pos, int64_t n_cand, int W,
                                      void *out_rows, uint8_t *out_codes, int threads)
{
    if (n_cand < 0 || n_rows < 0 || W < 1 || W > RS_MAX_W) { rs_set_error("rs_host_gather_windows: bad argument"); return RS_ERR_INVALID; }
    if (n_cand == 0) return RS_OK;
    if (!rows || !pos || !out_rows || !out_codes) { rs_set_error("rs_host_gather_windows: null buffer"); return RS_ERR_INVALID; }
    if (code_stride < 1 && code_stride != -4) code_stride = 1;
    if (rows_dtype == RS_F32)
        gather_impl((const float *)rows, n_rows, codes, code_stride, pos, n_cand, W, (float *)out_rows, out_codes, threads);
    else if (rows_dtype == RS_F64)
        gather_impl((const double *)rows, n_rows, codes, code_stride, pos, n_cand, W, (double *)out_rows, out_codes, threads);
    else { rs_set_error("rows_dtype must be RS_F32 or RS_F64"); return RS_ERR_INVALID; }
    return RS_OK;
}
