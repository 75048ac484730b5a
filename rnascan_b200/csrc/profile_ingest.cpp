// structure.<id>.txt profiles (pfmutil.format_pfm layout) -> float64 rows, many files at once, host threads.
//
// Replaces, for the plain files run_folding writes, the per-file pd.read_table + del struct['PO'] of
// /root/reference/rnascan/rnascan.py:296-297 (one pandas call and one DataFrame per profile).  Scores are
// only bit-identical to the reference's if the numbers are converted exactly as pandas converts them, and
// pandas' default converter is NOT correctly rounded: it is `precise_xstrtod` of pandas' C tokenizer
// (at most 17 significant digits accumulated in a double, then ONE multiplication or division by an exact
// power of ten).  parse_double() restates that algorithm; tests/test_host_cpu.py pins it against
// pandas.read_csv on millions of random tokens.  Anything outside the plain format (quotes, blanks, NaN
// words, ragged rows, duplicate or missing columns, very long integers ...) is reported per file and the
// caller parses that file with pandas itself -- no approximation of pandas' other behaviours is attempted.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "../../include/rnascan_b200.h"

namespace {

const double kPow10[] = {
    1e0,   1e1,   1e2,   1e3,   1e4,   1e5,   1e6,   1e7,   1e8,   1e9,   1e10,  1e11,  1e12,  1e13,  1e14,  1e15,
    1e16,  1e17,  1e18,  1e19,  1e20,  1e21,  1e22,  1e23,  1e24,  1e25,  1e26,  1e27,  1e28,  1e29,  1e30,  1e31,
    1e32,  1e33,  1e34,  1e35,  1e36,  1e37,  1e38,  1e39,  1e40,  1e41,  1e42,  1e43,  1e44,  1e45,  1e46,  1e47,
    1e48,  1e49,  1e50,  1e51,  1e52,  1e53,  1e54,  1e55,  1e56,  1e57,  1e58,  1e59,  1e60,  1e61,  1e62,  1e63,
    1e64,  1e65,  1e66,  1e67,  1e68,  1e69,  1e70,  1e71,  1e72,  1e73,  1e74,  1e75,  1e76,  1e77,  1e78,  1e79,
    1e80,  1e81,  1e82,  1e83,  1e84,  1e85,  1e86,  1e87,  1e88,  1e89,  1e90,  1e91,  1e92,  1e93,  1e94,  1e95,
    1e96,  1e97,  1e98,  1e99,  1e100, 1e101, 1e102, 1e103, 1e104, 1e105, 1e106, 1e107, 1e108, 1e109, 1e110, 1e111,
    1e112, 1e113, 1e114, 1e115, 1e116, 1e117, 1e118, 1e119, 1e120, 1e121, 1e122, 1e123, 1e124, 1e125, 1e126, 1e127,
    1e128, 1e129, 1e130, 1e131, 1e132, 1e133, 1e134, 1e135, 1e136, 1e137, 1e138, 1e139, 1e140, 1e141, 1e142, 1e143,
    1e144, 1e145, 1e146, 1e147, 1e148, 1e149, 1e150, 1e151, 1e152, 1e153, 1e154, 1e155, 1e156, 1e157, 1e158, 1e159,
    1e160, 1e161, 1e162, 1e163, 1e164, 1e165, 1e166, 1e167, 1e168, 1e169, 1e170, 1e171, 1e172, 1e173, 1e174, 1e175,
    1e176, 1e177, 1e178, 1e179, 1e180, 1e181, 1e182, 1e183, 1e184, 1e185, 1e186, 1e187, 1e188, 1e189, 1e190, 1e191,
    1e192, 1e193, 1e194, 1e195, 1e196, 1e197, 1e198, 1e199, 1e200, 1e201, 1e202, 1e203, 1e204, 1e205, 1e206, 1e207,
    1e208, 1e209, 1e210, 1e211, 1e212, 1e213, 1e214, 1e215, 1e216, 1e217, 1e218, 1e219, 1e220, 1e221, 1e222, 1e223,
    1e224, 1e225, 1e226, 1e227, 1e228, 1e229, 1e230, 1e231, 1e232, 1e233, 1e234, 1e235, 1e236, 1e237, 1e238, 1e239,
    1e240, 1e241, 1e242, 1e243, 1e244, 1e245, 1e246, 1e247, 1e248, 1e249, 1e250, 1e251, 1e252, 1e253, 1e254, 1e255,
    1e256, 1e257, 1e258, 1e259, 1e260, 1e261, 1e262, 1e263, 1e264, 1e265, 1e266, 1e267, 1e268, 1e269, 1e270, 1e271,
    1e272, 1e273, 1e274, 1e275, 1e276, 1e277, 1e278, 1e279, 1e280, 1e281, 1e282, 1e283, 1e284, 1e285, 1e286, 1e287,
    1e288, 1e289, 1e290, 1e291, 1e292, 1e293, 1e294, 1e295, 1e296, 1e297, 1e298, 1e299, 1e300, 1e301, 1e302, 1e303,
    1e304, 1e305, 1e306, 1e307, 1e308};

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// One numeric token [p, end), no surrounding blanks.  Returns false when the token is not a plain decimal
// number that pandas would convert with precise_xstrtod (the caller then hands the file to pandas).
bool parse_double(const char *p, const char *end, double *out)
{
    bool negative = false;
    if (p < end && (*p == '-' || *p == '+')) { negative = *p == '-'; p++; }
    double number = 0.0;
    int exponent = 0, num_digits = 0, num_decimals = 0;
    const int max_digits = 17;
    bool has_point = false, has_exp = false;
    int int_digits = 0;
    while (p < end && is_digit(*p)) {
        if (num_digits < max_digits) { number = number * 10.0 + (*p - '0'); num_digits++; }
        else ++exponent;
        int_digits++;
        p++;
    }
    if (p < end && *p == '.') {
        has_point = true;
        p++;
        while (num_digits < max_digits && p < end && is_digit(*p)) {
            number = number * 10.0 + (*p - '0');
            p++; num_digits++; num_decimals++;
        }
        if (num_digits >= max_digits)
            while (p < end && is_digit(*p)) ++p;                      // extra decimals are dropped
        exponent -= num_decimals;
    }
    if (num_digits == 0) return false;
    if (negative) number = -number;
    if (p < end && (*p == 'e' || *p == 'E')) {
        has_exp = true;
        p++;
        bool eneg = false;
        if (p < end && (*p == '-' || *p == '+')) { eneg = *p == '-'; p++; }
        int n = 0, nd = 0;
        while (p < end && is_digit(*p)) {
            if (nd >= 5) return false;                                // absurd exponent: leave it to pandas
            n = n * 10 + (*p - '0'); nd++; p++;
        }
        if (nd == 0) return false;
        exponent += eneg ? -n : n;
    }
    if (p != end) return false;
    // an all-digit token belongs to pandas' integer path (int64 column -> one rounding); identical to the
    // accumulation above only while the integer is exactly representable
    if (!has_point && !has_exp && int_digits > 15) return false;
    if (exponent > 308) {
        return false;                                                 // pandas: ERANGE -> not a float column
    } else if (exponent > 0) {
        number *= kPow10[exponent];
    } else if (exponent < -308) {
        if (exponent < -616) number = 0.0;
        else { number /= kPow10[-308 - exponent]; number /= kPow10[308]; }
    } else {
        number /= kPow10[-exponent];
    }
    if (number == HUGE_VAL || number == -HUGE_VAL) return false;      // pandas: ERANGE
    *out = number;
    return true;
}

struct ProfileFile {
    std::string text;
    int64_t rows = -1;          // data rows, -1 = not handled natively
    int n_cols = 0;             // columns in the header
    int col_of[7];              // header column of channel B,E,H,L,M,R,T
    size_t body = 0;            // offset of the first data line
};

struct Batch {
    std::vector<ProfileFile> files;
};

bool read_file(const char *path, std::string &out)
{
    FILE *fh = fopen(path, "rb");
    if (!fh) return false;
    if (fseek(fh, 0, SEEK_END) != 0) { fclose(fh); return false; }
    long size = ftell(fh);
    if (size < 0) { fclose(fh); return false; }
    rewind(fh);
    out.resize((size_t)size);
    const size_t got = size ? fread(&out[0], 1, (size_t)size, fh) : 0;
    fclose(fh);
    return got == (size_t)size;
}

// end of the line starting at `p` (exclusive of the terminator) and start of the next line; the C parser of
// pandas accepts \n, \r\n and a lone \r
inline void line_bounds(const std::string &t, size_t p, size_t &line_end, size_t &next)
{
    size_t q = p;
    while (q < t.size() && t[q] != '\n' && t[q] != '\r') q++;
    line_end = q;
    if (q < t.size()) {
        if (t[q] == '\r' && q + 1 < t.size() && t[q + 1] == '\n') q += 2;
        else q += 1;
    }
    next = q;
}

// header + row count; false => pandas
bool index_file(ProfileFile &f)
{
    const std::string &t = f.text;
    for (size_t i = 0; i < t.size(); i++) {
        const unsigned char c = (unsigned char)t[i];
        if (c == '"' || c == 0 || c >= 0x80) return false;            // quoting / binary / BOM: pandas decides
    }
    size_t p = 0, le, nx;
    // pandas skips blank lines, also before the header
    for (;;) {
        if (p >= t.size()) return false;
        line_bounds(t, p, le, nx);
        if (le > p) break;
        p = nx;
    }
    // header: tab-separated names; exactly one PO and one of each channel, no duplicates at all
    static const char chan[7] = {'B', 'E', 'H', 'L', 'M', 'R', 'T'};
    for (int c = 0; c < 7; c++) f.col_of[c] = -1;
    int col = 0, po = -1;
    std::vector<std::string> names;
    size_t a = p;
    for (size_t i = p; i <= le; i++) {
        if (i == le || t[i] == '\t') {
            std::string name = t.substr(a, i - a);
            for (const std::string &seen : names)
                if (seen == name) return false;
            names.push_back(name);
            if (name == "PO") po = col;
            else if (name.size() == 1)
                for (int c = 0; c < 7; c++)
                    if (name[0] == chan[c]) f.col_of[c] = col;
            col++;
            a = i + 1;
        }
    }
    if (po < 0) return false;
    for (int c = 0; c < 7; c++)
        if (f.col_of[c] < 0) return false;
    f.n_cols = col;
    f.body = nx;
    // data rows = non-blank lines
    int64_t rows = 0;
    p = nx;
    while (p < t.size()) {
        line_bounds(t, p, le, nx);
        if (le > p) rows++;
        p = nx;
    }
    f.rows = rows;
    return true;
}

// rows of file f -> out[rows][7]; false => pandas
bool fill_file(const ProfileFile &f, double *out)
{
    const std::string &t = f.text;
    int chan_of_col[64];
    if (f.n_cols > 64) return false;
    for (int c = 0; c < f.n_cols; c++) chan_of_col[c] = -1;
    for (int c = 0; c < 7; c++) chan_of_col[f.col_of[c]] = c;
    size_t p = f.body, le, nx;
    int64_t r = 0;
    while (p < t.size()) {
        line_bounds(t, p, le, nx);
        if (le > p) {
            int col = 0;
            size_t a = p;
            for (size_t i = p; i <= le; i++) {
                if (i == le || t[i] == '\t') {
                    if (col >= f.n_cols) return false;                // more fields than names: pandas re-indexes
                    const int ch = chan_of_col[col];
                    if (ch >= 0) {
                        if (!parse_double(t.data() + a, t.data() + i, &out[r * 7 + ch])) return false;
                    } else {
                        // PO and foreign columns are dropped by the caller, but a blank in them still changes how
                        // pandas reads the line
                        for (size_t k = a; k < i; k++)
                            if (t[k] == ' ') return false;
                        if (i == a) return false;
                    }
                    col++;
                    a = i + 1;
                }
            }
            if (col != f.n_cols) return false;                        // short row: pandas pads with NaN
            r++;
        }
        p = nx;
    }
    return r == f.rows;
}

template <typename F>
void parallel_for(int64_t n, int threads, F fn)
{
    if (threads < 1) threads = 1;
    if ((int64_t)threads > n) threads = (int)(n > 0 ? n : 1);
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
    };
    if (threads == 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < threads; k++) pool.emplace_back(worker);
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" int rs_host_profiles_open(const char *const *paths, int64_t n_files, int threads, void **handle,
                                     int64_t *rows_per_file)
{
    if (!paths || !handle || !rows_per_file || n_files < 0) return RS_ERR_INVALID;
    Batch *b = new Batch();
    b->files.resize((size_t)n_files);
    parallel_for(n_files, threads, [&](int64_t i) {
        ProfileFile &f = b->files[(size_t)i];
        if (!read_file(paths[i], f.text) || !index_file(f)) {
            f.rows = -1;
            std::string().swap(f.text);
        }
        rows_per_file[i] = f.rows;
    });
    *handle = b;
    return RS_OK;
}

extern "C" int rs_host_profiles_fill(void *handle, int threads, double *out, const int64_t *row_offsets, int *status)
{
    if (!handle || !out || !row_offsets || !status) return RS_ERR_INVALID;
    Batch *b = (Batch *)handle;
    parallel_for((int64_t)b->files.size(), threads, [&](int64_t i) {
        const ProfileFile &f = b->files[(size_t)i];
        if (f.rows < 0) { status[i] = RS_ERR_INVALID; return; }
        status[i] = fill_file(f, out + row_offsets[i] * 7) ? RS_OK : RS_ERR_INVALID;
    });
    return RS_OK;
}

extern "C" int rs_host_profiles_close(void *handle)
{
    delete (Batch *)handle;
    return RS_OK;
}

// the converter alone (tests pin it against pandas.read_csv): tokens separated by '\n'
extern "C" int rs_host_parse_doubles(const char *text, int64_t n_bytes, double *out, int64_t capacity, int64_t *n_out,
                                     uint8_t *ok)
{
    if (!text || !out || !n_out || !ok || n_bytes < 0) return RS_ERR_INVALID;
    int64_t k = 0;
    const char *p = text, *end = text + n_bytes;
    while (p < end) {
        const char *q = p;
        while (q < end && *q != '\n') q++;
        if (k >= capacity) return RS_ERR_WORKSPACE;
        out[k] = 0.0;
        ok[k] = parse_double(p, q, &out[k]) ? 1 : 0;
        k++;
        p = q + 1;
    }
    *n_out = k;
    return RS_OK;
}
