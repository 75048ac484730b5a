// hits.tab text straight from the hit arrays (host side, multi-threaded).
//
// The reference assembles one pandas DataFrame per record, concatenates them, merges the two
// modalities and calls DataFrame.to_csv(sep="\t", index=False) (rnascan.py:284-286,401-413,
// 416-434,555-567).  At every-position output (-m -inf, BASELINE configs 1 and 3) that is hundreds
// of millions of Python objects.  This formatter writes the same bytes from plain arrays:
//   * floats as Python's repr / numpy's shortest round-trip text (what to_csv emits for float
//     columns and for object columns of Python floats): fixed notation for 1e-4 <= |x| < 1e16 with
//     a trailing ".0" for integral values, d.ddde+XX otherwise;
//   * round(x, 3) of a Python float = correctly rounded decimal at 3 places (rnascan.py:273);
//   * csv QUOTE_MINIMAL quoting of fields that hold the delimiter, a quote or a line break.
// rnascan_b200/rnascan.py falls back to the DataFrame path for anything this does not cover and
// tests/test_cli_gpu.py checks both paths against the reference's golden stdout.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <charconv>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/rnascan_b200.h"

namespace {

struct Out {
    std::string s;
    void put(const char *p, size_t n) { s.append(p, n); }
    void put(char c) { s.push_back(c); }
    void put_int(int64_t v)
    {
        char b[24];
        auto r = std::to_chars(b, b + sizeof(b), v);
        s.append(b, r.ptr - b);
    }
    // csv.QUOTE_MINIMAL with delimiter '\t', quotechar '"', lineterminator '\n'
    void put_field(const char *p, int64_t n)
    {
        bool quote = false;
        for (int64_t i = 0; i < n; i++) {
            const char c = p[i];
            if (c == '\t' || c == '"' || c == '\n' || c == '\r') { quote = true; break; }
        }
        if (!quote) { s.append(p, (size_t)n); return; }
        s.push_back('"');
        for (int64_t i = 0; i < n; i++) {
            if (p[i] == '"') s.push_back('"');
            s.push_back(p[i]);
        }
        s.push_back('"');
    }
};

// repr(float): shortest round-trip digits, Python's layout
bool put_f64(Out &o, double x)
{
    if (x != x || isinf(x)) return false;                    // never a hit; let the caller fall back
    char b[40];
    const double a = fabs(x);
    if (a == 0.0 || (a >= 1e-4 && a < 1e16)) {
        auto r = std::to_chars(b, b + sizeof(b), x, std::chars_format::fixed);
        size_t n = r.ptr - b;
        o.put(b, n);
        if (!memchr(b, '.', n)) o.put(".0", 2);
    } else {
        auto r = std::to_chars(b, b + sizeof(b), x, std::chars_format::scientific);
        o.put(b, r.ptr - b);                                 // d.ddde-05 / 1e+16, as Python prints it
    }
    return true;
}

// numpy's text of a float32 in the range rnascan's rounded scores live in
bool put_f32(Out &o, float x)
{
    const float a = fabsf(x);
    if (!(a == 0.0f || (a >= 1e-4f && a < 1e7f))) return false;
    char b[40];
    auto r = std::to_chars(b, b + sizeof(b), x, std::chars_format::fixed);
    size_t n = r.ptr - b;
    o.put(b, n);
    if (!memchr(b, '.', n)) o.put(".0", 2);
    return true;
}

// Python's round(x, 3) on a float: correctly rounded decimal at 3 places, back to double
double round3(double x)
{
    if (x != x || isinf(x) || fabs(x) >= 1e15) return x;
    char b[64];
    snprintf(b, sizeof(b), "%.3f", x);
    return strtod(b, nullptr);
}

struct Strings {                 // n strings in one blob: string r = blob[off[r] .. off[r+1])
    const char *blob;
    const int64_t *off;
    void put(Out &o, int64_t r) const { o.put_field(blob + off[r], off[r + 1] - off[r]); }
};

template <typename F>
int run_rows(int64_t n_rows, char *out, int64_t capacity, int64_t *written, F format_rows)
{
    unsigned hw = std::thread::hardware_concurrency();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, n_rows / 20000));
    std::vector<Out> parts((size_t)nt);
    std::vector<int> ok((size_t)nt, 1);
    const int64_t per = (n_rows + nt - 1) / nt;
    auto work = [&](int64_t t) {
        const int64_t a = t * per, b = std::min(n_rows, a + per);
        if (a < b) {
            parts[(size_t)t].s.reserve((size_t)(b - a) * 96);
            ok[(size_t)t] = format_rows(parts[(size_t)t], a, b) ? 1 : 0;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int64_t t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    int64_t total = 0;
    for (int64_t t = 0; t < nt; t++) {
        if (!ok[(size_t)t]) return RS_ERR_INVALID;           // a value outside the covered text formats
        total += (int64_t)parts[(size_t)t].s.size();
    }
    *written = total;
    if (total > capacity) return RS_ERR_WORKSPACE;           // *written tells the caller how much is needed
    int64_t off = 0;
    for (int64_t t = 0; t < nt; t++) {
        memcpy(out + off, parts[(size_t)t].s.data(), parts[(size_t)t].s.size());
        off += (int64_t)parts[(size_t)t].s.size();
    }
    return RS_OK;
}

}  // namespace

// round(x, 3) delivered by the device as an int32 number of thousandths (rs_scores_dense_struct_milli): the text
// of repr(k / 1000.0) is the decimal itself -- integer part, '.', the thousandths without trailing zeros (at least
// one digit).  RS_MILLI_NEG0 is the -0.0 Python prints for scores in (-0.0005, 0].
bool put_milli(Out &o, int32_t k)
{
    if (k == RS_MILLI_NEG0) { o.put_field("-0.0", 4); return true; }
    if (k <= RS_MILLI_RANGE) return false;               // NaN / -inf / out of range: not a printable hit
    int64_t v = k;
    if (v < 0) { o.put('-'); v = -v; }
    o.put_int(v / 1000);
    o.put('.');
    const int frac = (int)(v % 1000);
    const char d0 = (char)('0' + frac / 100), d1 = (char)('0' + (frac / 10) % 10), d2 = (char)('0' + frac % 10);
    o.put(d0);
    if (d1 != '0' || d2 != '0') o.put(d1);
    if (d2 != '0') o.put(d2);
    return true;
}

// Rows of a single-modality scan (modes RNA and SS):
//   Sequence_ID  Description  Motif_ID  Start  End  Sequence  LogOdds  Match_ID
// rec[r] indexes the per-record id/description strings; Start = start0[r] + 1, End = start0[r] + width;
// the fragment is text[text_pos[r] .. +width).  score_kind: 0 = float32 values already rounded,
// printed as numpy float32 text; 1 = float32 values already rounded, printed as the Python float they
// widen to (object column, rnascan.py:408 concat with an empty frame); 2 = float64 values, round(x, 3)
// applied here, printed as Python floats; 4 = int32 thousandths of round(x, 3) computed on the device.
// Match_ID = match_id_first + r.
extern "C" int rs_host_format_hits(int64_t n_rows, int64_t match_id_first, const int64_t *rec,
                                   const char *id_blob, const int64_t *id_off, const char *desc_blob,
                                   const int64_t *desc_off, const char *motif_id, const int64_t *start0,
                                   int64_t width, const uint8_t *text, const int64_t *text_pos, int score_kind,
                                   const void *scores, char *out, int64_t capacity, int64_t *written)
{
    if (n_rows < 0 || !written || (n_rows > 0 && (!rec || !id_off || !desc_off || !start0 || !scores || !motif_id)))
        return RS_ERR_INVALID;
    // bit 8 of score_kind: Start and End as float64 text ("12.0") -- what pandas prints for a directory of averaged
    // profiles in which some file had no hit (the empty per-file frame turns the concatenated columns float)
    const bool float_pos = (score_kind & 0x100) != 0;
    score_kind &= 0xFF;
    if (score_kind < 0 || score_kind > 4) return RS_ERR_INVALID;
    const Strings ids{id_blob, id_off}, descs{desc_blob, desc_off};
    const size_t motif_len = strlen(motif_id);
    return run_rows(n_rows, out, capacity, written, [&](Out &o, int64_t a, int64_t b) {
        for (int64_t r = a; r < b; r++) {
            if (float_pos && start0[r] + width >= 1000000000000000LL) return false;   // repr() turns to exponents at 1e16
            ids.put(o, rec[r]); o.put('\t');
            descs.put(o, rec[r]); o.put('\t');
            o.put_field(motif_id, (int64_t)motif_len); o.put('\t');
            o.put_int(start0[r] + 1); if (float_pos) { o.put('.'); o.put('0'); } o.put('\t');
            o.put_int(start0[r] + width); if (float_pos) { o.put('.'); o.put('0'); } o.put('\t');
            if (text) o.put_field(reinterpret_cast<const char *>(text) + text_pos[r], width);
            else o.put('.');
            o.put('\t');
            bool fine;
            if (score_kind == 0) fine = put_f32(o, static_cast<const float *>(scores)[r]);
            else if (score_kind == 1) fine = put_f64(o, (double)static_cast<const float *>(scores)[r]);
            else if (score_kind == 2) fine = put_f64(o, round3(static_cast<const double *>(scores)[r]));
            else if (score_kind == 3) fine = put_f64(o, static_cast<const double *>(scores)[r]);
            else fine = put_milli(o, static_cast<const int32_t *>(scores)[r]);
            if (!fine) return false;
            o.put('\t');
            o.put_int(match_id_first + r);
            o.put('\n');
        }
        return true;
    });
}

// Rows of the combined mode (rnascan.py:416-434 after the inner join):
//   Sequence_ID Description.Seq Motif_ID.Seq Start End Sequence.Seq LogOdds.Seq
//   Description.Struct Motif_ID.Struct Sequence.Struct LogOdds.Struct LogOdds.SeqStruct Match_ID
// seq_kind as score_kind 0/1 above; struct_kind 2 = one-hot scores (round(x, 3) applied here),
// 3 = averaged-profile scores (unrounded float64, rnascan.py:311); struct_text NULL prints ".",
// struct descriptions NULL print empty fields (averaged mode).  LogOdds.SeqStruct is the float64
// sum of the two PRINTED values (rnascan.py:432-433).
extern "C" int rs_host_format_hits_combined(int64_t n_rows, int64_t match_id_first, const int64_t *rec,
                                            const char *id_blob, const int64_t *id_off, const char *desc_blob,
                                            const int64_t *desc_off, const char *sdesc_blob,
                                            const int64_t *sdesc_off, const char *motif_seq,
                                            const char *motif_struct, const int64_t *start0, int64_t width,
                                            const uint8_t *seq_text, const uint8_t *struct_text,
                                            const int64_t *text_pos, int seq_kind, const float *seq_scores,
                                            int struct_kind, const double *struct_scores, char *out,
                                            int64_t capacity, int64_t *written)
{
    if (n_rows < 0 || !written ||
        (n_rows > 0 && (!rec || !id_off || !desc_off || !start0 || !seq_scores || !struct_scores || !motif_seq ||
                        !motif_struct || !seq_text || !text_pos)))
        return RS_ERR_INVALID;
    if (seq_kind < 0 || seq_kind > 1 || struct_kind < 2 || struct_kind > 3) return RS_ERR_INVALID;
    const Strings ids{id_blob, id_off}, descs{desc_blob, desc_off}, sdescs{sdesc_blob, sdesc_off};
    const size_t ml1 = strlen(motif_seq), ml2 = strlen(motif_struct);
    return run_rows(n_rows, out, capacity, written, [&](Out &o, int64_t a, int64_t b) {
        for (int64_t r = a; r < b; r++) {
            ids.put(o, rec[r]); o.put('\t');
            descs.put(o, rec[r]); o.put('\t');
            o.put_field(motif_seq, (int64_t)ml1); o.put('\t');
            o.put_int(start0[r] + 1); o.put('\t');
            o.put_int(start0[r] + width); o.put('\t');
            o.put_field(reinterpret_cast<const char *>(seq_text) + text_pos[r], width); o.put('\t');
            const float q = seq_scores[r];
            if (!(seq_kind == 0 ? put_f32(o, q) : put_f64(o, (double)q))) return false;
            o.put('\t');
            if (sdesc_off) sdescs.put(o, rec[r]);
            o.put('\t');
            o.put_field(motif_struct, (int64_t)ml2); o.put('\t');
            if (struct_text) o.put_field(reinterpret_cast<const char *>(struct_text) + text_pos[r], width);
            else o.put('.');
            o.put('\t');
            const double s = struct_kind == 2 ? round3(struct_scores[r]) : struct_scores[r];
            if (!put_f64(o, s)) return false;
            o.put('\t');
            if (!put_f64(o, (double)q + s)) return false;
            o.put('\t');
            o.put_int(match_id_first + r);
            o.put('\n');
        }
        return true;
    });
}
