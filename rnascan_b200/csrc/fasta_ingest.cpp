// FASTA text -> packed symbol stream, host side (north-star subsystem 1: "FASTA is packed on the host
// into uint8 codes with an ambiguity mask").
//
// Replaces, for file inputs, the per-record Python objects of the reference's ingest:
// fileinput + Bio.SeqIO.parse(fin, 'fasta') (rnascan.py:170-174), Seq.transcribe()/upper()
// (rnascan.py:186-193) and the per-window character switch of _pwm.c:41-63.  Semantics restated from
// Biopython's SimpleFastaParser as the reference uses it through a text-mode handle:
//   * universal newlines: "\n", "\r\n" and a lone "\r" end a line;
//   * everything before the first line that starts with '>' is ignored;
//   * title  = header line without '>' and without trailing whitespace; id = its first word;
//   * sequence = the following lines, each stripped of trailing whitespace, concatenated, with all
//     blanks removed;
//   * RNA target alphabet: T->U, t->u, then ASCII upper-casing; structure alphabet: unchanged.
// The caller guarantees ASCII input (rnascan.py falls back to its Python parser otherwise).
#include <stdint.h>
#include <string.h>
#include "../../include/rnascan_b200.h"

namespace {

inline bool is_space(uint8_t c)          // str.rstrip() whitespace within ASCII
{
    return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f);
}

struct Luts {
    uint8_t rna_text[256], rna_code[256], ss_code[256];
    Luts()
    {
        for (int i = 0; i < 256; i++) {
            uint8_t c = (uint8_t)i;
            if (c == 'T' || c == 't') c = 'U';
            else if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
            rna_text[i] = c;
            rna_code[i] = RS_RNA_OTHER;
            ss_code[i] = RS_SS_OTHER;
        }
        const char *r = "ACGU";
        for (int k = 0; k < 4; k++) rna_code[(uint8_t)r[k]] = (uint8_t)k;
        const char *s = "BEHLMRT";
        for (int k = 0; k < 7; k++) { ss_code[(uint8_t)s[k]] = (uint8_t)k; ss_code[(uint8_t)(s[k] | 0x20)] = (uint8_t)(k | 8); }
    }
};
const Luts g;

// One pass over the text; `emit` callbacks are no-ops when only counting.
template <typename OnRecord, typename OnSymbols>
void walk(const uint8_t *buf, int64_t n, OnRecord on_record, OnSymbols on_symbols)
{
    int64_t i = 0;
    bool in_record = false;
    while (i < n) {
        int64_t e = i;                                   // line = [i, e)
        while (e < n && buf[e] != '\n' && buf[e] != '\r') e++;
        int64_t next = e;
        if (next < n) next += (buf[next] == '\r' && next + 1 < n && buf[next + 1] == '\n') ? 2 : 1;
        if (e > i && buf[i] == '>') {
            int64_t t1 = e;
            while (t1 > i + 1 && is_space(buf[t1 - 1])) t1--;
            on_record(i + 1, t1);                        // title span
            in_record = true;
        } else if (in_record) {
            int64_t l1 = e;
            while (l1 > i && is_space(buf[l1 - 1])) l1--;
            if (l1 > i) on_symbols(i, l1);               // blanks inside are dropped by the callee
        }
        i = next;
    }
}

}  // namespace

// Pass 1: number of records, symbols (sequence letters after blank removal) and title bytes.
extern "C" int rs_host_fasta_index(const uint8_t *buf, int64_t n, int64_t *n_records, int64_t *n_symbols,
                                   int64_t *title_bytes)
{
    if (n < 0 || (n > 0 && !buf) || !n_records || !n_symbols || !title_bytes) return RS_ERR_INVALID;
    int64_t recs = 0, syms = 0, tb = 0;
    walk(buf, n, [&](int64_t a, int64_t b) { recs++; tb += b - a; },
         [&](int64_t a, int64_t b) { for (int64_t k = a; k < b; k++) syms += buf[k] != ' '; });
    *n_records = recs; *n_symbols = syms; *title_bytes = tb;
    return RS_OK;
}

// Pass 2.  kind 0 = RNA target alphabet, 1 = structure contexts.  Outputs (caller-allocated):
//   text  [n_symbols + n_records]  pre-processed letters, '\n' after every record
//   codes [n_symbols + n_records]  symbol codes, RS_SEP after every record (may be NULL)
//   rec_off, rec_len [n_records]   position of each record in text/codes
//   titles [title_bytes], title_off [n_records + 1]
extern "C" int rs_host_fasta_fill(const uint8_t *buf, int64_t n, int kind, uint8_t *text, uint8_t *codes,
                                  int64_t *rec_off, int64_t *rec_len, char *titles, int64_t *title_off)
{
    if (n < 0 || (n > 0 && !buf) || kind < 0 || kind > 1 || !text || !rec_off || !rec_len || !title_off)
        return RS_ERR_INVALID;
    const uint8_t *tl = kind == 0 ? g.rna_text : nullptr;
    const uint8_t *cl = kind == 0 ? g.rna_code : g.ss_code;
    int64_t r = -1, w = 0, tw = 0;
    auto close_record = [&]() {
        if (r >= 0) {
            rec_len[r] = w - rec_off[r];
            text[w] = '\n';
            if (codes) codes[w] = RS_SEP;
            w++;
        }
    };
    title_off[0] = 0;
    walk(buf, n,
         [&](int64_t a, int64_t b) {
             close_record();
             r++;
             rec_off[r] = w;
             if (titles && b > a) memcpy(titles + tw, buf + a, (size_t)(b - a));
             tw += b - a;
             title_off[r + 1] = tw;
         },
         [&](int64_t a, int64_t b) {
             for (int64_t k = a; k < b; k++) {
                 const uint8_t c = buf[k];
                 if (c == ' ') continue;
                 const uint8_t t = tl ? tl[c] : c;
                 text[w] = t;
                 if (codes) codes[w] = cl[t];
                 w++;
             }
         });
    close_record();
    return RS_OK;
}
