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
//
// Both passes run on host threads: the buffer is cut at record starts ('>' at the beginning of a line)
// into one chunk per thread, the chunks are counted independently, and a prefix sum over the counts
// gives every chunk its place in the outputs.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
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

struct Chunk {
    int64_t begin, end;                  // byte range; every chunk but the first starts at a record's '>'
    int64_t recs, syms, title_bytes;     // counts of the chunk
};

// first record start at or after `from`: a '>' that begins a line
int64_t next_record_start(const uint8_t *buf, int64_t n, int64_t from)
{
    for (int64_t i = from; i < n; i++)
        if (buf[i] == '>' && (i == 0 || buf[i - 1] == '\n' || buf[i - 1] == '\r')) return i;
    return n;
}

std::vector<Chunk> make_chunks(const uint8_t *buf, int64_t n)
{
    int threads = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("RNASCAN_HOST_THREADS")) threads = atoi(e);
    if (threads < 1) threads = 1;
    if (threads > 32) threads = 32;
    int64_t min_chunk = 4 << 20;                                     // below this a thread is not worth starting
    if (const char *e = getenv("RNASCAN_FASTA_MIN_CHUNK")) min_chunk = atoll(e) > 0 ? atoll(e) : min_chunk;   // tests
    int64_t parts = n / min_chunk;
    if (parts > threads) parts = threads;
    if (parts < 1) parts = 1;
    std::vector<Chunk> chunks;
    int64_t begin = 0;
    for (int64_t k = 1; k <= parts && begin < n; k++) {
        int64_t end = k == parts ? n : next_record_start(buf, n, n / parts * k);
        if (end <= begin) continue;
        chunks.push_back(Chunk{begin, end, 0, 0, 0});
        begin = end;
    }
    if (chunks.empty()) chunks.push_back(Chunk{0, n, 0, 0, 0});
    return chunks;
}

template <typename F>
void run_chunks(std::vector<Chunk> &chunks, F fn)
{
    if (chunks.size() == 1) { fn(0); return; }
    std::vector<std::thread> pool;
    for (size_t k = 0; k < chunks.size(); k++) pool.emplace_back(fn, k);
    for (auto &t : pool) t.join();
}

// A chunk that starts at a record start is a FASTA text of its own; the first chunk keeps the file's
// leading junk, which walk() skips.
void count_chunks(const uint8_t *buf, std::vector<Chunk> &chunks)
{
    run_chunks(chunks, [&](size_t k) {
        Chunk &c = chunks[k];
        int64_t recs = 0, syms = 0, tb = 0;
        const uint8_t *b = buf + c.begin;
        walk(b, c.end - c.begin, [&](int64_t a, int64_t e) { recs++; tb += e - a; },
             [&](int64_t a, int64_t e) { for (int64_t i = a; i < e; i++) syms += b[i] != ' '; });
        c.recs = recs; c.syms = syms; c.title_bytes = tb;
    });
}

}  // namespace

// Pass 1: number of records, symbols (sequence letters after blank removal) and title bytes.
extern "C" int rs_host_fasta_index(const uint8_t *buf, int64_t n, int64_t *n_records, int64_t *n_symbols,
                                   int64_t *title_bytes)
{
    if (n < 0 || (n > 0 && !buf) || !n_records || !n_symbols || !title_bytes) return RS_ERR_INVALID;
    int64_t recs = 0, syms = 0, tb = 0;
    std::vector<Chunk> chunks = make_chunks(buf, n);
    count_chunks(buf, chunks);
    for (const Chunk &c : chunks) { recs += c.recs; syms += c.syms; tb += c.title_bytes; }
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
    std::vector<Chunk> chunks = make_chunks(buf, n);
    count_chunks(buf, chunks);
    // where each chunk writes: records, symbols (+1 separator per record), title bytes before it
    std::vector<int64_t> r0(chunks.size()), w0(chunks.size()), t0(chunks.size());
    int64_t racc = 0, wacc = 0, tacc = 0;
    for (size_t k = 0; k < chunks.size(); k++) {
        r0[k] = racc; w0[k] = wacc; t0[k] = tacc;
        racc += chunks[k].recs; wacc += chunks[k].syms + chunks[k].recs; tacc += chunks[k].title_bytes;
    }
    title_off[0] = 0;
    run_chunks(chunks, [&](size_t k) {
        const uint8_t *b = buf + chunks[k].begin;
        int64_t r = r0[k] - 1, w = w0[k], tw = t0[k];
        const int64_t first = r0[k];
        auto close_record = [&]() {
            if (r >= first) {
                rec_len[r] = w - rec_off[r];
                text[w] = '\n';
                if (codes) codes[w] = RS_SEP;
                w++;
            }
        };
        walk(b, chunks[k].end - chunks[k].begin,
             [&](int64_t a, int64_t e) {
                 close_record();
                 r++;
                 rec_off[r] = w;
                 if (titles && e > a) memcpy(titles + tw, b + a, (size_t)(e - a));
                 tw += e - a;
                 title_off[r + 1] = tw;
             },
             [&](int64_t a, int64_t e) {
                 for (int64_t i = a; i < e; i++) {
                     const uint8_t c = b[i];
                     if (c == ' ') continue;
                     const uint8_t t = tl ? tl[c] : c;
                     text[w] = t;
                     if (codes) codes[w] = cl[t];
                     w++;
                 }
             });
        close_record();
    });
    return RS_OK;
}
