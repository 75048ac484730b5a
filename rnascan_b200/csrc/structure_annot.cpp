// Dot-bracket -> structural-context annotation (B,E,H,L,M,R,T), host side.
//
// Restates /root/reference/scripts/parse_secondary_structure.cpp:65-221 (`parse`), the tool
// run_folding pipes every RNAfold centroid through to obtain the one-hot structure alphabet
// the scan consumes (BASELINE config 3's input).  The reference finds pairs with a nested
// forward scan (O(L^2), :13-43) and answers "is there an enclosing pair" by a linear search
// per unpaired run (:113-117); here pairs come from a stack and the enclosing-pair test is the
// nesting depth, so the whole annotation is O(L).  Same output for every balanced structure:
//   L/R  paired 5'/3' base            E  external (also a dangling 3' run, and a run between two
//   H    hairpin loop  "( ... )"         stems with no enclosing pair)
//   B    bulge: run between same-direction parens whose partners are adjacent
//   M    run between ")" and "(" inside an enclosing pair, plus the interior-looking runs that
//        flank such a junction (the 2017 multiloop edit, :157-217);   T  every other interior run
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "../../include/rnascan_b200.h"

namespace {

// returns 0 or RS_ERR_INVALID (unbalanced / foreign characters)
int annotate_one(const char *st, int64_t len, char *out)
{
    std::vector<int64_t> pairs((size_t)len, -1), stack, depth_after((size_t)len, 0);
    int64_t depth = 0;
    for (int64_t i = 0; i < len; i++) {
        const char c = st[i];
        if (c == '(') { stack.push_back(i); depth++; }
        else if (c == ')') {
            if (stack.empty()) return RS_ERR_INVALID;
            const int64_t o = stack.back();
            stack.pop_back();
            pairs[o] = i; pairs[i] = o;
            depth--;
        } else if (c != '.') return RS_ERR_INVALID;
        depth_after[i] = depth;                     // open pairs after position i
    }
    if (!stack.empty()) return RS_ERR_INVALID;

    int64_t i = 0;
    for (; i < len && st[i] == '.'; i++) out[i] = 'E';          // 5' external run (:74-78)
    for (int64_t j = i; j < len;) {
        if (st[j] != '.') { out[j] = st[j] == '(' ? 'L' : 'R'; j++; continue; }
        const int64_t k = j - 1;                                 // nearest paren on the left (j starts a run)
        int64_t m = j;
        while (m < len && st[m] == '.') m++;                     // nearest paren on the right, or len
        char a;
        if (m == len) a = 'E';
        else if (st[k] == '(' && st[m] == ')') a = 'H';
        else if (st[k] == ')' && st[m] == ')') a = pairs[m] + 1 == pairs[k] ? 'B' : 'N';
        else if (st[k] == ')' && st[m] == '(') a = depth_after[k] > 0 ? 'M' : 'E';   // enclosing pair <=> depth > 0
        else a = pairs[m] + 1 == pairs[k] ? 'B' : 'N';           // '(' ... '('
        for (; j < m; j++) out[j] = a;
    }
    // multiloop pass (:157-202): at every ")" directly followed by "(" or by an M run, the interior-
    // looking (N) runs just outside the closing partner of the ")" and just outside the partner of the
    // next "(" belong to the same multiloop
    std::vector<char> is_m((size_t)len, 0);
    for (int64_t j = 0; j < len; j++) {
        if (out[j] != 'R') continue;
        const char nxt = j + 1 < len ? out[j + 1] : '\0';
        if (nxt != 'L' && nxt != 'M') continue;
        int64_t m = j + 1;
        while (st[m] == '.') m++;                                // the "(" ending the switch
        for (int64_t s = pairs[j] - 1; s >= 0 && st[s] == '.'; s--)
            if (out[s] == 'N') is_m[s] = 1;
        for (int64_t e = pairs[m] + 1; e < len && st[e] == '.'; e++)
            if (out[e] == 'N') is_m[e] = 1;
    }
    for (int64_t j = 0; j < len; j++)
        if (out[j] == 'N') out[j] = is_m[j] ? 'M' : 'T';
    return RS_OK;
}

}  // namespace

// n_structs dot-bracket strings: structure r is text[offsets[r] .. offsets[r] + lengths[r]); the
// annotation is written to out at the same offsets.  status[r] (may be NULL) receives RS_OK or
// RS_ERR_INVALID per structure; the return value is RS_OK iff all are valid.
extern "C" int rs_host_annotate_structures(const char *text, const int64_t *offsets, const int64_t *lengths,
                                           int64_t n_structs, char *out, int *status)
{
    if (n_structs < 0 || (n_structs > 0 && (!text || !offsets || !lengths || !out))) return RS_ERR_INVALID;
    unsigned hw = std::thread::hardware_concurrency();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, n_structs / 64));
    std::vector<int> bad((size_t)nt, 0);
    auto work = [&](int64_t t) {
        for (int64_t r = t; r < n_structs; r += nt) {
            const int rc = annotate_one(text + offsets[r], lengths[r], out + offsets[r]);
            if (status) status[r] = rc;
            if (rc) bad[(size_t)t] = 1;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int64_t t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int b : bad) if (b) return RS_ERR_INVALID;
    return RS_OK;
}
