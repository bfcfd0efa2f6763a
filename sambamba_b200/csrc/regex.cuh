// regex.cuh -- the regular expressions of `-F` (`field =~ /pattern/flags`, queryparser.d:425-458; RegexpFieldFilter /
// RegexpTagFilter, filtering.d:305-345: "does the string contain a match").
//
// The reference hands the pattern to D's std.regex.  A filter only asks whether a match EXISTS, and for patterns without
// back-references or look-around that is a question about a regular language: the pattern is compiled on the host into a
// Thompson NFA of at most 64 instructions and the per-record evaluation is a breadth-first simulation with the set of
// live states in one 64-bit register (no backtracking, no captures; greedy / lazy quantifiers give the same answer).
// Supported: literals, `.`, classes `[a-z0-9_]` / `[^...]` with ranges and \d \w \s inside, \d \D \w \W \s \S, \t \n \r,
// \xHH, escaped punctuation, `^` `$` `\b` `\B`, groups `( )` `(?: )`, alternation, `* + ? {m} {m,} {m,n}` (also lazy),
// flags i (case-insensitive), s, g.  Refused at compile time (never evaluated differently): back-references, look-around,
// \p / unicode classes, class set operations, flags x / m / U, patterns that need more than 64 NFA states or 4 classes.
//
// __host__ __device__ evaluator; tests/test_emul_filter.py checks it against Python's `re` on the same strings.
#pragma once
#include <stdint.h>
#ifndef BD_HD
#if defined(__CUDACC__)
#define BD_HD __host__ __device__ __forceinline__
#else
#define BD_HD inline
#endif
#endif

namespace bdk {

constexpr int RX_MAX_INST = 64, RX_MAX_CLS = 4;
enum RxOp : uint8_t { RX_CHAR = 0, RX_CHARI, RX_ANY, RX_ANYNL, RX_CLS, RX_SPLIT, RX_JMP, RX_MATCH, RX_BOL, RX_EOL, RX_WB, RX_NWB };
struct RxInst { uint8_t op, c, x, y; };
struct RegexProg { uint8_t n, n_cls, pad[2]; RxInst in[RX_MAX_INST]; uint32_t cls[RX_MAX_CLS][8]; };

BD_HD bool rx_isword(int c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
BD_HD int rx_lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
BD_HD int rx_ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
// epsilon closure of `seeds` at a position whose neighbours are prev / next (-1: none)
BD_HD uint64_t rx_closure(const RegexProg& p, uint64_t seeds, int prev, int next) {
    uint64_t done = 0, out = 0, work = seeds;
    while (work) {
        int pc = rx_ctz64(work); work &= work - 1;
        if ((done >> pc) & 1) continue;
        done |= 1ull << pc;
        const RxInst in = p.in[pc];
        switch (in.op) {
        case RX_SPLIT: work |= ((1ull << in.x) | (1ull << in.y)) & ~done; break;
        case RX_JMP: work |= (1ull << in.x) & ~done; break;
        case RX_BOL: if (prev < 0) work |= (1ull << (pc + 1)) & ~done; break;
        case RX_EOL: if (next < 0) work |= (1ull << (pc + 1)) & ~done; break;
        case RX_WB: if ((prev >= 0 && rx_isword(prev)) != (next >= 0 && rx_isword(next))) work |= (1ull << (pc + 1)) & ~done; break;
        case RX_NWB: if ((prev >= 0 && rx_isword(prev)) == (next >= 0 && rx_isword(next))) work |= (1ull << (pc + 1)) & ~done; break;
        default: out |= 1ull << pc; break;      // consuming instruction or MATCH
        }
    }
    return out;
}
// Gen: bool next(int* c) -- the characters of the subject, one after the other
template <class Gen> BD_HD bool rx_search(const RegexProg& p, Gen& g) {
    if (!p.n) return false;
    uint64_t match_mask = 0; for (int i = 0; i < p.n; i++) if (p.in[i].op == RX_MATCH) match_mask |= 1ull << i;
    int prev = -1, c = -1; bool have = g.next(&c);
    uint64_t cur = rx_closure(p, 1ull, prev, have ? c : -1);
    while (true) {
        if (cur & match_mask) return true;
        if (!have) return false;
        uint64_t nxt = 0, w = cur & ~match_mask;
        while (w) {
            int pc = rx_ctz64(w); w &= w - 1; const RxInst in = p.in[pc]; bool ok;
            switch (in.op) {
            case RX_CHAR: ok = c == in.c; break;
            case RX_CHARI: ok = rx_lower(c) == in.c; break;
            case RX_ANY: ok = c != '\n' && c != '\r'; break;
            case RX_ANYNL: ok = true; break;
            case RX_CLS: ok = (p.cls[in.c][(c >> 5) & 7] >> (c & 31)) & 1; break;
            default: ok = false; break;
            }
            if (ok) nxt |= 1ull << (pc + 1);
        }
        prev = c; have = g.next(&c);
        cur = rx_closure(p, nxt | 1ull, prev, have ? c : -1);        // | 1: a match may begin at every position
    }
}

struct RxBytes { const uint8_t* p; uint32_t n, i; BD_HD bool next(int* c) { if (i >= n) return false; *c = p[i++]; return true; } };
struct RxSeq {      // decoded bases, read.d:364-383
    const uint8_t* seq; uint32_t n, i;
    BD_HD bool next(int* c) { if (i >= n) return false; uint8_t b = seq[i >> 1]; *c = "=ACMGRSVTWYHKDBN"[(i & 1) ? (b & 15u) : (b >> 4)]; i++; return true; }
};
struct RxCigar {    // cigarString(), read.d:265-276
    const uint8_t* cg; uint32_t n_ops, op_i; char buf[11]; int nd; bool started;
    BD_HD bool next(int* c) {
        if (nd == 0) {
            if (started) op_i++;
            if (op_i >= n_ops) return false;
            uint32_t raw = (uint32_t)cg[4 * op_i] | ((uint32_t)cg[4 * op_i + 1] << 8) | ((uint32_t)cg[4 * op_i + 2] << 16) | ((uint32_t)cg[4 * op_i + 3] << 24), len = raw >> 4;
            buf[0] = "MIDNSHP=X????????"[raw & 15u]; nd = 1;
            do { buf[nd++] = (char)('0' + len % 10u); len /= 10u; } while (len);
            started = true;
        }
        *c = (uint8_t)buf[--nd];
        return true;
    }
};

}  // namespace bdk
