// host_regex.hpp -- compiles the pattern of a `-F` regular expression into the NFA of regex.cuh (host only).
// Recursive descent over the subset listed in regex.cuh, Thompson construction, counted repetitions by cloning.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "regex.cuh"

namespace bdk {

struct RegexCompiler {
    struct Node { int kind; int c = 0; int a = -1, b = -1; int lo = 0, hi = 0; uint32_t set[8]; };      // kinds below
    enum { N_EMPTY, N_CHAR, N_ANY, N_SET, N_CAT, N_ALT, N_REP, N_BOL, N_EOL, N_WB, N_NWB };
    std::string pat, err; size_t pos = 0; bool icase = false, dotall = false;
    std::vector<Node> nodes;
    RegexProg* out = nullptr;

    int add(const Node& n) { nodes.push_back(n); return (int)nodes.size() - 1; }
    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }
    bool more() const { return pos < pat.size(); }

    static void set_add(uint32_t* s, int c) { s[(c >> 5) & 7] |= 1u << (c & 31); }
    static void set_range(uint32_t* s, int a, int b) { for (int c = a; c <= b; c++) set_add(s, c); }
    static void set_class(uint32_t* s, char k) {      // \d \w \s
        if (k == 'd') set_range(s, '0', '9');
        else if (k == 'w') { set_range(s, '0', '9'); set_range(s, 'a', 'z'); set_range(s, 'A', 'Z'); set_add(s, '_'); }
        else { for (int c : {' ', '\t', '\n', '\r', '\f', '\v'}) set_add(s, c); }
    }
    static void set_invert(uint32_t* s) { for (int i = 0; i < 8; i++) s[i] = ~s[i]; }

    int parse_alt() {
        int left = parse_cat(); if (left < 0) return -1;
        while (more() && pat[pos] == '|') { pos++; int r = parse_cat(); if (r < 0) return -1; Node n{N_ALT}; n.a = left; n.b = r; left = add(n); }
        return left;
    }
    int parse_cat() {
        int left = add(Node{N_EMPTY});
        while (more() && pat[pos] != '|' && pat[pos] != ')') { int r = parse_rep(); if (r < 0) return -1; Node n{N_CAT}; n.a = left; n.b = r; left = add(n); }
        return left;
    }
    int parse_rep() {
        int a = parse_atom(); if (a < 0) return -1;
        while (more()) {
            int lo, hi; char c = pat[pos];
            if (c == '*') { lo = 0; hi = -1; pos++; }
            else if (c == '+') { lo = 1; hi = -1; pos++; }
            else if (c == '?') { lo = 0; hi = 1; pos++; }
            else if (c == '{') {
                size_t q = pos + 1; int v = 0, n1 = 0; while (q < pat.size() && isdigit((unsigned char)pat[q])) { v = v * 10 + (pat[q] - '0'); q++; n1++; if (v > 1000) { fail("repetition count too large for the GPU engine"); return -1; } }
                if (!n1) { fail("malformed {m,n} in the regular expression"); return -1; }
                lo = hi = v;
                if (q < pat.size() && pat[q] == ',') { q++; int w = 0, n2 = 0; while (q < pat.size() && isdigit((unsigned char)pat[q])) { w = w * 10 + (pat[q] - '0'); q++; n2++; if (w > 1000) { fail("repetition count too large for the GPU engine"); return -1; } } hi = n2 ? w : -1; }
                if (q >= pat.size() || pat[q] != '}' || (hi >= 0 && hi < lo)) { fail("malformed {m,n} in the regular expression"); return -1; }
                pos = q + 1;
            } else break;
            if (more() && pat[pos] == '?') pos++;        // lazy: the same language
            int k = nodes[a].kind; if (k == N_BOL || k == N_EOL || k == N_WB || k == N_NWB) { fail("a quantifier after an anchor is not available in the GPU engine"); return -1; }
            Node n{N_REP}; n.a = a; n.lo = lo; n.hi = hi; a = add(n);
        }
        return a;
    }
    bool parse_escape_into_set(uint32_t* s, bool* single, int* ch) {      // after '\\'
        if (!more()) return fail("dangling backslash in the regular expression");
        char e = pat[pos++]; *single = false;
        switch (e) {
        case 'd': case 'w': case 's': set_class(s, e); return true;
        case 'D': case 'W': case 'S': { uint32_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; set_class(t, (char)(e + 32)); set_invert(t); for (int i = 0; i < 8; i++) s[i] |= t[i]; return true; }
        case 't': *single = true; *ch = '\t'; return true;
        case 'n': *single = true; *ch = '\n'; return true;
        case 'r': *single = true; *ch = '\r'; return true;
        case 'f': *single = true; *ch = '\f'; return true;
        case 'v': *single = true; *ch = '\v'; return true;
        case 'x': { int v = 0; for (int k = 0; k < 2; k++) { if (!more() || !isxdigit((unsigned char)pat[pos])) return fail("malformed \\xHH in the regular expression"); char h = pat[pos++]; v = v * 16 + (isdigit((unsigned char)h) ? h - '0' : (tolower(h) - 'a' + 10)); } *single = true; *ch = v; return true; }
        default:
            if (isalnum((unsigned char)e)) return fail(std::string("\\") + e + " in a regular expression is not available in the GPU engine");      // back-references, \p, \b inside a class, ...
            *single = true; *ch = (unsigned char)e; return true;
        }
    }
    int parse_atom() {
        char c = pat[pos];
        if (c == '(') {
            pos++;
            if (more() && pat[pos] == '?') { if (pos + 1 < pat.size() && pat[pos + 1] == ':') pos += 2; else { fail("look-around / named groups in regular expressions are not available in the GPU engine"); return -1; } }
            int a = parse_alt(); if (a < 0) return -1;
            if (!more() || pat[pos] != ')') { fail("unbalanced parenthesis in the regular expression"); return -1; }
            pos++; return a;
        }
        if (c == '*' || c == '+' || c == '?' || c == '{') { fail("quantifier without an operand in the regular expression"); return -1; }
        pos++;
        if (c == '.') return add(Node{N_ANY});
        if (c == '^') return add(Node{N_BOL});
        if (c == '$') return add(Node{N_EOL});
        if (c == '[') {
            Node n{N_SET}; memset(n.set, 0, sizeof n.set); bool neg = false;
            if (more() && pat[pos] == '^') { neg = true; pos++; }
            bool first = true;
            while (true) {
                if (!more()) { fail("unterminated character class in the regular expression"); return -1; }
                char d = pat[pos];
                if (d == ']' && !first) { pos++; break; }
                first = false;
                if (d == '[' || (d == '&' && pos + 1 < pat.size() && pat[pos + 1] == '&') || (d == '-' && pos + 1 < pat.size() && pat[pos + 1] == '-') || (d == '~' && pos + 1 < pat.size() && pat[pos + 1] == '~')) { fail("set operations / nested classes in regular expressions are not available in the GPU engine"); return -1; }
                int lo; bool single = true;
                if (d == '\\') { pos++; if (!parse_escape_into_set(n.set, &single, &lo)) return -1; if (!single) continue; }
                else { lo = (unsigned char)d; pos++; }
                if (more() && pat[pos] == '-' && pos + 1 < pat.size() && pat[pos + 1] != ']') {
                    pos++; int hi; char e = pat[pos];
                    if (e == '\\') { pos++; bool s2; if (!parse_escape_into_set(n.set, &s2, &hi)) return -1; if (!s2) { fail("malformed range in a character class"); return -1; } }
                    else { hi = (unsigned char)e; pos++; }
                    if (hi < lo) { fail("malformed range in a character class"); return -1; }
                    set_range(n.set, lo, hi);
                } else set_add(n.set, lo);
            }
            if (icase) for (int ch = 'a'; ch <= 'z'; ch++) { bool l = (n.set[ch >> 5] >> (ch & 31)) & 1, u = (n.set[(ch - 32) >> 5] >> ((ch - 32) & 31)) & 1; if (l || u) { set_add(n.set, ch); set_add(n.set, ch - 32); } }
            if (neg) set_invert(n.set);
            return add(n);
        }
        if (c == '\\') {
            if (more() && pat[pos] == 'b') { pos++; return add(Node{N_WB}); }
            if (more() && pat[pos] == 'B') { pos++; return add(Node{N_NWB}); }
            Node n{N_SET}; memset(n.set, 0, sizeof n.set); bool single; int ch;
            if (!parse_escape_into_set(n.set, &single, &ch)) return -1;
            if (single) { Node m{N_CHAR}; m.c = ch; return add(m); }
            return add(n);
        }
        Node m{N_CHAR}; m.c = (unsigned char)c; return add(m);
    }

    // ---- emission (Thompson): emit(node) appends code that falls through to the next instruction on success
    bool put(RxInst in) { if (out->n >= RX_MAX_INST - 1) return fail("regular expression too large for the GPU engine (more than 64 NFA states)"); out->in[out->n++] = in; return true; }
    bool emit(int i) {
        const Node n = nodes[i];
        switch (n.kind) {
        case N_EMPTY: return true;
        case N_CHAR: if (icase && isalpha(n.c)) return put(RxInst{RX_CHARI, (uint8_t)tolower(n.c), 0, 0}); return put(RxInst{RX_CHAR, (uint8_t)n.c, 0, 0});
        case N_ANY: return put(RxInst{(uint8_t)(dotall ? RX_ANYNL : RX_ANY), 0, 0, 0});
        case N_SET: {
            int k = -1; for (int j = 0; j < out->n_cls; j++) if (!memcmp(out->cls[j], n.set, 32)) k = j;
            if (k < 0) { if (out->n_cls >= RX_MAX_CLS) return fail("regular expression too large for the GPU engine (more than 4 character classes)"); k = out->n_cls++; memcpy(out->cls[k], n.set, 32); }
            return put(RxInst{RX_CLS, (uint8_t)k, 0, 0}); }
        case N_BOL: return put(RxInst{RX_BOL, 0, 0, 0});
        case N_EOL: return put(RxInst{RX_EOL, 0, 0, 0});
        case N_WB: return put(RxInst{RX_WB, 0, 0, 0});
        case N_NWB: return put(RxInst{RX_NWB, 0, 0, 0});
        case N_CAT: return emit(n.a) && emit(n.b);
        case N_ALT: {      // split L1 L2 ; L1: a ; jmp end ; L2: b ; end:
            int s = out->n; if (!put(RxInst{RX_SPLIT, 0, 0, 0})) return false;
            out->in[s].x = (uint8_t)out->n; if (!emit(n.a)) return false;
            int j = out->n; if (!put(RxInst{RX_JMP, 0, 0, 0})) return false;
            out->in[s].y = (uint8_t)out->n; if (!emit(n.b)) return false;
            out->in[j].x = (uint8_t)out->n; return true; }
        case N_REP: {
            for (int k = 0; k < n.lo; k++) if (!emit(n.a)) return false;
            if (n.hi < 0) {        // a* after the mandatory copies: L: split body end ; body: a ; jmp L ; end:
                int s = out->n; if (!put(RxInst{RX_SPLIT, 0, 0, 0})) return false;
                out->in[s].x = (uint8_t)out->n; if (!emit(n.a)) return false;
                if (!put(RxInst{RX_JMP, 0, (uint8_t)s, 0})) return false;
                out->in[s].y = (uint8_t)out->n; return true;
            }
            std::vector<int> splits;      // (a?){hi-lo}, nested so that skipping one skips the rest
            for (int k = n.lo; k < n.hi; k++) { splits.push_back(out->n); if (!put(RxInst{RX_SPLIT, 0, 0, 0})) return false; out->in[splits.back()].x = (uint8_t)out->n; if (!emit(n.a)) return false; }
            for (int s : splits) out->in[s].y = (uint8_t)out->n;
            return true; }
        }
        return fail("internal: unknown regular expression node");
    }

    // returns "" on success
    std::string compile(const std::string& pattern, const std::string& options, RegexProg& prog) {
        memset(&prog, 0, sizeof prog); out = &prog; pat = pattern; pos = 0; nodes.clear(); err.clear(); icase = dotall = false;
        for (char o : options) { if (o == 'i') icase = true; else if (o == 's') dotall = true; else if (o == 'g') {} else return std::string("regular expression flag '") + o + "' is not available in the GPU engine"; }
        for (unsigned char ch : pattern) if (ch >= 128) return "non-ASCII regular expressions are not available in the GPU engine";
        int root = pattern.empty() ? add(Node{N_EMPTY}) : parse_alt();
        if (root < 0) return err.empty() ? "cannot parse the regular expression" : err;
        if (more()) return pat[pos] == ')' ? "unbalanced parenthesis in the regular expression" : "cannot parse the regular expression";
        if (!emit(root)) return err;
        out->in[out->n++] = RxInst{RX_MATCH, 0, 0, 0};
        // jump targets may point one past the last emitted instruction: that is the MATCH just added
        return "";
    }
};

}  // namespace bdk
