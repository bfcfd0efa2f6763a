// host_filter.hpp -- compiles a `-F` query into the postfix program of filter.cuh (host only, control plane).
//
// Grammar restated from sambamba/utils/common/queryparser.d:271-480 and its Pratt parser
// (sambamba/utils/common/pratt_parser.d:49-64,281-300): tokens are the longest match over the symbol scanners at
// the current position (no word boundaries: "notpaired" is `not` `paired`), binding powers: comparison 110,
// `not` 100 (prefix), `and` 80, `or` 60, brackets.  Supported nodes: the 13 flag conditions, the integer fields
// (incl. avg_base_quality), [XX] tags against integers / strings / null, read_name and strand against strings,
// sequence and cigar against strings, ref_name / mate_ref_name == / != 'name' (folded into ref_id comparisons), and
// `=~ /regex/` on read_name, sequence, cigar and string tags (host_regex.hpp: the subset listed in regex.cuh).
// Refused with a message (never evaluated differently): regular expressions outside that subset or on reference names,
// ordering comparisons of reference names.  Unlike the reference, tokens left over after a complete expression are an error.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "filter.cuh"
#include "host_regex.hpp"

namespace bdk {

struct FilterCompiler {
    enum Kind { K_END, K_NULL, K_FLAG, K_IFIELD, K_SFIELD, K_TAG, K_INT, K_STR, K_REGEX, K_CMP, K_MATCH, K_AND, K_OR, K_NOT, K_OPEN, K_CLOSE };
    struct Tok { Kind k = K_END; std::string text; size_t pos = 0; };
    struct Node { Kind k; std::string s; long long v = 0; int a = -1, b = -1; bool cond = false; };

    const std::vector<std::string>& ref_names;
    std::string q; size_t pos = 0; Tok tok; std::string err;
    std::vector<Node> nodes;
    explicit FilterCompiler(const std::vector<std::string>& refs) : ref_names(refs) {}

    static const std::vector<std::string>& flagnames() {
        static const std::vector<std::string> v{"paired", "proper_pair", "unmapped", "mate_is_unmapped", "reverse_strand", "mate_is_reverse_strand", "first_of_pair",
                                                "second_of_pair", "secondary_alignment", "failed_quality_control", "duplicate", "supplementary", "chimeric"};
        return v;
    }
    static const std::vector<std::string>& int_fields() {
        static const std::vector<std::string> v{"ref_id", "position", "mapping_quality", "avg_base_quality", "sequence_length", "mate_ref_id", "mate_position", "template_length"};
        return v;
    }
    static const std::vector<std::string>& str_fields() {
        static const std::vector<std::string> v{"read_name", "sequence", "cigar", "strand", "ref_name", "mate_ref_name"};
        return v;
    }
    static int lbp(Kind k) { return k == K_CMP || k == K_MATCH ? 110 : k == K_NOT ? 100 : k == K_AND ? 80 : k == K_OR ? 60 : k == K_OPEN ? 0x7fffffff : k == K_END ? (int)0x80000000 : 0; }
    static bool is_white(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

    size_t scan_list(const std::vector<std::string>& l) const { for (auto& v : l) if (q.compare(pos, v.size(), v) == 0) return pos + v.size(); return pos; }   // first prefix, as makeScanner does
    size_t scan_lit(const char* s) const { size_t n = strlen(s); return q.compare(pos, n, s) == 0 ? pos + n : pos; }

    bool next() {       // TokenRange.popFront + front
        while (pos < q.size() && is_white(q[pos])) pos++;
        tok = Tok{}; tok.pos = pos;
        if (pos >= q.size()) { tok.k = K_END; return true; }
        size_t best = pos; Kind bk = K_END;
        auto offer = [&](size_t e, Kind k) { if (e > best || (bk != K_END && e == best && e > pos && lbp(k) > lbp(bk))) { best = e; bk = k; } };
        offer(scan_lit("null"), K_NULL);
        offer(scan_list(flagnames()), K_FLAG);
        offer(scan_list(int_fields()), K_IFIELD);
        offer(scan_list(str_fields()), K_SFIELD);
        if (q[pos] == '[' && q.size() - pos >= 4 && q[pos + 3] == ']') offer(pos + 4, K_TAG);
        {   // integer: optional sign, digits
            size_t i = pos; if (q[i] == '-' || q[i] == '+') i++;
            size_t d = i; while (i < q.size() && q[i] >= '0' && q[i] <= '9') i++;
            if (i > d) offer(i, K_INT);
        }
        if (q[pos] == '\'') {
            size_t i = pos + 1; bool closed = false;
            while (i < q.size()) { if (q[i] == '\\' && i + 1 < q.size() && q[i + 1] == '\'') i += 2; else if (q[i] == '\'') { i++; closed = true; break; } else i++; }
            if (closed) offer(i, K_STR);
        }
        if (q[pos] == '/') {
            size_t i = pos + 1; bool closed = false;
            while (i < q.size()) { if (q[i] == '\\' && i + 1 < q.size() && q[i + 1] == '/') i += 2; else if (q[i] == '/') { i++; closed = true; break; } else i++; }
            if (closed) { bool ok = true; while (i < q.size() && !is_white(q[i]) && q[i] != ')') { if (strchr("gixUms", q[i])) i++; else { ok = false; break; } } if (ok) offer(i, K_REGEX); }
        }
        for (const char* op : {">=", "<=", "==", "!=", ">", "<"}) offer(scan_lit(op), K_CMP);
        offer(scan_lit("=~"), K_MATCH);
        offer(scan_lit("and"), K_AND); offer(scan_lit("or"), K_OR); offer(scan_lit("not"), K_NOT);
        offer(scan_lit("("), K_OPEN); offer(scan_lit(")"), K_CLOSE);
        if (bk == K_END) { err = "invalid symbol in input stream at position " + std::to_string(pos); return false; }
        tok.k = bk; tok.text = q.substr(pos, best - pos); pos = best;
        return true;
    }
    int add(Node n) { nodes.push_back(n); return (int)nodes.size() - 1; }

    int parse(int rbp) {
        Tok t = tok; if (!next()) return -1;
        int left = nud(t); if (left < 0) return -1;
        while (rbp < lbp(tok.k)) {
            Tok op = tok; if (!next()) return -1;
            left = led(op, left); if (left < 0) return -1;
        }
        return left;
    }
    int nud(const Tok& t) {
        switch (t.k) {
        case K_NULL: return add(Node{K_NULL, ""});
        case K_FLAG: { Node n{K_FLAG, t.text}; n.cond = true; return add(n); }
        case K_IFIELD: return add(Node{K_IFIELD, t.text});
        case K_SFIELD: return add(Node{K_SFIELD, t.text});
        case K_TAG: return add(Node{K_TAG, t.text.substr(1, 2)});
        case K_INT: { Node n{K_INT, t.text}; n.v = strtoll(t.text.c_str(), nullptr, 10); return add(n); }
        case K_STR: { std::string s; for (size_t i = 1; i + 1 < t.text.size(); i++) { if (t.text[i] == '\\' && i + 2 < t.text.size() && t.text[i + 1] == '\'') { s += '\''; i++; } else s += t.text[i]; } return add(Node{K_STR, s}); }
        case K_REGEX: return add(Node{K_REGEX, t.text});
        case K_NOT: { int a = parse(100); if (a < 0) return -1; if (!nodes[a].cond) { err = "`not` needs a condition, got '" + nodes[a].s + "'"; return -1; } Node n{K_NOT, "not"}; n.a = a; n.cond = true; return add(n); }
        case K_OPEN: { int a = parse(0); if (a < 0) return -1; if (tok.k != K_CLOSE) { err = "unexpected character at position " + std::to_string(tok.pos); return -1; } if (!next()) return -1; return a; }
        default: err = "parsing error: unexpected '" + (t.k == K_END ? std::string("end of filter") : t.text) + "'"; return -1;
        }
    }
    int led(const Tok& op, int left) {
        if (op.k == K_AND || op.k == K_OR) {
            int r = parse(op.k == K_AND ? 80 : 60); if (r < 0) return -1;
            if (!nodes[left].cond || !nodes[r].cond) { err = "`" + op.text + "` needs two conditions"; return -1; }
            Node n{op.k, op.text}; n.a = left; n.b = r; n.cond = true; return add(n);
        }
        if (op.k == K_MATCH) {
            int r = parse(110); if (r < 0) return -1;
            if (nodes[r].k != K_REGEX) { err = "expected regular expression, not '" + nodes[r].s + "'"; return -1; }
            if (nodes[left].k != K_SFIELD && nodes[left].k != K_TAG) { err = "expected string field or tag name, not '" + nodes[left].s + "'"; return -1; }
            Node n{K_MATCH, "=~"}; n.a = left; n.b = r; n.cond = true; return add(n);
        }
        if (op.k == K_CMP) {
            int r = parse(110); if (r < 0) return -1;
            const Node& a = nodes[left]; const Node& b = nodes[r];
            if (b.k == K_INT) { if (a.k != K_TAG && a.k != K_IFIELD) { err = "expected tag or integer field name instead of '" + a.s + "'"; return -1; } }
            else if (b.k == K_STR) { if (a.k != K_TAG && a.k != K_SFIELD) { err = "expected tag or string field name instead of '" + a.s + "'"; return -1; } }
            else if (b.k == K_NULL && (op.text == "==" || op.text == "!=")) { if (a.k != K_TAG) { err = "only tag value can be compared with null"; return -1; } }
            else { err = "can't compare `" + a.s + "` and `" + b.s + "`"; return -1; }
            Node n{K_CMP, op.text}; n.a = left; n.b = r; n.cond = true; return add(n);
        }
        err = "parsing error: expected infix/postfix operator";
        return -1;
    }

    static uint8_t cmp_code(const std::string& op) { return op == ">" ? FC_GT : op == "<" ? FC_LT : op == ">=" ? FC_GE : op == "<=" ? FC_LE : op == "==" ? FC_EQ : FC_NE; }
    bool push(FilterProg& p, FilterOp o) { if (p.n >= (uint32_t)FILTER_MAX_OPS) { err = "filter expression too long for the GPU engine"; return false; } p.ops[p.n++] = o; return true; }
    bool pool(FilterProg& p, const std::string& s, FilterOp& o, size_t& used) {
        if (s.size() > 255 || used + s.size() > (size_t)FILTER_POOL) { err = "string literals of the filter are too long for the GPU engine"; return false; }
        o.s_off = (uint8_t)used; o.s_len = (uint8_t)s.size(); memcpy(p.pool + used, s.data(), s.size()); used += s.size();
        return true;
    }
    bool emit(int i, FilterProg& p, size_t& used) {
        const Node n = nodes[i];
        FilterOp o; memset(&o, 0, sizeof o);
        switch (n.k) {
        case K_FLAG: {
            static const struct { const char* nm; uint32_t bit; } F[] = {{"paired", 0x1}, {"proper_pair", 0x2}, {"unmapped", 0x4}, {"mate_is_unmapped", 0x8}, {"reverse_strand", 0x10}, {"mate_is_reverse_strand", 0x20},
                {"first_of_pair", 0x40}, {"second_of_pair", 0x80}, {"secondary_alignment", 0x100}, {"failed_quality_control", 0x200}, {"duplicate", 0x400}, {"supplementary", 0x800}};
            if (n.s == "chimeric") { o.op = FO_CHIMERIC; return push(p, o); }
            for (auto& f : F) if (n.s == f.nm) { o.op = FO_FLAG; o.imm = f.bit; return push(p, o); }
            err = "unknown flag '" + n.s + "'"; return false; }
        case K_MATCH: {
            const Node a = nodes[n.a], b = nodes[n.b];
            size_t d = b.s.size() - 1; while (d > 0 && b.s[d] != '/') d--;
            if (p.n_rx >= (uint32_t)FILTER_MAX_RX) { err = "more than two regular expressions in a filter are not available in the GPU engine"; return false; }
            RegexCompiler rc; std::string e = rc.compile(b.s.substr(1, d - 1), b.s.substr(d + 1), p.rx[p.n_rx]);
            if (!e.empty()) { err = e; return false; }
            o.op = FO_REGEX; o.s_off = (uint8_t)p.n_rx++;
            if (a.k == K_TAG) { o.a = FS_TAG; o.tag = (uint16_t)((uint8_t)a.s[0] | ((uint8_t)a.s[1] << 8)); }
            else if (a.s == "read_name") o.a = FS_NAME;
            else if (a.s == "sequence") o.a = FS_SEQ;
            else if (a.s == "cigar") o.a = FS_CIGAR;
            else { err = "regular expressions on `" + a.s + "` are not available in the GPU engine yet (compare ref_id instead)"; return false; }
            return push(p, o); }
        case K_NOT: if (!emit(n.a, p, used)) return false; o.op = FO_NOT; return push(p, o);
        case K_AND: case K_OR: if (!emit(n.a, p, used) || !emit(n.b, p, used)) return false; o.op = n.k == K_AND ? FO_AND : FO_OR; return push(p, o);
        case K_CMP: {
            const Node a = nodes[n.a], b = nodes[n.b]; o.cmp = cmp_code(n.s);
            if (a.k == K_TAG) {
                o.tag = (uint16_t)((uint8_t)a.s[0] | ((uint8_t)a.s[1] << 8));
                if (b.k == K_INT) { o.op = FO_INTTAG; o.imm = b.v; }
                else if (b.k == K_NULL) o.op = FO_TAGNULL;
                else { o.op = FO_STRTAG; if (!pool(p, b.s, o, used)) return false; }
                return push(p, o);
            }
            if (a.k == K_IFIELD) {
                o.imm = b.v;
                if (a.s == "avg_base_quality") { o.op = FO_AVGQ; return push(p, o); }
                o.op = FO_INTFIELD;
                o.a = a.s == "ref_id" ? FF_REF_ID : a.s == "position" ? FF_POSITION : a.s == "mapping_quality" ? FF_MAPQ : a.s == "sequence_length" ? FF_SEQ_LEN : a.s == "mate_ref_id" ? FF_MATE_REF_ID : a.s == "mate_position" ? FF_MATE_POSITION : FF_TLEN;
                return push(p, o);
            }
            // string fields
            if (a.s == "read_name") { o.op = FO_NAME; if (!pool(p, b.s, o, used)) return false; return push(p, o); }
            if (a.s == "strand") { if (b.s.empty()) { err = "strand must be compared with '+' or '-'"; return false; } o.op = FO_STRAND; if (!pool(p, b.s.substr(0, 1), o, used)) return false; return push(p, o); }
            if (a.s == "ref_name" || a.s == "mate_ref_name") {
                if (o.cmp != FC_EQ && o.cmp != FC_NE) { err = "ordering comparisons of reference names are not available in the GPU engine yet"; return false; }
                long id = -2; for (size_t k = 0; k < ref_names.size(); k++) if (ref_names[k] == b.s) { id = (long)k; break; }
                if (b.s == "*") id = -1;          // ref_name of an unplaced read (read.d: "*")
                if (id == -2) { o.op = FO_CONST; o.imm = o.cmp == FC_NE; return push(p, o); }
                o.op = FO_INTFIELD; o.a = a.s == "ref_name" ? FF_REF_ID : FF_MATE_REF_ID; o.imm = id; return push(p, o);
            }
            if (a.s == "sequence" || a.s == "cigar") { o.op = a.s == "sequence" ? FO_SEQ : FO_CIGAR; if (!pool(p, b.s, o, used)) return false; return push(p, o); }
            err = "comparisons of `" + a.s + "` are not available in the GPU engine yet"; return false; }
        default: err = "filter string must represent a condition"; return false;
        }
    }

    // returns "" on success
    std::string compile(const std::string& query, FilterProg& out) {
        memset(&out, 0, sizeof out); q = query; pos = 0; nodes.clear(); err.clear();
        if (!next()) return err;
        int root = parse(0);
        if (root < 0) return err.empty() ? "cannot parse the filter" : err;
        if (tok.k != K_END) return "unexpected '" + tok.text + "' at position " + std::to_string(tok.pos) + " of the filter";
        if (!nodes[root].cond) return "filter string must represent a condition";
        size_t used = 0;
        if (!emit(root, out, used)) return err;
        return "";
    }
};

}  // namespace bdk
