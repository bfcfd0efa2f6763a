// mates.cuh -- fix-mate-overlaps (`-m`) as a correction pass over the plain counters.
//
// Replaces: detectOverlappingMates (sambamba/depth.d:319-388), selectBetterMate (:391-399), the `-m` branches of
// PerBasePrinter.writeColumn (:521-530) and of PerRegionPrinter.push (:760-845: countPreviouslySeenMateOverlaps,
// uncountOverlappingMates :717-743, the per-column n_bases/coverage loop :808-841), CustomBamRead's FNV-1a name
// hash (:252-258) and, because it decides which mates meet at all, the overlap test of BamReadFilter
// (BioD/bio/std/hts/bam/randomaccessmanager.d:397-461).
//
// The reference sorts every column's reads by name hash and walks neighbours.  Here the plain counters (K3) are
// computed first and the mate logic only takes out what the reference would not have counted:
//   km_hash  (thread per read)    in-stream test (passes the filter, overlaps a -L region if there are any) and the
//                                 64-bit FNV-1a of the read name
//   km_link  (thread per read)    reads are sorted by start, so a read that overlaps read r starts before r ends:
//                                 scan forward over those, mark same-hash reads as "has a predecessor"
//   km_fix   (thread per leader)  a leader (same-hash successor, no predecessor) collects its component (the chain
//                                 of overlapping same-hash reads).  Two members (the normal pair): walk the columns
//                                 both cover, pick the better mate exactly as selectBetterMate does and take the
//                                 loser's contribution out of the counter planes; in region mode also derive the
//                                 region statistics the reference's per-column state machine would have produced
//                                 (see mate_pair_regions).  More members (supplementary / secondary alignments of
//                                 one name): replay the reference's detected/past state machine for that name.
//
// Everything here is __host__ __device__ and free of warp intrinsics: tests/emul/emul_mates.cpp runs the very same
// functions on the CPU, thread by thread (in several orders), against the oracle's faithful column sweep.  The
// product only ever calls them from the kernels at the end of this file.
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

constexpr int MATE_MAX_MEMBERS = 8;        // reads of one name hash present in ONE column (a component may have any number of members)
enum MateErr : int { MATE_OK = 0, MATE_ERR_TOO_MANY = 1, MATE_ERR_CROSS = 2, MATE_ERR_ZONE = 3 };
constexpr uint32_t MATE_NCL_FOREIGN = 1u << 30;    // ... of another rank's shard (kernels.cuh NCL_FOREIGN): seen, never the leader of a component this rank fixes
constexpr uint32_t MATE_NCL_GHOST = 1u << 31;      // RecordSoA.ncl bit of a re-read record of the previous batch (kernels.cuh NCL_GHOST)
enum MateFlag : uint32_t { MF_INSTREAM = 1u, MF_HAS_PRED = 2u, MF_HAS_SUCC = 4u };

struct MateParams {
    // records of the batch (the SoA K2 wrote) and the inflated bytes they point into
    const uint64_t* start; const uint32_t* span; const uint32_t* meta; const int64_t* off; const uint32_t* ncl; const int32_t* lseq;
    const uint8_t* u; uint32_t R;
    uint64_t* mhash; uint32_t* mflag;
    // -L: merged regions in linear coordinates, sorted, disjoint (n_flt == 0: every passing read is in the stream)
    const uint64_t* flt_s; const uint64_t* flt_e; uint32_t n_flt;
    // counter planes: [S][7][win_len], position cnt_base + i at index i
    uint32_t* counts; uint64_t cnt_base, win_len; uint32_t S; uint32_t minq;
    // region mode (n_seg == 0: base mode): segments sorted by start with prefix maxima of their ends, as k_read_segments
    // uses them; seg_reads is decremented for pairs the plain count saw as two reads, seg_mbases collects what n_bases
    // has on top of the A/C/G/T/N planes
    const uint64_t* seg_s; const uint64_t* seg_e; const uint64_t* seg_pmax; const uint32_t* seg_id; uint32_t n_seg;
    uint32_t* seg_reads; uint32_t* seg_mbases; uint32_t n_samples_out;
    // overlapping windows (nullptr / 0 otherwise): a ring slot is updated from seg_u[k] <= seg_s[k] on (the reference
    // updates all slots once position >= window, depth.d:215-226, so the per-column terms of a window start collecting
    // before the window does), and reference 0's first slots only ever count reads that start at or after seg_qmin[k]
    // (is_first_occurrence starts false, depth.d:1031-1032).  Same order as seg_s.
    const uint64_t* seg_u; const uint64_t* seg_qmin; uint64_t seg_ext_max;
    // K3's read index, for the one question that needs the other reads of a column (mate_follows): per 1024-position
    // tile the first passing short read overlapping it, and the list of long reads
    const uint32_t* tile_lo; uint64_t tiles_base; uint32_t n_tiles; const uint32_t* long_list; uint32_t n_long;
    // Several batches.  The first n_ghost rows are records of the previous batch, re-read (pass bit clear, ncl bit 31 set if
    // they pass the filter) because they can still meet a mate: a component (chain of overlapping same-hash reads) is
    // fixed in the batch in which it closes -- its reach does not go beyond s_last, the start of the batch's last record
    // (later records start there or further right) -- and a component made of ghosts only that had already closed in the
    // previous batch (reach <= prev_s_last) is left alone.  What has to be re-read next time -- the leaders of open
    // components and every read that ends after s_last -- lowers *open_off (offset relative to u).
    // km_cover then adds every read that reaches into the columns of something open, so that the next batch knows all
    // reads of those columns (covered_from = the previous batch's *open_start; mate_follows needs them).
    uint32_t n_ghost; uint64_t s_last, prev_s_last, covered_from; int last_batch; unsigned long long* open_off; unsigned long long* open_start;
    // Several ranks: a rank's stream is [zone of the previous shard | its shard | zone of the next shard]; a component is
    // fixed by the rank that owns its leader.  stream_cut: the stream ends before the file does (a component of an own
    // leader that is still open there is refused); *fix_max_end collects how far the fixed components reach (the halo
    // exchange must carry the counters up to there).
    int stream_cut; unsigned long long* fix_max_end; uint32_t n_right;      // n_right: the last rows are the next rank's zone
    int64_t old_below;           // records whose off lies below were part of the previous batch too (INT64_MIN in the first batch)
    int force_general;           // tests: send pairs through the state-machine path as well
    int* err;                    // err[0] = MateErr, err[1] = record index
    unsigned long long* stat;    // [0] pairs, [1] (pair, column) fixes, [2] components with more than two members
};

#if defined(__CUDA_ARCH__)
BD_HD void m_add(uint32_t* p, uint32_t v) { atomicAdd(p, v); }
BD_HD void m_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
BD_HD void m_stat(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
BD_HD void m_err(const MateParams& p, int code, uint32_t r) { if (atomicMax(p.err, code) < code) p.err[1] = (int)r; }
BD_HD void m_min(unsigned long long* p, unsigned long long v) { atomicMin(p, v); }
BD_HD void m_max(unsigned long long* p, unsigned long long v) { atomicMax(p, v); }
#else
BD_HD void m_add(uint32_t* p, uint32_t v) { *p += v; }
BD_HD void m_or(uint32_t* p, uint32_t v) { *p |= v; }
BD_HD void m_stat(unsigned long long* p, unsigned long long v) { *p += v; }
BD_HD void m_err(const MateParams& p, int code, uint32_t r) { if (p.err[0] < code) { p.err[0] = code; p.err[1] = (int)r; } }
BD_HD void m_min(unsigned long long* p, unsigned long long v) { if (v < *p) *p = v; }
BD_HD void m_max(unsigned long long* p, unsigned long long v) { if (v > *p) *p = v; }
#endif

BD_HD uint32_t m_ld32(const uint8_t* q) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); }
BD_HD bool m_rcons(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }      // cigar.d:116
BD_HD bool m_qcons(uint32_t op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }
BD_HD bool m_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }

struct MRead {
    uint64_t s, e; uint32_t span, mapq, sample, l_name, n_cigar, lseq;
    const uint8_t* name; const uint8_t* cg; const uint8_t* seq; const uint8_t* qual;
};
BD_HD void m_load(const MateParams& p, uint32_t r, MRead& m) {
    m.s = p.start[r]; m.span = p.span[r]; m.e = m.s + m.span;
    uint32_t mt = p.meta[r]; m.mapq = (mt >> 8) & 0xFFu; m.sample = (mt >> 2) & 63u;
    uint32_t ncl = p.ncl[r]; m.n_cigar = (ncl >> 8) & 0xFFFFu; m.l_name = ncl & 0xFFu;
    int32_t ls = p.lseq[r]; m.lseq = ls > 0 ? (uint32_t)ls : 0u;
    const uint8_t* rec = p.u + p.off[r];
    m.name = rec + 32; m.cg = m.name + m.l_name; m.seq = m.cg + 4u * m.n_cigar; m.qual = m.seq + (m.lseq + 1) / 2;
}
// the reference compares sample id and name (depth.d:352-353)
BD_HD bool m_same_name(const MRead& a, const MRead& b) {
    if (a.sample != b.sample || a.l_name != b.l_name) return false;
    for (uint32_t i = 0; i < a.l_name; i++) if (a.name[i] != b.name[i]) return false;
    return true;
}

// CIGAR cursor for column-by-column walks: m_at(x) with non-decreasing reference offsets x (relative to the read start)
struct MCur { uint32_t i = 0, rpos = 0, qpos = 0, len = 0, op = 0; bool loaded = false; };
enum { MK_NONE = 0, MK_BASE = 1, MK_DEL = 2, MK_SKIP = 3 };
BD_HD int m_at(const MRead& m, MCur& c, uint32_t x, uint32_t* q) {
    if (x >= m.span) return MK_NONE;
    for (;;) {
        if (c.i >= m.n_cigar) return MK_NONE;
        if (!c.loaded) { uint32_t raw = m_ld32(m.cg + 4u * c.i); c.len = raw >> 4; c.op = raw & 15u; c.loaded = true; }
        if (m_rcons(c.op)) {
            if (x < c.rpos + c.len) break;
            c.rpos += c.len; if (m_qcons(c.op)) c.qpos += c.len;
        } else if (m_qcons(c.op)) c.qpos += c.len;
        c.i++; c.loaded = false;
    }
    if (m_match(c.op)) { *q = c.qpos + (x - c.rpos); return MK_BASE; }
    return c.op == 2 ? MK_DEL : MK_SKIP;
}
// counter plane a read adds to in a column (what K3 counted for it), -1: nothing (base.d:186 for the nt16 -> nt5 map)
BD_HD int m_plane(const MRead& m, int kind, uint32_t q, uint32_t minq) {
    if (kind == MK_DEL) return 5;
    if (kind == MK_SKIP) return 6;
    if (kind != MK_BASE || q >= m.lseq || m.qual[q] < minq) return -1;
    uint32_t b = m.seq[q >> 1], nib = (q & 1) ? (b & 15u) : (b >> 4);
    return nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : 4;
}
// current_base_quality (pileup.d:125-134): 255 on D/N
BD_HD uint32_t m_qual(const MRead& m, int kind, uint32_t q) { return kind == MK_BASE ? (q < m.lseq ? m.qual[q] : 0u) : 255u; }
// selectBetterMate (depth.d:391-399); a is the earlier read of the pair in file order.  true: a wins
BD_HD bool m_first_wins(const MRead& a, int ka, uint32_t qa, const MRead& b, int kb, uint32_t qb) {
    if (ka != MK_BASE || kb != MK_BASE) return a.mapq > b.mapq;
    return m_qual(a, ka, qa) > m_qual(b, kb, qb);
}
BD_HD uint32_t* m_planes(const MateParams& p, uint32_t sample) { return p.counts + (uint64_t)(p.S > 1 ? sample : 0u) * 7u * p.win_len; }
BD_HD void m_count(const MateParams& p, uint32_t* planes, int plane, uint64_t g, uint32_t delta) {
    if (plane < 0 || g < p.cnt_base || g - p.cnt_base >= p.win_len) return;
    m_add(planes + (uint64_t)plane * p.win_len + (g - p.cnt_base), delta);
}

// ---------------------------------------------------------------------------------------- km_hash
BD_HD void mate_hash_one(const MateParams& p, uint32_t r) {
    uint32_t fl = 0; uint64_t h = 0;
    if ((p.meta[r] & 1u) || (p.ncl[r] & MATE_NCL_GHOST)) {
        bool in = true;
        if (p.n_flt) {      // first merged region that ends after the read starts; the read is kept iff it reaches it
            uint64_t s = p.start[r], e = s + p.span[r];
            uint32_t lo = 0, hi = p.n_flt;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (p.flt_e[mid] <= s) lo = mid + 1; else hi = mid; }
            in = lo < p.n_flt && p.flt_s[lo] < e;
        }
        if (in) {
            uint32_t l_name = p.ncl[r] & 0xFFu; const uint8_t* nm = p.u + p.off[r] + 32;
            h = 14695981039346656037ull;
            for (uint32_t i = 0; i + 1 < l_name; i++) { h ^= nm[i]; h *= 1099511628211ull; }
            fl = MF_INSTREAM;
        }
    }
    p.mhash[r] = h; p.mflag[r] = fl;
}

// ---------------------------------------------------------------------------------------- km_link
BD_HD void mate_link_one(const MateParams& p, uint32_t r) {
    if (!(p.mflag[r] & MF_INSTREAM)) return;
    const uint64_t e = p.start[r] + p.span[r], h = p.mhash[r];
    bool any = false;
    for (uint32_t k = r + 1; k < p.R && p.start[k] < e; k++)
        if (p.mhash[k] == h && (p.mflag[k] & MF_INSTREAM)) { m_or(&p.mflag[k], MF_HAS_PRED); any = true; }
    if (any) m_or(&p.mflag[r], MF_HAS_SUCC);
    if (!p.last_batch && e > p.s_last) { m_min(p.open_off, (unsigned long long)p.off[r]); m_min(p.open_start, p.start[r]); }      // a later batch may bring its mate
}

// ---------------------------------------------------------------------------------------- km_fix
// Number of M/=/X bases of a read inside [a, b) with quality >= minq (countOverlappingBases, depth.d:671-698)
BD_HD uint32_t mate_full(const MRead& m, uint64_t a, uint64_t b, uint32_t minq) {
    uint32_t rpos = 0, qpos = 0, n = 0;
    for (uint32_t i = 0; i < m.n_cigar; i++) {
        uint32_t raw = m_ld32(m.cg + 4u * i), len = raw >> 4, op = raw & 15u;
        if (m_match(op)) {
            uint64_t ma = m.s + rpos, mb = ma + len; if (mb > m.e) mb = m.e;
            uint64_t xa = ma > a ? ma : a, xb = mb < b ? mb : b;
            for (uint64_t g = xa; g < xb; g++) { uint32_t q = qpos + (uint32_t)(g - ma); if (q < m.lseq && m.qual[q] >= minq) n++; }
            rpos += len; qpos += len;
        } else if (op == 2 || op == 3) rpos += len;
        else if (m_qcons(op)) qpos += len;
    }
    return n;
}

// Region statistics of one pair (region mode).  Derivation (depth.d:760-845; states none/detected/fixed/past):
//   * a pair is `detected` in the first column both mates cover, d = start of the later mate, and `fixed` in the first
//     such column that lies in a region; from then on neither mate is counted through countRead any more, every column
//     adds one base for the better mate if its quality (255 on D/N) reaches -q, and when one mate ends the other is
//     `past` and adds one base per column under the same rule.  uncountOverlappingMates takes the bases both mates
//     had from that column on out again.  Net: n_bases of a region = sum over its columns of the -m A/C/G/T/N counters
//     (which the counter fix below produces) + one for every column where the counted read of the pair sits on a
//     D or N  [term a];
//   * a survivor that is already `past` in the first column of a region is still handed to countRead there (only
//     `fixed` reads are skipped, depth.d:804), so its bases in that region count twice  [term b];
//   * n_reads: the two mates are one read for every region that contains a column both cover
//     (countPreviouslySeenMateOverlaps / uncountOverlappingMates, depth.d:728-742).
BD_HD void mate_pair_regions(const MateParams& p, const MRead& A, const MRead& B) {
    const uint64_t sB = B.s, eMin = A.e < B.e ? A.e : B.e, eMax = A.e < B.e ? B.e : A.e;
    const MRead& Sv = A.e > B.e ? A : B;                     // the survivor (if the ends differ)
    const uint32_t samp = p.n_samples_out > 1 ? A.sample : 0u;
    uint32_t lo = 0, hi = p.n_seg;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (p.seg_s[mid] < eMax) lo = mid + 1; else hi = mid; }
    for (int64_t k = (int64_t)lo - 1; k >= 0; k--) {
        if (p.seg_pmax[k] <= sB) break;
        const uint64_t a = p.seg_s[k], b = p.seg_e[k];
        if (b <= sB || a >= eMax) continue;
        uint32_t extra = 0;
        const uint64_t pa = a > sB ? a : sB, pb = b < eMin ? b : eMin;          // columns of the region both mates cover
        const uint64_t ta = a > eMin ? a : eMin, tb = b < eMax ? b : eMax;      // columns only the survivor covers
        if (pa < pb) {
            MCur ca, cb;
            for (uint64_t g = pa; g < pb; g++) {
                uint32_t qa = 0, qb = 0; int ka = m_at(A, ca, (uint32_t)(g - A.s), &qa), kb = m_at(B, cb, (uint32_t)(g - B.s), &qb);
                bool fw = m_first_wins(A, ka, qa, B, kb, qb);
                int kw = fw ? ka : kb;
                if (kw == MK_DEL || kw == MK_SKIP) extra++;
            }
            if (mate_full(A, a, b, p.minq) > 0 && mate_full(B, a, b, p.minq) > 0) m_add(&p.seg_reads[(uint64_t)samp * p.n_seg + p.seg_id[k]], 0xFFFFFFFFu);
        }
        if (ta < tb) {
            MCur cs;
            for (uint64_t g = ta; g < tb; g++) { uint32_t q = 0; int ks = m_at(Sv, cs, (uint32_t)(g - Sv.s), &q); if (ks == MK_DEL || ks == MK_SKIP) extra++; }
            if (a >= eMin) extra += mate_full(Sv, a, b, p.minq);      // the region begins while the survivor is `past`
        }
        if (extra) m_add(&p.seg_mbases[(uint64_t)samp * p.n_seg + p.seg_id[k]], extra);
    }
}

// The normal case: two reads of one name.  In every column both cover only the better mate counts.
BD_HD void mate_fix_pair(const MateParams& p, const MRead& A, const MRead& B) {
    const uint64_t sB = B.s, eMin = A.e < B.e ? A.e : B.e;
    uint32_t* planes = m_planes(p, A.sample);
    MCur ca, cb; unsigned long long cols = 0;
    for (uint64_t g = sB; g < eMin; g++) {
        uint32_t qa = 0, qb = 0; int ka = m_at(A, ca, (uint32_t)(g - A.s), &qa), kb = m_at(B, cb, (uint32_t)(g - B.s), &qb);
        bool fw = m_first_wins(A, ka, qa, B, kb, qb);
        int pl = fw ? m_plane(B, kb, qb, p.minq) : m_plane(A, ka, qa, p.minq);      // the loser's contribution goes
        m_count(p, planes, pl, g, 0xFFFFFFFFu);
        cols++;
    }
    m_stat(&p.stat[0], 1); m_stat(&p.stat[1], cols);
    if (p.n_seg) mate_pair_regions(p, A, B);
}

// Is a read with a larger name hash present in column g?  Then a read of hash h is not the last entry of the column's
// hash-sorted array (depth.d:380-384 treats the last entry differently).  Same search as k3_gather: the short reads from
// the tile's first overlapping one up to the last that starts at or before g, plus the long reads.
BD_HD bool mate_follows(const MateParams& p, uint64_t h, uint64_t g, uint32_t who) {
    // the re-read records hold every read that reaches column covered_from or beyond (km_cover); left of it the column is not fully known
    if (p.n_ghost && g < p.covered_from) { m_err(p, MATE_ERR_CROSS, who); return true; }
    for (uint32_t k = 0; k < p.n_ghost && p.start[k] <= g; k++)           // re-read records are not in K3's index
        if ((p.mflag[k] & MF_INSTREAM) && p.start[k] + p.span[k] > g && p.mhash[k] > h) return true;
    for (uint32_t k = p.R - p.n_right; k < p.R && p.start[k] <= g; k++)   // nor are the records of the next rank's zone
        if ((p.mflag[k] & MF_INSTREAM) && p.start[k] + p.span[k] > g && p.mhash[k] > h) return true;
    if (g >= p.tiles_base) {
        uint64_t t = (g - p.tiles_base) >> 10;
        if (t < p.n_tiles) {
            uint32_t lo = p.tile_lo[t];
            if (lo != 0xFFFFFFFFu)
                for (uint32_t k = lo; k < p.R && p.start[k] <= g; k++)
                    if ((p.meta[k] & 3u) == 1u && (p.mflag[k] & MF_INSTREAM) && p.start[k] + p.span[k] > g && p.mhash[k] > h) return true;
        }
    }
    for (uint32_t i = 0; i < p.n_long; i++) {
        uint32_t k = p.long_list[i];
        if ((p.mflag[k] & MF_INSTREAM) && p.start[k] <= g && p.start[k] + p.span[k] > g && p.mhash[k] > h) return true;
    }
    return false;
}

// Three or more overlapping reads of one name hash: replay detectOverlappingMates for this name column by column.
// In a column the present reads of the name are adjacent in the hash-sorted array, in file order; neighbours pair up
// (first with second, third with fourth); a read that was flagged before and is alone again becomes `past`.  A read
// in state `detected` / `fixed` is skipped by the plain loop, every pair adds its better mate (depth.d:521-530) -- so a
// `past` read that pairs again counts twice, as in the reference.
//
// Region mode (n_seg > 0) replays PerRegionPrinter.push (depth.d:760-845) for the members as well and books the
// difference between what the reference accumulates and what the reducers will compute from the counters and the
// plain per-read count:
//   reference  countRead of a member in the first column of the region it is present in (not if `fixed` by then),
//              countPreviouslySeenMateOverlaps in the region's first column, uncountOverlappingMates for pairs not yet
//              `fixed`, one base per column for a `past` member / a pair's better mate whose quality reaches -q
//              (255 on D/N); pair members become `fixed` after a column that lies in a region;
//   reducers   n_bases = A/C/G/T/N counters of the region's columns, n_reads = members with a counted base in it.
// Nothing here needs reads outside the component: a member that started before a region is present in the region's
// first column, so that column is the region's first visited one.
//
// The component may be any length (a chain of supplementary alignments, a name that many reads share): the walk keeps only the
// members that are present in the current column -- slots are handed out when the column reaches a member's start and taken back
// when it passes its end; `pres` lists the occupied slots in file order, which is the order the reference pairs neighbours in.  A
// member that has ended is never looked at again (its state cannot matter any more), so nothing but the number of reads of one
// name in ONE column is bounded (MATE_MAX_MEMBERS).
BD_HD void mate_fix_group(const MateParams& p, uint32_t leader, uint64_t hi) {
    MRead M[MATE_MAX_MEMBERS]; MCur C[MATE_MAX_MEMBERS]; uint8_t st[MATE_MAX_MEMBERS];      // 0 none, 1 detected, 2 fixed, 3 past
    int pres[MATE_MAX_MEMBERS]; int np = 0; uint32_t used = 0;      // pres[0..np): the slots of the members present in column g, in file order
    const uint64_t lo = p.start[leader], h = p.mhash[leader];
    uint32_t next = leader;                                          // first record not yet looked at
    unsigned long long cols = 0;
    uint32_t seg_hi = 0;
    if (p.n_seg) { uint32_t l = 0, h2 = p.n_seg; while (l < h2) { uint32_t mid = (l + h2) >> 1; if (p.seg_s[mid] < hi + p.seg_ext_max) l = mid + 1; else h2 = mid; } seg_hi = l; }   // segments that are updated before the component ends
    // Per-base output with -L: PerBasePrinter.push hands a column to writeColumn -- and with it to detectOverlappingMates -- only when it
    // lies in a region (depth.d:567-591), so the states do not move on the columns in between (a `detected` read that is alone there is
    // still `detected` when the next region begins).  Region and window mode run the detection on every column (depth.d:760-770).
    const bool gated = p.n_flt != 0 && p.n_seg == 0;
    uint32_t fi = 0;
    if (gated) { uint32_t l = 0, h2 = p.n_flt; while (l < h2) { uint32_t mid = (l + h2) >> 1; if (p.flt_e[mid] <= lo) l = mid + 1; else h2 = mid; } fi = l; }
    for (uint64_t g = lo; g < hi; g++) {
        if (gated) {
            while (fi < p.n_flt && p.flt_e[fi] <= g) fi++;
            if (fi >= p.n_flt) break;                                   // no written column is left
            if (p.flt_s[fi] > g) { g = (p.flt_s[fi] < hi ? p.flt_s[fi] : hi) - 1; continue; }      // on to the next region's first column
        }
        { int w = 0; for (int i = 0; i < np; i++) { const int k = pres[i]; if (g < M[k].e) pres[w++] = k; else used &= ~(1u << k); } np = w; }      // members that ended
        for (; next < p.R && p.start[next] <= g; next++) {                                                                              // members that begin here
            if (p.mhash[next] != h || !(p.mflag[next] & MF_INSTREAM)) continue;
            if (p.start[next] + p.span[next] <= g) continue;             // (began and ended between two regions: never seen by a written column)
            if (np == MATE_MAX_MEMBERS) { m_err(p, MATE_ERR_TOO_MANY, next); return; }
            int k = 0; while ((used >> k) & 1u) k++;
            used |= 1u << k; m_load(p, next, M[k]); C[k] = MCur(); st[k] = 0; pres[np++] = k;
        }
        int kind[MATE_MAX_MEMBERS]; uint32_t q[MATE_MAX_MEMBERS];
        for (int i = 0; i < np; i++) { const int k = pres[i]; q[k] = 0; kind[k] = m_at(M[k], C[k], (uint32_t)(g - M[k].s), &q[k]); }
        int pa[MATE_MAX_MEMBERS / 2], pb[MATE_MAX_MEMBERS / 2], npairs = 0;
        for (int i = 0; i < np;) {
            if (i + 1 < np) {
                int a = pres[i], b = pres[i + 1];
                if (m_same_name(M[a], M[b])) { pa[npairs] = a; pb[npairs] = b; npairs++; if (!st[a]) st[a] = 1; if (!st[b]) st[b] = 1; i += 2; }
                else i += 1;                      // same hash, other name or sample: the reference moves on, states untouched
                continue;
            }
            int a = pres[i];
            if (st[a]) {
                // Alone again: `past` -- unless it is the very last entry of the column's sorted array and its predecessor has
                // the same hash; then the reference leaves it as it is (depth.d:380-384).
                if (!(np >= 2 && st[a] != 3 && !mate_follows(p, h, g, leader))) st[a] = 3;
            }
            i += 1;
        }
        int win[MATE_MAX_MEMBERS / 2];
        for (int i = 0; i < np; i++) { int k = pres[i]; if (st[k] == 1 || st[k] == 2) m_count(p, m_planes(p, M[k].sample), m_plane(M[k], kind[k], q[k], p.minq), g, 0xFFFFFFFFu); }
        for (int j = 0; j < npairs; j++) {
            int a = pa[j], b = pb[j]; bool fw = m_first_wins(M[a], kind[a], q[a], M[b], kind[b], q[b]); int w = fw ? a : b; win[j] = w;
            m_count(p, m_planes(p, M[w].sample), m_plane(M[w], kind[w], q[w], p.minq), g, 1u);
        }
        cols += (unsigned long long)npairs;
        if (p.n_seg) {
            bool in_region = false;
            for (int64_t k = (int64_t)seg_hi - 1; k >= 0; k--) {
                if (p.seg_pmax[k] <= g) break;
                const uint64_t a = p.seg_s[k], b = p.seg_e[k];
                const uint64_t ua = p.seg_u ? p.seg_u[k] : a;                            // first column in which the slot is updated
                const uint64_t qmin = p.seg_qmin ? p.seg_qmin[k] : 0;                    // != 0: the slot never sees a first occurrence
                if (g < ua || g >= b) continue;
                in_region = true;
                const bool in_w = g >= a, quirk = qmin != 0;
                const uint64_t slot = p.seg_id[k];
#define M_SAMP(X) ((uint64_t)(p.n_samples_out > 1 ? M[X].sample : 0u) * p.n_seg + slot)
                for (int i = 0; i < np; i++) {               // countRead where the member enters the region
                    int x = pres[i];
                    if (g != (ua > M[x].s ? ua : M[x].s)) continue;
                    if (quirk && M[x].s < qmin) continue;                                       // present before the slot's first update: never handed to countRead, not in the plain count either
                    uint32_t f = mate_full(M[x], a, b, p.minq);
                    if (f) m_add(&p.seg_reads[M_SAMP(x)], 0xFFFFFFFFu);                         // the plain count had it as a read of its own
                    if (quirk) m_add(&p.seg_mbases[M_SAMP(x)], 0u - f);                         // ... and, for such a slot, its bases (k_read_segments' own sum)
                    if (st[x] != 2) { m_add(&p.seg_mbases[M_SAMP(x)], f); if (f) m_add(&p.seg_reads[M_SAMP(x)], 1u); }
                }
                for (int j = 0; j < npairs; j++) {
                    int x = pa[j], y = pb[j];
                    uint32_t fx = mate_full(M[x], a, b, p.minq), fy = mate_full(M[y], a, b, p.minq);
                    if (g == ua && !quirk && st[x] == 2) { if (fx + fy) m_add(&p.seg_reads[M_SAMP(x)], 1u); }     // countPreviouslySeenMateOverlaps
                    if (!(st[x] == 2 && st[y] == 2)) {                                                     // uncountOverlappingMates
                        uint32_t nx = M[x].s == g ? fx : mate_full(M[x], a > g ? a : g, b, p.minq), ny = M[y].s == g ? fy : mate_full(M[y], a > g ? a : g, b, p.minq);
                        m_add(&p.seg_mbases[M_SAMP(x)], 0u - (nx + ny));
                        m_add(&p.seg_reads[M_SAMP(x)], 0u - (uint32_t)((fx > 0) + (fy > 0)) + (uint32_t)(fx + fy > 0));
                    }
                }
                for (int i = 0; i < np; i++) {               // per column: what the reference adds, minus what the counters hold
                    int x = pres[i]; int pl = m_plane(M[x], kind[x], q[x], p.minq);
                    if (st[x] == 3 && m_qual(M[x], kind[x], q[x]) >= p.minq) m_add(&p.seg_mbases[M_SAMP(x)], 1u);
                    if (in_w && !quirk && (st[x] == 0 || st[x] == 3) && pl >= 0 && pl <= 4) m_add(&p.seg_mbases[M_SAMP(x)], 0xFFFFFFFFu);
                }
                for (int j = 0; j < npairs; j++) {
                    int w = win[j]; int pl = m_plane(M[w], kind[w], q[w], p.minq);
                    if (m_qual(M[w], kind[w], q[w]) >= p.minq) m_add(&p.seg_mbases[M_SAMP(w)], 1u);
                    if (in_w && !quirk && pl >= 0 && pl <= 4) m_add(&p.seg_mbases[M_SAMP(w)], 0xFFFFFFFFu);
                }
#undef M_SAMP
            }
            if (in_region) for (int j = 0; j < npairs; j++) { st[pa[j]] = 2; st[pb[j]] = 2; }
        }
    }
    m_stat(&p.stat[1], cols); m_stat(&p.stat[2], 1);
}

BD_HD void mate_fix_one(const MateParams& p, uint32_t r) {
    if ((p.mflag[r] & (MF_INSTREAM | MF_HAS_PRED | MF_HAS_SUCC)) != (MF_INSTREAM | MF_HAS_SUCC)) return;
    // the component: every same-hash read that starts before the running end of the members found so far (any number of them)
    uint32_t second = r, last = r; int n = 1;
    uint64_t reach = p.start[r] + p.span[r]; const uint64_t h = p.mhash[r];
    for (uint32_t k = r + 1; k < p.R && p.start[k] < reach; k++) {
        if (p.mhash[k] != h || !(p.mflag[k] & MF_INSTREAM)) continue;
        if (n == 1) second = k;
        last = k; n++;
        uint64_t e = p.start[k] + p.span[k]; if (e > reach) reach = e;
    }
    if (!p.last_batch && reach > p.s_last) { m_min(p.open_off, (unsigned long long)p.off[r]); m_min(p.open_start, p.start[r]); return; }      // still open: the batch that closes it fixes it
    if (p.ncl[r] & MATE_NCL_FOREIGN) return;                                                                    // another rank owns the leader
    if (p.stream_cut && reach > p.s_last) { m_err(p, MATE_ERR_ZONE, r); return; }                               // runs out of the zone read behind the shard
    if (p.fix_max_end) m_max(p.fix_max_end, reach);
    if (p.off[last] < p.old_below && reach <= p.prev_s_last) return;                                             // only records the previous batch has seen, closed there: already fixed
    if (n == 2 && !p.force_general && !p.seg_u) {
        MRead A, B; m_load(p, r, A); m_load(p, second, B);
        if (m_same_name(A, B)) mate_fix_pair(p, A, B);
        return;
    }
    if (n < 2) return;
    mate_fix_group(p, r, reach);
}

// every read that reaches into the columns of an open component has to be re-read with it (after km_link and km_fix)
BD_HD void mate_cover_one(const MateParams& p, uint32_t r) {
    if ((p.mflag[r] & MF_INSTREAM) && p.start[r] + p.span[r] > *p.open_start) m_min(p.open_off, (unsigned long long)p.off[r]);
}

#if defined(__CUDACC__) || defined(BDEPTH_EMULATE_SHIM)
__global__ void km_cover(MateParams p) { uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; if (r < p.R) mate_cover_one(p, r); }
__global__ void km_hash(MateParams p) { uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; if (r < p.R) mate_hash_one(p, r); }
__global__ void km_link(MateParams p) { uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; if (r < p.R) mate_link_one(p, r); }
__global__ void km_fix(MateParams p) { uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; if (r < p.R) mate_fix_one(p, r); }
#endif

}  // namespace bdk
