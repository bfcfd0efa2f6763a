// host_bam.hpp -- host-side container parsing for libbdepth: BGZF block framing, BAM header text,
// BAI linear index.  Control plane only; no record or DEFLATE work happens on the CPU.
//
// Restates: BGZF framing  BioD/bio/core/bgzf/inputstream.d:54-199, constants.d:26-61
//           BAM header    BioD/bio/std/hts/bam/reader.d:101-125,580-598
//           SAM @HD/@RG   BioD/bio/std/hts/sam/header.d:461-530 ; sample table depth.d:1170-1181
//           BAI           BioD/bio/std/hts/bam/baifile.d:126-169 (parse), :77-82 (ioffsets)
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

namespace bdk {

struct HostBlock {
    uint64_t coff;        // offset of the block in the file
    uint32_t cdata_off;   // offset of the raw deflate payload inside the block
    uint32_t csize;       // payload bytes
    uint32_t isize;       // inflated bytes
    uint32_t bsize;       // whole block bytes
    uint64_t uoff;        // offset in the concatenated inflated stream
};

static inline uint16_t h_rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t h_rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t h_rd64(const uint8_t* p) { return (uint64_t)h_rd32(p) | ((uint64_t)h_rd32(p + 4) << 32); }

// Walk the BGZF members of [file, file+len).  Stops at the first block with ISIZE == 0 (EOF
// marker), as BgzfInputStream.fillNextBlock does (inputstream.d:393).  Returns "" or an error.
// frame_bgzf is the resumable form: it starts at *off_io / *uoff_io, appends at most max_blocks members that begin at or
// before stop_coff, leaves the position after the last one in *off_io / *uoff_io and sets *eof when the file (or the EOF
// marker) has been reached.  Region queries frame only the members inside their BAI chunks with it.
static inline std::string frame_bgzf(const uint8_t* file, size_t len, size_t* off_io, uint64_t* uoff_io, size_t max_blocks, uint64_t stop_coff, std::vector<HostBlock>& out, bool* eof) {
    size_t off = *off_io; uint64_t uoff = *uoff_io; size_t n_new = 0;
    char msg[256];
    *eof = false;
    while (true) {
        if (off >= len || len - off < 4) { *eof = true; break; }
        if (n_new >= max_blocks || (uint64_t)off > stop_coff) break;
        const uint8_t* p = file + off;
        if (!(p[0] == 0x1f && p[1] == 0x8b && p[2] == 0x08 && p[3] == 0x04)) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: wrong BGZF magic", off); return msg; }
        if (len - off < 12) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: stream error", off); return msg; }
        uint32_t xlen = h_rd16(p + 10);
        if (len - off < 12 + (size_t)xlen) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: stream error", off); return msg; }
        uint32_t l = 0, bsize = 0; bool found = false;
        while (l + 4 <= xlen) {
            uint32_t slen = h_rd16(p + 14 + l);
            if (p[12 + l] == 66 && p[13 + l] == 67) {
                if (slen != 2) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: wrong BC subfield length: %u; expected 2", off, slen); return msg; }
                if (found) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: duplicate field with block size", off); return msg; }
                bsize = h_rd16(p + 16 + l); found = true;
            }
            l += 4 + slen;
        }
        if (l != xlen) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: total length of subfields in bytes (%u) is not equal to gzip_extra_length (%u)", off, l, xlen); return msg; }
        if (!found) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: block size was not found in any subfield", off); return msg; }
        int64_t cdata = (int64_t)bsize - (int64_t)xlen - 19;
        if (cdata < 0 || cdata > 65536) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: compressed data size is more than 65536 bytes, which is not allowed by current BAM specification", off); return msg; }
        size_t total = (size_t)bsize + 1;
        if (len - off < total) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: stream error: not enough data in stream", off); return msg; }
        uint32_t isize = h_rd32(p + total - 4);
        if (isize == 0) { *eof = true; break; }
        if (isize > 65536) { snprintf(msg, sizeof msg, "Error reading BGZF block starting from offset %zu: input size is more than 65536", off); return msg; }
        out.push_back(HostBlock{(uint64_t)off, 12 + xlen, (uint32_t)cdata, isize, (uint32_t)total, uoff});
        uoff += isize; off += total; n_new++;
    }
    *off_io = off; *uoff_io = uoff;
    return "";
}
static inline std::string index_bgzf(const uint8_t* file, size_t len, std::vector<HostBlock>& out, uint64_t* total_u) {
    size_t off = 0; uint64_t uoff = 0; bool eof = false;
    std::string e = frame_bgzf(file, len, &off, &uoff, SIZE_MAX, UINT64_MAX, out, &eof);
    *total_u = uoff;
    return e;
}

struct BamHeader {
    std::string text;
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_len;
    std::vector<uint64_t> ref_lin0;      // linear coordinate of position 0 of each reference
    uint64_t total_len = 0;
    uint64_t first_rec_off = 0;          // offset of the first record in the inflated stream
    bool so_coordinate = false;
    std::vector<std::string> sample_names;  // distinct SM values in @RG order ("*" when none)
    std::vector<std::string> rg_ids; std::vector<int> rg_sample;
};

// Parse from the first n inflated bytes.  Returns 0 ok, 1 need more bytes, -1 format error.
static inline int parse_bam_header(const uint8_t* u, size_t n, BamHeader& h, std::string& err) {
    if (n < 12) return 1;
    if (memcmp(u, "BAM\1", 4)) { err = "Invalid file format: expected BAM\\1"; return -1; }
    uint32_t l_text = h_rd32(u + 4);
    if (8 + (size_t)l_text + 4 > n) return 1;
    h.text.assign((const char*)u + 8, l_text);
    size_t off = 8 + (size_t)l_text;
    int32_t n_ref = (int32_t)h_rd32(u + off); off += 4;
    if (n_ref < 0) { err = "invalid BAM header (n_ref < 0)"; return -1; }
    h.ref_names.clear(); h.ref_len.clear(); h.ref_lin0.clear();
    uint64_t lin = 0;
    for (int i = 0; i < n_ref; i++) {
        if (off + 4 > n) return 1;
        uint32_t l_name = h_rd32(u + off); off += 4;
        if (off + (size_t)l_name + 4 > n) return 1;
        h.ref_names.emplace_back((const char*)u + off, l_name ? l_name - 1 : 0); off += l_name;
        uint32_t L = h_rd32(u + off); off += 4;
        h.ref_len.push_back(L); h.ref_lin0.push_back(lin); lin += L;
    }
    h.total_len = lin; h.first_rec_off = off;
    // SAM text
    h.so_coordinate = false; h.sample_names.clear(); h.rg_ids.clear(); h.rg_sample.clear();
    const std::string& t = h.text; size_t p = 0;
    while (p < t.size()) {
        size_t e = t.find('\n', p); if (e == std::string::npos) e = t.size();
        if (e - p >= 3 && t[p] == '@') {
            bool is_hd = !t.compare(p, 3, "@HD"), is_rg = !t.compare(p, 3, "@RG");
            if (is_hd || is_rg) {
                std::string id, sm; size_t q = p + 3;
                while (q < e) {
                    if (t[q] == '\t') { q++; continue; }
                    size_t fe = t.find('\t', q); if (fe == std::string::npos || fe > e) fe = e;
                    if (fe - q >= 3 && t[q + 2] == ':') {
                        if (is_hd && !t.compare(q, 3, "SO:")) h.so_coordinate = (t.compare(q + 3, fe - q - 3, "coordinate") == 0);
                        if (is_rg && !t.compare(q, 3, "ID:")) id = t.substr(q + 3, fe - q - 3);
                        if (is_rg && !t.compare(q, 3, "SM:")) sm = t.substr(q + 3, fe - q - 3);
                    }
                    q = fe;
                }
                if (is_rg) {
                    int sid = -1;
                    for (size_t i = 0; i < h.sample_names.size(); i++) if (h.sample_names[i] == sm) sid = (int)i;
                    if (sid < 0) { sid = (int)h.sample_names.size(); h.sample_names.push_back(sm); }
                    h.rg_ids.push_back(id); h.rg_sample.push_back(sid);
                }
            }
        }
        p = e + 1;
    }
    if (h.sample_names.empty()) h.sample_names.push_back("*");
    return 0;
}

// BAI: only the linear index (ioffsets) and the per-reference chunk extents are needed here.
struct BaiChunk { uint64_t beg, end; };             // BGZF virtual offsets [beg, end)
struct BaiBin { uint32_t bin; std::vector<BaiChunk> chunks; };
struct BaiIndex {
    bool valid = false;
    std::vector<std::vector<uint64_t>> ioffsets;   // per reference
    std::vector<uint64_t> min_chunk_beg;           // per reference, smallest chunk_beg (UINT64_MAX if none)
    std::vector<std::vector<BaiBin>> bins;         // per reference (the metadata pseudo-bin 37450 is dropped)
};
static inline bool parse_bai(const uint8_t* b, size_t n, BaiIndex& idx) {
    if (n < 8 || memcmp(b, "BAI\1", 4)) return false;
    size_t off = 4; int32_t n_ref = (int32_t)h_rd32(b + off); off += 4;
    if (n_ref < 0) return false;
    idx.ioffsets.assign(n_ref, {}); idx.min_chunk_beg.assign(n_ref, UINT64_MAX); idx.bins.assign(n_ref, {});
    for (int r = 0; r < n_ref; r++) {
        if (off + 4 > n) return false;
        uint32_t n_bin = h_rd32(b + off); off += 4;
        for (uint32_t k = 0; k < n_bin; k++) {
            if (off + 8 > n) return false;
            uint32_t bin = h_rd32(b + off), n_chunk = h_rd32(b + off + 4); off += 8;
            if (off + 16ull * n_chunk > n) return false;
            if (bin != 37450) {
                BaiBin bb; bb.bin = bin; bb.chunks.resize(n_chunk);
                for (uint32_t c = 0; c < n_chunk; c++) { uint64_t beg = h_rd64(b + off + 16ull * c); bb.chunks[c] = BaiChunk{beg, h_rd64(b + off + 16ull * c + 8)}; if (beg < idx.min_chunk_beg[r]) idx.min_chunk_beg[r] = beg; }
                idx.bins[r].push_back(std::move(bb));
            }
            off += 16ull * n_chunk;
        }
        if (off + 4 > n) return false;
        uint32_t n_intv = h_rd32(b + off); off += 4;
        if (off + 8ull * n_intv > n) return false;
        idx.ioffsets[r].resize(n_intv);
        for (uint32_t k = 0; k < n_intv; k++) idx.ioffsets[r][k] = h_rd64(b + off + 8ull * k);
        off += 8ull * n_intv;
    }
    idx.valid = true;
    return true;
}


// The chunk list a region query has to read, as the reference computes it (getGroupChunks,
// BioD/bio/std/hts/bam/randomaccessmanager.d:247-294): the bins that can overlap a region (5 levels over the 16 kb
// leaves), their chunks that end after the linear-index offset of the first region, sorted, overlaps merged.
// `regions` = (ref, start, end) sorted by (ref, start), start < end.  A superset is always safe: reads that do
// not overlap a region contribute nothing to what is printed for the regions.
struct HostRegion { uint32_t ref, start, end; };
static inline std::vector<BaiChunk> region_chunks(const BaiIndex& bai, const std::vector<HostRegion>& regions) {
    std::vector<BaiChunk> cs;
    size_t i = 0;
    std::vector<uint8_t> sel(37450);
    while (i < regions.size()) {
        size_t j = i; uint32_t ref = regions[i].ref;
        while (j < regions.size() && regions[j].ref == ref) j++;
        if (ref < bai.bins.size() && !bai.bins[ref].empty()) {
            std::fill(sel.begin(), sel.end(), 0); sel[0] = 1;
            for (size_t r = i; r < j; r++) {
                uint32_t beg = regions[r].start, end = regions[r].end - 1; uint32_t k;
                for (k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) sel[k] = 1;
                for (k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) sel[k] = 1;
                for (k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) sel[k] = 1;
                for (k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) sel[k] = 1;
                for (k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) if (k < 37450) sel[k] = 1;
            }
            // linear index: no read that overlaps the first region starts before this offset (0 = unknown)
            uint64_t min_off = 0; const auto& lin = bai.ioffsets[ref]; size_t w = regions[i].start >> 14;
            if (w < lin.size()) min_off = lin[w];
            for (const BaiBin& b : bai.bins[ref]) if (b.bin < 37450 && sel[b.bin]) for (const BaiChunk& c : b.chunks) if (c.end > min_off && c.beg < c.end) cs.push_back(BaiChunk{std::max(c.beg, min_off), c.end});
        }
        i = j;
    }
    std::sort(cs.begin(), cs.end(), [](const BaiChunk& a, const BaiChunk& b) { return a.beg != b.beg ? a.beg < b.beg : a.end < b.end; });
    std::vector<BaiChunk> m;
    for (const BaiChunk& c : cs) { if (!m.empty() && c.beg <= m.back().end) m.back().end = std::max(m.back().end, c.end); else m.push_back(c); }
    return m;
}

}  // namespace bdk
