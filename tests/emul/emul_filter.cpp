// Host emulation harness (TEST ONLY): the `-F` compiler (host_filter.hpp, product host code) and the per-record
// evaluator k2_decode calls (filter.cuh) on the CPU.  Never linked into libbdepth.so.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "../../sambamba_b200/csrc/host_filter.hpp"
using namespace bdk;

// ref_names: '\n'-separated.  Returns 0 and fills *out, or 1 with the message in err.
extern "C" int emul_filter_compile(const char* query, const char* ref_names, FilterProg* out, char* err, size_t cap) {
    std::vector<std::string> refs; { std::string s(ref_names); size_t p = 0; while (p <= s.size() && !s.empty()) { size_t e = s.find('\n', p); if (e == std::string::npos) e = s.size(); refs.push_back(s.substr(p, e - p)); p = e + 1; if (e == s.size()) break; } }
    FilterCompiler fc(refs);
    std::string m = fc.compile(query, *out);
    if (!m.empty()) { snprintf(err, cap, "%s", m.c_str()); return 1; }
    return 0;
}
extern "C" size_t emul_filter_prog_size() { return sizeof(FilterProg); }
// offs[i]: offset of record i's refID field inside u; sizes[i]: its block_size
extern "C" void emul_filter_eval(const FilterProg* p, const uint8_t* u, const uint64_t* offs, const uint32_t* sizes, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; i++) out[i] = filter_eval(*p, u + offs[i], sizes[i]) ? 1 : 0;
}
// one regular expression against one subject: 1 / 0, or -1 when the pattern is outside the supported subset
extern "C" int emul_regex_search(const char* pattern, const char* options, const uint8_t* subject, uint32_t n) {
    RegexProg p; RegexCompiler rc;
    if (!rc.compile(pattern, options, p).empty()) return -1;
    RxBytes g{subject, n, 0u};
    return rx_search(p, g) ? 1 : 0;
}
