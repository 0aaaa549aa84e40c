// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// The reference's den-graph loader (src/ctc_crf/gpu_den/fst_read.cc) needs OpenFst 1.6.7, which
// is not vendored and cannot be downloaded here.  This shim provides the one symbol the unmodified
// reference den_calculate.cu leaves undefined -- ReadFst, declared at den_calculate.cu:275-285 --
// by parsing the OpenFst binary "vector"/"standard" container directly (layout: SURVEY.md 8c).
// Semantics follow fst_read.cc:23-59: per-state in/out arc lists in file order, weight = -tropical,
// label = ilabel-1, start weight 0 on Start(), end weight = -Final for non-Zero finals.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern int DEN_NUM_STATES;
extern int DEN_NUM_ARCS;

namespace {
struct Reader {
    std::vector<unsigned char> buf;
    size_t off = 0;
    template <typename T> T get() {
        if (off + sizeof(T) > buf.size()) { fprintf(stderr, "ref ReadFst shim: truncated file\n"); exit(2); }
        T v; memcpy(&v, buf.data() + off, sizeof(T)); off += sizeof(T); return v;
    }
    std::string str() {
        int32_t n = get<int32_t>();
        std::string s((const char *)buf.data() + off, n); off += n; return s;
    }
};
}  // namespace

void ReadFst(const char *fst_name,
             std::vector<std::vector<int> > &alpha_next,
             std::vector<std::vector<int> > &beta_next,
             std::vector<std::vector<int> > &alpha_ilabel,
             std::vector<std::vector<int> > &beta_ilabel,
             std::vector<std::vector<float> > &alpha_weight,
             std::vector<std::vector<float> > &beta_weight,
             std::vector<float> &start_weight,
             std::vector<float> &end_weight,
             int &num_states,
             int &num_arcs) {
    Reader r;
    FILE *f = fopen(fst_name, "rb");
    if (!f) { fprintf(stderr, "ref ReadFst shim: cannot open %s\n", fst_name); exit(2); }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    r.buf.resize(sz);
    if (fread(r.buf.data(), 1, sz, f) != (size_t)sz) { fprintf(stderr, "ref ReadFst shim: short read\n"); exit(2); }
    fclose(f);
    if ((uint32_t)r.get<int32_t>() != 0x7EB2FDD6u) { fprintf(stderr, "ref ReadFst shim: bad magic\n"); exit(2); }
    std::string ft = r.str(), at = r.str();
    if (ft != "vector" || at != "standard") { fprintf(stderr, "ref ReadFst shim: unsupported type\n"); exit(2); }
    r.get<int32_t>();                    // version
    int32_t flags = r.get<int32_t>();
    if (flags & 3) { fprintf(stderr, "ref ReadFst shim: symbol tables unsupported\n"); exit(2); }
    r.get<uint64_t>();                   // properties
    int64_t start = r.get<int64_t>();
    int64_t ns = r.get<int64_t>();
    r.get<int64_t>();                    // header arc count (not trusted)
    num_states = (int)ns;
    DEN_NUM_STATES = num_states;
    alpha_next.assign(ns, {}); beta_next.assign(ns, {});
    alpha_ilabel.assign(ns, {}); beta_ilabel.assign(ns, {});
    alpha_weight.assign(ns, {}); beta_weight.assign(ns, {});
    start_weight.assign(ns, -float(INFINITY));
    end_weight.assign(ns, -float(INFINITY));
    start_weight[start] = 0.f;
    num_arcs = 0;
    for (int64_t s = 0; s < ns; ++s) {
        float fin = r.get<float>();
        int64_t na = r.get<int64_t>();
        if (!std::isinf(fin)) end_weight[s] = -fin;
        for (int64_t a = 0; a < na; ++a) {
            int32_t il = r.get<int32_t>();
            r.get<int32_t>();            // olabel
            float w = r.get<float>();
            int32_t nx = r.get<int32_t>();
            beta_next[s].push_back(nx);       alpha_next[nx].push_back((int)s);
            beta_ilabel[s].push_back(il - 1); alpha_ilabel[nx].push_back(il - 1);
            beta_weight[s].push_back(-w);     alpha_weight[nx].push_back(-w);
            ++num_arcs;
        }
    }
    DEN_NUM_ARCS = num_arcs;
}
