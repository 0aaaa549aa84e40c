// Host-side den-graph loader and kernel plan.  See den_graph.h.
#include "den_graph.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>

namespace ccb {

namespace {

struct ByteReader {
    const unsigned char *p;
    size_t n, off = 0;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if (off + sizeof(T) > n) { ok = false; return v; }
        memcpy(&v, p + off, sizeof(T));
        off += sizeof(T);
        return v;
    }
    std::string str() {
        int32_t len = get<int32_t>();
        if (!ok || len < 0 || off + (size_t)len > n) { ok = false; return {}; }
        std::string s((const char *)p + off, (size_t)len);
        off += (size_t)len;
        return s;
    }
};

}  // namespace

// OpenFst binary container, "vector" fst of "standard" (tropical) arcs.  Layout: SURVEY.md 8c.
// Semantics: fst_read.cc:40-59 (weight negated, ilabel shifted by one, -inf for Zero finals).
bool ReadFstFile(const char *path, HostFst *out, std::string *err) {
    FILE *f = fopen(path, "rb");
    if (!f) { *err = std::string("cannot open den graph: ") + path; return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> buf((size_t)std::max(sz, 0L));
    size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if ((long)got != sz) { *err = std::string("short read: ") + path; return false; }

    ByteReader r{buf.data(), buf.size()};
    if ((uint32_t)r.get<int32_t>() != 0x7EB2FDD6u) { *err = std::string("not an OpenFst binary file: ") + path; return false; }
    std::string fst_type = r.str(), arc_type = r.str();
    if (!r.ok || fst_type != "vector" || arc_type != "standard") {
        *err = "unsupported fst/arc type '" + fst_type + "'/'" + arc_type + "' (need vector/standard)";
        return false;
    }
    (void)r.get<int32_t>();  // version
    int32_t flags = r.get<int32_t>();
    if (flags & 0x3) { *err = "den graph carries embedded symbol tables; strip them (fstsymbols --clear_*)"; return false; }
    (void)r.get<uint64_t>();  // properties
    int64_t start = r.get<int64_t>();
    int64_t ns = r.get<int64_t>();
    (void)r.get<int64_t>();   // header arc count: 0 in files written by fstcompile pipelines, not trusted
    if (!r.ok || ns <= 0 || ns > (int64_t)0x7fffffff || start < 0 || start >= ns) {
        *err = "corrupt den graph header";
        return false;
    }
    out->num_states = (int)ns;
    out->start = (int)start;
    out->final_logw.assign((size_t)ns, -std::numeric_limits<float>::infinity());
    out->src.clear(); out->dst.clear(); out->label.clear(); out->logw.clear();
    for (int64_t s = 0; s < ns; ++s) {
        float fin = r.get<float>();
        int64_t na = r.get<int64_t>();
        if (!r.ok || na < 0 || r.off + (size_t)na * 16 > r.n) { *err = "truncated den graph"; return false; }
        if (!std::isinf(fin)) out->final_logw[(size_t)s] = -fin;
        for (int64_t a = 0; a < na; ++a) {
            int32_t il = r.get<int32_t>();
            (void)r.get<int32_t>();  // olabel
            float w = r.get<float>();
            int32_t nx = r.get<int32_t>();
            if (nx < 0 || nx >= ns || il < 1) { *err = "den graph arc out of range (epsilon input labels are not allowed)"; return false; }
            out->src.push_back((int)s);
            out->dst.push_back(nx);
            out->label.push_back(il - 1);
            out->logw.push_back(-w);
        }
    }
    if (r.off != r.n) { *err = "trailing bytes in den graph"; return false; }
    if (out->src.size() > (size_t)0x3fffffff) { *err = "den graph too large"; return false; }
    return true;
}

namespace {

inline float NegateBits(float w) {
    uint32_t b;
    memcpy(&b, &w, 4);
    b |= 0x80000000u;
    memcpy(&w, &b, 4);
    return w;
}

// Cut rows into n_chunks contiguous chunks of near-equal cost and emit the chunk-major padded arc stream.
// cost(row) = quads(row) * 4 + kRowCost.
void Layout(const std::vector<std::vector<Arc>> &rows, const std::vector<int> &state_label, int n_ctas, int n_warps,
            PassPlan *pp) {
    constexpr int64_t kRowCost = 2;   // a row end costs about two arcs' worth of a (latency-bound) warp's time
    const int S = (int)rows.size();
    const int n_chunks = n_ctas * n_warps;
    auto quads = [&](int q) { return std::max<int64_t>(1, ((int64_t)rows[(size_t)q].size() + kQuad - 1) / kQuad); };
    std::vector<int64_t> prefix((size_t)S + 1, 0);
    pp->real_arcs = 0;
    for (int q = 0; q < S; ++q) {
        prefix[q + 1] = prefix[q] + quads(q) * kQuad + kRowCost;
        pp->real_arcs += (int)rows[(size_t)q].size();
    }
    const int64_t total = prefix[S];
    pp->chunk_state.assign((size_t)n_chunks + 1, 0);
    for (int c = 1; c < n_chunks; ++c) {
        int64_t target = (total * c + n_chunks / 2) / n_chunks;
        int q = (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
        q = std::min(std::max(q, pp->chunk_state[c - 1]), S);
        pp->chunk_state[c] = q;
    }
    pp->chunk_state[n_chunks] = S;

    pp->arcs.clear();
    pp->chunk_arc.assign((size_t)n_chunks + 1, 0);
    uint32_t last_peer = 0;
    for (int c = 0; c < n_chunks; ++c) {
        pp->chunk_arc[c] = (int)pp->arcs.size();
        for (int q = pp->chunk_state[c]; q < pp->chunk_state[c + 1]; ++q) {
            const auto &r = rows[(size_t)q];
            const size_t padded = (size_t)quads(q) * kQuad;
            // Padding arcs carry weight 0 but are still gathered: point them at a row this warp reads anyway (the
            // previous arc's peer, or the row itself).  Pointing them all at one fixed row would make every warp of
            // the grid hammer a single L2 line (measured: 2x frame time from that hot spot alone).
            uint32_t pad_peer = r.empty() ? (uint32_t)q : r.back().peer;
            for (size_t i = 0; i < padded; ++i) {
                Arc a = i < r.size() ? r[i] : Arc{pad_peer, 0.f};
                if (i + 1 == padded) a.w = NegateBits(a.w);   // last quad of the row
                pp->arcs.push_back(a);
            }
            last_peer = pad_peer;
        }
        while (pp->arcs.size() % kChunkArcPad) pp->arcs.push_back(Arc{last_peer, 0.f});
    }
    pp->chunk_arc[n_chunks] = (int)pp->arcs.size();
    pp->max_tile_arcs = 0;
    pp->max_tile_labels = 1;
    pp->max_tile_rows = 1;
    for (int c = 0; c < n_ctas; ++c) {
        int a0 = pp->chunk_arc[(size_t)c * n_warps], a1 = pp->chunk_arc[(size_t)(c + 1) * n_warps];
        pp->max_tile_arcs = std::max(pp->max_tile_arcs, a1 - a0);
        int s0 = pp->chunk_state[(size_t)c * n_warps], s1 = pp->chunk_state[(size_t)(c + 1) * n_warps];
        if (s1 > s0) pp->max_tile_labels = std::max(pp->max_tile_labels, state_label[s1 - 1] - state_label[s0] + 1);
        pp->max_tile_rows = std::max(pp->max_tile_rows, s1 - s0);
    }
}

}  // namespace

bool BuildDenPlan(const HostFst &fst, int n_ctas, int n_warps, DenPlan *plan, std::string *err) {
    const int S0 = fst.num_states;
    const size_t A0 = fst.src.size();
    if (n_ctas < 1 || n_warps < 1) { *err = "bad grid for den plan"; return false; }

    // 1. distinct in-labels per file state
    std::vector<std::vector<int>> in_labels((size_t)S0);
    for (size_t a = 0; a < A0; ++a) in_labels[(size_t)fst.dst[a]].push_back(fst.label[a]);
    size_t S = 0;
    int max_label = 0;
    for (int q = 0; q < S0; ++q) {
        auto &v = in_labels[(size_t)q];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        if (v.empty()) v.push_back(0);  // no in-arcs (e.g. the start state): the label is never used
        S += v.size();
        max_label = std::max(max_label, v.back());
    }
    // split-arc count and a sanity bound on the blow-up
    size_t A = 0;
    for (size_t a = 0; a < A0; ++a) A += in_labels[(size_t)fst.src[a]].size();
    if (S > (size_t)0x3fffffff || A > (size_t)0x3fffffff) {
        *err = "den graph too large after in-label state split (S=" + std::to_string(S) + ", A=" + std::to_string(A) + ")";
        return false;
    }

    // 2. renumber (label, file state) sorted by label
    struct Copy { int label, orig; };
    std::vector<Copy> copies;
    copies.reserve(S);
    for (int q = 0; q < S0; ++q) for (int k : in_labels[(size_t)q]) copies.push_back(Copy{k, q});
    std::stable_sort(copies.begin(), copies.end(), [](const Copy &a, const Copy &b) { return a.label < b.label; });
    // first_copy[q] + index of label within in_labels[q] -> new id
    std::vector<std::vector<int>> new_id((size_t)S0);
    for (int q = 0; q < S0; ++q) new_id[(size_t)q].assign(in_labels[(size_t)q].size(), -1);
    plan->state_label.resize(S);
    plan->orig_state.resize(S);
    plan->final_lin.resize(S);
    for (size_t i = 0; i < S; ++i) {
        const Copy &c = copies[i];
        auto &labs = in_labels[(size_t)c.orig];
        size_t j = (size_t)(std::lower_bound(labs.begin(), labs.end(), c.label) - labs.begin());
        new_id[(size_t)c.orig][j] = (int)i;
        plan->state_label[i] = c.label;
        plan->orig_state[i] = c.orig;
        float fw = fst.final_logw[(size_t)c.orig];
        plan->final_lin[i] = std::isinf(fw) ? 0.f : std::exp(fw);
    }
    plan->file_states = S0;
    plan->file_arcs = (int)A0;
    plan->num_states = (int)S;
    plan->num_labels = max_label + 1;
    plan->start = new_id[(size_t)fst.start][0];  // all start mass on one copy (copies share out-arcs)
    plan->n_ctas = n_ctas;
    plan->n_warps = n_warps;

    // 3. rows
    std::vector<std::vector<Arc>> in_rows(S), out_rows(S);
    for (size_t a = 0; a < A0; ++a) {
        const int p = fst.src[a], q = fst.dst[a];
        auto &labs = in_labels[(size_t)q];
        size_t j = (size_t)(std::lower_bound(labs.begin(), labs.end(), fst.label[a]) - labs.begin());
        const int qn = new_id[(size_t)q][j];
        const float w = std::exp(fst.logw[a]);
        for (int pn : new_id[(size_t)p]) {
            in_rows[(size_t)qn].push_back(Arc{(uint32_t)pn, w});
            out_rows[(size_t)pn].push_back(Arc{(uint32_t)qn, w});
        }
    }
    // gather locality: visit peers in ascending order
    for (auto &r : in_rows) std::sort(r.begin(), r.end(), [](const Arc &a, const Arc &b) { return a.peer < b.peer; });
    for (auto &r : out_rows) std::sort(r.begin(), r.end(), [](const Arc &a, const Arc &b) { return a.peer < b.peer; });

    for (auto &r : in_rows) for (auto &a : r) if (!(a.w >= 0.f) || std::isinf(a.w)) { *err = "den graph arc weight is not a finite probability-like value"; return false; }
    Layout(in_rows, plan->state_label, n_ctas, n_warps, &plan->fwd);
    Layout(out_rows, plan->state_label, n_ctas, n_warps, &plan->bwd);
    return true;
}

}  // namespace ccb
