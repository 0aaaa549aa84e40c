// Host-side den-graph loader and kernel plan (no CUDA in this header).
//
// Replaces the reference's loader src/ctc_crf/gpu_den/fst_read.cc:11-61 (OpenFst) and the CSR flattening
// in den_calculate.cu:309-355 with a native parser plus a plan laid out for the persistent kernels in
// den_kernels.cu.  See DESIGN.md "Data layout".
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ccb {

// One arc as the kernels read it: 8 bytes, broadcast-loaded from shared memory two at a time (LDS.128).
//   peer : state id the kernel gathers from (source state for the forward pass, destination state for the
//          backward pass).
//   w    : arc weight in the LINEAR domain, exp(-tropical weight), always >= 0.  Rows are padded with
//          zero-weight arcs (peer 0) to whole QUADS of 4 arcs; the SIGN BIT of the 4th weight of a quad is set
//          when the quad is the last one of its row (tested once per quad, applied as |w| in the FMA).
struct alignas(8) Arc {
    uint32_t peer;
    float w;
};
static constexpr int kQuad = 4;          // arcs per quad (row padding granule)
static constexpr int kChunkArcPad = 16;  // every warp chunk's arc count is padded to a multiple of this

// Den graph as stored in the file, in the view fst_read.cc:40-59 gives the kernels.
struct HostFst {
    int num_states = 0;
    int start = 0;
    std::vector<int> src, dst, label;   // label = ilabel-1
    std::vector<float> logw;            // -tropical
    std::vector<float> final_logw;      // -Final, -inf for non-final states
};

// Returns false and fills err on failure (never exits, unlike den_calculate.cu:16-25,331-334).
bool ReadFstFile(const char *path, HostFst *out, std::string *err);

// Pass-specific half of the plan: rows are states (in the renumbered order), a row's arcs are the arcs the
// pass sums over (in-arcs for forward, out-arcs for backward).  chunk_* cut the rows into
// n_ctas*n_warps contiguous chunks of near-equal cost; CTA c owns chunks [c*n_warps, (c+1)*n_warps).
struct PassPlan {
    std::vector<Arc> arcs;              // chunk-major: rows as whole quads, chunk tails padded with unflagged zero quads
    std::vector<int> chunk_state;       // [n_chunks+1] first state of each chunk
    std::vector<int> chunk_arc;         // [n_chunks+1] first arc of each chunk (multiples of kChunkArcPad)
    int real_arcs = 0;                  // arcs of the graph in this pass (before padding)
    int max_tile_arcs = 0;              // max arcs owned by one CTA
    int max_tile_labels = 0;            // max (label range + 1) over CTAs
    int max_tile_rows = 0;              // max rows owned by one CTA
};

struct DenPlan {
    int file_states = 0, file_arcs = 0;
    int num_states = 0;                 // after in-label split
    int start = 0;                      // renumbered start state
    int num_labels = 0;                 // max label + 1
    int n_ctas = 0, n_warps = 0;
    std::vector<int> state_label;       // [S] the single label carried by every arc INTO the state
    std::vector<float> final_lin;       // [S] exp(final_logw) (0 for non-final)
    std::vector<int> orig_state;        // [S] state id in the file
    PassPlan fwd, bwd;
};

// Build the plan:
//  1. split every state by the label of its incoming arcs so that each state has exactly one in-label
//     (a no-op for T-compose-LM graphs: build_ctc_topo.py:49-60 gives every arc into token state i the
//     ilabel i+1); this lets the kernels hoist the emission term out of the arc loop:
//         alpha_t(q) = y_{t-1}[lab(q)] * sum_p w_pq alpha_{t-1}(p)
//     and turns the per-arc atomic gradient of den_calculate.cu:219-223 into a per-state product
//         gamma_t[k] = sum_{q: lab(q)=k} alpha_{t+1}(q) beta_{t+1}(q) / Z.
//  2. renumber states sorted by label (so a CTA tile spans very few labels);
//  3. build in-arc and out-arc rows with linear-domain weights;
//  4. cut rows into cost-balanced chunks for an n_ctas x n_warps persistent grid.
bool BuildDenPlan(const HostFst &fst, int n_ctas, int n_warps, DenPlan *plan, std::string *err);

}  // namespace ccb
