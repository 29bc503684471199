// TEST INFRASTRUCTURE ONLY -- the product path (recommendersystems_b200/, librwr_b200.so) never links, loads or calls this.
//
// C ABI over the reference's OWN classes: oracle/_ref/reference_rwr.hpp is Recommenders/RWRBased/{Graph,Model,Recommender}.cs
// as oracle/cs2cpp.py respells them for a C++ compiler (built from the sources where they lie under /root/reference; the
// header is git-ignored and never committed).  This file only does what the reference's callers do
// (TweetRecommender/DataLoader.cs:60-77 fills `allNodes` / `allLinks`, Experiment.cs:104-109 builds the graph and asks for a
// recommendation) and copies the results out; it contains no arithmetic of the path.
//
// Build: `make -C oracle ref`  ->  oracle/_ref/libref.so   (g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math)
#include <cstdint>
#include <cstring>

#include "_ref/reference_rwr.hpp"

using namespace Recommenders_RWRBased;

namespace {

enum { REF_OK = 0, REF_E_INVALID = -1, REF_E_BADSEED = -2, REF_E_ALREADY_BUILT = -3, REF_E_BADINDEX = -4, REF_E_NOT_BUILT = -5 };

struct RefGraph {
    Dictionary<int, Node> nodes;
    Dictionary<int, List<ForwardLink>> edges;
    Graph* graph = nullptr;
    bool built = false;
    int32_t n = 0;
    ~RefGraph() { delete graph; }
};

template <typename F>
int guarded(const RefGraph* g, F f) {
    try {
        f();
        return REF_OK;
    } catch (const bcl::KeyNotFoundException&) {
        return g->built ? REF_E_BADSEED : REF_E_NOT_BUILT;      // graph.graph[i] before buildGraph() (Model.cs:79) / edges[seed] (Recommender.cs:21)
    } catch (const bcl::ArgumentException&) {
        return REF_E_ALREADY_BUILT;                             // graph.Add(i, ..) twice (Graph.cs:86)
    } catch (const bcl::IndexOutOfRangeException&) {
        return REF_E_BADINDEX;                                  // nextRank[link.targetNode] (Model.cs:87)
    } catch (...) {
        return REF_E_INVALID;
    }
}

}  // namespace

extern "C" {

// Links arrive flattened in (source ascending, insertion) order; `has_entry` (nullable, int8[n]) marks sources whose
// `edges` entry exists even when it holds no link (DataLoader never creates one; the tests do).
void* ref_graph_create(int32_t n, const int64_t* node_id, const int32_t* node_type, int64_t n_links, const int32_t* src,
                       const int32_t* dst, const int32_t* etype, const double* w, const int8_t* has_entry) {
    if (n < 0 || n_links < 0) return nullptr;
    RefGraph* g = new RefGraph();
    g->n = n;
    for (int32_t i = 0; i < n; i++) g->nodes.Add(i, Node(node_id[i], (NodeType)node_type[i]));       // DataLoader.cs:40-41
    if (has_entry)
        for (int32_t i = 0; i < n; i++)
            if (has_entry[i]) g->edges.Add(i, List<ForwardLink>());
    for (int64_t k = 0; k < n_links; k++) {                                                            // DataLoader.cs:61-62, :73-74
        if (!g->edges.ContainsKey(src[k])) g->edges.Add(src[k], List<ForwardLink>());
        g->edges[src[k]].Add(ForwardLink(dst[k], (EdgeType)etype[k], w[k]));
    }
    g->graph = new Graph(g->nodes, g->edges);                                                          // Experiment.cs:104
    return g;
}

void ref_graph_destroy(void* h) { delete (RefGraph*)h; }

int ref_graph_build(void* h) {
    RefGraph* g = (RefGraph*)h;
    const int rc = guarded(g, [&] { g->graph->buildGraph(); });                                        // Experiment.cs:105
    if (rc == REF_OK) g->built = true;
    return rc;
}

int64_t ref_graph_nnz(void* h) {
    RefGraph* g = (RefGraph*)h;
    if (!g->built) return -1;
    int64_t nnz = 0;
    for (int32_t i = 0; i < g->n; i++) {
        Array<ForwardLink> row = g->graph->graph[i];
        if (row != nullptr) nnz += row.Length();
    }
    return nnz;
}

// Graph.graph as CSR: row_ptr[n + 1], col / val / type [nnz] (type nullable); a null row is an empty row
int ref_graph_get_csr(void* h, int64_t* row_ptr, int32_t* col, double* val, int32_t* type) {
    RefGraph* g = (RefGraph*)h;
    if (!g->built) return REF_E_NOT_BUILT;
    int64_t p = 0;
    for (int32_t i = 0; i < g->n; i++) {
        row_ptr[i] = p;
        Array<ForwardLink> row = g->graph->graph[i];
        if (row == nullptr) continue;
        for (int k = 0; k < row.Length(); k++, p++) {
            col[p] = row[k].targetNode;
            val[p] = row[k].weight;
            if (type) type[p] = (int32_t)row[k].type;
        }
    }
    row_ptr[g->n] = p;
    return REF_OK;
}

// seed >= 0: Model(graph, damping, seed); seed == -1: Model(graph, damping).
// mode 0: run(int n_iter).  mode 2: run() (default threshold).  mode 3: run(double thr) -- the reference's own loops.
// mode 1: the loop of run(double) driven from here through the public methods, so that the deliverRanks calls can be
//         counted and capped (`max_iter`, 0 = none): deliver; converged?; update.
int ref_model_run(void* h, int32_t seed, double damping, int32_t mode, int32_t n_iter, double thr, int64_t max_iter,
                  double* rank_out, int64_t* iters_out) {
    RefGraph* g = (RefGraph*)h;
    int64_t iters = 0;
    const int rc = guarded(g, [&] {
        Model model = seed >= 0 ? Model(g->graph, damping, seed) : Model(g->graph, damping);
        if (mode == 0) { model.run(n_iter); iters = n_iter > 0 ? n_iter : 0; }
        else if (mode == 2) model.run();
        else if (mode == 3) model.run(thr);
        else {
            while (true) {
                model.deliverRanks();
                iters++;
                const bool converged = model.checkConvergence(thr);
                model.updateRanks();
                if (converged || (max_iter > 0 && iters >= max_iter)) break;
            }
        }
        for (int32_t i = 0; i < g->n; i++) rank_out[i] = model.rank[i];
    });
    if (iters_out) *iters_out = (mode == 2 || mode == 3) ? -1 : iters;
    return rc;
}

// Recommender.Recommendation(seed, float damping, n_iter[, top_n]) -> number of pairs written (the whole list must fit `cap`)
int64_t ref_recommend(void* h, int32_t seed, float damping, int32_t n_iter, int32_t has_top_n, int32_t top_n, int64_t* ids,
                      double* scores, int64_t cap) {
    RefGraph* g = (RefGraph*)h;
    int64_t count = 0;
    const int rc = guarded(g, [&] {
        Recommender recommender(g->graph);                                                             // Experiment.cs:108
        auto list = has_top_n ? recommender.Recommendation(seed, damping, n_iter, top_n)
                              : recommender.Recommendation(seed, damping, n_iter);                     // Experiment.cs:109
        count = list.Count();
        if (count > cap) throw 0;                                   // -> REF_E_INVALID
        for (int64_t i = 0; i < count; i++) { ids[i] = list[i].Key; scores[i] = list[i].Value; }
    });
    return rc == REF_OK ? count : rc;
}

}  // extern "C"
