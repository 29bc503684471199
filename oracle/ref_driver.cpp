// TEST INFRASTRUCTURE ONLY -- the product path (recommendersystems_b200/, librwr_b200.so) never links, loads or calls this.
//
// C ABI over the reference's OWN classes: reference_rwr.hpp is Recommenders/RWRBased/{Graph,Model,Recommender}.cs and
// reference_experiment.hpp is TweetRecommender/DataLoader.cs + the k-fold loop of Experiment.cs, as oracle/cs2cpp.py respells
// them for a C++ compiler (generated from the sources where they lie under /root/reference into a temporary directory that
// the Makefile deletes after the build: reference text never stays in the tree).  This file only does what the reference's callers do
// (TweetRecommender/DataLoader.cs:60-77 fills `allNodes` / `allLinks`, Experiment.cs:104-109 builds the graph and asks for a
// recommendation, Experiment.cs:61-66 sets up the result dictionary) and copies the results out; it contains no arithmetic
// of the path.
//
// Build: `make -C oracle ref`  ->  oracle/_ref/libref.so   (g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math)
#include <cstdint>
#include <cstring>

#include "reference_rwr.hpp"           // generated into a temporary directory by `make -C oracle ref` (-I), deleted after the build
#include "reference_experiment.hpp"

using namespace Recommenders_RWRBased;
using TweetRecommender::DataLoader;
using TweetRecommender::EvaluationMetric;
using TweetRecommender::Methodology;

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

// ---------------------------------------------------------------------------------------------------------------------
// The callers: DataLoader over in-memory tables, and the k-fold loop of Experiment.runKFoldCrossValidation.

// the tables of one ego network's database, registered under `path` (`<ego user id>.sqlite`, DataLoader.cs:32)
int ref_db_reset(const char* path) {
    bcl::mem_dbs()[path] = bcl::MemDb();
    return REF_OK;
}

// rows (a[i], b[i]) appended to `table` in order: follow(source, target), tweet(id, author), retweet / quote / favorite(user, tweet),
// mention(source, target)
int ref_db_add(const char* path, const char* table, const int64_t* a, const int64_t* b, int64_t n) {
    auto it = bcl::mem_dbs().find(path);
    if (it == bcl::mem_dbs().end()) return REF_E_INVALID;
    auto& rows = it->second.tables[table];
    for (int64_t i = 0; i < n; i++) rows.emplace_back((long long)a[i], (long long)b[i]);
    return REF_OK;
}

struct RefLoader {
    DataLoader* loader = nullptr;
    int32_t n = 0;
    int64_t e = 0;
    ~RefLoader() { delete loader; }
};

// checkEgoNetworkValidation (DataLoader.cs:79-92) and the two counts it is made of
int ref_loader_validation(const char* path, int32_t n_folds, int32_t* valid, int32_t* cnt_likes, int32_t* cnt_friends) {
    try {
        DataLoader loader(path, n_folds);
        *cnt_likes = loader.getLikeCountOfEgoUser();
        *cnt_friends = loader.getFriendsCountOfEgoUser();
        *valid = loader.checkEgoNetworkValidation() ? 1 : 0;
        return REF_OK;
    } catch (...) {
        return REF_E_INVALID;
    }
}

// `new DataLoader(path, nFolds)` + `graphConfiguration(methodology, fold)` (Experiment.cs:71, :77)
void* ref_loader_run(const char* path, int32_t n_folds, int32_t methodology, int32_t fold) {
    RefLoader* h = new RefLoader();
    try {
        h->loader = new DataLoader(path, n_folds);
        h->loader->graphConfiguration((Methodology)methodology, fold);
        h->n = h->loader->allNodes.Count();
        for (int32_t i = 0; i < h->n; i++)
            if (h->loader->allLinks.ContainsKey(i)) h->e += h->loader->allLinks[i].Count();
        return h;
    } catch (...) {
        delete h;
        return nullptr;
    }
}

void ref_loader_destroy(void* h) { delete (RefLoader*)h; }

void ref_loader_sizes(void* hv, int32_t* n_nodes, int64_t* n_links, int64_t* n_test) {
    RefLoader* h = (RefLoader*)hv;
    *n_nodes = h->n;
    *n_links = h->e;
    *n_test = h->loader->testSet.Count();
}

// allNodes / allLinks flattened as `for i in 0..N-1: foreach l in allLinks[i]`, has_entry[i] = allLinks.ContainsKey(i),
// testSet in its enumeration order
void ref_loader_copy(void* hv, int64_t* node_id, int32_t* node_type, int8_t* has_entry, int32_t* src, int32_t* dst, int32_t* etype,
                     double* w, int64_t* test_ids) {
    RefLoader* h = (RefLoader*)hv;
    int64_t p = 0;
    for (int32_t i = 0; i < h->n; i++) {
        node_id[i] = h->loader->allNodes[i].id;
        node_type[i] = (int32_t)h->loader->allNodes[i].type;
        has_entry[i] = h->loader->allLinks.ContainsKey(i) ? 1 : 0;
        if (!has_entry[i]) continue;
        for (ForwardLink l : h->loader->allLinks[i]) {
            src[p] = i; dst[p] = l.targetNode; etype[p] = (int32_t)l.type; w[p] = l.weight;
            p++;
        }
    }
    int64_t t = 0;
    for (long long id : h->loader->testSet) test_ids[t++] = id;
}

// Experiment.runKFoldCrossValidation for one database and one methodology: the result dictionary as Experiment.cs:61-66 sets it
// up, then the reference's own k-fold loop (runFolds).  Out: finalResult[HIT], finalResult[AVGPRECISION] (the SUM over the folds;
// the result row divides it by nFolds, Experiment.cs:150), cntLikes; valid = 0 when checkEgoNetworkValidation made it return.
int ref_experiment_run(const char* path, int32_t n_folds, int32_t n_iterations, int32_t methodology, double* hit,
                       double* avg_precision_sum, int32_t* cnt_likes, int32_t* valid) {
    try {
        Dictionary<EvaluationMetric, double> finalResult;
        finalResult.Add(EvaluationMetric::HIT, 0.0);                       // foreach (metric in Enum.GetValues(..)) finalResult.Add(metric, 0d)
        finalResult.Add(EvaluationMetric::AVGPRECISION, 0.0);
        List<EvaluationMetric> metrics;                                    // new List<EvaluationMetric>(finalResult.Keys)
        for (EvaluationMetric m : finalResult.Keys()) metrics.Add(m);
        int cntLikes = 0;
        *valid = TweetRecommender::runFolds(path, n_folds, n_iterations, (Methodology)methodology, finalResult, metrics, cntLikes) ? 1 : 0;
        *hit = finalResult[EvaluationMetric::HIT];
        *avg_precision_sum = finalResult[EvaluationMetric::AVGPRECISION];
        *cnt_likes = cntLikes;
        return REF_OK;
    } catch (const bcl::KeyNotFoundException&) {
        return REF_E_BADSEED;                                              // the ego user has no `edges` entry (Recommender.cs:21)
    } catch (...) {
        return REF_E_INVALID;
    }
}

}  // extern "C"
