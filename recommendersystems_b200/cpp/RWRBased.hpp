// RWRBased.hpp -- C++ host side of the drop-in: the public surface of the reference's `Recommenders.RWRBased`
// (Recommenders/RWRBased/Graph.cs, Model.cs, Recommender.cs) over the C ABI of librwr_b200.so (include/rwr_b200.h).
//
// The reference is compiled code (C#) and this image has no C# toolchain, so next to the P/Invoke shim shipped as source
// (../csharp/) the same surface exists in C++: same type and member names, same argument meaning, the .NET exceptions the
// reference's callers can see as C++ exception types of the same name.  A caller written against the reference
// (TweetRecommender/Experiment.cs:104-109) reads the same here:
//
//     Graph graph(nodes, edges);                 // Dictionary<int, Node>, Dictionary<int, List<ForwardLink>>
//     graph.buildGraph();
//     Recommender recommender(graph);
//     auto recommendation = recommender.Recommendation(0, 0.15f, nIterations);      // List<KeyValuePair<long, double>>
//
// (or, spelt the way oracle/cs2cpp.py respells C# class references: `Graph* graph = new Graph(nodes, edges); graph->buildGraph();
// Recommender* recommender = new Recommender(graph);` -- tests/cpp/experiment_caller.cpp is ONE such caller, compiled against this
// header and against the reference's own sources; the two programs must print the same lists.)
//
// Header only; link with -lrwr_b200.  Nothing is computed here: flattening `edges` in `for i in 0..N-1: foreach l in edges[i]`
// order (rwr_graph_create's input contract), widening `float dampingFactor` to double exactly as Recommender.cs:16 does, and
// mapping status codes to exceptions.  There is no CPU fallback: without a CUDA device every call throws RwrException(RWR_E_CUDA).
#pragma once

#include <cstdint>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/rwr_b200.h"

namespace Recommenders {
namespace RWRBased {

enum class NodeType : int { UNDEFINED, USER, ITEM, ETC };                                               // Recommender.cs:4
enum class EdgeType : int { UNDEFINED, LIKE, FRIENDSHIP, FOLLOW, MENTION, AUTHORSHIP, PURCHASE, ETC };   // Recommender.cs:5

struct Node {                                                                                            // Graph.cs:4-17
    long long id = 0;
    NodeType type = NodeType::UNDEFINED;
    Node() {}
    Node(long long id) : id(id) {}
    Node(long long id, NodeType type) : id(id), type(type) {}
};

struct ForwardLink {                                                                                     // Graph.cs:19-35
    int targetNode = 0;
    EdgeType type = EdgeType::UNDEFINED;
    double weight = 0;
    ForwardLink() {}
    ForwardLink(int targetNode, double weight) : targetNode(targetNode), weight(weight) {}
    ForwardLink(int targetNode, EdgeType type, double weight) : targetNode(targetNode), type(type), weight(weight) {}
};

// what the reference's callers can catch
struct KeyNotFoundException : std::runtime_error { using std::runtime_error::runtime_error; };        // Recommender.cs:21, Model.cs:79
struct ArgumentException : std::runtime_error { using std::runtime_error::runtime_error; };           // Graph.cs:86 (buildGraph twice)
struct IndexOutOfRangeException : std::runtime_error { using std::runtime_error::runtime_error; };    // Model.cs:87
struct RwrException : std::runtime_error {                                                            // everything else (CUDA, NCCL, ...)
    int code;
    RwrException(int code, const std::string& what) : std::runtime_error(what), code(code) {}
};

// the containers of the reference's signatures, with the members its callers use (DataLoader.cs:60-77, Experiment.cs:123):
// std::map / std::vector underneath, so any standard algorithm works on them as well
template <typename K, typename V>
struct Dictionary : std::map<K, V> {
    void Add(const K& k, const V& v) {
        if (!this->emplace(k, v).second) throw ArgumentException("An item with the same key has already been added");
    }
    bool ContainsKey(const K& k) const { return this->find(k) != this->end(); }
    int Count() const { return (int)this->size(); }
    V& operator[](const K& k) {                      // the indexer's getter: KeyNotFoundException, never an insertion
        auto it = this->find(k);
        if (it == this->end()) throw KeyNotFoundException("The given key was not present in the dictionary");
        return it->second;
    }
};
template <typename T>
struct List : std::vector<T> {
    using std::vector<T>::vector;
    void Add(const T& x) { this->push_back(x); }
    int Count() const { return (int)this->size(); }
};
template <typename K, typename V>
struct KeyValuePair {
    K Key;
    V Value;
    KeyValuePair() : Key(), Value() {}
    KeyValuePair(const K& k, const V& v) : Key(k), Value(v) {}
};

inline void check(int rc) {
    if (rc == RWR_OK) return;
    const std::string msg = rwr_last_error();
    switch (rc) {
        case RWR_E_BADSEED:
        case RWR_E_NOT_BUILT: throw KeyNotFoundException(msg);
        case RWR_E_ALREADY_BUILT: throw ArgumentException(msg);
        case RWR_E_BADINDEX: throw IndexOutOfRangeException(msg);
        default: throw RwrException(rc, msg);
    }
}

class Graph {                                                                                            // Graph.cs:37-94
public:
    // Graph information: the caller's dictionaries, kept by reference and read at buildGraph() like the reference does
    Dictionary<int, Node>& nodes;
    Dictionary<int, List<ForwardLink>>& edges;

    Graph(Dictionary<int, Node>& nodes, Dictionary<int, List<ForwardLink>>& edges, int device = -1)
        : nodes(nodes), edges(edges), device_(device) {}
    Graph(const Graph&) = delete;
    Graph& operator=(const Graph&) = delete;
    ~Graph() { rwr_graph_destroy(h_); }

    void buildGraph() {                                                                                  // Graph.cs:51-88
        if (h_) {                                   // graph.Add(i, ..) on an existing key
            check(rwr_graph_build(h_));
            return;
        }
        const int n = (int)nodes.size();
        std::vector<int64_t> id((size_t)n);
        std::vector<int32_t> type((size_t)n), src, dst, et;
        std::vector<double> w;
        for (int i = 0; i < n; i++) {               // keys are the dense indices 0..N-1 (DataLoader.cs:41, :54)
            auto it = nodes.find(i);
            if (it == nodes.end()) throw KeyNotFoundException("nodes has no entry for index " + std::to_string(i));
            id[(size_t)i] = it->second.id;
            type[(size_t)i] = (int32_t)it->second.type;
            auto e = edges.find(i);                 // `if (edges.ContainsKey(i))`, Graph.cs:55
            if (e == edges.end()) continue;
            for (const ForwardLink& l : e->second) {
                src.push_back(i); dst.push_back(l.targetNode); et.push_back((int32_t)l.type); w.push_back(l.weight);
            }
        }
        rwr_opts o{};
        o.device = device_;
        o.hub_entries = -1;
        o.empty_seed_ok = 1;                        // an `edges` entry that exists but is empty is served; Recommender checks the key itself
        check(rwr_graph_create(n, id.data(), type.data(), (int64_t)src.size(), src.data(), dst.data(), et.data(), w.data(), &o, &h_));
        check(rwr_graph_build(h_));
    }

    int size() { return (int)nodes.size(); }                                                            // Graph.cs:91

    // `graph[i]` (Graph.cs:43): the adjusted forward links of node i; empty for a null (dangling) row
    List<ForwardLink> graph(int i) {
        load_csr();
        List<ForwardLink> row;
        for (int64_t k = row_ptr_[(size_t)i]; k < row_ptr_[(size_t)i + 1]; k++)
            row.Add(ForwardLink(col_[(size_t)k], (EdgeType)etype_[(size_t)k], val_[(size_t)k]));
        return row;
    }

    rwr_graph* handle() {
        if (!h_) throw KeyNotFoundException("buildGraph() has not run (graph.graph[i], Model.cs:79)");
        return h_;
    }

private:
    void load_csr() {
        if (!row_ptr_.empty()) return;
        rwr_graph_info info{};
        check(rwr_graph_get_info(handle(), &info));
        row_ptr_.assign((size_t)info.n_nodes + 1, 0);
        col_.assign((size_t)info.nnz, 0); val_.assign((size_t)info.nnz, 0.0); etype_.assign((size_t)info.nnz, 0);
        check(rwr_graph_get_csr(h_, row_ptr_.data(), col_.data(), val_.data()));
        check(rwr_graph_get_csr_types(h_, etype_.data()));
    }
    rwr_graph* h_ = nullptr;
    int device_;
    std::vector<int64_t> row_ptr_;
    std::vector<int32_t> col_, etype_;
    std::vector<double> val_;
};

class Model {                                                                                            // Model.cs:5-116
public:
    Graph* graph;
    std::vector<double> rank;
    int nNodes;
    double dampingFactor;
    int nIterations = 0;            // deliverRanks() calls so far

    Model(Graph& graph, double dampingFactor) : graph(&graph), nNodes(graph.size()), dampingFactor(dampingFactor), seed_(-1) {
        rank.assign((size_t)nNodes, 1.0);                                                                // Model.cs:24
    }
    Model(Graph& graph, double dampingFactor, int targetNode)
        : graph(&graph), nNodes(graph.size()), dampingFactor(dampingFactor), seed_(targetNode) {
        rank.assign((size_t)nNodes, 0.0);                                                                // Model.cs:44
        if (targetNode >= 0 && targetNode < nNodes) rank[(size_t)targetNode] = nNodes;
    }
    // a C# class variable is a reference: `new Model(graph, c, seed)` with `Graph* graph` reads the same here
    Model(Graph* graph, double dampingFactor) : Model(*graph, dampingFactor) {}
    Model(Graph* graph, double dampingFactor, int targetNode) : Model(*graph, dampingFactor, targetNode) {}

    void run() { run_threshold(0.0); }                                    // Model.cs:52-55: (1 / double.MaxValue) * N, chosen by the library
    void run(double threshold) { run_threshold(threshold); }              // Model.cs:57-66
    void run(int nIterations) {                                           // Model.cs:68-73; successive calls accumulate, as in the reference
        total_ += nIterations > 0 ? nIterations : 0;
        rwr_result* r = nullptr;
        check(rwr_run_fixed(graph->handle(), &seed_, 1, dampingFactor, total_, RWR_FP64, &r));
        fetch(r);
        this->nIterations = total_;
    }

private:
    void run_threshold(double thr) {
        rwr_result* r = nullptr;
        int32_t iters = 0;
        check(rwr_run_threshold(graph->handle(), &seed_, 1, dampingFactor, thr, 0, RWR_FP64, &iters, &r));
        fetch(r);
        nIterations = iters;
    }
    void fetch(rwr_result* r) {
        const int rc = rwr_scores(r, 0, rank.data());
        rwr_result_destroy(r);
        check(rc);
    }
    int32_t seed_;
    int total_ = 0;
};

class Recommender {                                                                                      // Recommender.cs:7-52
public:
    Recommender(Graph& graph) : graph(graph) {}
    Recommender(Graph* graph) : graph(*graph) {}

    List<KeyValuePair<long long, double>> Recommendation(int idxTargetUser, float dampingFactor, int nIteration) {
        require_entry(idxTargetUser);
        rwr_result* r = nullptr;
        const int32_t seed = idxTargetUser;
        check(rwr_run_fixed(graph.handle(), &seed, 1, (double)dampingFactor, nIteration, RWR_FP64, &r));     // float -> double, :16
        const size_t cap = (size_t)graph.size();
        std::vector<int64_t> ids(cap ? cap : 1);
        std::vector<double> scores(cap ? cap : 1);
        int64_t count = 0;
        const int rc = rwr_rank_all(r, 0, ids.data(), scores.data(), (int64_t)cap, &count);
        rwr_result_destroy(r);
        check(rc);
        List<KeyValuePair<long long, double>> recommendation;
        recommendation.reserve((size_t)count);
        for (int64_t i = 0; i < count; i++) recommendation.Add(KeyValuePair<long long, double>((long long)ids[(size_t)i], scores[(size_t)i]));
        return recommendation;
    }

    List<KeyValuePair<long long, double>> Recommendation(int idxTargetUser, float dampingFactor, int nIteration, int topN) {
        if (topN <= 0 || topN > 16) {               // `Count == topN` never fires for topN <= 0 (Recommender.cs:47): the whole list
            auto recommendation = Recommendation(idxTargetUser, dampingFactor, nIteration);
            if (topN > 0 && (size_t)topN < recommendation.size()) recommendation.resize((size_t)topN);
            return recommendation;
        }
        require_entry(idxTargetUser);
        const int32_t seed = idxTargetUser;
        int64_t ids[16];
        double scores[16];
        int32_t count = 0;
        check(rwr_recommend(graph.handle(), &seed, 1, (double)dampingFactor, nIteration, RWR_FP64, topN, ids, scores, &count, nullptr));
        List<KeyValuePair<long long, double>> topNRecommendation;
        for (int i = 0; i < count; i++) topNRecommendation.Add(KeyValuePair<long long, double>((long long)ids[i], scores[i]));
        return topNRecommendation;
    }

private:
    void require_entry(int idx) {                   // `graph.edges[idxTargetUser]`, Recommender.cs:21
        if (graph.edges.find(idx) == graph.edges.end())
            throw KeyNotFoundException("edges has no entry for node " + std::to_string(idx) + " (Recommender.cs:21)");
    }
    Graph& graph;
};

}  // namespace RWRBased
}  // namespace Recommenders
