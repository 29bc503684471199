// TEST INFRASTRUCTURE ONLY -- the product path (recommendersystems_b200/, librwr_b200.so) must
// never link, load or call anything in this file.  Allowed users: tests/, __graft_entry__.smoke(),
// and bench.py's cpu_baseline / --impl reference legs.
//
// CPU restatement (strict IEEE double, sequential loops in the reference's order) of
//   /root/reference/Recommenders/RWRBased/Graph.cs:51-88        buildGraph
//   /root/reference/Recommenders/RWRBased/Model.cs:14-50        both constructors
//   /root/reference/Recommenders/RWRBased/Model.cs:52-73        run() / run(double) / run(int)
//   /root/reference/Recommenders/RWRBased/Model.cs:76-115       deliverRanks / updateRanks / checkConvergence
//   /root/reference/Recommenders/RWRBased/Recommender.cs:14-51  Recommendation (+ topN overload)
//   /root/reference/TweetRecommender/Experiment.cs:121-128      hits / average precision
//
// PINNING.  The reference holds no tests, fixtures or golden vectors for this path and no C# runtime exists in this
// image, so it cannot be executed as a .NET assembly.  Its sources are compiled instead: `make -C oracle ref` respells the
// declarations of Graph.cs / Model.cs / Recommender.cs for g++ (oracle/cs2cpp.py, syntactic rules only, every statement
// and expression untouched) from where they lie under /root/reference -> oracle/_ref/libref.so.  tests/test_reference_pin.py
// holds this restatement to that library bit for bit (golden vectors, random graphs with dangling / multi-edge / NaN / Inf
// rows, exceptions, the C1 graph), next to (a) the hand-derived known-answer vector of SURVEY.md section 8c and (b) an
// independent pure-Python restatement (oracle/rwr_literal.py).  What the compiled sources cannot show is the .NET runtime
// itself: double arithmetic is IEEE binary64 on both sides (RyuJIT x64 / SSE2), List.Sort's algorithm differs but sorts a
// total order on distinct ids, Dictionary / List / array semantics are oracle/ref_shim.hpp's.
//
// Two iteration forms:
//   literal   -- the O(N^2) restart loops exactly as written (Model.cs:92-93, :96-97);
//   collapsed -- `next[seed] += rank_restart` at the same point of the i-loop.  For a one-hot
//                restart vector every other addend is `x * 0.0 == +-0.0` and `y + 0.0 == y`, so the
//                two forms are bit-identical while x is finite; a NaN or Inf rank (a row whose
//                weights sum to 0, Graph.cs:81) makes `x * 0.0` NaN for EVERY node -- the whole
//                next vector is NaN in the reference -- and the collapsed form then runs the
//                literal loop for that source.  The tests assert the identity, NaN cases included.
//
// Also here: the CPU side of the deterministic synthetic graph generator (this repo's own spec,
// include/rwr_b200.h `rwr_synth_spec`; it replaces DataLoader.cs:256-436) written independently of
// the CUDA one so the two can be compared bit for bit.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared (see Makefile)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

enum { NT_UNDEFINED = 0, NT_USER = 1, NT_ITEM = 2, NT_ETC = 3 };                       // Recommender.cs:4
enum { ET_UNDEFINED = 0, ET_LIKE = 1, ET_FRIENDSHIP = 2, ET_FOLLOW = 3, ET_MENTION = 4,
       ET_AUTHORSHIP = 5, ET_PURCHASE = 6, ET_ETC = 7 };                               // Recommender.cs:5

enum { ORC_OK = 0, ORC_E_INVALID = -1, ORC_E_BADSEED = -2, ORC_E_ALREADY_BUILT = -3,
       ORC_E_BADINDEX = -4, ORC_E_NOT_BUILT = -5 };

struct OrcGraph {
    int32_t n = 0;
    std::vector<int64_t> node_id;
    std::vector<int32_t> node_type;
    // raw `edges`: per-source lists in insertion order (Graph.cs:40); a source without links has no key
    std::vector<int64_t> raw_ptr;
    std::vector<int32_t> raw_dst, raw_type;
    std::vector<double> raw_w;
    // built `graph` (Graph.cs:43): null rows have row_ptr[i] == row_ptr[i+1]
    bool built = false;
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> col;
    std::vector<double> val;
};

// ---------------------------------------------------------------- Graph.buildGraph, Graph.cs:51-88
int build_graph(OrcGraph& g) {
    if (g.built) return ORC_E_ALREADY_BUILT;            // Dictionary.Add on an existing key (Graph.cs:86)
    const int32_t n = g.n;
    g.row_ptr.assign((size_t)n + 1, 0);
    g.col.clear();
    g.val.clear();
    for (int32_t i = 0; i < n; i++) {
        g.row_ptr[i] = (int64_t)g.col.size();
        const int64_t b = g.raw_ptr[i], e = g.raw_ptr[i + 1];
        if (e > b) {                                    // edges.ContainsKey(i), :55
            int nExplicitLinks = 0;
            for (int64_t k = b; k < e; k++)             // :58-61
                if (g.raw_type[k] != ET_UNDEFINED) nExplicitLinks += 1;
            if (nExplicitLinks > 0) {                   // :64
                const size_t first = g.col.size();
                double sumWeights = 0;                  // :70
                for (int64_t k = b; k < e; k++) {       // :71-77
                    if (g.raw_type[k] != ET_UNDEFINED) {
                        g.col.push_back(g.raw_dst[k]);
                        g.val.push_back(g.raw_w[k]);
                        sumWeights += g.raw_w[k];       // :75
                    }
                }
                for (size_t f = first; f < g.col.size(); f++)   // :80-81
                    g.val[f] /= sumWeights;
            }
        }
    }
    g.row_ptr[n] = (int64_t)g.col.size();
    g.built = true;
    return ORC_OK;
}

// ---------------------------------------------------------------- Model, Model.cs:5-116
struct Model {
    const OrcGraph& g;
    int32_t nNodes;
    double dampingFactor;
    std::vector<double> rank, nextRank, restart;
    int32_t seed;           // -1: uniform constructor
    bool literal;
    int64_t nDeliver = 0;

    Model(const OrcGraph& graph, double d, int32_t targetNode, bool lit)
        : g(graph), nNodes(graph.n), dampingFactor(d), seed(targetNode), literal(lit) {
        rank.assign(nNodes, 0.0);
        nextRank.assign(nNodes, 0.0);
        restart.assign(nNodes, 0.0);
        if (targetNode < 0) {                                   // Model.cs:14-31
            for (int i = 0; i < nNodes; i++) {
                rank[i] = 1.0;
                nextRank[i] = 0;
                restart[i] = 1.0 / nNodes;
            }
        } else {                                                // Model.cs:33-50
            for (int i = 0; i < nNodes; i++) {
                rank[i] = (i == targetNode) ? (double)nNodes : 0.0;
                nextRank[i] = 0;
                restart[i] = (i == targetNode) ? 1.0 : 0.0;
            }
        }
    }

    void deliverRanks() {                                       // Model.cs:76-100
        nDeliver++;
        const int64_t* rp = g.row_ptr.data();
        const int32_t* col = g.col.data();
        const double* val = g.val.data();
        const bool collapse = !literal && seed >= 0 && seed < nNodes;
        const bool skip_restart = !literal && seed >= nNodes;   // out-of-range target: restart == 0 everywhere
        for (int i = 0; i < nNodes; i++) {
            const int64_t b = rp[i], e = rp[i + 1];
            if (e > b) {
                double rank_randomWalk = (1 - dampingFactor) * rank[i];         // :84
                for (int64_t w = b; w < e; w++)
                    nextRank[col[w]] += rank_randomWalk * val[w];               // :87
                double rank_restart = rank[i] - rank_randomWalk;                // :91
                if (collapse && std::isfinite(rank_restart)) {
                    nextRank[seed] += rank_restart * 1.0;
                } else if (!skip_restart) {                                     // literal form, or a NaN / Inf that `* 0` spreads
                    for (int r = 0; r < nNodes; r++)
                        nextRank[r] += rank_restart * restart[r];               // :93
                }
            } else {
                if (collapse && std::isfinite(rank[i])) {
                    nextRank[seed] += rank[i] * 1.0;
                } else if (!skip_restart) {
                    for (int r = 0; r < nNodes; r++)
                        nextRank[r] += rank[i] * restart[r];                    // :97
                }
            }
        }
    }

    void updateRanks() {                                        // Model.cs:103-108
        for (int i = 0; i < nNodes; i++) {
            rank[i] = nextRank[i];
            nextRank[i] = 0;
        }
    }

    bool checkConvergence(double threshold) {                   // Model.cs:110-115
        double diff = 0;
        for (int i = 0; i < nNodes; i++)
            diff += (rank[i] > nextRank[i]) ? (rank[i] - nextRank[i]) : (nextRank[i] - rank[i]);
        return diff < threshold;
    }

    void run_fixed(int nIterations) {                           // Model.cs:68-73
        for (int n = 0; n < nIterations; n++) {
            deliverRanks();
            updateRanks();
        }
    }

    // Model.cs:57-66; `max_iter` is a safety net the reference does not have (returns false if hit)
    bool run_threshold(double threshold, int64_t max_iter) {
        while (true) {
            deliverRanks();
            if (checkConvergence(threshold)) {
                updateRanks();
                return true;
            }
            updateRanks();
            if (max_iter > 0 && nDeliver >= max_iter) return false;
        }
    }
};

// System.Double.CompareTo
inline int compare_to(double x, double y) {
    if (x < y) return -1;
    if (x > y) return 1;
    if (x == y) return 0;
    if (std::isnan(x)) return std::isnan(y) ? 0 : -1;
    return 1;
}

struct Scored {
    int64_t id;
    double score;
};

// Recommender.cs:19-38 applied to an already computed rank vector
int rank_candidates(const OrcGraph& g, int32_t idxTargetUser, const double* rank, std::vector<Scored>& out) {
    if (idxTargetUser < 0 || idxTargetUser >= g.n) return ORC_E_BADSEED;
    const int64_t b = g.raw_ptr[idxTargetUser], e = g.raw_ptr[idxTargetUser + 1];
    if (e == b) return ORC_E_BADSEED;                           // graph.edges[idx] -> KeyNotFoundException (:21)
    std::vector<char> liked((size_t)g.n, 0);                    // List<int>.Contains, as a bitmap
    for (int64_t k = b; k < e; k++)
        if (g.raw_type[k] == ET_LIKE) liked[g.raw_dst[k]] = 1;  // :22-23
    out.clear();
    for (int i = 0; i < g.n; i++)                               // :28-31
        if (g.node_type[i] == NT_ITEM && !liked[i]) out.push_back({g.node_id[i], rank[i]});
    std::sort(out.begin(), out.end(), [](const Scored& one, const Scored& another) {   // :35-38
        int result = compare_to(one.score, another.score) * -1;
        if (result != 0) return result < 0;
        return one.id > another.id;
    });
    return ORC_OK;
}

// ---------------------------------------------------------------- small thread helpers (no OpenMP in this image)
int hw_threads() {
    unsigned h = std::thread::hardware_concurrency();
    return h == 0 ? 1 : (int)std::min(h, 32u);
}
template <class F>
void parallel_chunks(int64_t n, F fn) {          // fn(begin, end) on contiguous chunks
    int P = (int)std::min<int64_t>(hw_threads(), std::max<int64_t>(1, n / 65536));
    if (P <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < P; t++) pool.emplace_back([=]() { fn(n * t / P, n * (t + 1) / P); });
    for (auto& th : pool) th.join();
}
void parallel_sort(std::vector<uint64_t>& a) {
    const int64_t n = (int64_t)a.size();
    int P = 1;
    while (P * 2 <= hw_threads() && n / (P * 2) >= 65536) P *= 2;
    if (P == 1) { std::sort(a.begin(), a.end()); return; }
    auto cut = [&](int t) { return a.begin() + n * t / P; };
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < P; t++) pool.emplace_back([&, t]() { std::sort(cut(t), cut(t + 1)); });
        for (auto& th : pool) th.join();
    }
    for (int w = 1; w < P; w *= 2) {
        std::vector<std::thread> pool;
        for (int t = 0; t + w < P; t += 2 * w)
            pool.emplace_back([&, t, w]() { std::inplace_merge(cut(t), cut(t + w), cut(std::min(P, t + 2 * w))); });
        for (auto& th : pool) th.join();
    }
}

// ---------------------------------------------------------------- synthetic generator (repo spec)
struct SynthSpec {          // must mirror rwr_synth_spec in include/rwr_b200.h
    uint64_t seed;
    int32_t n_users, n_items, n_third;
    int32_t authorship_per_mille;
    int64_t n_like, n_friend, n_follow, n_mention;
    int32_t undefined_per_mille;
    int32_t scramble;
    int32_t p1_byte;
    int32_t reserved;
};

inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
    z ^= z >> 27; z *= 0x94d049bb133111ebULL;
    z ^= z >> 31;
    return z;
}
inline uint64_t H(uint64_t seed, uint64_t j, uint64_t k) {
    return mix64(mix64(seed + 0x9E3779B97F4A7C15ULL * (j + 1)) + 0xD1B54A32D192ED03ULL * (k + 1));
}
inline int ceil_log2(uint64_t n) {
    int L = 0;
    while ((1ULL << L) < n) L++;
    return L;
}
inline uint64_t draw(const SynthSpec& s, uint64_t j, int which, uint64_t range) {
    const int L = ceil_log2(range);
    uint64_t v = 0, h = 0;
    for (int l = 0; l < L; l++) {
        if ((l & 7) == 0) h = H(s.seed, j, (uint64_t)(which * 4 + (l >> 3)));
        uint64_t byte = (h >> (8 * (l & 7))) & 255;
        v |= (uint64_t)(byte < (uint64_t)s.p1_byte) << l;
    }
    return v % range;
}
inline uint64_t perm(const SynthSpec& s, uint64_t x, uint64_t range, uint64_t salt) {
    if (!s.scramble) return x;
    return (x * 2654435761ULL + mix64(s.seed ^ salt) % range) % range;
}
constexpr uint64_t SALT_U = 0x1111111111111111ULL, SALT_T = 0x2222222222222222ULL;
constexpr uint64_t SALT_UNDEF = 0xF1E2D3C4B5A69788ULL, SALT_MENTION = 0xA5A5A5A55A5A5A5AULL;
constexpr uint64_t INVALID_KEY = ~0ULL;
enum { CLS_LIKE = 0, CLS_FRIEND = 1, CLS_FOLLOW = 2, CLS_AUTHOR = 3, CLS_MENTION = 4 };

inline uint64_t make_key(uint64_t src, int cls, uint64_t dst) { return (src << 31) | ((uint64_t)cls << 28) | dst; }

struct SynthOut {
    int32_t n = 0;
    std::vector<int64_t> node_id;
    std::vector<int32_t> node_type;
    std::vector<int32_t> src, dst, etype;
    std::vector<double> w;
};

int synth_generate(const SynthSpec& s, SynthOut& o) {
    const uint64_t U = (uint64_t)s.n_users, T = (uint64_t)s.n_items, X = (uint64_t)s.n_third;
    if (s.n_users < 1 || s.n_items < 0 || s.n_third < 0) return ORC_E_INVALID;
    const uint64_t N = U + T + X;
    if (N >= (1ULL << 28)) return ORC_E_INVALID;
    if ((s.n_like > 0 && T == 0) || (s.n_follow > 0 && X == 0)) return ORC_E_INVALID;
    const uint64_t nrel = T + (uint64_t)s.n_like + (uint64_t)s.n_friend + (uint64_t)s.n_follow + (uint64_t)s.n_mention;
    std::vector<uint64_t> keys(2 * nrel, INVALID_KEY);
    const uint64_t r_like = T, r_friend = r_like + (uint64_t)s.n_like, r_follow = r_friend + (uint64_t)s.n_friend,
                   r_mention = r_follow + (uint64_t)s.n_follow;
    parallel_chunks((int64_t)nrel, [&](int64_t jb, int64_t je) {
    for (int64_t jj = jb; jj < je; jj++) {
        const uint64_t j = (uint64_t)jj;
        uint64_t k0 = INVALID_KEY, k1 = INVALID_KEY;
        if (j < r_like) {                                        // authorship of item j
            if ((H(s.seed, j, 15) % 1000) < (uint64_t)s.authorship_per_mille) {
                uint64_t a = perm(s, draw(s, j, 0, U), U, SALT_U), it = U + j;
                k0 = make_key(a, CLS_AUTHOR, it);
                k1 = make_key(it, CLS_AUTHOR, a);
            }
        } else if (j < r_friend) {                               // like
            uint64_t u = perm(s, draw(s, j, 0, U), U, SALT_U), it = U + perm(s, draw(s, j, 1, T), T, SALT_T);
            k0 = make_key(u, CLS_LIKE, it);
            k1 = make_key(it, CLS_LIKE, u);
        } else if (j < r_follow) {                               // friendship (possibly retyped UNDEFINED)
            uint64_t u = perm(s, draw(s, j, 0, U), U, SALT_U), v = perm(s, draw(s, j, 1, U), U, SALT_U);
            if (u != v) {
                k0 = make_key(u, CLS_FRIEND, v);
                k1 = make_key(v, CLS_FRIEND, u);
            }
        } else if (j < r_mention) {                              // follow on a third-party user
            uint64_t u = perm(s, draw(s, j, 0, U), U, SALT_U), x = U + T + draw(s, j, 1, X);
            k0 = make_key(u, CLS_FOLLOW, x);
            k1 = make_key(x, CLS_FOLLOW, u);
        } else {                                                 // mention, directed
            uint64_t u = perm(s, draw(s, j, 0, U), U, SALT_U), v = perm(s, draw(s, j, 1, U), U, SALT_U);
            if (u != v) k0 = make_key(u, CLS_MENTION, v);
        }
        keys[2 * j] = k0;
        keys[2 * j + 1] = k1;
    }
    });
    parallel_sort(keys);
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    while (!keys.empty() && keys.back() == INVALID_KEY) keys.pop_back();

    o.n = (int32_t)N;
    o.node_id.resize(N);
    o.node_type.resize(N);
    for (uint64_t i = 0; i < N; i++) {
        if (i < U) { o.node_id[i] = 1000000000LL + (int64_t)i; o.node_type[i] = NT_USER; }
        else if (i < U + T) { o.node_id[i] = 5000000000000LL + (int64_t)(i - U); o.node_type[i] = NT_ITEM; }
        else { o.node_id[i] = 2000000000LL + (int64_t)(i - U - T); o.node_type[i] = NT_ETC; }
    }
    const size_t E = keys.size();
    o.src.resize(E); o.dst.resize(E); o.etype.resize(E); o.w.resize(E);
    parallel_chunks((int64_t)E, [&](int64_t eb, int64_t eend) {
    for (int64_t ee = eb; ee < eend; ee++) {
        const uint64_t k = keys[ee];
        const uint64_t sN = k >> 31, dN = k & ((1ULL << 28) - 1);
        const int cls = (int)((k >> 28) & 7);
        int et = ET_UNDEFINED;
        double wt = 1.0;
        switch (cls) {
            case CLS_LIKE: et = ET_LIKE; break;
            case CLS_FRIEND: {
                uint64_t lo = std::min(sN, dN), hi = std::max(sN, dN);
                bool undef = (mix64(s.seed ^ SALT_UNDEF ^ ((lo << 32) | hi)) % 1000) < (uint64_t)s.undefined_per_mille;
                et = undef ? ET_UNDEFINED : ET_FRIENDSHIP;
            } break;
            case CLS_FOLLOW: et = ET_FOLLOW; break;
            case CLS_AUTHOR: et = ET_AUTHORSHIP; break;
            case CLS_MENTION:
                et = ET_MENTION;
                wt = (double)(1 + (mix64(s.seed ^ SALT_MENTION ^ ((sN << 32) | dN)) & 127)) / 32.0;
                break;
        }
        o.src[ee] = (int32_t)sN; o.dst[ee] = (int32_t)dN; o.etype[ee] = et; o.w[ee] = wt;
    }
    });
    return ORC_OK;
}

}  // namespace

// ======================================================================= C entry points (ctypes)
extern "C" {

// Links may arrive in any order; they are grouped by source keeping their relative order
// (== per-source insertion order of Dictionary<int, List<ForwardLink>>).
void* orc_graph_create(int32_t n_nodes, const int64_t* node_id, const int32_t* node_type, int64_t n_links,
                       const int32_t* src, const int32_t* dst, const int32_t* etype, const double* w) {
    if (n_nodes < 0 || n_links < 0) return nullptr;
    for (int64_t k = 0; k < n_links; k++)
        if (src[k] < 0 || src[k] >= n_nodes) return nullptr;
    OrcGraph* g = new OrcGraph();
    g->n = n_nodes;
    g->node_id.assign(node_id, node_id + n_nodes);
    g->node_type.assign(node_type, node_type + n_nodes);
    g->raw_ptr.assign((size_t)n_nodes + 1, 0);
    for (int64_t k = 0; k < n_links; k++) g->raw_ptr[src[k] + 1]++;
    for (int32_t i = 0; i < n_nodes; i++) g->raw_ptr[i + 1] += g->raw_ptr[i];
    g->raw_dst.resize(n_links); g->raw_type.resize(n_links); g->raw_w.resize(n_links);
    std::vector<int64_t> cur(g->raw_ptr.begin(), g->raw_ptr.end() - 1);
    for (int64_t k = 0; k < n_links; k++) {
        int64_t p = cur[src[k]]++;
        g->raw_dst[p] = dst[k]; g->raw_type[p] = etype[k]; g->raw_w[p] = w[k];
    }
    return g;
}

void orc_graph_destroy(void* h) { delete (OrcGraph*)h; }

int orc_graph_build(void* h) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g) return ORC_E_INVALID;
    // IndexOutOfRangeException at Model.cs:87 in the reference; reported at build time here
    for (size_t k = 0; k < g->raw_dst.size(); k++)
        if (g->raw_type[k] != ET_UNDEFINED && (g->raw_dst[k] < 0 || g->raw_dst[k] >= g->n)) return ORC_E_BADINDEX;
    return build_graph(*g);
}

int64_t orc_graph_nnz(void* h) {
    OrcGraph* g = (OrcGraph*)h;
    return (g && g->built) ? (int64_t)g->col.size() : (int64_t)ORC_E_NOT_BUILT;
}

int orc_graph_get_csr(void* h, int64_t* row_ptr, int32_t* col, double* val) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built) return ORC_E_NOT_BUILT;
    if (row_ptr) std::memcpy(row_ptr, g->row_ptr.data(), sizeof(int64_t) * g->row_ptr.size());
    if (col && !g->col.empty()) std::memcpy(col, g->col.data(), sizeof(int32_t) * g->col.size());
    if (val && !g->val.empty()) std::memcpy(val, g->val.data(), sizeof(double) * g->val.size());
    return ORC_OK;
}

// mode: 0 = run(int n_iter), 1 = run(double thr), 2 = run() [thr = (1/double.MaxValue) * N]
// seed < 0 selects the uniform-restart constructor (literal form only).
int orc_model_run(void* h, int32_t seed, double damping, int32_t mode, int32_t n_iter, double thr, int32_t literal,
                  int64_t max_iter, double* rank_out, int64_t* iters_out) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built) return ORC_E_NOT_BUILT;
    Model m(*g, damping, seed, literal != 0 || seed < 0);
    int rc = ORC_OK;
    if (mode == 0) {
        m.run_fixed(n_iter);
    } else {
        double threshold = (mode == 2) ? (1 / std::numeric_limits<double>::max()) * g->n : thr;   // Model.cs:53
        if (!m.run_threshold(threshold, max_iter)) rc = 1;
    }
    if (rank_out && g->n) std::memcpy(rank_out, m.rank.data(), sizeof(double) * g->n);
    if (iters_out) *iters_out = m.nDeliver;
    return rc;
}

// NOT the reference: the same seeded fixed-count iteration (collapsed form) with every accumulator in x87 extended
// precision (64-bit significand) and only the final ranks rounded to double -- a yardstick for how much of a difference
// between two double-precision evaluations is rounding noise of the summation order (oracle/rounding_study.py).
int orc_model_run_extended(void* h, int32_t seed, double damping, int32_t n_iter, double* rank_out) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built || seed < 0 || seed >= g->n) return ORC_E_INVALID;
    const int n = g->n;
    std::vector<long double> rank(n, 0.0L), next(n, 0.0L);
    rank[seed] = (long double)n;
    const long double omc = 1.0L - (long double)damping;
    for (int it = 0; it < n_iter; it++) {
        for (int i = 0; i < n; i++) {
            const int64_t b = g->row_ptr[i], e = g->row_ptr[i + 1];
            if (e > b) {
                const long double rw = omc * rank[i];
                for (int64_t w = b; w < e; w++) next[g->col[w]] += rw * (long double)g->val[w];
                next[seed] += rank[i] - rw;
            } else {
                next[seed] += rank[i];
            }
        }
        for (int i = 0; i < n; i++) { rank[i] = next[i]; next[i] = 0.0L; }
    }
    for (int i = 0; i < n; i++) rank_out[i] = (double)rank[i];
    return ORC_OK;
}

// Recommendation(idx, c, nIter[, topN]); `damping` is the already widened float (Recommender.cs:16).
// top_n < 0: the 3-argument overload (full ranking).  The 4-argument overload returns the WHOLE list when
// topN <= 0 (the `Count == topN` test at Recommender.cs:47 never fires) -- kept.
// Returns the number of pairs (writes at most `cap`), or a negative error.
int64_t orc_recommend(void* h, int32_t seed, double damping, int32_t n_iter, int32_t literal, int32_t has_top_n,
                      int32_t top_n, int64_t* ids, double* scores, int64_t cap) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built) return ORC_E_NOT_BUILT;
    Model m(*g, damping, seed, literal != 0);
    m.run_fixed(n_iter);
    std::vector<Scored> rec;
    int rc = rank_candidates(*g, seed, m.rank.data(), rec);
    if (rc != ORC_OK) return rc;
    int64_t count = (int64_t)rec.size();
    if (has_top_n && top_n > 0 && count > top_n) count = top_n;
    for (int64_t i = 0; i < count && i < cap; i++) {
        if (ids) ids[i] = rec[i].id;
        if (scores) scores[i] = rec[i].score;
    }
    return count;
}

// Recommender.cs:19-38 on a caller-supplied rank vector (lets tests rank GPU scores with reference rules)
int64_t orc_rank_scores(void* h, int32_t seed, const double* rank, int64_t* ids, double* scores, int64_t cap) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g) return ORC_E_INVALID;
    std::vector<Scored> rec;
    int rc = rank_candidates(*g, seed, rank, rec);
    if (rc != ORC_OK) return rc;
    for (int64_t i = 0; i < (int64_t)rec.size() && i < cap; i++) {
        if (ids) ids[i] = rec[i].id;
        if (scores) scores[i] = rec[i].score;
    }
    return (int64_t)rec.size();
}

// Thread-per-seed fan-out mirroring Program.cs:11 / :61-66 (<= 10 concurrent experiments).  Collapsed form.
int orc_recommend_many(void* h, const int32_t* seeds, int32_t n_seeds, double damping, int32_t n_iter, int32_t top_n,
                       int32_t n_threads, int64_t* ids, double* scores, int32_t* counts) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built) return ORC_E_NOT_BUILT;
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    std::vector<int> rcs(n_threads, ORC_OK);
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t]() {
            for (int32_t s = t; s < n_seeds; s += n_threads) {
                int64_t c = orc_recommend(h, seeds[s], damping, n_iter, 0, 1, top_n, ids + (int64_t)s * top_n,
                                          scores + (int64_t)s * top_n, top_n);
                if (c < 0) { rcs[t] = (int)c; counts[s] = 0; } else counts[s] = (int32_t)c;
            }
        });
    }
    for (auto& th : pool) th.join();
    for (int rc : rcs) if (rc != ORC_OK) return rc;
    return ORC_OK;
}

// Model.run(n_iter) for several seeds at once, one thread per seed in flight (<= n_threads), collapsed form.
// checksum[s] = rank[seed_s] so the work cannot be optimised away.
int orc_run_many(void* h, const int32_t* seeds, int32_t n_seeds, double damping, int32_t n_iter, int32_t n_threads,
                 double* checksum) {
    OrcGraph* g = (OrcGraph*)h;
    if (!g || !g->built) return ORC_E_NOT_BUILT;
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t]() {
            for (int32_t s = t; s < n_seeds; s += n_threads) {
                Model m(*g, damping, seeds[s], false);
                m.run_fixed(n_iter);
                checksum[s] = (seeds[s] >= 0 && seeds[s] < g->n) ? m.rank[seeds[s]] : 0.0;
            }
        });
    }
    for (auto& th : pool) th.join();
    return ORC_OK;
}

// Experiment.cs:121-128; `test` must be sorted ascending (HashSet<long>.Contains as a binary search)
void orc_evaluate(const int64_t* ids, int64_t n, const int64_t* test_sorted, int64_t n_test, int32_t* hits, double* avg_precision) {
    int nHits = 0;
    double sumPrecision = 0;
    for (int64_t i = 0; i < n; i++) {
        if (std::binary_search(test_sorted, test_sorted + n_test, ids[i])) {
            nHits += 1;
            sumPrecision += (double)nHits / (i + 1);
        }
    }
    *hits = nHits;
    *avg_precision = (nHits == 0) ? 0 : sumPrecision / nHits;   // Experiment.cs:136
}

// ---- synthetic generator: create -> sizes -> copy -> destroy
void* orc_synth_create(const void* spec) {
    SynthSpec s;
    std::memcpy(&s, spec, sizeof(s));
    SynthOut* o = new SynthOut();
    if (synth_generate(s, *o) != ORC_OK) { delete o; return nullptr; }
    return o;
}
void orc_synth_sizes(void* h, int32_t* n_nodes, int64_t* n_links) {
    SynthOut* o = (SynthOut*)h;
    *n_nodes = o->n;
    *n_links = (int64_t)o->src.size();
}
void orc_synth_copy(void* h, int64_t* node_id, int32_t* node_type, int32_t* src, int32_t* dst, int32_t* etype, double* w) {
    SynthOut* o = (SynthOut*)h;
    const size_t N = o->node_id.size(), E = o->src.size();
    if (node_id) std::memcpy(node_id, o->node_id.data(), 8 * N);
    if (node_type) std::memcpy(node_type, o->node_type.data(), 4 * N);
    if (E) {
        if (src) std::memcpy(src, o->src.data(), 4 * E);
        if (dst) std::memcpy(dst, o->dst.data(), 4 * E);
        if (etype) std::memcpy(etype, o->etype.data(), 4 * E);
        if (w) std::memcpy(w, o->w.data(), 8 * E);
    }
}
void orc_synth_destroy(void* h) { delete (SynthOut*)h; }

int orc_sizeof_synth_spec(void) { return (int)sizeof(SynthSpec); }

}  // extern "C"
