// experiment_caller.cpp -- ONE caller of the RWR path, compiled against two implementations of the same surface:
//
//   caller_b200       (default)          RWRBased.hpp: the drop-in over librwr_b200.so -- needs a CUDA device, no CPU fallback
//   caller_reference  (-DUSE_REFERENCE)  oracle/_ref/reference_rwr.hpp: the reference's OWN Graph.cs / Model.cs / Recommender.cs as
//                                        oracle/cs2cpp.py respells them for g++ (test infrastructure; CPU)
//
// The body below is what TweetRecommender does with the library, in the C++ spelling cs2cpp gives C# (class variables are
// pointers): DataLoader.cs:40-41 / :60-77 fill `nodes` and `edges`, Experiment.cs:104-109 builds the graph and asks for a
// recommendation, :121-128 walks it.  Nothing in it knows which implementation it runs on; tests/test_cpp_host.py runs both
// on the same graphs and compares what they print (ids identical, scores to 1e-12).
//
//   caller_x <graph file> <seed> <nIterations> [topN]
//   graph file (text): N, then N lines `id type`, then E, then E lines `src dst type weight` (weight as a C99 hex double),
//   links grouped by source in insertion order.  Output: `count`, then one `id score(hex)` line per recommended item, then
//   the rank vector of Model(graph, c, seed).run(nIterations) as `rank i hex`.
#include <cstdio>
#include <cstdlib>
#include <exception>

#ifdef USE_REFERENCE
#include "../../oracle/_ref/reference_rwr.hpp"
using namespace Recommenders_RWRBased;
using namespace bcl;
#else
#include "RWRBased.hpp"
using namespace Recommenders::RWRBased;
#endif

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <graph file> <seed> <nIterations> [topN]\n", argv[0]);
        return 2;
    }
    std::FILE* f = std::fopen(argv[1], "r");
    if (!f) { std::perror(argv[1]); return 2; }
    const int idxTargetUser = std::atoi(argv[2]), nIterations = std::atoi(argv[3]);
    try {
        // ---- DataLoader: allNodes / allLinks
        Dictionary<int, Node> nodes;
        Dictionary<int, List<ForwardLink>> edges;
        int nNodes = 0;
        long long nLinks = 0;
        if (std::fscanf(f, "%d", &nNodes) != 1) return 2;
        for (int i = 0; i < nNodes; i++) {
            long long id;
            int type;
            if (std::fscanf(f, "%lld %d", &id, &type) != 2) return 2;
            nodes.Add(i, Node(id, (NodeType)type));                                  // DataLoader.cs:40-41
        }
        if (std::fscanf(f, "%lld", &nLinks) != 1) return 2;
        for (long long k = 0; k < nLinks; k++) {
            int idxSourceNode, idxTargetNode, type;
            double weight;
            if (std::fscanf(f, "%d %d %d %la", &idxSourceNode, &idxTargetNode, &type, &weight) != 4) return 2;
            if (!edges.ContainsKey(idxSourceNode))                                   // DataLoader.cs:61-62
                edges.Add(idxSourceNode, List<ForwardLink>());
            ForwardLink link = ForwardLink(idxTargetNode, (EdgeType)type, weight);    // DataLoader.cs:73-74
            edges[idxSourceNode].Add(link);
        }
        std::fclose(f);

        // ---- Experiment.cs:104-109
        Graph* graph = new Graph(nodes, edges);
        graph->buildGraph();
        Recommender* recommender = new Recommender(graph);
        auto recommendation = argc > 4 ? recommender->Recommendation(idxTargetUser, 0.15f, nIterations, std::atoi(argv[4]))
                                       : recommender->Recommendation(idxTargetUser, 0.15f, nIterations);
        std::printf("%d\n", recommendation.Count());
        for (int i = 0; i < recommendation.Count(); i++)                             // Experiment.cs:123
            std::printf("%lld %a\n", (long long)recommendation[i].Key, recommendation[i].Value);

        // ---- a caller that drives the Model itself (Model.cs:33-50, :68-73)
        Model* model = new Model(graph, (double)0.15f, idxTargetUser);
        model->run(nIterations);
        for (int i = 0; i < model->nNodes; i++) std::printf("rank %d %a\n", i, (double)model->rank[i]);
        return 0;
    } catch (const KeyNotFoundException& e) {
        std::fprintf(stderr, "KeyNotFoundException: %s\n", e.what());
        return 3;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
