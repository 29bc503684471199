// Model.cs -- drop-in for Recommenders/RWRBased/Model.cs: constructors, run() / run(double) / run(int) and `rank`.
// deliverRanks / updateRanks / checkConvergence run on the device inside the run* calls.  Source only.
using System;
using Recommenders.RWRBased.Native;

namespace Recommenders.RWRBased {
    public class Model : IDisposable {
        public Graph graph;
        public int nNodes;
        public double dampingFactor;
        public int nIterations;                 // deliverRanks() calls performed by the last run
        readonly int seed;                      // -1: uniform restart (Model.cs:14-31)
        ResultHandle result;
        double[] rankCache;

        public Model(Graph graph, double dampingFactor) : this(graph, dampingFactor, -1) { }
        public Model(Graph graph, double dampingFactor, int targetNode) {
            this.graph = graph; this.nNodes = graph.size(); this.dampingFactor = dampingFactor; this.seed = targetNode;
        }

        public void run() { runThreshold(0.0); }                    // thr <= 0 selects (1/double.MaxValue) * N, Model.cs:53
        public void run(double threshold) { runThreshold(threshold); }
        public void run(int nIterations) {
            Release();
            RwrNative.Check(RwrNative.rwr_run_fixed(graph.handle, new[] { seed }, 1, dampingFactor, nIterations, RwrNative.FP64, out result));
            this.nIterations = nIterations;
        }
        void runThreshold(double thr) {
            Release();
            var iters = new int[1];
            RwrNative.Check(RwrNative.rwr_run_threshold(graph.handle, new[] { seed }, 1, dampingFactor, thr, 0, RwrNative.FP64, iters, out result));
            nIterations = iters[0];
        }

        public double[] rank {
            get {
                if (rankCache != null) return rankCache;
                var r = new double[nNodes];
                if (result == null) {               // constructor state, Model.cs:24 / :44
                    for (int i = 0; i < nNodes; i++) r[i] = seed < 0 ? 1.0 : (i == seed ? nNodes : 0.0);
                    return r;
                }
                RwrNative.Check(RwrNative.rwr_scores(result, 0, r));
                return rankCache = r;
            }
        }

        void Release() { rankCache = null; if (result != null) { result.Dispose(); result = null; } }
        public void Dispose() { Release(); }
    }
}
