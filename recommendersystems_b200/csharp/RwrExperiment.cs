// RwrExperiment.cs -- the callers' side of the hot path for TweetRecommender/Experiment.cs: methodology masks, the k-fold
// hold-out and the hit / average-precision walk on the device.  With these, the body of the fold loop
// (Experiment.cs:69-138) needs ONE DataLoader.graphConfiguration(Methodology.ALL, -1) per ego network instead of one
// SQLite load per methodology and fold.  Source only; see INTEGRATION.md.
using System;
using System.Collections.Generic;
using Recommenders.RWRBased.Native;

namespace Recommenders.RWRBased {
    public static class RwrExperiment {
        // DataLoader.cs:142-219 + Experiment.cs:84-101 as link-type masks over the graph with every relation loaded
        public static RwrOpts OptionsFor(int methodology) {
            int features, undef, zero;
            RwrNative.Check(RwrNative.rwr_methodology_masks(methodology, out features, out undef, out zero));
            RwrOpts o = RwrOpts.Default();
            o.undefined_type_mask = undef;
            o.zero_weight_type_mask = zero;
            return o;
        }

        // DataLoader.splitLikeHistory (DataLoader.cs:122-140) for `users`, applied by the next buildGraph() of `graph`;
        // testSets[i] receives users[i]'s held-out tweet ids (`loader.testSet`).  A held-out tweet nobody else likes is no node
        // of the graph DataLoader would have built (:291-303, :355-356): the native side drops its remaining links and makes
        // it no candidate, so HIT and AVGPRECISION are those of a reload
        public static void HoldOut(Graph graph, int[] users, int nFolds, int fold, List<long>[] testSets) {
            graph.beforeBuild = g => {
                var ptr = new long[users.Length + 1];
                long total;
                long cap = 0;
                foreach (int u in users) { List<ForwardLink> l; if (g.edges.TryGetValue(u, out l)) cap += l.Count; }
                var ids = new long[Math.Max(cap, 1)];
                RwrNative.Check(RwrNative.rwr_graph_hold_out(g.handle, users, users.Length, nFolds, fold, ptr, ids, cap, out total));
                for (int i = 0; i < users.Length; i++) {
                    testSets[i] = new List<long>();
                    for (long p = ptr[i]; p < ptr[i + 1]; p++) testSets[i].Add(ids[p]);
                }
            };
        }

        // Experiment.cs:121-128, :136 for the held-out users of `graph` (after buildGraph()): nHits and
        // sumPrecision / nHits over the full ranking of Recommendation(u, dampingFactor, nIteration)
        public static void Evaluate(Graph graph, int nUsers, float dampingFactor, int nIteration, int[] hits, double[] avgPrecision) {
            double c = dampingFactor;
            var atK = new int[nUsers]; var nTest = new int[nUsers];
            RwrRunInfo info;
            RwrNative.Check(RwrNative.rwr_evaluate_users(graph.handle, null, nUsers, null, null, c, nIteration, RwrNative.FP64, 10, hits,
                                                         avgPrecision, atK, nTest, out info));
        }
    }
}
