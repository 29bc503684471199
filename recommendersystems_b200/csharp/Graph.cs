// Graph.cs -- drop-in for Recommenders/RWRBased/Graph.cs: same public surface (Node, ForwardLink, Graph with
// nodes / edges / graph, buildGraph(), size()), the work forwarded to librwr_b200.  Source only; see INTEGRATION.md.
using System;
using System.Collections.Generic;
using Recommenders.RWRBased.Native;

namespace Recommenders.RWRBased {
    public struct Node {
        public long id;
        public NodeType type;
        public Node(long id) { this.id = id; this.type = NodeType.UNDEFINED; }
        public Node(long id, NodeType type) { this.id = id; this.type = type; }
    }

    public struct ForwardLink {
        public int targetNode;
        public EdgeType type;
        public double weight;
        public ForwardLink(int targetNode, double weight) { this.targetNode = targetNode; this.type = EdgeType.UNDEFINED; this.weight = weight; }
        public ForwardLink(int targetNode, EdgeType type, double weight) { this.targetNode = targetNode; this.type = type; this.weight = weight; }
    }

    public class Graph : IDisposable {
        public Dictionary<int, Node> nodes;
        public Dictionary<int, List<ForwardLink>> edges;
        Dictionary<int, ForwardLink[]> graphCache;      // materialised lazily from the device CSR
        internal GraphHandle handle;
        // options applied at buildGraph(): Methodology masks (RwrExperiment.OptionsFor), device, ...
        public RwrOpts options = RwrOpts.Default();
        // test users of the k-fold split, applied between the upload and the build (RwrExperiment.HoldOut)
        internal Action<Graph> beforeBuild;

        // The reference's constructor only stores the two dictionaries (Graph.cs:45-48); `edges` is read when buildGraph()
        // runs, so callers may still edit it in between (Experiment.cs:84-101 does, before constructing the graph).
        public Graph(Dictionary<int, Node> nodes, Dictionary<int, List<ForwardLink>> edges) {
            this.nodes = nodes;
            this.edges = edges;
        }

        public void buildGraph() {
            if (handle != null) RwrNative.Check((int)RwrStatus.E_ALREADY_BUILT);      // ArgumentException, Graph.cs:86
            // Flatten `for i in 0..N-1: foreach l in edges[i]` into SoA arrays; the native side copies them to the device.
            int n = nodes.Count;
            long nLinks = 0;
            for (int i = 0; i < n; i++) { List<ForwardLink> l; if (edges.TryGetValue(i, out l)) nLinks += l.Count; }
            var nodeId = new long[n]; var nodeType = new int[n];
            var src = new int[nLinks]; var dst = new int[nLinks]; var et = new int[nLinks]; var w = new double[nLinks];
            long p = 0;
            for (int i = 0; i < n; i++) {
                nodeId[i] = nodes[i].id; nodeType[i] = (int)nodes[i].type;
                List<ForwardLink> l;
                if (!edges.TryGetValue(i, out l)) continue;
                foreach (ForwardLink f in l) { src[p] = i; dst[p] = f.targetNode; et[p] = (int)f.type; w[p] = f.weight; p++; }
            }
            RwrNative.Check(RwrNative.rwr_graph_create(n, nodeId, nodeType, nLinks, src, dst, et, w, ref options, out handle));
            if (beforeBuild != null) beforeBuild(this);
            RwrNative.Check(RwrNative.rwr_graph_build(handle));
        }

        public int size() { return nodes.Count; }

        // KeyNotFoundException of `graph.edges[idxTargetUser]` (Recommender.cs:21): only a MISSING key throws; a key with an
        // empty list is served (the native side runs with empty_seed_ok)
        internal void RequireEdgesEntry(int idx) {
            if (!edges.ContainsKey(idx)) throw new KeyNotFoundException("edges has no entry for node " + idx);
        }
        internal void RequireBuilt() {
            if (handle == null) throw new KeyNotFoundException("buildGraph() has not run (graph.graph[i], Model.cs:79)");
        }

        // `Graph.graph` of the reference (row -> normalised ForwardLink[], null for a row without explicit links); every
        // link keeps its EdgeType (Graph.cs:73-74 copies the whole ForwardLink, only the weight is rewritten at :81).
        public Dictionary<int, ForwardLink[]> graph {
            get {
                if (graphCache != null) return graphCache;
                RequireBuilt();
                int n = nodes.Count;
                var deg = new int[n];
                RwrNative.Check(RwrNative.rwr_graph_get_degrees(handle, deg, null));
                long nnz = 0; foreach (int d in deg) nnz += d;
                var rowPtr = new long[n + 1]; var col = new int[nnz]; var val = new double[nnz]; var typ = new int[nnz];
                RwrNative.Check(RwrNative.rwr_graph_get_csr(handle, rowPtr, col, val));
                RwrNative.Check(RwrNative.rwr_graph_get_csr_types(handle, typ));
                var g = new Dictionary<int, ForwardLink[]>(n);
                for (int i = 0; i < n; i++) {
                    if (deg[i] == 0) { g.Add(i, null); continue; }
                    var row = new ForwardLink[deg[i]];
                    for (int k = 0; k < deg[i]; k++)
                        row[k] = new ForwardLink(col[rowPtr[i] + k], (EdgeType)typ[rowPtr[i] + k], val[rowPtr[i] + k]);
                    g.Add(i, row);
                }
                return graphCache = g;
            }
        }

        public void Dispose() { if (handle != null) handle.Dispose(); }
    }
}
