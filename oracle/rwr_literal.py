"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Pure-Python, object-for-object restatement of the reference's RWR library
(`Recommenders/RWRBased/{Graph,Model,Recommender}.cs`).  Python floats are IEEE-754
binary64 and every statement below is executed in the reference's order, so results
are bit-identical to a strict-double execution of the C# code.  It is used
  * to re-derive the known-answer vector of SURVEY.md section 8c,
  * to generate the fixtures under tests/golden/ (see oracle/make_golden.py),
  * as an independent cross-check of the C++ oracle (oracle/rwr_oracle.cpp).

PINNING: the reference ships no tests, fixtures or golden vectors and no C# runtime exists in
this image.  Since round 2 the golden files this module writes are checked against the reference's
own sources compiled for g++ (oracle/cs2cpp.py -> oracle/_ref/libref.so, tests/test_reference_pin.py):
every vector is reproduced bit for bit.  This file stays as the third, independent restatement.

Only small graphs: the restart loops are the reference's literal O(N^2) form.
"""
from __future__ import annotations

import math
from functools import cmp_to_key

# Recommender.cs:4-5 -- enum integer values are part of the ABI
UNDEFINED_NODE, USER, ITEM, ETC_NODE = 0, 1, 2, 3
E_UNDEFINED, LIKE, FRIENDSHIP, FOLLOW, MENTION, AUTHORSHIP, PURCHASE, E_ETC = range(8)

DOUBLE_MAX = 1.7976931348623157e308


def widen_float(x: float) -> float:
    """`float dampingFactor` widened to double (Recommender.cs:14 -> :16)."""
    import struct
    return struct.unpack("<f", struct.pack("<f", x))[0]


class Node:
    """Graph.cs:4-17"""
    __slots__ = ("id", "type")

    def __init__(self, id: int, type: int = UNDEFINED_NODE):
        self.id = id
        self.type = type


class ForwardLink:
    """Graph.cs:19-35 (a C# struct: copies are by value)"""
    __slots__ = ("targetNode", "type", "weight")

    def __init__(self, targetNode: int, type: int = E_UNDEFINED, weight: float = 1.0):
        self.targetNode = targetNode
        self.type = type
        self.weight = weight

    def copy(self) -> "ForwardLink":
        return ForwardLink(self.targetNode, self.type, self.weight)


class Graph:
    """Graph.cs:37-94"""

    def __init__(self, nodes: dict, edges: dict):
        self.nodes = nodes          # Dictionary<int, Node>
        self.edges = edges          # Dictionary<int, List<ForwardLink>>
        self.graph = {}             # Dictionary<int, ForwardLink[]>

    def buildGraph(self) -> None:   # Graph.cs:51-88
        for i in range(len(self.nodes)):
            forwardLinks = None
            if i in self.edges:                                   # :55
                nExplicitLinks = 0
                for forwardLink in self.edges[i]:                 # :58
                    if forwardLink.type != E_UNDEFINED:
                        nExplicitLinks += 1
                if nExplicitLinks > 0:                            # :64
                    forwardLinks = [None] * nExplicitLinks
                    idx = 0
                    sumWeights = 0.0
                    for link in self.edges[i]:                    # :71
                        if link.type != E_UNDEFINED:
                            forwardLinks[idx] = link.copy()       # struct copy
                            idx += 1
                            sumWeights += link.weight             # :75
                    for f in range(nExplicitLinks):               # :80
                        # C# `/=` on doubles: IEEE division, 0/0 -> NaN, x/0 -> +-Inf, no exception
                        forwardLinks[f].weight = _div(forwardLinks[f].weight, sumWeights)
            if i in self.graph:                                   # Dictionary.Add throws on a duplicate key
                raise ValueError("ArgumentException: buildGraph() called twice")
            self.graph[i] = forwardLinks                          # :86

    def size(self) -> int:          # Graph.cs:91
        return len(self.nodes)


def _div(a: float, b: float) -> float:
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0.0 or math.isnan(a):
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


class Model:
    """Model.cs:5-116"""

    def __init__(self, graph: Graph, dampingFactor: float, targetNode: int | None = None):
        self.graph = graph
        self.nNodes = graph.size()
        self.dampingFactor = dampingFactor
        n = self.nNodes
        self.rank = [0.0] * n
        self.nextRank = [0.0] * n
        self.restart = [0.0] * n
        if targetNode is None:                                    # Model.cs:14-31
            for i in range(n):
                self.rank[i] = 1.0
                self.nextRank[i] = 0.0
                self.restart[i] = 1.0 / n
        else:                                                     # Model.cs:33-50
            for i in range(n):
                self.rank[i] = float(n) if i == targetNode else 0.0
                self.nextRank[i] = 0.0
                self.restart[i] = 1.0 if i == targetNode else 0.0
        self.nDeliver = 0           # not in the reference: counts deliverRanks() calls

    def run(self, arg=None):
        if arg is None:                                           # Model.cs:52-55
            threshold = (1 / DOUBLE_MAX) * self.graph.size()
            return self.run(threshold)
        if isinstance(arg, bool):
            raise TypeError
        if isinstance(arg, int):                                  # Model.cs:68-73
            for _ in range(arg):
                self.deliverRanks()
                self.updateRanks()
            return
        threshold = float(arg)                                    # Model.cs:57-66
        while True:
            self.deliverRanks()
            if self.checkConvergence(threshold):
                self.updateRanks()
                return
            self.updateRanks()

    def deliverRanks(self) -> None:                               # Model.cs:76-100
        self.nDeliver += 1
        forwardLinks = self.graph.graph
        nNodes = self.nNodes
        rank, nextRank, restart = self.rank, self.nextRank, self.restart
        for i in range(nNodes):
            links = forwardLinks[i]
            if links is not None and len(links) > 0:
                rank_randomWalk = (1 - self.dampingFactor) * rank[i]          # :84
                for link in links:
                    nextRank[link.targetNode] += rank_randomWalk * link.weight  # :87
                rank_restart = rank[i] - rank_randomWalk                        # :91
                for r in range(nNodes):
                    nextRank[r] += rank_restart * restart[r]                    # :93
            else:
                for r in range(nNodes):
                    nextRank[r] += rank[i] * restart[r]                         # :97

    def updateRanks(self) -> None:                                # Model.cs:103-108
        for i in range(self.nNodes):
            self.rank[i] = self.nextRank[i]
            self.nextRank[i] = 0.0

    def checkConvergence(self, threshold: float) -> bool:         # Model.cs:110-115
        diff = 0.0
        for i in range(self.nNodes):
            diff += (self.rank[i] - self.nextRank[i]) if self.rank[i] > self.nextRank[i] \
                else (self.nextRank[i] - self.rank[i])
        return diff < threshold


def _compare_to(x: float, y: float) -> int:
    """System.Double.CompareTo: NaN is smaller than everything and equal to NaN."""
    if x < y:
        return -1
    if x > y:
        return 1
    if x == y:
        return 0
    if math.isnan(x):
        return 0 if math.isnan(y) else -1
    return 1


class Recommender:
    """Recommender.cs:7-52"""

    def __init__(self, graph: Graph):
        self.graph = graph
        self.lastModel = None       # not in the reference: lets tests look at model.rank

    def Recommendation(self, idxTargetUser: int, dampingFactor: float, nIteration: int, topN: int | None = None):
        if topN is not None:                                      # Recommender.cs:42-51
            recommendation = self.Recommendation(idxTargetUser, dampingFactor, nIteration)
            top = []
            for i in range(len(recommendation)):
                top.append(recommendation[i])
                if len(top) == topN:
                    break
            return top
        graph = self.graph
        model = Model(graph, widen_float(dampingFactor), idxTargetUser)  # :16 (float -> double)
        model.run(int(nIteration))                                        # :17
        self.lastModel = model
        linksOfTargetUser = []
        for link in graph.edges[idxTargetUser]:                   # :21 KeyError == KeyNotFoundException
            if link.type == LIKE:
                linksOfTargetUser.append(link.targetNode)
        recommendation = []
        for i in range(model.nNodes):                             # :28
            if graph.nodes[i].type == ITEM and i not in linksOfTargetUser:
                recommendation.append((graph.nodes[i].id, model.rank[i]))

        def comparison(one, another):                             # :35-38
            result = _compare_to(one[1], another[1]) * -1
            if result != 0:
                return result
            return ((one[0] > another[0]) - (one[0] < another[0])) * -1

        recommendation.sort(key=cmp_to_key(comparison))
        return recommendation


def evaluate(recommendation, testSet):
    """Experiment.cs:121-128 and :131-138 -> (nHits, averagePrecision)."""
    nHits = 0
    sumPrecision = 0.0
    for i in range(len(recommendation)):
        if recommendation[i][0] in testSet:
            nHits += 1
            sumPrecision += nHits / (i + 1)
    return nHits, (0.0 if nHits == 0 else sumPrecision / nHits)


def graph_from_flat(node_id, node_type, src, dst, etype, w) -> Graph:
    """Builds the reference's dictionaries from the flattened SoA the C ABI uses
    (links in (source, insertion) order; a source with no links has no key)."""
    nodes = {i: Node(int(node_id[i]), int(node_type[i])) for i in range(len(node_id))}
    edges = {}
    for s, d, t, ww in zip(src, dst, etype, w):
        edges.setdefault(int(s), []).append(ForwardLink(int(d), int(t), float(ww)))
    return Graph(nodes, edges)


def kat_graph_8c():
    """The 8-node known-answer graph of SURVEY.md section 8c."""
    types = [USER, USER, ITEM, ITEM, ITEM, ITEM, ITEM, ETC_NODE]
    ids = [1000, 1001, 5002, 5003, 5004, 5005, 5006, 1007]
    links = [  # (src, dst, type, w) in global insertion order
        (0, 1, FRIENDSHIP, 1.0), (1, 0, FRIENDSHIP, 1.0), (0, 2, LIKE, 1.0), (2, 0, LIKE, 1.0),
        (1, 2, LIKE, 1.0), (2, 1, LIKE, 1.0), (1, 3, LIKE, 1.0), (3, 1, LIKE, 1.0),
        (1, 4, LIKE, 1.0), (1, 2, AUTHORSHIP, 1.0), (2, 1, AUTHORSHIP, 1.0), (1, 0, MENTION, 0.5),
        (0, 7, E_UNDEFINED, 1.0), (7, 0, FOLLOW, 1.0),
    ]
    nodes = {i: Node(ids[i], types[i]) for i in range(8)}
    edges = {}
    for s, d, t, ww in links:
        edges.setdefault(s, []).append(ForwardLink(d, t, ww))
    return nodes, edges
