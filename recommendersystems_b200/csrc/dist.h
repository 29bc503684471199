// dist.h -- row-partitioned mode: NCCL communicator handle and the per-iteration exchange (dist.cu).
#pragma once

#include "graph.h"

int dist_rank(const rwr_comm* c);          // 0 when c is null
int dist_n_ranks(const rwr_comm* c);       // 1 when c is null
// After an iteration on a row slice: every rank's x_next slice to all ranks (NCCL, in place; skipped when x_next is
// null because the epilogue kernel already stored the slice into the peers' buffers), then the sum over ranks of the
// two scalars (restart mass S, L1 residual) at `two_doubles` -- which is also the cross-rank barrier of the iteration.
void dist_exchange(rwr_graph* g, void* x_next, size_t elt, double* two_doubles);
// allocates the two peer-mapped gather vectors of a partitioned graph and maps every peer's pair (CUDA IPC)
void dist_setup_p2p(rwr_graph* g);
void dist_release_p2p(rwr_graph* g);
// every rank's slice of `vec` (elements of `elt` bytes, indexed by global row) to all ranks, in place
void dist_allgather_rows(rwr_graph* g, void* vec, size_t elt);
