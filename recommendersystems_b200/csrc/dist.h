// dist.h -- row-partitioned mode: NCCL communicator handle and the per-iteration exchange (dist.cu).
#pragma once

#include "graph.h"

int dist_rank(const rwr_comm* c);          // 0 when c is null
int dist_n_ranks(const rwr_comm* c);       // 1 when c is null
// After an iteration on a row slice: every rank's x_next slice to all ranks (allGather over NVLink, in place), then
// the sum over ranks of the two scalars (restart mass S, L1 residual) at `two_doubles`.
void dist_exchange(rwr_graph* g, void* x_next, size_t elt, double* two_doubles);
// every rank's slice of `vec` (elements of `elt` bytes, indexed by global row) to all ranks, in place
void dist_allgather_rows(rwr_graph* g, void* vec, size_t elt);
