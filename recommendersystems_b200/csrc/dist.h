// dist.h -- row-partitioned mode: NCCL communicator handle and the per-iteration exchange (dist.cu).
#pragma once

#include "graph.h"

bool dist_is_fake(const rwr_comm* c);      // RWR_FAKE_COMM probe: one slice on one GPU, no exchange
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

// ---- overlapped exchange (default from 5 ranks on with peer mapping; RWR_DIST_LEGACY=1 keeps the peer stores of the
// epilogue kernel, RWR_DIST_OVERLAP=1 forces it at 2 ranks).  The edge stream of every rank is cut into slice-aligned
// blocks (stream.cu), k_finish_ws writes the rank's slice of the next x locally only, and the NEXT k_spmv_ws carries it
// to the peers itself: a 17th warp per CTA drives TMA bulk copies into the peers' vectors (peer rank+1 first, then the
// 8-byte arrival tag) while the other warps gather -- from the rank's own slice first, then block by block from the
// slices whose tags have arrived.
struct DistSync {
    unsigned long long arrive[8];     // arrive[s]: tag of the newest slice of rank s that is complete in this rank's vectors
};
bool dist_overlap_wanted(const rwr_graph* g);
constexpr size_t DIST_PUSH_SMEM_BYTES = 13 * 1024;     // shared memory the push warp needs behind the hub table (3 x 4 KB + barriers)
void dist_barrier(rwr_graph* g);

// ---- collectives of the partitioned build (graph.cu): in-place sums over the ranks and the all-to-all of the transpose
enum { DIST_U32 = 3, DIST_I64 = 4, DIST_F64 = 8 };                        // ncclDataType_t values
void dist_allreduce_sum(rwr_graph* g, void* buf, size_t count, int dtype);
void dist_allreduce_max_u32(rwr_graph* g, u32* buf, size_t count);
// rank r sends send_cnt[d] elements (of `elt` bytes) starting at send_off[d] of `send` to every rank d and receives
// recv_cnt[s] elements from every rank s at recv_off[s] of `recv` (one grouped set of ncclSend / ncclRecv)
void dist_alltoallv(rwr_graph* g, const void* send, const size_t* send_off, const size_t* send_cnt, void* recv,
                    const size_t* recv_off, const size_t* recv_cnt, size_t elt);
