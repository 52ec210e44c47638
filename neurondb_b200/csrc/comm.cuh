// comm.cuh -- the process's NCCL communicator (comm.cu) as the other translation units see it.
#pragma once
#include "common.cuh"

namespace ndb {

// element types of comm_allreduce_sum (values of ncclDataType_t are resolved inside comm.cu)
enum CommType { COMM_F32 = 0, COMM_I32 = 1, COMM_F64 = 2, COMM_I64 = 3 };

void comm_at_shutdown();
bool comm_ready();
int comm_rank();
int comm_nranks();
// collectives on device buffers, queued on `s`; with one rank (or no communicator) they degenerate
// to a device copy / nothing, so single-process callers need no special case
int comm_allgather(const void *send_dev, void *recv_dev, size_t bytes_per_rank, cudaStream_t s);
int comm_allreduce_sum(void *buf_dev, size_t count, CommType t, cudaStream_t s);
int comm_broadcast(void *buf_dev, size_t bytes, int root, cudaStream_t s);

}  // namespace ndb
