// Multi-GPU communicator of a bin-sharded run: NCCL reached through dlopen (the library loads without it; only
// gnb_comm_* and sharded sessions need it).  One process per GPU; see include/ganon_b200.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "gnb_internal.h"

struct gnb_comm
{
    int   rank = 0, n_ranks = 1, device = 0;
    void *nccl = nullptr;    // ncclComm_t of the compute stream's collectives (tuple exchange), used in GPU-turn order
    void *nccl_in = nullptr; // ncclComm_t of the ingest stream's collectives (read block slices), used in submission order
    int   nccl_version = 0;
    ~gnb_comm();
};

namespace gnb
{
// all ranks contribute `bytes_per_rank` bytes at send (may alias recv + rank * bytes_per_rank); recv holds n_ranks * bytes_per_rank
int comm_all_gather(void *nccl_comm, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st);
} // namespace gnb
