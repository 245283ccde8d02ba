// az_comm.h -- internal: the point-to-point plumbing of the spatial domain decomposition.
// The reference has no communication layer at all (no NCCL/MPI, device 0 hard-coded);
// this is new functionality.  One implementation ships: NCCL over NVLink/NVSwitch
// (az_comm_nccl.cpp, libnccl loaded with dlopen so that single-GPU use needs no NCCL).  The emulator
// build of tests/emu links an in-process mailbox instead (test infrastructure).
#pragma once
#include <stddef.h>
#include <vector>
#include <string>

#ifdef AZ_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

struct AzP2P { // one exchange with a peer; either side may be empty
    int peer;
    const void* send; size_t send_bytes;
    void* recv; size_t recv_bytes;
};

struct AzComm {
    int rank = 0, nranks = 1;
    std::string err;
    virtual ~AzComm() {}
    // All sends are posted in the order given; all receives in the order given.  Between one pair of
    // ranks the k-th send matches the k-th receive.  Asynchronous on `st`.
    virtual int exchange(const std::vector<AzP2P>& ops, cudaStream_t st) = 0;
    virtual int allreduce(double* dev, int n, bool take_max, cudaStream_t st) = 0; // in place: sum or max
    // Peer memory (NVLink / NVSwitch inside one box): every rank offers `nbuf` device buffers; on return
    // remote[r * nbuf + b] is a pointer THIS rank's kernels can store through to reach buffer b of rank r (its own buffers
    // for r == rank).  Collective.  AZ_ERR_UNSUPPORTED: the transport has no peer memory (the caller keeps send/recv).
    virtual int share(void* const* local, int nbuf, std::vector<void*>* remote) { (void)local; (void)nbuf; (void)remote; return -4; }
};

// id: 128 bytes shared by all ranks (ncclUniqueId for the NCCL implementation).
int az_comm_create(const void* id, int rank, int nranks, int device, AzComm** out, std::string* err);
int az_comm_make_id(void* id128, std::string* err);
