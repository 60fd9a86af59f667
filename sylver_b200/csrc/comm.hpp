// Communicator of libsylver_b200.so.  See comm.cpp.
//
// Two transports behind one interface:
//   * NCCL (one process per GPU; production): point-to-point ncclSend/ncclRecv for the
//     contribution blocks of cross-GPU tree edges, ncclBroadcast on sub-communicators
//     (ncclCommSplit) for the panels of fronts that are split over a rank group.
//   * local fabric (several ranks as threads of ONE process sharing ONE device): the same
//     calls served by device-to-device copies ordered with CUDA events.  It exists so that the
//     multi-rank schedule (partition, exchanges, delayed-pivot hand-over, distributed fronts)
//     is exercised on a single-GPU box by `pytest -m gpu`.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstddef>

namespace sylver_b200 {

struct LocalFabric;

struct Comm {
   int rank = 0;
   int world = 1;
   void* nccl = nullptr;          // ncclComm_t; null for world == 1, virtual and local comms
   LocalFabric* fabric = nullptr; // non-null: thread-per-rank transport
};

const Comm& comm();
int comm_unique_id(void* out128);
int comm_init(int rank, int world, const void* id128);
void comm_set_virtual(int rank, int world);
// Join (creating it on first use) the in-process fabric `fabric_id` as `rank` of `world`.
// Thread local: every rank thread calls this once, and comm_finalize() before it exits.
int comm_init_local(int rank, int world, int fabric_id);
void comm_finalize();

int comm_group_start();
int comm_group_end();
int comm_send(const double* buf, size_t count, int peer, cudaStream_t s);
int comm_recv(double* buf, size_t count, int peer, cudaStream_t s);
int comm_send_int(const int* buf, size_t count, int peer, cudaStream_t s);
int comm_recv_int(int* buf, size_t count, int peer, cudaStream_t s);
int comm_allreduce_sum(double* buf, size_t count, cudaStream_t s);
int comm_allreduce_max_int(int* buf, size_t count, cudaStream_t s);
int comm_allreduce_sum_int(int* buf, size_t count, cudaStream_t s);

// Rank groups [r0, r0 + size) for fronts that are split over several GPUs.  Every rank of
// the world calls comm_subgroup with the same sequence of (r0, size) (the partition is
// deterministic); the returned handle is valid on members only (-1 elsewhere, 0 = world).
int comm_subgroup(int r0, int size);
// Broadcast `count` doubles from world rank `root` to the members of group `gid`.
int comm_bcast(double* buf, size_t count, int root, int gid, cudaStream_t s);

}  // namespace sylver_b200
