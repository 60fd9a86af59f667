// Process-wide communicator (one process per GPU).  See comm.cpp.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstddef>

namespace sylver_b200 {

struct Comm {
   int rank = 0;
   int world = 1;
   void* nccl = nullptr;     // ncclComm_t; null for world == 1 and for virtual (planning-only) comms
};

const Comm& comm();
int comm_unique_id(void* out128);
int comm_init(int rank, int world, const void* id128);
void comm_set_virtual(int rank, int world);
void comm_finalize();

int comm_group_start();
int comm_group_end();
int comm_send(const double* buf, size_t count, int peer, cudaStream_t s);
int comm_recv(double* buf, size_t count, int peer, cudaStream_t s);
int comm_allreduce_sum(double* buf, size_t count, cudaStream_t s);
int comm_allreduce_max_int(int* buf, size_t count, cudaStream_t s);

}  // namespace sylver_b200
