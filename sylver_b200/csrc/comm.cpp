// One-process-per-GPU communicator of libsylver_b200.so: a thin NCCL layer for the
// contribution-block exchange between GPUs (ncclSend/ncclRecv over NVLink 5 / NVSwitch) and
// the small reductions of the distributed solve.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch.distributed process this
// resolves to the NCCL torch already loaded, in a plain C program to the system library, and
// the shared library keeps loading on machines without NCCL (single-GPU use, CPU-only ABI
// tests).  The reference has no distributed layer at all ("Distributed
// communication backend: none"); this replaces StarPU's host-staged PCIe tile transfers.
#include "comm.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>

namespace sylver_b200 {

namespace {
struct NcclApi {
   void* handle = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;
Comm g_comm;

bool load_nccl() {
   if (g_api.handle) return true;
   const char* names[] = {"libnccl.so.2", "libnccl.so"};
   for (const char* nm : names) {
      g_api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (g_api.handle) break;
   }
   if (!g_api.handle) {
      fprintf(stderr, "sylver_b200: cannot load NCCL (%s)\n", dlerror());
      return false;
   }
#define SYM(field, name)                                                      \
   g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(g_api.handle, name)); \
   if (!g_api.field) { fprintf(stderr, "sylver_b200: NCCL symbol %s missing\n", name); return false; }
   SYM(GetUniqueId, "ncclGetUniqueId")
   SYM(CommInitRank, "ncclCommInitRank")
   SYM(CommDestroy, "ncclCommDestroy")
   SYM(Send, "ncclSend")
   SYM(Recv, "ncclRecv")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
   SYM(AllReduce, "ncclAllReduce")
   SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
   return true;
}

int check(ncclResult_t r, const char* what) {
   if (r == ncclSuccess) return 0;
   fprintf(stderr, "sylver_b200: NCCL error in %s: %s\n", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
   return -1;
}
}  // namespace

const Comm& comm() { return g_comm; }

int comm_unique_id(void* out128) {
   if (!load_nccl()) return -1;
   ncclUniqueId id;
   if (check(g_api.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
   memcpy(out128, &id, sizeof(id));
   return 0;
}

int comm_init(int rank, int world, const void* id128) {
   if (g_comm.nccl) comm_finalize();
   if (world <= 1) {
      g_comm = Comm{};
      return 0;
   }
   if (!load_nccl()) return -1;
   ncclUniqueId id;
   memcpy(&id, id128, sizeof(id));
   ncclComm_t c = nullptr;
   if (check(g_api.CommInitRank(&c, world, id, rank), "ncclCommInitRank")) return -1;
   g_comm.rank = rank;
   g_comm.world = world;
   g_comm.nccl = c;
   return 0;
}

// host-only "communicator" for planning tests: no NCCL object, collectives must not be called
void comm_set_virtual(int rank, int world) {
   if (g_comm.nccl) comm_finalize();
   g_comm.rank = rank;
   g_comm.world = world;
   g_comm.nccl = nullptr;
}

void comm_finalize() {
   if (g_comm.nccl && g_api.CommDestroy) g_api.CommDestroy(static_cast<ncclComm_t>(g_comm.nccl));
   g_comm = Comm{};
}

int comm_group_start() { return check(g_api.GroupStart(), "ncclGroupStart"); }
int comm_group_end() { return check(g_api.GroupEnd(), "ncclGroupEnd"); }
int comm_send(const double* buf, size_t count, int peer, cudaStream_t s) {
   return check(g_api.Send(buf, count, ncclDouble, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclSend");
}
int comm_recv(double* buf, size_t count, int peer, cudaStream_t s) {
   return check(g_api.Recv(buf, count, ncclDouble, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclRecv");
}
int comm_allreduce_sum(double* buf, size_t count, cudaStream_t s) {
   return check(g_api.AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), s),
                "ncclAllReduce");
}
int comm_allreduce_max_int(int* buf, size_t count, cudaStream_t s) {
   return check(g_api.AllReduce(buf, buf, count, ncclInt, ncclMax, static_cast<ncclComm_t>(g_comm.nccl), s),
                "ncclAllReduce");
}

}  // namespace sylver_b200
