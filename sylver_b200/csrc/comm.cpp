// Communicator of libsylver_b200.so: a thin NCCL layer for the contribution-block exchange
// between GPUs (ncclSend/ncclRecv over NVLink 5 / NVSwitch), the panel broadcasts of fronts
// split over a rank group (ncclBroadcast on ncclCommSplit sub-communicators) and the small
// reductions of the level protocol -- plus an in-process "local fabric" that serves the same
// calls between rank THREADS sharing one device (test transport, see comm.hpp).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch.distributed process this
// resolves to the NCCL torch already loaded, in a plain C program to the system library, and
// the shared library keeps loading on machines without NCCL (single-GPU use, CPU-only ABI
// tests).  The reference has no distributed layer at all ("Distributed
// communication backend: none"); this replaces StarPU's host-staged PCIe tile transfers.
#include "comm.hpp"

#include <dlfcn.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace sylver_b200 {

// ===========================================================================
// local fabric (threads of one process, one device)
// ===========================================================================
struct LocalFabric {
   int world = 1;
   int refs = 0;
   std::mutex mu;
   std::condition_variable cv;
   struct Msg {
      const void* ptr;
      size_t bytes;
      cudaEvent_t ready = nullptr, done = nullptr;
      bool consumed = false;
      bool bad = false;
   };
   std::vector<std::deque<Msg*>> box;   // box[src * world + dst], FIFO per ordered pair
   struct Coll {
      int arrived = 0;
      long gen = 0;
      std::vector<const void*> slot;
   };
   std::map<int, Coll> coll;                      // keyed by group id
   std::vector<std::pair<int, int>> groups;       // gid -> (r0, size); gid 0 = world
};

namespace {
struct NcclApi {
   void* handle = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;
   ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_api;
Comm g_comm;                                   // process communicator (NCCL or virtual)
thread_local Comm t_comm;                      // rank thread of a local fabric
thread_local bool t_local = false;

struct SubGroup {
   int r0, size;
   ncclComm_t nccl;      // null when this rank is not a member (or non-NCCL transports)
};
std::vector<SubGroup> g_groups;                // process communicator's groups; [0] = world

std::mutex g_fabric_mu;
std::map<int, LocalFabric*> g_fabrics;

struct PendingOp {
   bool send;
   const void* sbuf;
   void* rbuf;
   size_t bytes;
   int peer;
   cudaStream_t s;
};
thread_local std::vector<PendingOp> t_pending;
thread_local int t_group_depth = 0;

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
   SYM(CommSplit, "ncclCommSplit")
   SYM(Send, "ncclSend")
   SYM(Recv, "ncclRecv")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
   SYM(AllReduce, "ncclAllReduce")
   SYM(Broadcast, "ncclBroadcast")
   SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
   return true;
}

int check(ncclResult_t r, const char* what) {
   if (r == ncclSuccess) return 0;
   fprintf(stderr, "sylver_b200: NCCL error in %s: %s\n", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
   return -1;
}

int cuda_ok(cudaError_t e, const char* what) {
   if (e == cudaSuccess) return 0;
   fprintf(stderr, "sylver_b200: CUDA error in local fabric (%s): %s\n", what, cudaGetErrorName(e));
   return -1;
}

// ---- local fabric: point-to-point ----
int local_flush() {
   LocalFabric* fb = t_comm.fabric;
   const int me = t_comm.rank, W = fb->world;
   int rc = 0;
   std::vector<std::pair<LocalFabric::Msg*, cudaStream_t>> mine;
   for (const PendingOp& op : t_pending) {
      if (!op.send) continue;
      auto* m = new LocalFabric::Msg();
      m->ptr = op.sbuf;
      m->bytes = op.bytes;
      rc |= cuda_ok(cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming), "event");
      rc |= cuda_ok(cudaEventRecord(m->ready, op.s), "record");
      {
         std::lock_guard<std::mutex> lk(fb->mu);
         fb->box[(size_t)me * W + op.peer].push_back(m);
      }
      fb->cv.notify_all();
      mine.emplace_back(m, op.s);
   }
   for (const PendingOp& op : t_pending) {
      if (op.send) continue;
      LocalFabric::Msg* m = nullptr;
      {
         std::unique_lock<std::mutex> lk(fb->mu);
         auto& q = fb->box[(size_t)op.peer * W + me];
         fb->cv.wait(lk, [&] { return !q.empty(); });
         m = q.front();
         q.pop_front();
      }
      bool bad = false;
      if (m->bytes != op.bytes) {
         fprintf(stderr, "sylver_b200: local fabric size mismatch %d <- %d: recv %zu B, send %zu B\n", me, op.peer,
                 op.bytes, m->bytes);
         bad = true;
         rc = -1;
      } else {
         rc |= cuda_ok(cudaStreamWaitEvent(op.s, m->ready, 0), "wait ready");
         if (op.bytes) rc |= cuda_ok(cudaMemcpyAsync(op.rbuf, m->ptr, op.bytes, cudaMemcpyDeviceToDevice, op.s), "copy");
      }
      cudaEvent_t done = nullptr;
      rc |= cuda_ok(cudaEventCreateWithFlags(&done, cudaEventDisableTiming), "event");
      rc |= cuda_ok(cudaEventRecord(done, op.s), "record");
      {
         std::lock_guard<std::mutex> lk(fb->mu);
         m->done = done;
         m->bad = bad;
         m->consumed = true;
      }
      fb->cv.notify_all();
   }
   for (auto& pr : mine) {
      LocalFabric::Msg* m = pr.first;
      {
         std::unique_lock<std::mutex> lk(fb->mu);
         fb->cv.wait(lk, [&] { return m->consumed; });
      }
      if (m->bad) rc = -1;
      rc |= cuda_ok(cudaStreamWaitEvent(pr.second, m->done, 0), "wait done");
      cudaEventDestroy(m->ready);
      cudaEventDestroy(m->done);
      delete m;
   }
   t_pending.clear();
   return rc;
}

int local_p2p(bool send, const void* sbuf, void* rbuf, size_t bytes, int peer, cudaStream_t s) {
   t_pending.push_back(PendingOp{send, sbuf, rbuf, bytes, peer, s});
   if (t_group_depth == 0) return local_flush();
   return 0;
}

// ---- local fabric: host-staged collectives over the members [r0, r0 + size) of group gid ----
void local_barrier(LocalFabric* fb, int gid, int size) {
   std::unique_lock<std::mutex> lk(fb->mu);
   LocalFabric::Coll& c = fb->coll[gid];
   const long g = c.gen;
   if (++c.arrived == size) {
      c.arrived = 0;
      ++c.gen;
      lk.unlock();
      fb->cv.notify_all();
   } else {
      fb->cv.wait(lk, [&] { return c.gen != g; });
   }
}

template <typename T, typename Op>
int local_allreduce(T* buf, size_t count, cudaStream_t s, Op op) {
   LocalFabric* fb = t_comm.fabric;
   const int W = fb->world, me = t_comm.rank;
   std::vector<T> h(count), r(count);
   int rc = cuda_ok(cudaMemcpyAsync(h.data(), buf, count * sizeof(T), cudaMemcpyDeviceToHost, s), "d2h");
   rc |= cuda_ok(cudaStreamSynchronize(s), "sync");
   {
      std::lock_guard<std::mutex> lk(fb->mu);
      LocalFabric::Coll& c = fb->coll[0];
      if ((int)c.slot.size() != W) c.slot.assign(W, nullptr);
      c.slot[me] = h.data();
   }
   local_barrier(fb, 0, W);
   {
      std::vector<const T*> src(W);
      {
         std::lock_guard<std::mutex> lk(fb->mu);
         for (int q = 0; q < W; ++q) src[q] = static_cast<const T*>(fb->coll[0].slot[q]);
      }
      for (size_t i = 0; i < count; ++i) {
         T v = src[0][i];
         for (int q = 1; q < W; ++q) v = op(v, src[q][i]);
         r[i] = v;
      }
   }
   local_barrier(fb, 0, W);
   rc |= cuda_ok(cudaMemcpyAsync(buf, r.data(), count * sizeof(T), cudaMemcpyHostToDevice, s), "h2d");
   rc |= cuda_ok(cudaStreamSynchronize(s), "sync");
   return rc;
}
}  // namespace

const Comm& comm() { return t_local ? t_comm : g_comm; }

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
   g_groups.clear();
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
   g_comm.fabric = nullptr;
   g_groups.push_back(SubGroup{0, world, c});
   return 0;
}

// host-only "communicator" for planning tests: no NCCL object, collectives must not be called
void comm_set_virtual(int rank, int world) {
   if (g_comm.nccl) comm_finalize();
   g_groups.clear();
   g_comm.rank = rank;
   g_comm.world = world;
   g_comm.nccl = nullptr;
   g_comm.fabric = nullptr;
   g_groups.push_back(SubGroup{0, world, nullptr});
}

int comm_init_local(int rank, int world, int fabric_id) {
   if (world < 1 || rank < 0 || rank >= world) return -1;
   std::lock_guard<std::mutex> lk(g_fabric_mu);
   LocalFabric*& fb = g_fabrics[fabric_id];
   if (!fb) {
      fb = new LocalFabric();
      fb->world = world;
      fb->box.resize((size_t)world * world);
      fb->groups.emplace_back(0, world);
   }
   if (fb->world != world) return -1;
   ++fb->refs;
   t_comm.rank = rank;
   t_comm.world = world;
   t_comm.nccl = nullptr;
   t_comm.fabric = fb;
   t_local = true;
   t_pending.clear();
   t_group_depth = 0;
   return 0;
}

void comm_finalize() {
   if (t_local) {
      std::lock_guard<std::mutex> lk(g_fabric_mu);
      LocalFabric* fb = t_comm.fabric;
      if (fb && --fb->refs == 0) {
         for (auto it = g_fabrics.begin(); it != g_fabrics.end(); ++it)
            if (it->second == fb) { g_fabrics.erase(it); break; }
         delete fb;
      }
      t_comm = Comm{};
      t_local = false;
      return;
   }
   for (size_t i = 1; i < g_groups.size(); ++i)
      if (g_groups[i].nccl && g_api.CommDestroy) g_api.CommDestroy(g_groups[i].nccl);
   g_groups.clear();
   if (g_comm.nccl && g_api.CommDestroy) g_api.CommDestroy(static_cast<ncclComm_t>(g_comm.nccl));
   g_comm = Comm{};
}

int comm_group_start() {
   if (comm().fabric) { ++t_group_depth; return 0; }
   return check(g_api.GroupStart(), "ncclGroupStart");
}
int comm_group_end() {
   if (comm().fabric) {
      if (--t_group_depth > 0) return 0;
      t_group_depth = 0;
      return local_flush();
   }
   return check(g_api.GroupEnd(), "ncclGroupEnd");
}
int comm_send(const double* buf, size_t count, int peer, cudaStream_t s) {
   if (comm().fabric) return local_p2p(true, buf, nullptr, count * sizeof(double), peer, s);
   return check(g_api.Send(buf, count, ncclDouble, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclSend");
}
int comm_recv(double* buf, size_t count, int peer, cudaStream_t s) {
   if (comm().fabric) return local_p2p(false, nullptr, buf, count * sizeof(double), peer, s);
   return check(g_api.Recv(buf, count, ncclDouble, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclRecv");
}
int comm_send_int(const int* buf, size_t count, int peer, cudaStream_t s) {
   if (comm().fabric) return local_p2p(true, buf, nullptr, count * sizeof(int), peer, s);
   return check(g_api.Send(buf, count, ncclInt, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclSend");
}
int comm_recv_int(int* buf, size_t count, int peer, cudaStream_t s) {
   if (comm().fabric) return local_p2p(false, nullptr, buf, count * sizeof(int), peer, s);
   return check(g_api.Recv(buf, count, ncclInt, peer, static_cast<ncclComm_t>(g_comm.nccl), s), "ncclRecv");
}
int comm_allreduce_sum(double* buf, size_t count, cudaStream_t s) {
   if (comm().fabric) return local_allreduce(buf, count, s, [](double a, double b) { return a + b; });
   return check(g_api.AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), s),
                "ncclAllReduce");
}
int comm_allreduce_max_int(int* buf, size_t count, cudaStream_t s) {
   if (comm().fabric) return local_allreduce(buf, count, s, [](int a, int b) { return a > b ? a : b; });
   return check(g_api.AllReduce(buf, buf, count, ncclInt, ncclMax, static_cast<ncclComm_t>(g_comm.nccl), s),
                "ncclAllReduce");
}
int comm_allreduce_sum_int(int* buf, size_t count, cudaStream_t s) {
   if (comm().fabric) return local_allreduce(buf, count, s, [](int a, int b) { return a + b; });
   return check(g_api.AllReduce(buf, buf, count, ncclInt, ncclSum, static_cast<ncclComm_t>(g_comm.nccl), s),
                "ncclAllReduce");
}

int comm_subgroup(int r0, int size) {
   const Comm& c = comm();
   if (r0 == 0 && size == c.world) return 0;
   const bool member = c.rank >= r0 && c.rank < r0 + size;
   if (c.fabric) {
      LocalFabric* fb = c.fabric;
      std::lock_guard<std::mutex> lk(fb->mu);
      for (size_t i = 0; i < fb->groups.size(); ++i)
         if (fb->groups[i].first == r0 && fb->groups[i].second == size) return member ? (int)i : -1;
      fb->groups.emplace_back(r0, size);
      return member ? (int)fb->groups.size() - 1 : -1;
   }
   for (size_t i = 0; i < g_groups.size(); ++i)
      if (g_groups[i].r0 == r0 && g_groups[i].size == size) return member ? (int)i : -1;
   ncclComm_t sub = nullptr;
   if (c.nccl) {
      // collective over the world communicator: non-members pass NCCL_SPLIT_NOCOLOR
      if (check(g_api.CommSplit(static_cast<ncclComm_t>(c.nccl), member ? 1 : NCCL_SPLIT_NOCOLOR, c.rank, &sub, nullptr),
                "ncclCommSplit"))
         return -2;
   }
   g_groups.push_back(SubGroup{r0, size, sub});
   return member ? (int)g_groups.size() - 1 : -1;
}

int comm_bcast(double* buf, size_t count, int root, int gid, cudaStream_t s) {
   const Comm& c = comm();
   if (gid < 0) return -1;
   if (c.fabric) {
      const std::pair<int, int> g = [&] {
         std::lock_guard<std::mutex> lk(c.fabric->mu);
         return c.fabric->groups[gid];
      }();
      int rc = comm_group_start();
      if (c.rank == root) {
         for (int q = g.first; q < g.first + g.second; ++q)
            if (q != root) rc |= comm_send(buf, count, q, s);
      } else {
         rc |= comm_recv(buf, count, root, s);
      }
      rc |= comm_group_end();
      return rc;
   }
   const SubGroup& g = g_groups[gid];
   return check(g_api.Broadcast(buf, buf, count, ncclDouble, root - g.r0, g.nccl, s), "ncclBroadcast");
}

}  // namespace sylver_b200
