// Mapping of the assembly tree onto the GPUs of one box (host code, deterministic: every
// rank computes the same map from the same symbolic tree, nothing is communicated).
//
// The reference balances pruned subtrees over its workers by flops with a greedy
// "heaviest first onto the least loaded" rule (prune_tree,
// src/spldlt_analyse_mod.F90:1435-1656, weights compute_flops :1744-1764) and keeps the top
// of the tree on the CPU.  Here the whole tree lives on GPUs: a proportional
// (subtree-to-subcube) mapping hands each child subtree a share of its parent's rank group
// in proportion to its flops; subtrees that are too light to deserve a rank of their own are
// packed whole onto the least loaded rank of the group (the reference's greedy rule); the
// fronts above the cut are owned by the least loaded rank of their group.  NVSwitch gives
// uniform bandwidth between any two GPUs, so the mapping optimises load only.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "engine.hpp"

namespace sylver_b200 {

namespace {
struct Mapper {
   const SymbolicTree& st;
   std::vector<double> wsub;      // flops of the subtree rooted at each front
   std::vector<double> wown;      // flops of the front itself
   std::vector<double> load;      // per rank: subtrees and fronts assigned so far
   std::vector<double> expect;    // per rank: share of the multi-rank subtrees still to be mapped
   std::vector<int>& owner;
   std::vector<std::pair<int, std::pair<int, int>>> top;   // (front, rank group) above the cut, top-down

   void assign_subtree(int f, int r) {
      // iterative DFS: the whole subtree goes to rank r
      std::vector<int> stack{f};
      while (!stack.empty()) {
         const int g = stack.back();
         stack.pop_back();
         owner[g] = r;
         for (int ci = st.child_ptr[g]; ci < st.child_ptr[g + 1]; ++ci) stack.push_back(st.child_list[ci]);
      }
      load[r] += wsub[f];
   }

   int least_loaded(int r0, int r1) const {
      int best = r0;
      for (int r = r0 + 1; r < r1; ++r)
         if (load[r] + expect[r] < load[best] + expect[best]) best = r;
      return best;
   }

   // distribute the children of `f` (or the roots when f == nnodes) over ranks [r0, r1)
   void map_children(int f, int r0, int r1) {
      struct Item { int f, r0, r1; };
      std::vector<Item> work{{f, r0, r1}};
      while (!work.empty()) {
         const Item it = work.back();
         work.pop_back();
         const int nr = it.r1 - it.r0;
         // the subtree below this front is being mapped now: its children re-enter `load`
         // (whole subtrees) or `expect` (multi-rank children); the front itself stays expected
         // until the fronts above the cut get their owners
         if (it.f < st.nnodes)
            for (int r = it.r0; r < it.r1; ++r) expect[r] -= (wsub[it.f] - wown[it.f]) / nr;
         std::vector<int> ch(st.child_list.begin() + st.child_ptr[it.f], st.child_list.begin() + st.child_ptr[it.f + 1]);
         if (ch.empty()) continue;
         std::stable_sort(ch.begin(), ch.end(), [&](int a, int b) { return wsub[a] > wsub[b]; });
         double tot = 0;
         for (int c : ch) tot += wsub[c];
         if (nr == 1 || tot <= 0) {
            for (int c : ch) assign_subtree(c, it.r0);
            continue;
         }
         // children heavy enough for ranks of their own ("big"), at most nr of them
         std::vector<int> big;
         for (int c : ch)
            if ((int)big.size() < nr && wsub[c] * nr >= 0.75 * tot) big.push_back(c);
         if (big.empty()) {
            for (int c : ch) assign_subtree(c, least_loaded(it.r0, it.r1));
            continue;
         }
         // ranks per big child: one each, then every further rank goes to the child whose
         // per-rank load (flops / ranks) is currently the largest -- the apportionment that
         // minimises the maximum load (largest remainders can leave a child with 1.4x the
         // average when two fractional parts nearly tie)
         std::vector<int> share(big.size(), 1);
         for (int used = (int)big.size(); used < nr; ++used) {
            size_t best = 0;
            for (size_t i = 1; i < big.size(); ++i)
               if (wsub[big[i]] / share[i] > wsub[big[best]] / share[best]) best = i;
            share[best]++;
         }
         int r = it.r0;
         for (size_t i = 0; i < big.size(); ++i) {
            const int c = big[i];
            if (share[i] == 1) {
               assign_subtree(c, r);
            } else {
               top.push_back({c, {r, r + share[i]}});
               work.push_back({c, r, r + share[i]});
               // work is a LIFO: this child's recursion has not run yet, so its ranks would look
               // idle to the packing of the light children below -- charge its expected share
               for (int q = r; q < r + share[i]; ++q) expect[q] += wsub[c] / share[i];
            }
            r += share[i];
         }
         // the light children are packed now, onto the rank with the least assigned + expected load
         for (int c : ch)
            if (std::find(big.begin(), big.end(), c) == big.end()) assign_subtree(c, least_loaded(it.r0, it.r1));
      }
   }
};
}  // namespace

// owner[f] in [0, world) for every front.  world == 1 maps everything to rank 0.
void partition_tree(const SymbolicTree& st, int world, std::vector<int>& owner, std::vector<int>* grp0,
                    std::vector<int>* grpn) {
   const int N = st.nnodes;
   owner.assign(N, 0);
   if (grp0) grp0->assign(N, 0);
   if (grpn) grpn->assign(N, 1);
   if (world <= 1 || N == 0) return;
   std::vector<double> wown(N), wsub(N);
   for (int f = 0; f < N; ++f) {
      const double mm = st.nrow[f] - st.ncol[f], n = st.ncol[f];
      // sum_{j=1..n} (mm + j)^2
      wown[f] = n * mm * mm + mm * n * (n + 1) + n * (n + 1) * (2 * n + 1) / 6.0;
      wsub[f] = wown[f];
   }
   for (int f = 0; f < N; ++f) {      // children precede parents
      const int p = st.parent[f];
      if (p < N) wsub[p] += wsub[f];
   }
   std::vector<int> own(N, -1);
   Mapper mp{st, wsub, wown, std::vector<double>(world, 0.0), std::vector<double>(world, 0.0), own, {}};
   mp.map_children(N, 0, world);       // children of the virtual root
   // fronts above the cut: bottom-up (reverse of discovery is not level order; sort by index,
   // children have smaller indices), least loaded rank of the group
   std::sort(mp.top.begin(), mp.top.end());
   for (auto& t : mp.top) {
      const int g0 = t.second.first, g1 = t.second.second;
      for (int q = g0; q < g1; ++q) mp.expect[q] -= wown[t.first] / (g1 - g0);
      const int r = mp.least_loaded(g0, g1);
      own[t.first] = r;
      mp.load[r] += wown[t.first];
   }
   for (int f = 0; f < N; ++f) owner[f] = own[f] < 0 ? 0 : own[f];
   // rank group of every front: [owner, owner + 1) below the cut, the proportional-mapping
   // group above it (the candidates for a front split over several GPUs)
   if (grp0 && grpn) {
      for (int f = 0; f < N; ++f) { (*grp0)[f] = owner[f]; (*grpn)[f] = 1; }
      for (auto& t : mp.top) {
         (*grp0)[t.first] = t.second.first;
         (*grpn)[t.first] = t.second.second - t.second.first;
      }
   }
}

}  // namespace sylver_b200
