// Symbolic analysis: see analyse.hpp.  All working arrays are 1-based (index 0
// unused) so that every step can be checked against the Fortran it restates.
#include "analyse.hpp"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <stdexcept>

namespace sylver_b200 {

namespace {

typedef std::vector<int> ivec;
typedef std::vector<long> lvec;

// Liu's elimination-tree algorithm with path compression through a virtual
// forest.  spral/src/core_analyse.f90:173-224 (find_etree).
void find_etree(int n, const long* ptr, const int* row, const ivec& perm,
                const ivec& invp, ivec& parent) {
   ivec vforest(n + 2, n + 1);
   for (int piv = 1; piv <= n; ++piv) {
      int col = invp[piv];
      for (long i = ptr[col]; i < ptr[col + 1]; ++i) {
         int j = perm[row[i]];
         if (j >= piv) continue;
         int k = j;
         while (vforest[k] < piv) {
            int l = vforest[k];
            vforest[k] = piv;
            k = l;
         }
         if (vforest[k] == piv) continue;
         parent[k] = piv;
         vforest[k] = piv;
      }
      parent[piv] = n + 1;
   }
}

// Depth-first relabelling of the elimination tree; empty columns are numbered
// last.  spral/src/core_analyse.f90:233-353 (find_postorder).
void find_postorder(int n, int& realn, const long* ptr, ivec& perm, ivec& invp,
                    ivec& parent) {
   realn = n;
   ivec chead(n + 2, -1), cnext(n + 2, -1);
   for (int i = n; i >= 1; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }
   ivec map(n + 2), stack(n + 2);
   int shead = 1;
   stack[shead] = n + 1;
   int id = n + 1;
   while (shead != 0) {
      int node = stack[shead--];
      map[node] = id--;
      if (node == n + 1) {
         for (int i = chead[node]; i != -1; i = cnext[i]) {
            if (ptr[invp[i] + 1] - ptr[invp[i]] == 0) continue;
            stack[++shead] = i;
         }
         for (int i = chead[node]; i != -1; i = cnext[i]) {
            if (ptr[invp[i] + 1] - ptr[invp[i]] != 0) continue;
            --realn;
            stack[++shead] = i;
         }
      } else {
         for (int i = chead[node]; i != -1; i = cnext[i]) stack[++shead] = i;
      }
   }
   for (int i = 1; i <= n; ++i) stack[i] = invp[i];
   for (int i = 1; i <= n; ++i) invp[map[i]] = stack[i];
   for (int i = 1; i <= n; ++i) perm[invp[i]] = i;
   for (int i = 1; i <= n; ++i) stack[i] = map[parent[i]];
   for (int i = 1; i <= n; ++i) parent[map[i]] = stack[i];
}

int forest_find(ivec& vforest, int u) {
   int prev = -1, current = u;
   while (vforest[current] != 0) {
      prev = current;
      current = vforest[current];
      if (vforest[current] != 0) vforest[prev] = vforest[current];
   }
   return current;
}

// Gilbert-Ng-Peyton column counts.  spral/src/core_analyse.f90:387-523.
void find_col_counts(int n, const long* ptr, const int* row, const ivec& perm,
                     const ivec& invp, const ivec& parent, ivec& cc) {
   ivec first(n + 2);
   for (int i = 1; i <= n + 1; ++i) first[i] = i;
   for (int i = 1; i <= n; ++i) {
      int par = parent[i];
      first[par] = std::min(first[i], first[par]);
      cc[i] = (first[i] == i) ? 1 : 0;
   }
   cc[n + 1] = n + 1;
   ivec vforest(n + 2, 0), last_p(n + 2, 0), last_nbr(n + 2, 0);
   for (int piv = 1; piv <= n; ++piv) {
      int col = invp[piv];
      for (long ii = ptr[col]; ii < ptr[col + 1]; ++ii) {
         int u = perm[row[ii]];
         if (u <= piv) continue;
         if (first[piv] > last_nbr[u]) {
            cc[piv] += 1;
            int pp = last_p[u];
            if (pp != 0) {
               int lca = forest_find(vforest, pp);
               cc[lca] -= 1;
            }
            last_p[u] = piv;
         }
         last_nbr[u] = piv;
      }
      int par = parent[piv];
      cc[par] += cc[piv] - 1;
      vforest[piv] = par;
   }
}

// Stable sort of idx[1..n] into decreasing val: insertion sort below 16 items,
// merge sort otherwise -- both stable with ">=" as in
// spral/src/core_analyse.f90:712-800, so any stable descending sort agrees.
void sort_by_val(int n, int* idx, const ivec& val) {
   std::stable_sort(idx, idx + n, [&](int a, int b) { return val[a] > val[b]; });
}

// Relaxed supernode amalgamation.  spral/src/core_analyse.f90:536-705
// (find_supernodes, do_merge:806-819, merge_nodes:824-853).
void find_supernodes(int n, int realn, const ivec& parent, const ivec& cc,
                     ivec& sperm, int& nnodes, ivec& sptr, ivec& sparent,
                     ivec& scc, int nemin) {
   ivec nelim(n + 2, 1), nvert(n + 2, 1), vhead(n + 2, -1), vnext(n + 2, -1);
   ivec stack(n + 2), map(n + 2), npar(n + 2);
   std::vector<char> mark(n + 2, 0);
   lvec ezero(n + 2, 0);
   const long HUGE_L = LONG_MAX;
   ezero[n + 1] = HUGE_L;
   nelim[n + 1] = n + 1 + nemin;

   ivec chead(n + 2, -1), cnext(n + 2, -1), child(n + 2);
   for (int i = realn; i >= 1; --i) {
      int j = parent[i];
      cnext[i] = chead[j];
      chead[j] = i;
   }
   for (int par = 1; par <= n + 1; ++par) {
      int nchild = 0;
      for (int node = chead[par]; node != -1; node = cnext[node]) child[nchild++] = node;
      sort_by_val(nchild, child.data(), cc);
      for (int j = 0; j < nchild; ++j) {
         int node = child[j];
         bool merge = false;
         if (ezero[par] != HUGE_L) {
            merge = ((cc[par] == cc[node] - 1) && (nelim[par] == 1)) ||
                    ((nelim[par] < nemin) && (nelim[node] < nemin));
         }
         if (merge) {
            vnext[node] = vhead[par];
            vhead[par] = node;
            ezero[par] += ezero[node] +
                          ((long)cc[par] - 1 + nelim[par] - cc[node] + 1) * nelim[par];
            nelim[par] += nelim[node];
            nvert[par] += nvert[node];
            mark[node] = 0;
         } else {
            mark[node] = 1;
         }
      }
   }
   int v = 1;
   nnodes = 0;
   for (int node = 1; node <= realn; ++node) {
      if (!mark[node]) continue;
      ++nnodes;
      sptr[nnodes] = v;
      npar[nnodes] = parent[node];
      scc[nnodes] = cc[node] + nelim[node] - 1;
      v += nvert[node];
      int k = v;
      int shead = 1;
      stack[shead] = node;
      while (shead > 0) {
         int i = stack[shead--];
         --k;
         sperm[i] = k;
         map[i] = nnodes;
         if (vnext[i] != -1) stack[++shead] = vnext[i];
         if (vhead[i] != -1) stack[++shead] = vhead[i];
      }
   }
   sptr[nnodes + 1] = v;
   map[n + 1] = nnodes + 1;
   npar[nnodes + 1] = n + 1;
   for (int i = realn + 1; i <= n; ++i) sperm[i] = i;
   for (int node = 1; node <= nnodes; ++node) sparent[node] = map[npar[node]];
}

// spral/src/core_analyse.f90:1069-1100 (apply_perm).
void apply_perm(int n, const ivec& perm, ivec& order, ivec& invp, ivec& cc) {
   for (int i = 1; i <= n; ++i) order[i] = cc[i];
   for (int i = 1; i <= n; ++i) cc[perm[i]] = order[i];
   for (int i = 1; i <= n; ++i) order[i] = invp[i];
   for (int i = 1; i <= n; ++i) invp[perm[i]] = order[i];
   for (int i = 1; i <= n; ++i) order[invp[i]] = i;
}

// spral/src/core_analyse.f90:911-1003 (find_row_lists).
void find_row_lists(int n, const long* ptr, const int* row, const ivec& perm,
                    const ivec& invp, int nnodes, const ivec& sptr,
                    const ivec& sparent, const ivec& scc, lvec& rptr, ivec& rlist) {
   ivec seen(n + 2, 0), chead(nnodes + 2, -1), cnext(nnodes + 2, -1);
   for (int node = nnodes; node >= 1; --node) {
      int i = sparent[node];
      cnext[node] = chead[i];
      chead[i] = node;
   }
   rptr[1] = 1;
   for (int node = 1; node <= nnodes; ++node) {
      rptr[node + 1] = rptr[node] + scc[node];
      long idx = rptr[node];
      for (int piv = sptr[node]; piv < sptr[node + 1]; ++piv) {
         seen[piv] = node;
         rlist[idx++] = piv;
      }
      for (int child = chead[node]; child != -1; child = cnext[child]) {
         for (long i = rptr[child]; i < rptr[child + 1]; ++i) {
            int j = rlist[i];
            if (j < sptr[node]) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
      }
      for (int piv = sptr[node]; piv < sptr[node + 1]; ++piv) {
         int col = invp[piv];
         for (long i = ptr[col]; i < ptr[col + 1]; ++i) {
            int j = perm[row[i]];
            if (j < piv) continue;
            if (seen[j] == node) continue;
            seen[j] = node;
            rlist[idx++] = j;
         }
      }
   }
}

// Double-transpose sort of each node's row list.
// spral/src/core_analyse.f90:1007-1065 (dbl_tr_sort).
void dbl_tr_sort(int n, int nnodes, const lvec& rptr, ivec& rlist) {
   lvec ptr(n + 3, 0);
   for (int node = 1; node <= nnodes; ++node)
      for (long ii = rptr[node]; ii < rptr[node + 1]; ++ii) ptr[rlist[ii] + 2]++;
   ptr[1] = ptr[2] = 1;
   for (int i = 1; i <= n; ++i) ptr[i + 2] += ptr[i + 1];
   long tot = ptr[n + 2] - 1;
   ivec col(tot + 1);
   for (int node = 1; node <= nnodes; ++node)
      for (long ii = rptr[node]; ii < rptr[node + 1]; ++ii) {
         int j = rlist[ii];
         col[ptr[j + 1]++] = node;
      }
   lvec nptr(nnodes + 1);
   for (int node = 1; node <= nnodes; ++node) nptr[node] = rptr[node];
   for (int i = 1; i <= n; ++i)
      for (long jj = ptr[i]; jj < ptr[i + 1]; ++jj) {
         int node = col[jj];
         rlist[nptr[node]++] = i;
      }
}

// SyLVER's A -> L scatter map.  src/spldlt_analyse_mod.F90:130-232 (build_map):
// for each node, first the entries reached through the transpose of the lower
// triangle, then the lower-triangle entries themselves.
void build_map(int n, const long* ptr, const int* row, const ivec& perm,
               const ivec& invp, int nnodes, const ivec& sptr, const lvec& rptr,
               const ivec& rlist, lvec& nptr, lvec& nlist) {
   long nz = ptr[n + 1] - 1;
   ivec map(n + 2, 0), row2(nz + 1);
   lvec ptr2(n + 4, 0), origin(nz + 1);
   for (int i = 1; i <= n; ++i)
      for (long jj = ptr[i]; jj < ptr[i + 1]; ++jj) {
         int k = row[jj];
         if (k == i) continue;
         ptr2[k + 2]++;
      }
   ptr2[1] = ptr2[2] = 1;
   for (int i = 1; i <= n; ++i) ptr2[i + 2] += ptr2[i + 1];
   for (int i = 1; i <= n; ++i)
      for (long jj = ptr[i]; jj < ptr[i + 1]; ++jj) {
         int k = row[jj];
         if (k == i) continue;
         row2[ptr2[k + 1]] = i;
         origin[ptr2[k + 1]] = jj;
         ptr2[k + 1]++;
      }
   long pp = 1;
   for (int node = 1; node <= nnodes; ++node) {
      long blkm = rptr[node + 1] - rptr[node];
      nptr[node] = pp;
      for (long jj = rptr[node]; jj < rptr[node + 1]; ++jj)
         map[rlist[jj]] = (int)(jj - rptr[node] + 1);
      for (int j = sptr[node]; j < sptr[node + 1]; ++j) {
         int col = invp[j];
         for (long i = ptr2[col]; i < ptr2[col + 1]; ++i) {
            int k = std::abs(perm[row2[i]]);
            if (k < j) continue;
            nlist[2 * (pp - 1) + 1] = (long)(j - sptr[node]) * blkm + map[k];
            nlist[2 * (pp - 1) + 0] = origin[i];
            ++pp;
         }
      }
      for (int j = sptr[node]; j < sptr[node + 1]; ++j) {
         int col = invp[j];
         for (long ii = ptr[col]; ii < ptr[col + 1]; ++ii) {
            int k = std::abs(perm[row[ii]]);
            if (k < j) continue;
            nlist[2 * (pp - 1) + 1] = (long)(j - sptr[node]) * blkm + map[k];
            nlist[2 * (pp - 1) + 0] = ii;
            ++pp;
         }
      }
   }
   nptr[nnodes + 1] = pp;
}

}  // namespace

void expand_pattern(int n, long nz, const long* ptr0, const int* row0,
                    std::vector<long>& aptr, std::vector<int>& arow) {
   // ptr0/row0 are 0-based C arrays holding 1-based values.
   const long* ptr = ptr0 - 1;
   const int* row = row0 - 1;
   aptr.assign(n + 2, 0);
   arow.assign(2 * nz + 1, 0);
   for (int j = 1; j <= n; ++j)
      for (long kk = ptr[j]; kk < ptr[j + 1]; ++kk) {
         int i = row[kk];
         aptr[i]++;
         if (j == i) continue;
         aptr[j]++;
      }
   for (int j = 2; j <= n; ++j) aptr[j] += aptr[j - 1];
   aptr[n + 1] = aptr[n] + 1;
   for (int j = 1; j <= n; ++j)
      for (long kk = ptr[j]; kk < ptr[j + 1]; ++kk) {
         int i = row[kk];
         arow[aptr[i]--] = j;
         if (j == i) continue;
         arow[aptr[j]--] = i;
      }
   for (int j = 1; j <= n; ++j) aptr[j]++;
}

int analyse(int n, const long* ptr0, const int* row0, const int* user_order,
            int nemin, Symbolic& sym) {
   if (n < 0) return ANAL_ERROR_A_N_OOR;
   sym = Symbolic();
   sym.n = n;
   if (n == 0) {
      sym.sptr.assign(1, 1);
      sym.rptr.assign(1, 1);
      sym.nptr.assign(1, 1);
      return ANAL_SUCCESS;
   }
   if (nemin < 1) nemin = 32;  // sylver_nemin_default, src/sylver_datatypes_mod.F90:9
   if (!user_order) return ANAL_ERROR_ORDER;
   int flag = ANAL_SUCCESS;
   long nz = ptr0[n] - 1;

   // check_order (spral/src/ssids/anal.f90:147-199)
   ivec perm(n + 2), invp(n + 2, 0);
   for (int i = 1; i <= n; ++i) {
      int j = std::abs(user_order[i - 1]);
      if (j <= 0 || j > n || invp[j] != 0) return ANAL_ERROR_ORDER;
      perm[i] = j;
      invp[j] = i;
   }

   lvec aptr;
   ivec arow;
   expand_pattern(n, nz, ptr0, row0, aptr, arow);
   const long* ptr2 = aptr.data();
   const int* row2 = arow.data();

   // basic_analyse (spral/src/core_analyse.f90:38-150)
   ivec parent(n + 2);
   find_etree(n, ptr2, row2, perm, invp, parent);
   int realn;
   find_postorder(n, realn, ptr2, perm, invp, parent);
   if (realn != n) flag = ANAL_WARNING_ANAL_SINGULAR;
   ivec cc(n + 2);
   find_col_counts(n, ptr2, row2, perm, invp, parent, cc);
   ivec tperm(n + 2), sptr(n + 2), sparent(n + 2), scc(n + 2);
   int nnodes;
   find_supernodes(n, realn, parent, cc, tperm, nnodes, sptr, sparent, scc, nemin);
   apply_perm(n, tperm, perm, invp, cc);
   long nrl = 0;
   for (int i = 1; i <= nnodes; ++i) nrl += scc[i];
   lvec rptr(nnodes + 2);
   ivec rlist(nrl + 2);
   find_row_lists(n, ptr2, row2, perm, invp, nnodes, sptr, sparent, scc, rptr, rlist);
   // calc_stats (spral/src/core_analyse.f90:862-905)
   long nfact = 0, nflops = 0;
   for (int node = 1; node <= nnodes; ++node) {
      long ne = sptr[node + 1] - sptr[node];
      long m = scc[node] - ne;
      nfact += (ne * (ne + 1)) / 2 + ne * m;
      for (long j = 1; j <= ne; ++j) nflops += (m + j) * (m + j);
   }
   dbl_tr_sort(n, nnodes, rptr, rlist);

   // analyse_core (src/spldlt_analyse_mod.F90:307-580): invp from order, unused
   // variables flagged with order 0, then the A->L map on the lower triangle.
   for (int i = 1; i <= n; ++i) invp[perm[i]] = i;
   for (int j = sptr[nnodes + 1]; j <= n; ++j) perm[invp[j]] = 0;
   lvec nptr(nnodes + 2), nlist(2 * nz + 2);
   build_map(n, ptr0 - 1, row0 - 1, perm, invp, nnodes, sptr, rptr, rlist, nptr, nlist);

   sym.nnodes = nnodes;
   sym.sptr.assign(sptr.begin() + 1, sptr.begin() + nnodes + 2);
   sym.sparent.assign(sparent.begin() + 1, sparent.begin() + nnodes + 1);
   sym.rptr.assign(rptr.begin() + 1, rptr.begin() + nnodes + 2);
   sym.rlist.assign(rlist.begin() + 1, rlist.begin() + 1 + nrl);
   sym.nptr.assign(nptr.begin() + 1, nptr.begin() + nnodes + 2);
   long nmap = nptr[nnodes + 1] - 1;
   sym.nlist.assign(nlist.begin(), nlist.begin() + 2 * nmap);
   sym.order.assign(perm.begin() + 1, perm.begin() + n + 1);
   sym.invp.assign(invp.begin() + 1, invp.begin() + n + 1);
   sym.num_factor = nfact;
   sym.num_flops = nflops;
   sym.realn = realn;
   // inform fields (src/spldlt_analyse_mod.F90:548-560)
   ivec level(nnodes + 2, 0);
   for (int i = nnodes; i >= 1; --i) {
      int blkn = sptr[i + 1] - sptr[i];
      level[i] = level[sparent[i]] + 1;
      sym.maxfront = std::max(sym.maxfront, blkn);
      sym.maxdepth = std::max(sym.maxdepth, level[i]);
   }
   sym.matrix_rank = sptr[nnodes + 1] - 1;
   return flag;
}

}  // namespace sylver_b200
