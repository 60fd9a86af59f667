// See ordering.hpp.
#include "ordering.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <new>
#include <vector>

#include "scaling.hpp"

#ifdef SYLVER_HAVE_METIS
extern "C" {
// METIS 5 API, idx_t = int64_t in the CUDA toolkit's build
int METIS_SetDefaultOptions(int64_t* options);
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options, int64_t* perm,
                 int64_t* iperm);
}
#endif

namespace sylver_b200 {

bool metis_available() {
#ifdef SYLVER_HAVE_METIS
   return true;
#else
   return false;
#endif
}

int metis_order(int n, const long* ptr, const int* row, int* perm, int* invp) {
#ifndef SYLVER_HAVE_METIS
   (void)n; (void)ptr; (void)row; (void)perm; (void)invp;
   return -2;
#else
   if (n < 1) return -99;
   if (n == 1) { perm[0] = 1; invp[0] = 1; return 0; }      // metis5_wrapper.f90:134-137
   try {
      // half_to_full_drop_diag64_32 (metis5_wrapper.f90:210-251), 1-based like the Fortran: the
      // adjacency lists are filled back to front, which fixes their order (METIS's result
      // depends on it)
      const long nin = ptr[n] - 1;
      std::vector<int64_t> ptr2(n + 2, 0), row2(2 * nin + 1, 0);
      for (int j = 1; j <= n; ++j)
         for (long k = ptr[j - 1]; k <= ptr[j] - 1; ++k) {
            const int i = row[k - 1];
            if (j != i) { ptr2[i] += 1; ptr2[j] += 1; }
         }
      for (int j = 2; j <= n; ++j) ptr2[j] = ptr2[j - 1] + ptr2[j];
      ptr2[n + 1] = ptr2[n] + 1;
      for (int j = 1; j <= n; ++j)
         for (long k = ptr[j - 1]; k <= ptr[j] - 1; ++k) {
            const int i = row[k - 1];
            if (j != i) {
               row2[ptr2[i]] = j;
               row2[ptr2[j]] = i;
               ptr2[i] -= 1;
               ptr2[j] -= 1;
            }
         }
      for (int j = 1; j <= n; ++j) ptr2[j] += 1;
      // C numbering for METIS (the reference sets METIS_OPTION_NUMBERING = 1 on the same lists;
      // METIS converts those to C numbering itself before doing anything else)
      std::vector<int64_t> xadj(n + 1), adjncy(ptr2[n + 1] - 1);
      for (int j = 0; j <= n; ++j) xadj[j] = ptr2[j + 1] - 1;
      for (size_t e = 0; e < adjncy.size(); ++e) adjncy[e] = row2[e + 1] - 1;
      int64_t options[40];
      METIS_SetDefaultOptions(options);
      std::vector<int64_t> mperm(n), miperm(n);
      int64_t nv = n;
      // metis_order passes (invp, perm) as METIS's (perm, iperm): iperm[old] = new position
      const int rc = METIS_NodeND(&nv, xadj.data(), adjncy.data(), nullptr, options, mperm.data(), miperm.data());
      if (rc == -2) return -1;      // METIS_ERROR_MEMORY
      if (rc != 1) return -99;      // anything but METIS_OK
      for (int i = 0; i < n; ++i) {
         perm[i] = (int)miperm[i] + 1;
         invp[i] = (int)mperm[i] + 1;
      }
      return 0;
   } catch (std::bad_alloc&) {
      return -1;
   }
#endif
}

namespace {

// mo_match (match_order.f90:495-629): matching + symmetric scaling (log scale) of the full
// matrix (ptr2, row2, val2 = |a|, zeros dropped), 1-based.  perm(i) = row matched to column i
// (-1 when unmatched).  Returns 0 or 1 (WARNING_SINGULAR).
int mo_match(int n, std::vector<long>& ptr2, std::vector<int>& row2, std::vector<double>& val2,
             std::vector<double>& scale, std::vector<int>& perm) {
   std::vector<int> cperm(n + 1, 0);
   std::vector<double> dualu(n + 1, 0.0), dualv(n + 1, 0.0), cmax(n + 1, 0.0);
   for (int i = 1; i <= n; ++i) {
      double colmax = 0.0;      // max(0, maxval(...))
      for (long j = ptr2[i]; j <= ptr2[i + 1] - 1; ++j) colmax = std::max(colmax, val2[j]);
      if (colmax != 0.0) colmax = std::log(colmax);
      cmax[i] = colmax;
   }
   for (int i = 1; i <= n; ++i)
      for (long j = ptr2[i]; j <= ptr2[i + 1] - 1; ++j) val2[j] = cmax[i] - std::log(val2[j]);
   int rank = 0;
   hungarian_match(n, n, ptr2, row2, val2, cperm, rank, dualu, dualv);
   if (rank == n) {
      for (int i = 1; i <= n; ++i) scale[i] = (dualu[i] + dualv[i] - cmax[i]) / 2;
      for (int i = 1; i <= n; ++i) perm[i] = cperm[i];
      return 0;
   }
   // structurally singular: match again on the rows/columns of the first matching
   std::vector<int> old_to_new(n + 1), new_to_old(n + 1);
   int k = 0;
   for (int i = 1; i <= n; ++i) {
      if (cperm[i] < 0) {
         old_to_new[i] = -1;
      } else {
         k = k + 1;
         old_to_new[i] = k;
         new_to_old[k] = i;
      }
   }
   long nne = 0;
   k = 0;
   long j2 = 1;
   ptr2[1] = 1;
   for (int i = 1; i <= n; ++i) {
      const long j1 = j2;
      j2 = ptr2[i + 1];
      if (cperm[i] < 0) continue;
      k = k + 1;
      for (long jl = j1; jl <= j2 - 1; ++jl) {
         const int jj = row2[jl];
         if (cperm[jj] < 0) continue;
         nne = nne + 1;
         row2[nne] = old_to_new[jj];
         val2[nne] = val2[jl];
      }
      ptr2[k + 1] = nne + 1;
   }
   const int nn = k;
   hungarian_match(nn, nn, ptr2, row2, val2, cperm, rank, dualu, dualv);
   for (int i = 1; i <= n; ++i) {
      const int j = old_to_new[i];
      if (j < 0) scale[i] = -DBL_MAX;
      else scale[i] = (dualu[j] + dualv[j] - cmax[i]) / 2;
   }
   for (int i = 1; i <= n; ++i) perm[i] = -1;
   for (int i = 1; i <= nn; ++i) perm[new_to_old[i]] = new_to_old[cperm[i]];
   return 1;
}

}  // namespace

int match_order_metis(int n, const long* ptr, const int* row, const double* val, int* order_out, double* scale_out,
                      int* pairs_out) {
   if (!metis_available()) return -2;
   if (n < 0) return -99;
   if (n == 0) return 0;
   try {
      // ---- expand_matrix (spral/src/ssids/anal.f90:87-138): lower -> full with values, lists
      // filled back to front ----
      const long nz = ptr[n] - 1;
      std::vector<long> aptr(n + 2, 0);
      std::vector<int> arow(2 * nz + 1, 0);
      std::vector<double> aval(2 * nz + 1, 0.0);
      for (int j = 1; j <= n; ++j)
         for (long kk = ptr[j - 1]; kk <= ptr[j] - 1; ++kk) {
            const int i = row[kk - 1];
            aptr[i] += 1;
            if (j == i) continue;
            aptr[j] += 1;
         }
      for (int j = 2; j <= n; ++j) aptr[j] = aptr[j - 1] + aptr[j];
      aptr[n + 1] = aptr[n] + 1;
      for (int j = 1; j <= n; ++j)
         for (long kk = ptr[j - 1]; kk <= ptr[j] - 1; ++kk) {
            const int i = row[kk - 1];
            const double atemp = val[kk - 1];
            const long ipos = aptr[i];
            arow[ipos] = j;
            aval[ipos] = atemp;
            aptr[i] = ipos - 1;
            if (j == i) continue;
            const long jpos = aptr[j];
            arow[jpos] = i;
            aval[jpos] = atemp;
            aptr[j] = jpos - 1;
         }
      for (int j = 1; j <= n; ++j) aptr[j] += 1;
      // ---- match_order_metis (:135-208): drop zeros, absolute values ----
      const long ne = aptr[n + 1] - 1;
      std::vector<long> ptr2(n + 2, 0);
      std::vector<int> row2(ne + 1, 0);
      std::vector<double> val2(ne + 1, 0.0);
      long k = 1;
      for (int i = 1; i <= n; ++i) {
         ptr2[i] = k;
         for (long j = aptr[i]; j <= aptr[i + 1] - 1; ++j) {
            if (aval[j] == 0.0) continue;
            row2[k] = arow[j];
            val2[k] = std::fabs(aval[j]);
            ++k;
         }
      }
      ptr2[n + 1] = k;
      // ---- mo_scale (:406-485): its own copy (zeros dropped again: none left), then mo_match.
      // (The Duff-Pralet correction below the call is dead code in the reference: struct_rank
      // is never changed from n, so unmatched rows keep scale = -huge, i.e. exp(scale) = 0.) ----
      std::vector<long> mptr(ptr2);
      std::vector<int> mrow(row2);
      std::vector<double> mval(val2);
      std::vector<double> scale(n + 1, 0.0);
      std::vector<int> cperm(n + 1, 0);
      const int mflag = mo_match(n, mptr, mrow, mval, scale, cperm);
      // ---- mo_split (:220-396): 1x1 and 2x2 pivots from the cycles of the matching ----
      std::vector<int> iwork(n + 1, 0), old_to_new(n + 1, 0), new_to_old(n + 1, 0);
      for (int i = 1; i <= n; ++i) {
         if (iwork[i] != 0) continue;
         int j = i;
         for (;;) {
            if (cperm[j] == -1) { iwork[j] = -2; break; }          // unmatched
            else if (cperm[j] == i) { iwork[j] = -1; break; }      // end of an odd cycle: 1x1
            const int jj = cperm[j];
            iwork[j] = jj;                                        // 2x2 pivot (j, jj)
            iwork[jj] = j;
            j = cperm[jj];
            if (j == i) break;
         }
      }
      for (int i = 1; i <= n; ++i) cperm[i] = iwork[i];
      int kk = 1;
      for (int i = 1; i <= n; ++i) {
         const int j = cperm[i];
         if (j < i && j > 0) continue;
         old_to_new[i] = kk;
         new_to_old[kk] = i;
         if (j > 0) old_to_new[j] = kk;
         ++kk;
      }
      const int ncomp_matched = kk - 1;
      // compressed pattern (both columns of a pair merged), then its lower triangle
      std::vector<long> ptr3(n + 2, 0);
      std::vector<int> row3(ne + 1, 0);
      std::fill(iwork.begin(), iwork.end(), 0);
      ptr3[1] = 1;
      int ncomp = 1;
      long jj = 1;
      for (int i = 1; i <= n; ++i) {
         const int j = cperm[i];
         if (j < i && j > 0) continue;
         for (long kl = ptr2[i]; kl <= ptr2[i + 1] - 1; ++kl) {
            const int krow = old_to_new[row2[kl]];
            if (iwork[krow] == i) continue;
            if (krow > ncomp_matched) continue;
            row3[jj] = krow;
            ++jj;
            iwork[krow] = i;
         }
         if (j > 0) {
            for (long kl = ptr2[j]; kl <= ptr2[j + 1] - 1; ++kl) {
               const int krow = old_to_new[row2[kl]];
               if (iwork[krow] == i) continue;
               if (krow > ncomp_matched) continue;
               row3[jj] = krow;
               ++jj;
               iwork[krow] = i;
            }
         }
         ptr3[ncomp + 1] = jj;
         ++ncomp;
      }
      ncomp = ncomp - 1;
      ptr3[1] = 1;
      jj = 1;
      long j1 = 1;
      for (int i = 1; i <= ncomp; ++i) {
         const long j2 = ptr3[i + 1];
         for (long kl = j1; kl <= j2 - 1; ++kl) {
            const int krow = row3[kl];
            if (krow < i) continue;
            row3[jj] = krow;
            ++jj;
         }
         ptr3[i + 1] = jj;
         j1 = j2;
      }
      // METIS on the compressed lower triangle (metis_order takes 0-based C arrays of 1-based values)
      std::vector<int> corder(ncomp), cinvp(ncomp);
      {
         std::vector<long> p3(ptr3.begin() + 1, ptr3.begin() + 2 + ncomp);
         std::vector<int> r3(row3.begin() + 1, row3.begin() + 1 + (ptr3[ncomp + 1] - 1));
         if (r3.empty()) r3.push_back(0);
         const int mf = metis_order(ncomp, p3.data(), r3.data(), corder.data(), cinvp.data());
         if (mf != 0) return mf == -1 ? -1 : -99;
      }
      // expand: iwork(position) = compressed variable
      for (int i = 1; i <= ncomp; ++i) iwork[corder[i - 1]] = i;
      std::vector<int> order(n + 1, 0);
      kk = 1;
      for (int i = 1; i <= ncomp; ++i) {
         int j = new_to_old[iwork[i]];
         order[j] = kk;
         ++kk;
         if (cperm[j] > 0) {
            j = cperm[j];
            order[j] = kk;
            ++kk;
         }
      }
      for (int i = 1; i <= n; ++i) {
         order_out[i - 1] = order[i];
         scale_out[i - 1] = std::exp(scale[i]);
         if (pairs_out) pairs_out[i - 1] = cperm[i];
      }
      return mflag;
   } catch (std::bad_alloc&) {
      return -1;
   }
}

}  // namespace sylver_b200