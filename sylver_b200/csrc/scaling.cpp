// Host-side scalings: see scaling.hpp.  Index arithmetic is kept 1-based (arrays carry a dummy
// element 0) so that every loop reads like the Fortran it restates.
#include "scaling.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <new>
#include <vector>

namespace sylver_b200 {

int equilib_scale_sym(int n, const long* ptr, const int* row, const double* val, double* scaling) {
   // equilib_options defaults (scaling.f90:29-32): max_iterations = 10, tol = 1e-8 (default REAL)
   const int max_iterations = 10;
   const double tol = (double)1e-8f;
   std::vector<double> maxentry(n);
   for (int i = 0; i < n; ++i) scaling[i] = 1.0;
   int itr = 1;
   for (; itr <= max_iterations; ++itr) {
      std::fill(maxentry.begin(), maxentry.end(), 0.0);
      for (int c = 0; c < n; ++c)
         for (long j = ptr[c] - 1; j < ptr[c + 1] - 1; ++j) {
            const int r = row[j] - 1;
            const double v = std::fabs(scaling[r] * val[j] * scaling[c]);
            maxentry[r] = std::max(maxentry[r], v);
            maxentry[c] = std::max(maxentry[c], v);
         }
      double dev = 0.0;
      for (int i = 0; i < n; ++i) {
         if (maxentry[i] > 0) scaling[i] = scaling[i] / std::sqrt(maxentry[i]);
         dev = std::max(dev, std::fabs(1 - maxentry[i]));
      }
      if (dev < tol) break;
   }
   return itr - 1;      // Fortran: the loop variable is max_iterations + 1 when the loop runs out
}

namespace {

// spral/src/matrix_util.f90:3167-3302 (half_to_full_int64 with values, Fortran base): expands
// the lower triangle held in (ptr, row, a) to the full matrix in place.  In column i the
// entries that come from the upper triangle are in increasing row order and precede the
// lower-triangle entries, which keep their input order.  All arrays 1-based.
void half_to_full(int n, std::vector<int>& row, std::vector<long>& ptr, std::vector<double>& a) {
   std::vector<int> iw(n + 1, 0);
   const long oldtau = ptr[n + 1] - 1;
   int ndiag = 0;
   for (int j = 1; j <= n; ++j) {
      const long i1 = ptr[j], i2 = ptr[j + 1] - 1;
      iw[j] += (int)(i2 - i1) + 1;
      for (long ii = i1; ii <= i2; ++ii) {
         const int i = row[ii];
         if (i != j) iw[i] += 1;
         else ++ndiag;
      }
   }
   const long newtau = 2 * oldtau - ndiag;
   long ipkp1 = oldtau + 1;
   long ckp1 = newtau + 1;
   for (int j = n; j >= 1; --j) {
      const long i1 = ptr[j];
      long i2 = ipkp1;
      const int lenk = (int)(i2 - i1);
      long jstart = ckp1;
      ipkp1 = i1;
      i2 = i2 - 1;
      for (long ii = i2; ii >= i1; --ii) {
         --jstart;
         a[jstart] = a[ii];
         row[jstart] = row[ii];
      }
      ptr[j] = jstart;
      ckp1 = ckp1 - iw[j];
      iw[j] = lenk;
   }
   for (int j = n; j >= 1; --j) {
      const long i1 = ptr[j], i2 = ptr[j] + iw[j] - 1;
      for (long ii = i1; ii <= i2; ++ii) {
         const int i = row[ii];
         if (i == j) continue;
         ptr[i] = ptr[i] - 1;
         const long ipos = ptr[i];
         a[ipos] = a[ii];
         row[ipos] = j;
      }
   }
   ptr[n + 1] = newtau + 1;
}

// scaling.f90:1351-1489.  match(j) = i: column j matched to row i.  All arrays 1-based.
void auction_match_core(int m, int n, const std::vector<long>& ptr, const std::vector<int>& row,
                        const std::vector<double>& val, std::vector<int>& match, std::vector<double>& dualu,
                        std::vector<double>& dualv, AuctionInform& inform) {
   // auction_options defaults (scaling.f90:33-38); min_proportion and eps_initial are default REALs
   const int max_iterations = 30000;
   const int max_unchanged[3] = {10, 100, 100};
   const float min_proportion[3] = {0.90f, 0.0f, 0.0f};
   const float eps_initial = 0.01f;
   inform.flag = 0;
   inform.unmatchable = 0;
   std::vector<int> owner(m + 1, 0), next(n + 1);
   const int minmn = std::min(m, n);
   int unmatched = minmn;
   for (int i = 1; i <= n; ++i) match[i] = 0;
   for (int i = 1; i <= m; ++i) dualu[i] = 0;
   int prev = -1, nunchanged = 0;
   int tail = n;
   for (int i = 1; i <= n; ++i) next[i] = i;
   double eps = (double)eps_initial;
   int itr = 1;
   for (; itr <= max_iterations; ++itr) {
      if (unmatched == 0) break;
      if (unmatched != prev) nunchanged = 0;
      prev = unmatched;
      nunchanged = nunchanged + 1;
      const float prop = (float)(minmn - unmatched) / (float)minmn;      // real(minmn-unmatched)/minmn
      if (nunchanged >= max_unchanged[0] && prop >= min_proportion[0]) break;
      if (nunchanged >= max_unchanged[1] && prop >= min_proportion[1]) break;
      if (nunchanged >= max_unchanged[2] && prop >= min_proportion[2]) break;
      eps = std::min(1.0, eps + 1.0 / (n + 1));
      int insert = 0;
      for (int cptr = 1; cptr <= tail; ++cptr) {
         const int col = next[cptr];
         if (match[col] != 0) continue;
         if (ptr[col] == ptr[col + 1]) continue;
         long j = ptr[col];
         int bestr = row[j];
         double bestu = val[j] - dualu[bestr];
         double bestv = -DBL_MAX;
         for (j = ptr[col] + 1; j <= ptr[col + 1] - 1; ++j) {
            const double u = val[j] - dualu[row[j]];
            if (u > bestu) {
               bestv = bestu;
               bestr = row[j];
               bestu = u;
            } else if (u > bestv) {
               bestv = u;
            }
         }
         if (bestv == -DBL_MAX) bestv = 0.0;
         if (bestu > 0) {
            dualu[bestr] = dualu[bestr] + bestu - bestv + eps;
            dualv[col] = bestv - eps;
            match[col] = bestr;
            unmatched = unmatched - 1;
            const int k = owner[bestr];
            owner[bestr] = col;
            if (k != 0) {
               match[k] = 0;
               unmatched = unmatched + 1;
               insert = insert + 1;
               next[insert] = k;
            }
         } else {
            match[col] = -1;
            unmatched = unmatched - 1;
            inform.unmatchable = inform.unmatchable + 1;
         }
      }
      tail = insert;
   }
   inform.iterations = itr - 1;
   for (int i = 1; i <= n; ++i)
      if (match[i] == -1) match[i] = 0;
}

}  // namespace

int auction_scale_sym(int n, const long* ptr_in, const int* row_in, const double* val_in, double* scaling,
                      int* match_out, AuctionInform* inform_out) {
   AuctionInform inform;
   try {
      const int m = n;
      // ---- auction_match (expand = .true.), scaling.f90:1504-1609 ----
      const long ne = 2 * (ptr_in[n] - 1);
      std::vector<long> ptr2(n + 2);
      std::vector<int> row2(ne + 1);
      std::vector<double> val2(ne + 1), cmax(n + 1);
      std::vector<int> cmatch(n + 1);
      std::vector<double> rscaling(m + 1), cscaling(n + 1);
      // expand matrix, drop explicit zeroes and take log absolute values
      long k = 1;
      for (int i = 1; i <= n; ++i) {
         ptr2[i] = k;
         for (long j = ptr_in[i - 1]; j <= ptr_in[i] - 1; ++j) {
            if (val_in[j - 1] == 0.0) continue;
            row2[k] = row_in[j - 1];
            val2[k] = std::fabs(val_in[j - 1]);
            ++k;
         }
         for (long j = ptr2[i]; j <= k - 1; ++j) val2[j] = std::log(val2[j]);
      }
      ptr2[n + 1] = k;
      half_to_full(n, row2, ptr2, val2);
      // column maximums
      for (int i = 1; i <= n; ++i) {
         if (ptr2[i + 1] <= ptr2[i]) { cmax[i] = 0.0; continue; }
         double colmax = val2[ptr2[i]];
         for (long j = ptr2[i] + 1; j <= ptr2[i + 1] - 1; ++j) colmax = std::max(colmax, val2[j]);
         cmax[i] = colmax;
         for (long j = ptr2[i]; j <= ptr2[i + 1] - 1; ++j) val2[j] = colmax - val2[j];
      }
      double maxentry = -DBL_MAX;      // maxval of an empty array is -huge
      for (long j = 1; j <= ptr2[n + 1] - 1; ++j) maxentry = std::max(maxentry, val2[j]);
      // 2*maxentry+1 prefers high cardinality matchings (+1 avoids 0 cols)
      maxentry = 2 * maxentry + 1;
      for (long j = 1; j <= ptr2[n + 1] - 1; ++j) val2[j] = maxentry - val2[j];
      for (int i = 1; i <= n; ++i) cscaling[i] = -cmax[i];
      auction_match_core(m, n, ptr2, row2, val2, cmatch, rscaling, cscaling, inform);
      inform.matched = 0;
      for (int i = 1; i <= n; ++i)
         if (cmatch[i] != 0) ++inform.matched;
      // undo the pre-processing
      for (int i = 1; i <= m; ++i) rscaling[i] = -rscaling[i] + maxentry;
      for (int i = 1; i <= n; ++i) cscaling[i] = -cscaling[i] - cmax[i];
      // row->col matching into col->row
      if (match_out) {
         for (int i = 0; i < m; ++i) match_out[i] = 0;
         for (int i = 1; i <= n; ++i)
            if (cmatch[i] != 0) match_out[cmatch[i] - 1] = i;
      }
      // ---- match_postproc, square case (scaling.f90:1631-1638) ----
      if (n > 0) {
         double rsum = 0.0, csum = 0.0;
         for (int i = 1; i <= m; ++i) rsum += rscaling[i];
         for (int i = 1; i <= n; ++i) csum += cscaling[i];
         const double ravg = rsum / m, cavg = csum / n;
         const double adjust = (ravg - cavg) / 2;
         for (int i = 1; i <= m; ++i) rscaling[i] = rscaling[i] - adjust;
         for (int i = 1; i <= n; ++i) cscaling[i] = cscaling[i] + adjust;
      }
      // ---- auction_scale_sym: symmetric scaling from the average (scaling.f90:307-308) ----
      for (int i = 1; i <= n; ++i) scaling[i - 1] = std::exp((rscaling[i] + cscaling[i]) / 2);
   } catch (std::bad_alloc&) {
      inform.flag = -1;
   }
   if (inform_out) *inform_out = inform;
   return inform.flag;
}

}  // namespace sylver_b200
