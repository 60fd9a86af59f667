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


// ---------------------------------------------------------------------------------------------
// Hungarian algorithm (MC64-like), scaling.f90:586-1325.  All arrays 1-based.
// ---------------------------------------------------------------------------------------------
const double RINF = DBL_MAX;      // huge(rinf)

// scaling.f90:1206-1239: the value of idx has decreased, move it towards the root
void heap_update(int idx, std::vector<int>& Q, const std::vector<double>& val, std::vector<int>& L) {
   int pos = L[idx];
   if (pos <= 1) {
      Q[pos] = idx;
      return;
   }
   const double v = val[idx];
   while (pos > 1) {
      const int parent_pos = pos / 2;
      const int parent_idx = Q[parent_pos];
      if (v >= val[parent_idx]) break;
      Q[pos] = parent_idx;
      L[parent_idx] = pos;
      pos = parent_pos;
   }
   Q[pos] = idx;
   L[idx] = pos;
}

// scaling.f90:1268-1325: delete the element in position pos0
void heap_delete(int pos0, int& qlen, std::vector<int>& Q, const std::vector<double>& D, std::vector<int>& L) {
   if (qlen == pos0) {
      qlen = qlen - 1;
      return;
   }
   const int idx = Q[qlen];
   const double v = D[idx];
   qlen = qlen - 1;
   int pos = pos0;
   if (pos > 1) {
      for (;;) {
         const int parent = pos / 2;
         const int qk = Q[parent];
         if (v >= D[qk]) break;
         Q[pos] = qk;
         L[qk] = pos;
         pos = parent;
         if (pos <= 1) break;
      }
   }
   Q[pos] = idx;
   L[idx] = pos;
   if (pos != pos0) return;
   for (;;) {
      int child = 2 * pos;
      if (child > qlen) break;
      double dk = D[Q[child]];
      if (child < qlen) {
         const double dr = D[Q[child + 1]];
         if (dk > dr) {
            child = child + 1;
            dk = dr;
         }
      }
      if (v <= dk) break;
      const int qk = Q[child];
      Q[pos] = qk;
      L[qk] = pos;
      pos = child;
   }
   Q[pos] = idx;
   L[idx] = pos;
}

int heap_pop(int& qlen, std::vector<int>& Q, const std::vector<double>& val, std::vector<int>& L) {
   const int top = Q[1];
   heap_delete(1, qlen, Q, val, L);
   return top;
}

// scaling.f90:810-929
void hungarian_init_heuristic(int m, int n, const std::vector<long>& ptr, const std::vector<int>& row,
                              const std::vector<double>& val, int& num, std::vector<int>& iperm,
                              std::vector<long>& jperm, std::vector<double>& dualu, std::vector<double>& d,
                              std::vector<long>& l, std::vector<long>& search_from) {
   for (int i = 1; i <= m; ++i) { dualu[i] = RINF; l[i] = 0; }
   for (int j = 1; j <= n; ++j)
      for (long k = ptr[j]; k <= ptr[j + 1] - 1; ++k) {
         const int i = row[k];
         if (val[k] > dualu[i]) continue;
         dualu[i] = val[k];
         iperm[i] = j;
         l[i] = k;
      }
   for (int i = 1; i <= m; ++i) {
      const int j = iperm[i];
      if (j == 0) continue;
      iperm[i] = 0;
      if (jperm[j] != 0) continue;
      if ((ptr[j + 1] - ptr[j] > m / 10) && (m > 50)) continue;
      num = num + 1;
      iperm[i] = j;
      jperm[j] = l[i];
   }
   if (num == std::min(m, n)) return;
   for (int j = 1; j <= n; ++j) { d[j] = 0.0; search_from[j] = ptr[j]; }
   for (int j = 1; j <= n; ++j) {
      if (jperm[j] != 0) continue;
      if (ptr[j] > ptr[j + 1] - 1) continue;
      int i0 = row[ptr[j]];
      double vj = val[ptr[j]] - dualu[i0];
      long k0 = ptr[j];
      for (long k = ptr[j] + 1; k <= ptr[j + 1] - 1; ++k) {
         const int i = row[k];
         const double di = val[k] - dualu[i];
         if (di > vj) continue;
         if ((di == vj) && (di != RINF)) {
            if ((iperm[i] != 0) || (iperm[i0] == 0)) continue;
         }
         vj = di;
         i0 = i;
         k0 = k;
      }
      d[j] = vj;
      if (iperm[i0] == 0) {
         num = num + 1;
         jperm[j] = k0;
         iperm[i0] = j;
         search_from[j] = k0 + 1;
         continue;
      }
      bool augmented = false;
      for (long k = k0; k <= ptr[j + 1] - 1 && !augmented; ++k) {
         const int i = row[k];
         if ((val[k] - dualu[i]) > vj) continue;
         const int jj = iperm[i];
         for (long kk = search_from[jj]; kk <= ptr[jj + 1] - 1; ++kk) {
            const int ii = row[kk];
            if (iperm[ii] > 0) continue;
            if ((val[kk] - dualu[ii]) <= d[jj]) {
               jperm[jj] = kk;
               iperm[ii] = jj;
               search_from[jj] = kk + 1;
               num = num + 1;
               jperm[j] = k;
               iperm[i] = j;
               search_from[j] = k + 1;
               augmented = true;
               break;
            }
         }
         if (!augmented) search_from[jj] = ptr[jj + 1];
      }
   }
}

}  // namespace

// scaling.f90:938-1194.  iperm(i) = column matched to row i (negative completion when singular)
void hungarian_match(int m, int n, const std::vector<long>& ptr, const std::vector<int>& row,
                     const std::vector<double>& val, std::vector<int>& iperm, int& num, std::vector<double>& dualu,
                     std::vector<double>& dualv) {
   std::vector<long> jperm(n + 1, 0), out(n + 1, 0), longwork(m + 1, 0);
   std::vector<int> pr(n + 1, 0), q(m + 2, 0), l(m + 1, 0);
   std::vector<double> d(std::max(m, n) + 1, 0.0);
   num = 0;
   for (int i = 1; i <= m; ++i) iperm[i] = 0;
   hungarian_init_heuristic(m, n, ptr, row, val, num, iperm, jperm, dualu, d, longwork, out);
   if (num != std::min(m, n)) {
      for (int i = 1; i <= m; ++i) { d[i] = RINF; l[i] = 0; }
      long isp = -1;
      int jsp = -1;
      for (int jord = 1; jord <= n; ++jord) {
         if (jperm[jord] != 0) continue;
         double dmin = RINF;
         int qlen = 0;
         int low = m + 1;
         int up = m + 1;
         double csp = RINF;
         int j = jord;
         pr[j] = -1;
         for (long klong = ptr[j]; klong <= ptr[j + 1] - 1; ++klong) {
            const int i = row[klong];
            const double dnew = val[klong] - dualu[i];
            if (dnew >= csp) continue;
            if (iperm[i] == 0) {
               csp = dnew;
               isp = klong;
               jsp = j;
            } else {
               if (dnew < dmin) dmin = dnew;
               d[i] = dnew;
               qlen = qlen + 1;
               longwork[qlen] = klong;
            }
         }
         int q0 = qlen;
         qlen = 0;
         for (int kk = 1; kk <= q0; ++kk) {
            const long klong = longwork[kk];
            const int i = row[klong];
            if (csp <= d[i]) {
               d[i] = RINF;
               continue;
            }
            if (d[i] <= dmin) {
               low = low - 1;
               q[low] = i;
               l[i] = low;
            } else {
               qlen = qlen + 1;
               l[i] = qlen;
               heap_update(i, q, d, l);
            }
            const int jj = iperm[i];
            out[jj] = klong;
            pr[jj] = j;
         }
         for (int jdum = 1; jdum <= num; ++jdum) {
            if (low == up) {
               if (qlen == 0) break;
               int i = q[1];
               if (d[i] >= csp) break;
               dmin = d[i];
               while (qlen > 0) {
                  i = q[1];
                  if (d[i] > dmin) break;
                  i = heap_pop(qlen, q, d, l);
                  low = low - 1;
                  q[low] = i;
                  l[i] = low;
               }
            }
            q0 = q[up - 1];
            const double dq0 = d[q0];
            if (dq0 >= csp) break;
            up = up - 1;
            j = iperm[q0];
            const double vj = dq0 - val[jperm[j]] + dualu[q0];
            for (long klong = ptr[j]; klong <= ptr[j + 1] - 1; ++klong) {
               const int i = row[klong];
               if (l[i] >= up) continue;
               const double dnew = vj + val[klong] - dualu[i];
               if (dnew >= csp) continue;
               if (iperm[i] == 0) {
                  csp = dnew;
                  isp = klong;
                  jsp = j;
               } else {
                  const double di = d[i];
                  if (di <= dnew) continue;
                  if (l[i] >= low) continue;
                  d[i] = dnew;
                  if (dnew <= dmin) {
                     const int lpos = l[i];
                     if (lpos != 0) heap_delete(lpos, qlen, q, d, l);
                     low = low - 1;
                     q[low] = i;
                     l[i] = low;
                  } else {
                     if (l[i] == 0) {
                        qlen = qlen + 1;
                        l[i] = qlen;
                     }
                     heap_update(i, q, d, l);
                  }
                  const int jj = iperm[i];
                  out[jj] = klong;
                  pr[jj] = j;
               }
            }
         }
         if (csp != RINF) {
            num = num + 1;
            int i = row[isp];
            iperm[i] = jsp;
            jperm[jsp] = isp;
            j = jsp;
            for (int jdum = 1; jdum <= num; ++jdum) {
               const int jj = pr[j];
               if (jj == -1) break;
               const long klong = out[j];
               i = row[klong];
               iperm[i] = jj;
               jperm[jj] = klong;
               j = jj;
            }
            for (int kk = up; kk <= m; ++kk) {
               i = q[kk];
               dualu[i] = dualu[i] + d[i] - csp;
            }
         }
         for (int kk = low; kk <= m; ++kk) {      // label 190
            const int i = q[kk];
            d[i] = RINF;
            l[i] = 0;
         }
         for (int kk = 1; kk <= qlen; ++kk) {
            const int i = q[kk];
            d[i] = RINF;
            l[i] = 0;
         }
      }
   }
   // label 1000: dual column variables
   for (int j = 1; j <= n; ++j) {
      const long klong = jperm[j];
      if (klong != 0) dualv[j] = val[klong] - dualu[row[klong]];
      else dualv[j] = 0.0;
   }
   for (int i = 1; i <= m; ++i)
      if (iperm[i] == 0) dualu[i] = 0.0;
   if (num == std::min(m, n)) return;
   // structurally singular: complete iperm
   for (int j = 1; j <= n; ++j) jperm[j] = 0;
   int k = 0;
   for (int i = 1; i <= m; ++i) {
      if (iperm[i] == 0) {
         k = k + 1;
         out[k] = i;
      } else {
         jperm[iperm[i]] = i;
      }
   }
   k = 0;
   for (int j = 1; j <= n; ++j) {
      if (jperm[j] != 0) continue;
      k = k + 1;
      const int jdum = (int)out[k];
      iperm[jdum] = -j;
   }
}

namespace {

// match_postproc for a square matrix (scaling.f90:1631-1638)
void match_postproc_square(int n, std::vector<double>& rscaling, std::vector<double>& cscaling) {
   if (n <= 0) return;
   double rsum = 0.0, csum = 0.0;
   for (int i = 1; i <= n; ++i) rsum += rscaling[i];
   for (int i = 1; i <= n; ++i) csum += cscaling[i];
   const double adjust = (rsum / n - csum / n) / 2;
   for (int i = 1; i <= n; ++i) rscaling[i] = rscaling[i] - adjust;
   for (int i = 1; i <= n; ++i) cscaling[i] = cscaling[i] + adjust;
}

}  // namespace

// hungarian_scale_sym (scaling.f90:134-170) -> hungarian_wrapper(sym = .true.) (:596-801)
int hungarian_scale_sym(int n, const long* ptr_in, const int* row_in, const double* val_in, double* scaling,
                        int* match_out, bool scale_if_singular, HungarianInform* inform_out) {
   HungarianInform inform;
   try {
      const int m = n;
      const long ne = 2 * (ptr_in[n] - 1);
      std::vector<long> ptr2(n + 2);
      std::vector<int> row2(ne + 1), match(m + 1, 0);
      std::vector<double> val2(ne + 1), dualu(m + 1, 0.0), dualv(n + 1, 0.0), cmax(n + 1, 0.0);
      std::vector<double> rscaling(m + 1, 0.0), cscaling(n + 1, 0.0);
      long klong = 1;
      for (int i = 1; i <= n; ++i) {
         ptr2[i] = klong;
         for (long j = ptr_in[i - 1]; j <= ptr_in[i] - 1; ++j) {
            if (val_in[j - 1] == 0.0) continue;
            row2[klong] = row_in[j - 1];
            val2[klong] = std::fabs(val_in[j - 1]);
            ++klong;
         }
         for (long j = ptr2[i]; j <= klong - 1; ++j) val2[j] = std::log(val2[j]);
      }
      ptr2[n + 1] = klong;
      half_to_full(n, row2, ptr2, val2);
      for (int i = 1; i <= n; ++i) {
         double colmax = -DBL_MAX;      // maxval of an empty section
         for (long j = ptr2[i]; j <= ptr2[i + 1] - 1; ++j) colmax = std::max(colmax, val2[j]);
         cmax[i] = colmax;
         for (long j = ptr2[i]; j <= ptr2[i + 1] - 1; ++j) val2[j] = colmax - val2[j];
      }
      hungarian_match(m, n, ptr2, row2, val2, match, inform.matched, dualu, dualv);
      if (inform.matched != std::min(m, n)) {
         if (scale_if_singular) inform.flag = 1;      // WARNING_SINGULAR
         else inform.flag = -2;                       // ERROR_SINGULAR (identity scaling, overwritten below as in the reference)
      }
      if (inform.matched == n) {
         for (int i = 1; i <= m; ++i) rscaling[i] = dualu[i];
         for (int i = 1; i <= n; ++i) cscaling[i] = dualv[i] - cmax[i];
         match_postproc_square(n, rscaling, cscaling);
      } else {
         // structurally rank deficient: matching on the full-rank submatrix (Duff and Pralet)
         std::vector<int> old_to_new(n + 1), new_to_old(n + 1), cperm(n + 1, 0);
         int j = inform.matched + 1;
         int k = 0;
         for (int i = 1; i <= m; ++i) {
            if (match[i] < 0) {
               old_to_new[i] = -j;
               j = j + 1;
            } else {
               k = k + 1;
               old_to_new[i] = k;
               new_to_old[k] = i;
            }
         }
         long nent = 0;
         k = 0;
         long j2 = 1;
         ptr2[1] = 1;
         for (int i = 1; i <= n; ++i) {
            const long j1 = j2;
            j2 = ptr2[i + 1];
            if (match[i] < 0) continue;
            k = k + 1;
            for (long jl = j1; jl <= j2 - 1; ++jl) {
               const int jj = row2[jl];
               if (match[jj] < 0) continue;
               nent = nent + 1;
               row2[nent] = old_to_new[jj];
               val2[nent] = val2[jl];
            }
            ptr2[k + 1] = nent + 1;
         }
         const int nn = k;
         hungarian_match(nn, nn, ptr2, row2, val2, cperm, inform.matched, dualu, dualv);
         for (int i = 1; i <= n; ++i) {
            const int jn = old_to_new[i];
            if (jn < 0) rscaling[i] = -DBL_MAX;
            else rscaling[i] = (dualu[jn] + dualv[jn] - cmax[i]) / 2;
         }
         for (int i = 1; i <= n; ++i) match[i] = -1;
         for (int i = 1; i <= nn; ++i) match[new_to_old[i]] = cperm[i];
         for (int i = 1; i <= n; ++i)
            if (match[i] == -1) match[i] = old_to_new[i];
         std::vector<double> cscale(rscaling);
         for (int i = 1; i <= n; ++i)
            for (long jl = ptr_in[i - 1]; jl <= ptr_in[i] - 1; ++jl) {
               const int kr = row_in[jl - 1];
               if (cscale[i] == -DBL_MAX && cscale[kr] != -DBL_MAX)
                  rscaling[i] = std::max(rscaling[i], std::log(std::fabs(val_in[jl - 1])) + rscaling[kr]);
               if (cscale[kr] == -DBL_MAX && cscale[i] != -DBL_MAX)
                  rscaling[kr] = std::max(rscaling[kr], std::log(std::fabs(val_in[jl - 1])) + rscaling[i]);
            }
         for (int i = 1; i <= n; ++i) {
            if (cscale[i] != -DBL_MAX) continue;
            if (rscaling[i] == -DBL_MAX) rscaling[i] = 0.0;
            else rscaling[i] = -rscaling[i];
         }
         for (int i = 1; i <= n; ++i) cscaling[i] = rscaling[i];
      }
      if (match_out)
         for (int i = 1; i <= n; ++i) match_out[i - 1] = match[i];
      for (int i = 1; i <= n; ++i) scaling[i - 1] = std::exp((rscaling[i] + cscaling[i]) / 2);
   } catch (std::bad_alloc&) {
      inform.flag = -1;
   }
   if (inform_out) *inform_out = inform;
   return inform.flag;
}

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
      match_postproc_square(n, rscaling, cscaling);
      // ---- auction_scale_sym: symmetric scaling from the average (scaling.f90:307-308) ----
      for (int i = 1; i <= n; ++i) scaling[i - 1] = std::exp((rscaling[i] + cscaling[i]) / 2);
   } catch (std::bad_alloc&) {
      inform.flag = -1;
   }
   if (inform_out) *inform_out = inform;
   return inform.flag;
}

}  // namespace sylver_b200
