// See ordering.hpp.
#include "ordering.hpp"

#include <cstdint>
#include <new>
#include <vector>

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

}  // namespace sylver_b200
