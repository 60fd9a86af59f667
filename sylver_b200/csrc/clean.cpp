// See clean.hpp.  1-based index arithmetic (arrays carry a dummy element 0).
#include "clean.hpp"

#include <algorithm>
#include <new>

namespace sylver_b200 {

namespace {

// pushdown64 with a map (matrix_util.f90, "map only, not val" branch): everything below `root`
// is a max-heap on array; sift array(root) down
void pushdown(long root, long last, int* array, long* map) {
   const int root_idx = array[root];
   const long root_map = map[root];
   long insert = root;
   long test = 2 * insert;
   while (test <= last) {
      if (test != last) {
         if (array[test + 1] > array[test]) test = test + 1;
      }
      if (array[test] <= root_idx) break;
      array[insert] = array[test];
      map[insert] = map[test];
      insert = test;
      test = 2 * insert;
   }
   array[insert] = root_idx;
   map[insert] = root_map;
}

// sort64 (matrix_util.f90:2729-2777): heap sort of array(1:n), same permutation applied to map
void heap_sort(int* array, long n, long* map) {
   if (n <= 1) return;
   for (long root = n / 2; root >= 1; --root) pushdown(root, n, array, map);
   for (long i = n; i >= 2; --i) {
      const int t = array[1]; array[1] = array[i]; array[i] = t;
      const long lt = map[1]; map[1] = map[i]; map[i] = lt;
      pushdown(1, i - 1, array, map);
   }
}

}  // namespace

int clean_cscl_oop_sym_indef(int n, const long* ptr_in0, const int* row_in0, CleanMatrix& out) {
   out = CleanMatrix();
   try {
      if (n < 0) { out.flag = -3; return out.flag; }                     // ERROR_N_OOR
      const long* ptr_in = ptr_in0 - 1;      // 1-based views
      const int* row_in = row_in0 - 1;
      if (ptr_in[1] < 1) { out.flag = -5; return out.flag; }             // ERROR_PTR_1
      const int m = n;
      out.ptr.assign(n + 2, 0);
      const long nin = ptr_in[n + 1] - 1;
      std::vector<int> row_out(std::max<long>(nin, 0) + 1);
      std::vector<long> map(2 * std::max<long>(nin, 0) + 2);
      struct Dup { long src, dest; };
      std::vector<Dup> dups;      // the reference's linked list, newest first: read back to front
      int idup = 0, ioor = 0, idiag = 0;
      long k = 1;
      for (int col = 1; col <= n; ++col) {
         out.ptr[col] = k;
         if (ptr_in[col + 1] < ptr_in[col]) { out.flag = -6; return out.flag; }      // ERROR_PTR_MONO
         const int minidx = col;      // symmetric: lower triangle only
         for (long i = ptr_in[col]; i <= ptr_in[col + 1] - 1; ++i) {
            const int j = row_in[i];
            if (j < minidx || j > m) { ++ioor; continue; }
            row_out[k] = j;
            map[k] = i;
            ++k;
         }
         long cnt = k - out.ptr[col];
         if (cnt == 0 && ptr_in[col + 1] - ptr_in[col] != 0) { out.flag = -10; return out.flag; }      // ERROR_ALL_OOR
         if (cnt != 0) {
            heap_sort(row_out.data() + out.ptr[col] - 1, cnt, map.data() + out.ptr[col] - 1);
            const long last = k - 1;
            k = out.ptr[col] + 1;
            if (row_out[out.ptr[col]] == col) ++idiag;
            for (long i = out.ptr[col] + 1; i <= last; ++i) {
               if (row_out[i] == row_out[i - 1]) {
                  ++idup;
                  dups.push_back(Dup{map[i], k - 1});
                  continue;
               }
               if (row_out[i] == col) ++idiag;
               row_out[k] = row_out[i];
               map[k] = map[i];
               ++k;
            }
         }
      }
      out.ptr[n + 1] = k;
      long lmap = k - 1;
      for (size_t d = dups.size(); d-- > 0;) {      // head of the list = last one found
         ++idup;                                     // (counted again, as in the reference)
         map[lmap + 1] = dups[d].dest;
         map[lmap + 2] = dups[d].src;
         lmap += 2;
      }
      // warnings (matrix_util.f90:1372-1387)
      if (ioor > 0 || idup > 0 || idiag < n) {
         if (ioor > 0) out.flag = 1;                       // WARNING_IDX_OOR
         if (idup > 0) out.flag = 2;                       // WARNING_DUP_IDX
         if (idup > 0 && ioor > 0) out.flag = 3;           // WARNING_DUP_AND_OOR
         if (idiag < n && ioor > 0) out.flag = 5;          // WARNING_MISS_DIAG_OORDUP
         else if (idiag < n && idup > 0) out.flag = 5;
         else if (idiag < n) out.flag = 4;                 // WARNING_MISSING_DIAGONAL
      }
      out.noor = ioor;
      out.ndup = idup;
      out.lmap = lmap;
      const long ne = k - 1;
      out.row.assign(row_out.begin() + 1, row_out.begin() + 1 + ne);
      out.map.assign(map.begin() + 1, map.begin() + 1 + lmap);
      out.ptr.erase(out.ptr.begin());      // back to a plain n + 1 array of 1-based values
   } catch (std::bad_alloc&) {
      out.flag = -1;
   }
   return out.flag;
}

void apply_conversion_map(const CleanMatrix& cm, const double* val, double* val_out) {
   const long ne = (long)cm.row.size();
   for (long i = 0; i < ne; ++i) val_out[i] = val[cm.map[i] - 1];
   for (long i = ne; i + 1 < cm.lmap; i += 2) {
      const long j = cm.map[i], k = cm.map[i + 1];
      val_out[j - 1] = val_out[j - 1] + val[k - 1];
   }
}

}  // namespace sylver_b200
