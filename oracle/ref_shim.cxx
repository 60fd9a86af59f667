// oracle/ref_shim.cxx -- TEST INFRASTRUCTURE ONLY.
//
// Glue compiled into oracle/_ref/liboracle.so next to the *unmodified* SPRAL
// SSIDS CPU sources under /root/reference/spral/src (see oracle/Makefile).
// It provides
//   (1) the two Fortran-defined externs SSIDS's C++ expects
//       (spral/src/ssids/contrib.h:16-21); they are never reached when the
//       whole tree is a single subtree (ncontrib == 0), and
//   (2) flat C entry points for the dense single-front kernels so that Python
//       (ctypes) can drive ldlt_app_factor / ldlt_tpp_factor / cholesky_factor
//       exactly as spral/src/ssids/cpu/factor.hxx:35-176 does.
// Nothing here is product code; the product (libsylver_b200.so) never links it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>
#include <memory>

#include "ssids/cpu/cpu_iface.hxx"
#include "ssids/cpu/BuddyAllocator.hxx"
#include "ssids/cpu/Workspace.hxx"
#include "ssids/cpu/ThreadStats.hxx"
#include "ssids/cpu/kernels/ldlt_app.hxx"
#include "ssids/cpu/kernels/ldlt_tpp.hxx"
#include "ssids/cpu/kernels/cholesky.hxx"
#include "ssids/cpu/kernels/calc_ld.hxx"
#include "ssids/cpu/kernels/wrappers.hxx"

using namespace spral::ssids::cpu;

extern "C" {

void spral_ssids_contrib_get_data(const void *const, int *const, const double **const,
      int *const, const int **const, int *const, const int **const,
      const double **const, int *const) {
   fprintf(stderr, "oracle: spral_ssids_contrib_get_data called (ncontrib must be 0)\n");
   abort();
}
void spral_ssids_contrib_free_dbl(void *const) {
   fprintf(stderr, "oracle: spral_ssids_contrib_free_dbl called (ncontrib must be 0)\n");
   abort();
}

// Task-parallel timing (liboracle_omp.so only; a plain call in the sequential build): SSIDS's
// C++ creates OpenMP tasks but relies on its Fortran caller for the parallel region
// (spral/src/ssids/fkeep.f90:226-230: "!$omp parallel proc_bind(close) default(shared)" +
// "!$omp single" around the subtree factorization).  Same two directives here.
void *spral_ssids_cpu_create_num_subtree_dbl(bool posdef, void const *symbolic_subtree, const double *aval,
      const double *scaling, void **child_contrib, struct cpu_factor_options const *options, ThreadStats *stats);
void *oracle_create_num_subtree_parallel(bool posdef, void const *symbolic_subtree, const double *aval,
      const double *scaling, struct cpu_factor_options const *options, int *stats_out) {
   void *res = nullptr;
   #pragma omp parallel proc_bind(close) default(shared)
   {
      #pragma omp single
      res = spral_ssids_cpu_create_num_subtree_dbl(posdef, symbolic_subtree, aval, scaling, nullptr, options,
                                                   reinterpret_cast<ThreadStats *>(stats_out));
   }
   return res;
}

// align_lda as the reference was compiled (depends on -march): lets Python size buffers.
long oracle_align_lda(long lda) { return (long) align_lda<double>((size_t) lda); }

// Dense front, indefinite: follows factor_node_indef (spral/src/ssids/cpu/factor.hxx:35-131).
// a: m x n panel, lda = oracle_align_lda(m), followed by nothing (d separate, 2*n).
// contrib: (m-n)^2, ld m-n. stats: int[8] laid out as ThreadStats.
// returns nelim (or negative flag).
int oracle_factor_front_indef(int m, int n, int *perm, double *a, int lda, double *d,
      double *contrib, struct cpu_factor_options const *options, int *stats_out) {
   typedef BuddyAllocator<double, std::allocator<double>> PoolAlloc;
   ThreadStats stats;
   std::vector<Workspace> work;
   work.emplace_back(8*1024*1024);
   PoolAlloc pool((size_t) m * align_lda<double>(m) + 64);
   int nelim = 0;
   if(options->pivot_method != PivotMethod::tpp) {
      nelim = ldlt_app_factor(m, n, perm, a, lda, d, 0.0, contrib, m-n, *options, work, pool);
      if(nelim < 0) return nelim;
   }
   if(nelim < n) {
      int nelim1 = nelim;
      if(options->pivot_method != PivotMethod::tpp) stats.not_first_pass += n - nelim;
      if(m==n || options->pivot_method==PivotMethod::tpp ||
            options->failed_pivot_method==FailedPivotMethod::tpp) {
         double *ld = work[0].get_ptr<double>(2*(m-nelim1));
         nelim += ldlt_tpp_factor(m-nelim1, n-nelim1, &perm[nelim1], &a[nelim1*((size_t)lda+1)], lda,
               &d[2*nelim1], ld, m-nelim1, options->action, options->u, options->small,
               nelim1, &a[nelim1], lda);
         if(m-n>0 && nelim>nelim1) {
            int nelim2 = nelim - nelim1;
            int ldld = align_lda<double>(m-n);
            double *ld2 = work[0].get_ptr<double>((size_t)nelim2*ldld);
            calcLD<OP_N>(m-n, nelim2, &a[(size_t)nelim1*lda+n], lda, &d[2*nelim1], ld2, ldld);
            double rbeta = (nelim1==0) ? 0.0 : 1.0;
            host_gemm<double>(OP_N, OP_T, m-n, m-n, nelim2, -1.0, &a[(size_t)nelim1*lda+n], lda,
                  ld2, ldld, rbeta, contrib, m-n);
         }
         if(options->pivot_method==PivotMethod::tpp) stats.not_first_pass += n - nelim;
         else stats.not_second_pass += n - nelim;
      }
   }
   stats.num_delay += n - nelim;
   if(nelim==0 && m>n) memset(contrib, 0, sizeof(double)*(size_t)(m-n)*(m-n));
   // inertia walk as NumericSubtree does (spral/src/ssids/cpu/NumericSubtree.hxx count of D)
   for(int i=0; i<nelim; ) {
      double a11 = d[2*i], a21 = d[2*i+1];
      if(i+1==nelim || std::isfinite(d[2*i+2])) {
         if(a11 == 0.0) stats.num_zero++;
         if(a11 < 0.0) stats.num_neg++;
         i++;
      } else {
         double a22 = d[2*i+3];
         stats.num_two++;
         double det = a11*a22 - a21*a21;
         double trace = a11 + a22;
         if(det < 0) stats.num_neg++;
         else if(trace < 0) stats.num_neg += 2;
         i += 2;
      }
   }
   memcpy(stats_out, &stats, sizeof(ThreadStats));
   return nelim;
}

// Dense front, posdef: follows factor_node_posdef (factor.hxx:133-162). info=-1 on success.
void oracle_factor_front_posdef(int m, int n, double *a, int lda, double *contrib, int blksz, int *info) {
   cholesky_factor(m, n, a, lda, 0.0, contrib, m-n, blksz, info);
}

int oracle_ldlt_tpp_factor(int m, int n, int *perm, double *a, int lda, double *d,
      double *ld, int ldld, bool action, double u, double small) {
   return ldlt_tpp_factor(m, n, perm, a, lda, d, ld, ldld, action, u, small);
}

void oracle_ldlt_solve(int m, int n, double const *l, int ldl, double const *d, int nrhs, double *x, int ldx) {
   ldlt_app_solve_fwd<double>(m, n, l, ldl, nrhs, x, ldx);
   ldlt_app_solve_diag<double>(n, d, nrhs, x, ldx);
   ldlt_app_solve_bwd<double>(m, n, l, ldl, nrhs, x, ldx);
}
void oracle_chol_solve(int m, int n, double const *l, int ldl, int nrhs, double *x, int ldx) {
   cholesky_solve_fwd(m, n, l, ldl, nrhs, x, ldx);
   cholesky_solve_bwd(m, n, l, ldl, nrhs, x, ldx);
}

int oracle_sizeof_threadstats(void) { return (int) sizeof(ThreadStats); }
int oracle_sizeof_options(void) { return (int) sizeof(cpu_factor_options); }

} // extern "C"
