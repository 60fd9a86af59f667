// C ABI of libsylver_b200.so: the public SyLVER API and the Fortran->C++ seam,
// re-stated in C++ (the reference implements them in Fortran:
// src/interfaces/C/sylver_ciface.F90:304-826, src/spldlt_analyse_mod.F90:587-890,
// src/spldlt_factorize_mod.F90:473-898,901-1060).  See include/sylver_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sylver_b200.h"
#include "analyse.hpp"
#include "comm.hpp"
#include "clean.hpp"
#include "engine.hpp"
#include "ordering.hpp"
#include "scaling.hpp"

using namespace sylver_b200;

namespace sylver_b200 {
void symbolic_tree_forget(const SymbolicTree* st);
}

namespace {

struct AKeep {
   Symbolic sym;
   SymbolicTree* tree = nullptr;
   sylver_inform_t inform{};
   bool analysed = false;
   // bumped by every spldlt_analyse on this handle: a numeric tree built against an earlier
   // analysis (its SymbolicTree is gone) must be rebuilt, never refactored
   // (the reference reuses akeep/fkeep across problems, sylver_ciface.F90:436-443,601-608)
   unsigned long generation = 0;
   bool check = false;        // analyse(check = true): the matrix below replaces the caller's
   CleanMatrix clean;         // cleaned structure + conversion map (akeep%ptr/row/map/lmap)
   std::vector<double> match_scaling;   // options.ordering = 2: scaling saved for options.scaling = 3
};

struct FKeep {
   NumericTree* tree = nullptr;
   AKeep* akeep = nullptr;
   unsigned long generation = 0;   // akeep->generation the numeric tree was built against
   bool posdef = false;
   std::vector<double> scaling;   // in elimination order (empty if none)
   sylver_inform_t inform{};
};

int g_ngpu = 1;

sylver_inform_t inform_default() {
   sylver_inform_t inf;
   std::memset(&inf, 0, sizeof(inf));
   return inf;
}

// src/sylver_ciface_mod.F90:37-55 (copy_options_f2c)
sylver_options_c options_to_c(const sylver_options_t* o) {
   sylver_options_c c{};
   c.print_level = o->print_level;
   c.action = o->action;
   c.small = o->small;
   c.u = o->u;
   c.multiplier = 1.1;   // undocumented Fortran-side default, not in the C struct
   c.small_subtree_threshold = o->small_subtree_threshold;
   c.nb = o->nb;
   c.pivot_method = o->pivot_method;
   c.failed_pivot_method = o->failed_pivot_method;
   c.cpu_topology = o->cpu_topology;
   return c;
}

// src/sylver_ciface_mod.F90:59-78 (copy_inform_c2f) + flag merge of
// src/sylver_ciface.cxx:52-64.
void fold_stats(const sylver_inform_c& s, sylver_inform_t* inf) {
   if (s.flag < 0)
      inf->flag = std::min(inf->flag, s.flag);
   else if (inf->flag >= 0)
      inf->flag = std::max(inf->flag, s.flag);
   inf->num_delay = s.num_delay;
   inf->num_neg = s.num_neg;
   inf->num_two = s.num_two;
   inf->matrix_rank = inf->matrix_rank - s.num_zero;
   inf->maxfront = std::max(inf->maxfront, s.maxfront);
}

}  // namespace

extern "C" {

const char* sylver_b200_version(void) { return "sylver_b200 0.1 (sm_100a)"; }
int sylver_b200_device_count(void) { return device_count(); }

void sylver_init(int ncpu, int ngpu) {
   (void)ncpu;
   g_ngpu = std::max(1, ngpu);
}
void sylver_finalize(void) {}

void sylver_default_options(sylver_options_t* o) {
   // src/sylver_datatypes_mod.F90:97-198 via src/interfaces/C/sylver_ciface.F90:304-370
   std::memset(o, 0, sizeof(*o));
   o->array_base = 0;
   o->print_level = 0;
   o->unit_diagnostics = 6;
   o->unit_error = 6;
   o->unit_warning = 6;
   o->ordering = 1;
   o->nemin = 32;
   o->prune_tree = true;
   o->min_gpu_work = 5000000000L;
   o->scaling = 0;
   o->pivot_method = 2;
   o->small = 1e-20;
   o->u = 0.01;
   o->small_subtree_threshold = 4000000L;
   o->nb = 256;
   o->cpu_topology = 1;
   o->action = true;
   o->use_gpu = true;
   o->gpu_perf_coeff = 1.0;
   o->failed_pivot_method = 1;
   o->scheduler = 1;
}

void spldlt_analyse(int n, int* order, long const* ptr, int const* row, double const* val, void** akeep_p,
                    bool check, sylver_options_t const* options, sylver_inform_t* inform) {
   (void)val;
   *inform = inform_default();
   AKeep* ak = static_cast<AKeep*>(*akeep_p);
   if (!ak) {
      ak = new (std::nothrow) AKeep();
      if (!ak) { inform->flag = SYLVER_ERROR_ALLOCATION; return; }
      *akeep_p = ak;
   } else {
      if (ak->tree) { symbolic_tree_forget(ak->tree); delete ak->tree; ak->tree = nullptr; }
      ak->sym = Symbolic();
      ak->analysed = false;
      ak->clean = CleanMatrix();
      ak->match_scaling.clear();
   }
   ++ak->generation;
   ak->check = check;
   if (n < 0) { inform->flag = SYLVER_ERROR_A_N_OOR; ak->inform = *inform; return; }
   if (!ptr || !row) { inform->flag = SYLVER_ERROR_PTR_ROW; ak->inform = *inform; return; }
   if (options->ordering < 0 || options->ordering > 2) { inform->flag = SYLVER_ERROR_ORDER; ak->inform = *inform; return; }
   if (options->ordering == 2 && !val) { inform->flag = SYLVER_ERROR_VAL; ak->inform = *inform; return; }
   if (options->ordering != 0 && !metis_available()) {
      // METIS (1) and the matching-based ordering (2) need the static METIS of the CUDA toolkit
      // at build time (csrc/ordering.cpp)
      inform->flag = SYLVER_ERROR_UNIMPLEMENTED;
      ak->inform = *inform;
      return;
   }
   if (options->ordering == 0 && n > 0 && !order) { inform->flag = SYLVER_ERROR_ORDER; ak->inform = *inform; return; }
   if (check) {
      // out-of-range entries dropped, duplicates summed, rows sorted; warnings keep matrix_util's
      // numbering (src/spldlt_analyse_mod.F90:707-739)
      const int mu_flag = clean_cscl_oop_sym_indef(n, ptr, row, ak->clean);
      if (mu_flag < 0) {
         inform->flag = mu_flag == -1 ? SYLVER_ERROR_ALLOCATION
                                      : (mu_flag == -10 ? SYLVER_ERROR_A_ALL_OOR : SYLVER_ERROR_A_PTR);
         ak->inform = *inform;
         return;
      }
      inform->matrix_outrange = ak->clean.noor;
      inform->matrix_dup = ak->clean.ndup;
      ptr = ak->clean.ptr.data();
      row = ak->clean.row.data();
   }
   const int clean_flag = check ? ak->clean.flag : 0;
   std::vector<int> metis_perm;
   int* user_order = order;
   if (options->ordering == 1 && n > 0) {
      // METIS nested dissection on the (cleaned) pattern (src/spldlt_analyse_mod.F90:748-757 ->
      // spral metis_order); the order goes back to the caller if an array was passed
      metis_perm.resize(n);
      std::vector<int> invp(n);
      const int mf = metis_order(n, ptr, row, metis_perm.data(), invp.data());
      if (mf != 0) {
         inform->flag = mf == -1 ? SYLVER_ERROR_ALLOCATION : SYLVER_ERROR_UNKNOWN;
         ak->inform = *inform;
         return;
      }
      order = metis_perm.data();
   }
   int match_flag = 0;
   if (options->ordering == 2 && n > 0) {
      // matching-based ordering on the (cleaned) matrix; its scaling is kept for
      // options.scaling = 3 (src/spldlt_analyse_mod.F90:788-817)
      std::vector<double> vtmp, vclean;
      long nin = ptr[n] - 1;
      if (check) {
         nin = 0;
         for (long i = 0; i < ak->clean.lmap; ++i) nin = std::max(nin, ak->clean.map[i]);
      }
      const double* hval = values_on_host(val, (size_t)nin, vtmp);
      if (!hval) { inform->flag = SYLVER_ERROR_CUDA_UNKNOWN; ak->inform = *inform; return; }
      if (check) {
         vclean.resize(ak->clean.row.size() + 1);
         apply_conversion_map(ak->clean, hval, vclean.data());
         hval = vclean.data();
      }
      metis_perm.resize(n);
      ak->match_scaling.assign(n, 1.0);
      match_flag = match_order_metis(n, ptr, row, hval, metis_perm.data(), ak->match_scaling.data(), nullptr);
      if (match_flag < 0) {
         inform->flag = match_flag == -1 ? SYLVER_ERROR_ALLOCATION : SYLVER_ERROR_UNKNOWN;
         ak->inform = *inform;
         return;
      }
      order = metis_perm.data();
   }
   int flag;
   try {
      flag = analyse(n, ptr, row, order, options->nemin, ak->sym);
   } catch (std::bad_alloc&) {
      flag = ANAL_ERROR_ALLOCATION;
   }
   if (flag < 0) { inform->flag = flag; ak->inform = *inform; return; }
   inform->flag = (flag > 0 || match_flag == 1) ? ANAL_WARNING_ANAL_SINGULAR : clean_flag;      // the singularity warning replaces a cleaning warning (:805-810)
   Symbolic& s = ak->sym;
   if (n > 0) {
      int tflag = 0;
      ak->tree = symbolic_tree_create(n, s.nnodes, s.sptr.data(), s.sparent.data(), s.rptr.data(), s.rlist.data(),
                                      s.nptr.data(), s.nlist.data(), &tflag);
      if (!ak->tree) { inform->flag = tflag ? tflag : SYLVER_ERROR_UNKNOWN; ak->inform = *inform; return; }
      if (user_order)
         for (int i = 0; i < n; ++i) user_order[i] = std::abs(s.order[i]);
   }
   inform->num_factor = s.num_factor;
   inform->num_flops = s.num_flops;
   inform->maxfront = s.maxfront;
   inform->maxdepth = s.maxdepth;
   inform->matrix_rank = s.matrix_rank;
   inform->num_sup = s.nnodes;
   ak->analysed = true;
   ak->inform = *inform;
}

void spldlt_factorize(bool posdef, long const* ptr, int const* row, double const* val, double* scale,
                      void* akeep_v, void** fkeep_p, sylver_options_t const* options, sylver_inform_t* inform) {
   (void)ptr;
   (void)row;
   AKeep* ak = static_cast<AKeep*>(akeep_v);
   if (!ak || !ak->analysed || ak->inform.flag < 0) {
      *inform = inform_default();
      inform->flag = SYLVER_ERROR_CALL_SEQUENCE;
      return;
   }
   *inform = ak->inform;
   FKeep* fk = static_cast<FKeep*>(*fkeep_p);
   if (!fk) {
      fk = new (std::nothrow) FKeep();
      if (!fk) { inform->flag = SYLVER_ERROR_ALLOCATION; return; }
      *fkeep_p = fk;
   }
   const int n = ak->sym.n;
   if (!options->action && n != ak->inform.matrix_rank) { inform->flag = SYLVER_ERROR_SINGULAR; fk->inform = *inform; return; }
   if (ak->sym.nnodes == 0) { inform->flag = SYLVER_SUCCESS; inform->matrix_rank = 0; fk->inform = *inform; return; }
   if (!val) { inform->flag = SYLVER_ERROR_VAL; fk->inform = *inform; return; }
   // analyse(check = true): a clean copy of the values through the conversion map
   // (apply_conversion_map, src/spldlt_factorize_mod.F90:672-679); the structure is akeep's
   std::vector<double> val_clean, val_host_tmp;
   if (ak->check) {
      long nin = 0;
      for (long i = 0; i < ak->clean.lmap; ++i) nin = std::max(nin, ak->clean.map[i]);
      const double* hval = values_on_host(val, (size_t)nin, val_host_tmp);
      if (!hval) { inform->flag = SYLVER_ERROR_CUDA_UNKNOWN; fk->inform = *inform; return; }
      val_clean.resize(ak->clean.row.size() + 1);
      apply_conversion_map(ak->clean, hval, val_clean.data());
      val = val_clean.data();
      ptr = ak->clean.ptr.data();
      row = ak->clean.row.data();
   }
   if (options->scaling == 3 && ak->match_scaling.empty()) {
      // no scaling saved by a matching-based ordering at analyse (spldlt_factorize_mod.F90:797-802)
      inform->flag = SYLVER_ERROR_NO_SAVED_SCALING;
      fk->inform = *inform;
      return;
   }
   // a scaling was computed with the matching-based ordering and is being ignored (:720-724)
   if (!ak->match_scaling.empty() && options->scaling != 3) inform->flag = SYLVER_WARNING_MATCH_ORD_NO_SCALE;
   const bool had_scaling = !fk->scaling.empty();
   fk->scaling.clear();
   if (options->scaling == 3) {
      fk->scaling.resize(n);
      for (int i = 0; i < n; ++i) fk->scaling[i] = ak->match_scaling[ak->sym.invp[i] - 1];
   } else if (options->scaling > 0) {
      // computed here: Hungarian matching (1, spldlt_factorize_mod.F90:738-769), auction matching
      // (2, :771-795) or norm equilibration (>= 4, :804-831); permuted to elimination order,
      // handed back in `scale`
      if (!ptr || !row) { inform->flag = SYLVER_ERROR_PTR_ROW; fk->inform = *inform; return; }
      std::vector<double> tmp, scaling(n);
      const double* hval = values_on_host(val, (size_t)(ptr[n] - 1), tmp);
      if (!hval) { inform->flag = SYLVER_ERROR_CUDA_UNKNOWN; fk->inform = *inform; return; }
      if (options->scaling == 1) {
         // hsoptions%scale_if_singular = options%action; -2: structurally singular and !action
         const int hf = hungarian_scale_sym(n, ptr, row, hval, scaling.data(), nullptr, options->action, nullptr);
         if (hf == -1 || hf == -2) {
            inform->flag = hf == -1 ? SYLVER_ERROR_ALLOCATION : SYLVER_ERROR_SINGULAR;
            fk->inform = *inform;
            return;
         }
      } else if (options->scaling == 2) {
         if (auction_scale_sym(n, ptr, row, hval, scaling.data(), nullptr, nullptr) != 0) {
            inform->flag = SYLVER_ERROR_ALLOCATION;
            fk->inform = *inform;
            return;
         }
      } else {
         equilib_scale_sym(n, ptr, row, hval, scaling.data());
      }
      fk->scaling.resize(n);
      for (int i = 0; i < n; ++i) fk->scaling[i] = scaling[ak->sym.invp[i] - 1];
      if (scale)
         for (int i = 0; i < n; ++i) scale[i] = scaling[i];
   } else if (scale) {
      // user supplied scaling, permuted to elimination order (spldlt_factorize_mod.F90:744-749)
      fk->scaling.resize(n);
      for (int i = 0; i < n; ++i) fk->scaling[i] = scale[ak->sym.invp[i] - 1];
   }
   sylver_options_c copt = options_to_c(options);
   sylver_inform_c stats{};
   const double* sc = fk->scaling.empty() ? nullptr : fk->scaling.data();
   // a numeric tree is built with or without a scaling buffer: rebuild it when that changes
   if (fk->tree && had_scaling != !fk->scaling.empty()) { numeric_tree_destroy(fk->tree); fk->tree = nullptr; }
   if (fk->tree && fk->akeep == ak && fk->generation == ak->generation && fk->posdef == posdef) {
      numeric_tree_refactor(fk->tree, val, sc, &copt, &stats);
   } else {
      if (fk->tree) { numeric_tree_destroy(fk->tree); fk->tree = nullptr; }
      fk->tree = numeric_tree_create(posdef, ak->tree, val, sc, &copt, &stats);
   }
   fk->akeep = ak;
   fk->generation = ak->generation;
   fk->posdef = posdef;
   fold_stats(stats, inform);
   if (inform->flag >= 0 && n != inform->matrix_rank)
      inform->flag = options->action ? SYLVER_WARNING_FACT_SINGULAR : SYLVER_ERROR_SINGULAR;
   fk->inform = *inform;
}

void spldlt_solve(int job, int nrhs, double* x, int ldx, void* akeep_v, void* fkeep_v,
                  sylver_options_t const* options, sylver_inform_t* inform) {
   (void)options;
   AKeep* ak = static_cast<AKeep*>(akeep_v);
   FKeep* fk = static_cast<FKeep*>(fkeep_v);
   if (!ak || !fk || !ak->analysed) {
      *inform = inform_default();
      inform->flag = SYLVER_ERROR_CALL_SEQUENCE;
      return;
   }
   *inform = fk->inform;
   if (fk->inform.flag < 0) { inform->flag = SYLVER_ERROR_CALL_SEQUENCE; return; }
   const int n = ak->sym.n;
   if (n == 0 || ak->sym.nnodes == 0) return;
   // factors of an earlier analysis on this akeep (re-analysed since): not usable
   if (!fk->tree || fk->akeep != ak || fk->generation != ak->generation) { inform->flag = SYLVER_ERROR_CALL_SEQUENCE; return; }
   if (ldx < n || nrhs < 1) { inform->flag = SYLVER_ERROR_X_SIZE; return; }
   if (job < 0 || job > 4) { inform->flag = SYLVER_ERROR_JOB_OOR; return; }
   if (fk->posdef && (job == 2 || job == 4)) { inform->flag = SYLVER_ERROR_JOB_OOR; return; }
   const std::vector<int>& invp = ak->sym.invp;
   const bool sc = !fk->scaling.empty();
   std::vector<double> x2((size_t)n * nrhs);
   for (int r = 0; r < nrhs; ++r)
      for (int i = 0; i < n; ++i) {
         double v = x[(size_t)r * ldx + invp[i] - 1];
         if (sc && (job == 0 || job == 1)) v *= fk->scaling[i];
         x2[(size_t)r * n + i] = v;
      }
   int flag = numeric_tree_solve(fk->tree, job, nrhs, x2.data(), n);
   if (flag < 0) { inform->flag = flag; return; }
   for (int r = 0; r < nrhs; ++r)
      for (int i = 0; i < n; ++i) {
         double v = x2[(size_t)r * n + i];
         if (sc && (job == 0 || job == 3 || job == 4)) v *= fk->scaling[i];
         x[(size_t)r * ldx + invp[i] - 1] = v;
      }
}

void spldlt_free_akeep(void** akeep_p) {
   if (!akeep_p || !*akeep_p) return;
   AKeep* ak = static_cast<AKeep*>(*akeep_p);
   if (ak->tree) { symbolic_tree_forget(ak->tree); delete ak->tree; }
   delete ak;
   *akeep_p = nullptr;
}

void spldlt_free_fkeep(void** fkeep_p) {
   if (!fkeep_p || !*fkeep_p) return;
   FKeep* fk = static_cast<FKeep*>(*fkeep_p);
   if (fk->tree) numeric_tree_destroy(fk->tree);
   delete fk;
   *fkeep_p = nullptr;
}

// ------------------------------- seam ------------------------------------

void* spldlt_create_symbolic_tree(void* akeep, int n, int nnodes, int const* sptr, int const* sparent,
                                  long const* rptr, int const* rlist, long const* nptr, long const* nlist,
                                  int nsubtrees, int const* subtrees, int const* small, int const* contrib_dest,
                                  int const* exec_loc) {
   // the subtree partition (prune_tree) is accepted and ignored: the arrays describe every node
   // and the engine factorizes the whole tree itself (include/sylver_b200.h)
   (void)akeep; (void)nsubtrees; (void)subtrees; (void)small; (void)contrib_dest; (void)exec_loc;
   int flag = 0;
   try {
      return symbolic_tree_create(n, nnodes, sptr, sparent, rptr, rlist, nptr, nlist, &flag);
   } catch (std::bad_alloc&) {
      return nullptr;
   }
}
void spldlt_destroy_symbolic_tree(void* t) {
   SymbolicTree* st = static_cast<SymbolicTree*>(t);
   if (!st) return;
   symbolic_tree_forget(st);
   delete st;
}

void* spldlt_create_numeric_tree_dbl(bool posdef, void* fkeep, void* symbolic_tree, double* aval,
                                     const double* scaling, void** child_contrib, sylver_options_c* options,
                                     sylver_inform_c* stats) {
   (void)fkeep; (void)child_contrib;
   return numeric_tree_create(posdef, static_cast<SymbolicTree*>(symbolic_tree), aval, scaling, options, stats);
}
void* spldlt_create_numeric_tree_posdef_dbl(void* fkeep, void* symbolic_tree, double* aval, const double* scaling,
                                            void** child_contrib, sylver_options_c* options,
                                            sylver_inform_c* stats) {
   return spldlt_create_numeric_tree_dbl(true, fkeep, symbolic_tree, aval, scaling, child_contrib, options, stats);
}
void spldlt_destroy_numeric_tree_dbl(bool posdef, void* tree) {
   (void)posdef;
   numeric_tree_destroy(static_cast<NumericTree*>(tree));
}
void spldlt_destroy_numeric_tree_posdef_dbl(void* tree) { numeric_tree_destroy(static_cast<NumericTree*>(tree)); }

int spldlt_tree_solve_fwd_dbl(bool, void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 1, nrhs, x, ldx);
}
int spldlt_tree_solve_bwd_dbl(bool, void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 3, nrhs, x, ldx);
}
int spldlt_tree_solve_diag_dbl(bool, void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 2, nrhs, x, ldx);
}
int spldlt_tree_solve_diag_bwd_dbl(bool, void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 4, nrhs, x, ldx);
}
int spldlt_tree_solve_fwd_posdef_dbl(void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 1, nrhs, x, ldx);
}
int spldlt_tree_solve_bwd_posdef_dbl(void const* t, int nrhs, double* x, int ldx) {
   return numeric_tree_solve(static_cast<const NumericTree*>(t), 3, nrhs, x, ldx);
}

// ------------------------------ helpers ----------------------------------

int sylver_b200_akeep_view(void* akeep, sylver_b200_symbolic_view* v) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   if (!ak || !ak->analysed) return -1;
   const Symbolic& s = ak->sym;
   v->n = s.n; v->nnodes = s.nnodes;
   v->sptr = s.sptr.data(); v->sparent = s.sparent.data(); v->rptr = s.rptr.data();
   v->rlist = s.rlist.data(); v->nptr = s.nptr.data(); v->nlist = s.nlist.data();
   v->order = s.order.data(); v->invp = s.invp.data();
   v->num_factor = s.num_factor; v->num_flops = s.num_flops;
   return 0;
}

void* sylver_b200_akeep_tree(void* akeep) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   return ak ? ak->tree : nullptr;
}

int sylver_b200_symbolic_tree_cmap(void* symbolic_tree, long const** cptr, int const** cmap) {
   SymbolicTree* st = static_cast<SymbolicTree*>(symbolic_tree);
   if (!st) return -1;
   *cptr = st->ref_cmapoff.data();      // maps of the reference structure (bit-exact output)
   *cmap = st->ref_cmap.data();
   return 0;
}

int sylver_b200_symbolic_tree_view(void* symbolic_tree, int* nnodes, int const** nrow, int const** ncol,
                                   int const** parent, int const** node_map) {
   SymbolicTree* st = static_cast<SymbolicTree*>(symbolic_tree);
   if (!st) return -1;
   if (nnodes) *nnodes = st->nnodes;
   if (nrow) *nrow = st->nrow.data();
   if (ncol) *ncol = st->ncol.data();
   if (parent) *parent = st->parent.data();
   if (node_map) *node_map = st->node_map.data();
   return st->ref_nnodes;
}

void* sylver_b200_fkeep_tree(void* fkeep) {
   FKeep* fk = static_cast<FKeep*>(fkeep);
   return fk ? fk->tree : nullptr;
}

int sylver_b200_numeric_tree_timings(void const* tree, double* out4) {
   if (!tree) return -1;
   numeric_tree_timings(static_cast<const NumericTree*>(tree), out4);
   return 0;
}

int sylver_b200_numeric_tree_split_info(void const* tree, int* out3) {
   if (!tree) return -1;
   numeric_tree_split_info(static_cast<const NumericTree*>(tree), out3);
   return 0;
}

int sylver_b200_numeric_tree_profile(void const* tree, double* out, int cap) {
   if (!tree) return -1;
   return numeric_tree_profile(static_cast<const NumericTree*>(tree), out, cap);
}

int sylver_b200_numeric_tree_profile_levels(void const* tree, double* out, int cap) {
   if (!tree) return -1;
   return numeric_tree_profile_levels(static_cast<const NumericTree*>(tree), out, cap);
}

long sylver_b200_numeric_tree_bytes(void const* tree, long* factor_bytes, long* contrib_bytes) {
   if (!tree) return -1;
   return numeric_tree_bytes(static_cast<const NumericTree*>(tree), factor_bytes, contrib_bytes);
}

void sylver_b200_set_stream(void* cuda_stream, int enable) { set_user_stream(cuda_stream, enable != 0); }

int sylver_b200_numeric_tree_get_front(void const* tree, int node, int* m, int* n, double* l, double* contrib) {
   if (!tree) return -1;
   return numeric_tree_get_front(static_cast<const NumericTree*>(tree), node, m, n, l, contrib);
}

// ------------------------- multi-GPU (one process per GPU) -------------------------
int sylver_b200_comm_unique_id(void* out128) { return comm_unique_id(out128); }
int sylver_b200_comm_init(int rank, int world, void const* id128) {
   if (world < 1 || rank < 0 || rank >= world) return -1;
   return comm_init(rank, world, id128);
}
void sylver_b200_comm_set_virtual(int rank, int world) { comm_set_virtual(rank, world); }
int sylver_b200_comm_init_local(int rank, int world, int fabric_id) { return comm_init_local(rank, world, fabric_id); }
void sylver_b200_comm_finalize(void) { comm_finalize(); }
int sylver_b200_comm_rank(void) { return comm().rank; }
int sylver_b200_comm_world(void) { return comm().world; }

int sylver_b200_partition(void* akeep, int world, int* owner) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   if (!ak || !ak->analysed || !ak->tree) return -1;
   std::vector<int> own;
   partition_tree(*ak->tree, world, own);
   // reported per REFERENCE node (the engine works on the chain-coarsened tree)
   const SymbolicTree& st = *ak->tree;
   for (int i = 0; i < st.ref_nnodes; ++i) owner[i] = own[st.node_map[i]];
   return st.ref_nnodes;
}

int sylver_b200_plan_exchanges(void* akeep, int rank, int world, int cap, int* out) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   if (!ak || !ak->analysed || !ak->tree) return -1;
   std::vector<int> own;
   partition_tree(*ak->tree, world, own);
   std::vector<std::vector<Xfer>> sends, recvs;
   plan_exchanges(*ak->tree, own, rank, sends, recvs);
   int cnt = 0;
   for (int dir = 0; dir < 2; ++dir) {
      const auto& lists = dir == 0 ? sends : recvs;
      for (size_t l = 0; l < lists.size(); ++l)
         for (const Xfer& x : lists[l]) {
            if (4 * cnt + 3 < cap) {
               out[4 * cnt] = (int)l; out[4 * cnt + 1] = ak->tree->ref_top[x.f]; out[4 * cnt + 2] = x.peer;
               out[4 * cnt + 3] = dir;
            }
            ++cnt;
         }
   }
   return cnt;
}

long sylver_b200_plan_levels(void* akeep, int rank, int world, long cap, long* out) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   if (!ak || !ak->analysed || !ak->tree || world < 1 || rank < 0 || rank >= world) return -1;
   return numeric_plan_levels(ak->tree, rank, world, cap, out);
}

int sylver_b200_equilib_scale(int n, long const* ptr, int const* row, double const* val, double* scaling) {
   if (n < 0 || !ptr || !row || !val || !scaling) return -1;
   return equilib_scale_sym(n, ptr, row, val, scaling);
}

int sylver_b200_auction_scale(int n, long const* ptr, int const* row, double const* val, double* scaling,
                              int* match, int* inform4) {
   if (n < 0 || !ptr || !row || !val || !scaling) return -1;
   AuctionInform inf;
   const int flag = auction_scale_sym(n, ptr, row, val, scaling, match, &inf);
   if (inform4) { inform4[0] = inf.flag; inform4[1] = inf.matched; inform4[2] = inf.iterations; inform4[3] = inf.unmatchable; }
   return flag;
}

int sylver_b200_clean_matrix(int n, long const* ptr, int const* row, int cap, long* ptr_out, int* row_out,
                             long* map, long* counts5) {
   if (!ptr || !row || !counts5) return -99;
   CleanMatrix cm;
   const int flag = clean_cscl_oop_sym_indef(n, ptr, row, cm);
   counts5[0] = flag; counts5[1] = cm.noor; counts5[2] = cm.ndup;
   counts5[3] = (long)cm.row.size(); counts5[4] = cm.lmap;
   if (flag < 0) return flag;
   if (ptr_out) std::copy(cm.ptr.begin(), cm.ptr.end(), ptr_out);
   if (row_out && (long)cm.row.size() <= cap) std::copy(cm.row.begin(), cm.row.end(), row_out);
   if (map && cm.lmap <= 2L * cap) std::copy(cm.map.begin(), cm.map.end(), map);
   return flag;
}

int sylver_b200_apply_conversion_map(long ne, long lmap, long const* map, double const* val, double* val_out) {
   if (ne < 0 || lmap < ne || !map || !val || !val_out) return -1;
   CleanMatrix cm;
   cm.row.resize((size_t)ne);
   cm.map.assign(map, map + lmap);
   cm.lmap = lmap;
   apply_conversion_map(cm, val, val_out);
   return 0;
}

int sylver_b200_hungarian_scale(int n, long const* ptr, int const* row, double const* val, double* scaling,
                                int* match, int scale_if_singular, int* inform2) {
   if (n < 0 || !ptr || !row || !val || !scaling) return -1;
   HungarianInform inf;
   const int flag = hungarian_scale_sym(n, ptr, row, val, scaling, match, scale_if_singular != 0, &inf);
   if (inform2) { inform2[0] = inf.flag; inform2[1] = inf.matched; }
   return flag;
}

int sylver_b200_match_order(int n, long const* ptr, int const* row, double const* val, int* order, double* scale,
                            int* pairs) {
   if (n < 0 || !ptr || !row || !val || !order || !scale) return -99;
   return match_order_metis(n, ptr, row, val, order, scale, pairs);
}

int sylver_b200_metis_order(int n, long const* ptr, int const* row, int* order, int* invp) {
   if (n < 0 || !ptr || !row || !order || !invp) return -99;
   if (n == 0) return 0;
   return metis_order(n, ptr, row, order, invp);
}

int sylver_b200_plan_split(void* akeep, int rank, int world, long* out8, int cap, long* pieces) {
   AKeep* ak = static_cast<AKeep*>(akeep);
   if (!ak || !ak->analysed || !ak->tree || world < 1 || rank < 0 || rank >= world) return -1;
   return numeric_plan_split(ak->tree, rank, world, out8, cap, pieces);
}

int sylver_b200_numeric_tree_get_front_indef(void const* tree, int node, int* nelim, double* d, int* perm) {
   if (!tree) return -1;
   return numeric_tree_get_front_indef(static_cast<const NumericTree*>(tree), node, nelim, d, perm);
}

}  // extern "C"
