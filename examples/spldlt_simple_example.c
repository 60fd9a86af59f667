/* The reference's C example (examples/C/spldlt_simple_example_c.c) against libsylver_b200.so.
 * Same calls, same matrix, same checked analyse; the one change is that the elimination order
 * is supplied (options.ordering = 0) because METIS is not part of this library.
 *
 *   gcc -std=c11 -I include examples/spldlt_simple_example.c -L sylver_b200 -lsylver_b200 \
 *       -Wl,-rpath,$PWD/sylver_b200 -o example && ./example        (needs a GPU)
 *
 * Expected output: the solution (1.5, 2, 1.5) of
 *     [  2 -1  0 ]       [ 1 ]
 *     [ -1  2 -1 ] x  =  [ 1 ]
 *     [  0 -1  2 ]       [ 1 ]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sylver_b200.h"

int main(void) {
   void *akeep = NULL, *fkeep = NULL;
   sylver_inform_t inform;
   sylver_options_t options;
   const int n = 3, nnz = 5, nrhs = 1;
   long ptr[4] = {1, 3, 5, 6};                      /* 1-based, lower triangle */
   int row[5] = {1, 2, 2, 3, 3};
   double val[5] = {2.0, -1.0, 2.0, -1.0, 2.0};
   int order[3] = {1, 2, 3};
   double x[3] = {1.0, 1.0, 1.0};
   (void)nnz;

   sylver_default_options(&options);
   options.ordering = 0;                            /* order supplied by the caller */
   options.scaling = 0;
   sylver_init(1, 1);

   spldlt_analyse(n, order, ptr, row, val, &akeep, true, &options, &inform);
   if (inform.flag < 0) { printf("analyse failed: %d\n", inform.flag); return 1; }
   spldlt_factorize(false, NULL, NULL, val, NULL, akeep, &fkeep, &options, &inform);
   if (inform.flag < 0) { printf("factorize failed: %d\n", inform.flag); return 1; }
   spldlt_solve(0, nrhs, x, n, akeep, fkeep, &options, &inform);
   if (inform.flag < 0) { printf("solve failed: %d\n", inform.flag); return 1; }
   printf("x = %.15g %.15g %.15g  (num_neg %d, num_delay %d)\n", x[0], x[1], x[2], inform.num_neg, inform.num_delay);

   spldlt_free_fkeep(&fkeep);
   spldlt_free_akeep(&akeep);
   sylver_finalize();
   return 0;
}
