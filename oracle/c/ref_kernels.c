/* see ref_kernels.h */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include "ref_kernels.h"

int sipref_threads(void) { return omp_get_max_threads(); }
/* launchers such as torchrun export OMP_NUM_THREADS=1: the caller states how many threads the baseline may use */
void sipref_set_threads(int n) { if (n >= 1) omp_set_num_threads(n); }

#define TF float
#define KEY uint32_t
#define SFX(name) sipref_##name##_f32
#include "ref_kernels_body.inc"
#undef TF
#undef KEY
#undef SFX

#define TF double
#define KEY uint64_t
#define SFX(name) sipref_##name##_f64
#include "ref_kernels_body.inc"
