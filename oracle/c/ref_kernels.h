/* ref_kernels — threaded C restatement of the vector phases of the reference's PARSDMM iteration.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (CPU baseline of bench.py): the product never links or calls this.
 * Every routine keeps the reference's structure — one pass per CDS diagonal (CDS_MVp_MT.jl:9-25), separate
 * BLAS-1 style passes in the CG (cg.jl:85-114), sort-based l1 projection (project_l1_Duchi!.jl:33-49) — and is
 * threaded with OpenMP over rows / columns, which is at least what the reference threads with Threads.@threads.
 * The body below is included twice (TF = float, TF = double); symbols carry the suffix _f32 / _f64.
 */
#ifndef SIP_REF_KERNELS_H
#define SIP_REF_KERNELS_H
#include <stdint.h>
#endif
