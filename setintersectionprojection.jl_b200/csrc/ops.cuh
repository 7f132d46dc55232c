// ops.cuh — matrix-free transform-domain operators (forward rows / adjoint columns).
//
// The reference materialises D_x, D_y, D_z, TV, D_xz and the identity as SparseMatrixCSC via
// Kronecker products (get_discrete_Grad.jl:16-37,51-76; get_TD_operator.jl:12-95) and applies them
// with SparseArrays mul! (update_y_l.jl:43, rhs_compose.jl:28).  Here an operator is a small POD
// descriptor and both A*x and A'*v are evaluated on the fly with exactly the reference's term
// order:   (A*x)[r]  = left fold over the stored columns of row r in ascending column index,
//          (A'*v)[c] = left fold over the stored rows   of column c in ascending row index,
// each product rounded separately (no FMA), starting from zero.
#pragma once
#include "common.cuh"
#include "../../include/sipb200.h"

namespace sipb {

struct OpDev {
  int kind;          // SIPB_OP_*
  int mode;          // SIPB_BLOCK_*
  int nblk;          // row blocks: 1, or ndim for TV
  int axis[3];       // storage axis differenced by each block, in row order
  unsigned n[3];     // model grid (n[2] == 1 in 2-D)
  i64 npts;          // n0*n1*n2
  i64 rows;          // rows of the operator
  i64 cols;          // npts, or 2*npts for Minkowski block modes
  i64 row_start[4];  // first row of each block
  double ih[3];      // 1/h per storage axis, value already rounded to T
  double a_xz;       // fl(ih[1]*ih[0]) for D_xz
};

// strides of the storage axes
__device__ __forceinline__ i64 op_stride(const OpDev& op, int a) {
  return a == 0 ? 1 : (a == 1 ? (i64)op.n[0] : (i64)op.n[0] * op.n[1]);
}

// ---- forward: add the terms of row r (acting on one N-block of x) to acc, in column order -----
template <typename T>
__device__ __forceinline__ T op_fwd_terms(const OpDev& op, i64 r, const T* __restrict__ x, T acc) {
  switch (op.kind) {
    case SIPB_OP_IDENTITY:
      return acc + x[r];
    case SIPB_OP_DXZ: {
      const unsigned w = op.n[0] - 1u;
      const unsigned q = (unsigned)r;
      const unsigned j = q / w, i = q - j * w;
      const i64 c = (i64)i + (i64)op.n[0] * j;
      const T a = (T)op.a_xz;
      acc = acc + a * x[c];
      acc = acc + (-a) * x[c + 1];
      acc = acc + (-a) * x[c + op.n[0]];
      acc = acc + a * x[c + op.n[0] + 1];
      return acc;
    }
    default: {
      int b = 0;
      if (op.nblk > 1) {
        b = (r >= op.row_start[1]) ? 1 : 0;
        if (op.nblk > 2 && r >= op.row_start[2]) b = 2;
      }
      const int a = op.axis[b];
      const unsigned q = (unsigned)(r - op.row_start[b]);
      i64 c;
      if (a == 0) {
        c = (i64)q + (i64)(q / (op.n[0] - 1u));
      } else if (a == 1) {
        const unsigned plane = op.n[0] * (op.n[1] - 1u);
        c = (i64)q + (i64)op.n[0] * (q / plane);
      } else {
        c = (i64)q;
      }
      const T ih = (T)op.ih[a];
      acc = acc + (-ih) * x[c];
      acc = acc + ih * x[c + op_stride(op, a)];
      return acc;
    }
  }
}

template <typename T>
__device__ __forceinline__ T op_forward(const OpDev& op, i64 r, const T* __restrict__ x) {
  T acc = (T)0;
  switch (op.mode) {
    case SIPB_BLOCK_RIGHT: return op_fwd_terms<T>(op, r, x + op.npts, acc);
    case SIPB_BLOCK_BOTH:
      acc = op_fwd_terms<T>(op, r, x, acc);
      return op_fwd_terms<T>(op, r, x + op.npts, acc);
    default: return op_fwd_terms<T>(op, r, x, acc);
  }
}

// The adjoint (A' v)[c] is evaluated by op_adjoint_f in kernels.cuh (generic over the value functor).

}  // namespace sipb
