// ops.cuh — matrix-free transform-domain operators (forward rows / adjoint columns).
//
// The reference materialises D_x, D_y, D_z, TV, D_xz and the identity as SparseMatrixCSC via
// Kronecker products (get_discrete_Grad.jl:16-37,51-76; get_TD_operator.jl:12-95) and applies them
// with SparseArrays mul! (update_y_l.jl:43, rhs_compose.jl:28).  Here an operator is a small POD
// descriptor and both A*x and A'*v are evaluated on the fly with exactly the reference's term
// order:   (A*x)[r]  = left fold over the stored columns of row r in ascending column index,
//          (A'*v)[c] = left fold over the stored rows   of column c in ascending row index,
// each product rounded separately (no FMA), starting from zero.
//
// The kernels are bandwidth bound only if the index arithmetic stays cheap (ncu, round 1: the first
// generic per-element version spent ~120 instructions per row, mostly 64-bit IMAD and divisions, and
// ran at 30 % of HBM peak).  Therefore: 32-bit indices, W rows / columns per thread with one
// division per group and carries for the rest, and the operator-kind switch hoisted to the group.
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
  unsigned rs[4];    // first row of each block (32-bit mirror of row_start)
  unsigned kofs;     // slabs: global index of the first owned plane of the slowest axis (0 otherwise)
  unsigned nlast;    // slabs: GLOBAL extent of the slowest axis (== n[2] otherwise)
  i64 npts;          // n0*n1*n2
  i64 rows;          // rows of the operator
  i64 cols;          // npts, or 2*npts for Minkowski block modes
  i64 row_start[4];  // first row of each block
  double ih[3];      // 1/h per storage axis, value already rounded to T
  double a_xz;       // fl(ih[1]*ih[0]) for D_xz
};

__device__ __forceinline__ unsigned op_stride32(const OpDev& op, int a) {
  return a == 0 ? 1u : (a == 1 ? op.n[0] : op.n[0] * op.n[1]);
}

template <typename T, int W>
__device__ __forceinline__ void load_any(const T* p, T (&v)[W]) {
  constexpr int VW = Vec<T>::W;
  if constexpr (W >= VW && W % VW == 0) {
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
#pragma unroll
      for (int h = 0; h < W / VW; ++h) {
        T t[VW];
        vload<T>(p + h * VW, t);
#pragma unroll
        for (int e = 0; e < VW; ++e) v[h * VW + e] = t[e];
      }
      return;
    }
  }
#pragma unroll
  for (int e = 0; e < W; ++e) v[e] = p[e];
}

// ---- forward, one row (generic; used for block-straddling groups and tiny grids) ---------------
template <typename T>
__device__ __forceinline__ T op_fwd_terms1(const OpDev& op, unsigned r, const T* __restrict__ x, T acc) {
  switch (op.kind) {
    case SIPB_OP_IDENTITY:
      return acc + x[r];
    case SIPB_OP_DXZ: {
      const unsigned w = op.n[0] - 1u;
      const unsigned j = r / w, i = r - j * w;
      const unsigned c = i + op.n[0] * j;
      const T a = (T)op.a_xz;
      acc = acc + a * x[c];
      acc = acc + (-a) * x[c + 1u];
      acc = acc + (-a) * x[c + op.n[0]];
      acc = acc + a * x[c + op.n[0] + 1u];
      return acc;
    }
    default: {
      int b = 0;
      if (op.nblk > 1) {
        b = (r >= op.rs[1]) ? 1 : 0;
        if (op.nblk > 2 && r >= op.rs[2]) b = 2;
      }
      const int a = op.axis[b];
      const unsigned q = r - op.rs[b];
      unsigned c;
      if (a == 0) c = q + q / (op.n[0] - 1u);
      else if (a == 1) c = q + op.n[0] * (q / (op.n[0] * (op.n[1] - 1u)));
      else c = q;
      const T ih = (T)op.ih[a];
      acc = acc + (-ih) * x[c];
      acc = acc + ih * x[c + op_stride32(op, a)];
      return acc;
    }
  }
}

// ---- forward, W consecutive rows r0..r0+W-1 acting on one N-block of x; terms are ADDED to acc ----
template <typename T, int W>
__device__ __forceinline__ void op_fwd_terms(const OpDev& op, unsigned r0, const T* __restrict__ x, T (&acc)[W]) {
  if (op.kind == SIPB_OP_IDENTITY) {
    T xv[W];
    load_any<T, W>(x + r0, xv);
#pragma unroll
    for (int e = 0; e < W; ++e) acc[e] = acc[e] + xv[e];
    return;
  }
  if (op.kind != SIPB_OP_DXZ && W > 1) {
    int b = 0;
    if (op.nblk > 1) {
      b = (r0 >= op.rs[1]) ? 1 : 0;
      if (op.nblk > 2 && r0 >= op.rs[2]) b = 2;
    }
    const unsigned bend = (b + 1 < op.nblk) ? op.rs[b + 1] : (unsigned)op.rows;
    const int a = op.axis[b];
    const unsigned q0 = r0 - op.rs[b];
    if (r0 + W <= bend) {       // whole group inside one block: one division, then carries
      const T ih = (T)op.ih[a];
      const T nih = -ih;
      if (a == 2 || (a == 1 && op.n[2] == 1u)) {     // slowest axis: rows and columns coincide
        const unsigned st = op_stride32(op, a);
        T x0[W], x1[W];
        load_any<T, W>(x + q0, x0);
        load_any<T, W>(x + q0 + st, x1);
#pragma unroll
        for (int e = 0; e < W; ++e) {
          acc[e] = acc[e] + nih * x0[e];
          acc[e] = acc[e] + ih * x1[e];
        }
        return;
      }
      if (a == 1) {
        const unsigned plane = op.n[0] * (op.n[1] - 1u);
        if (plane >= (unsigned)W) {
          const unsigned k0 = q0 / plane;
          const unsigned rem0 = q0 - k0 * plane;
          const unsigned st = op.n[0];
#pragma unroll
          for (int e = 0; e < W; ++e) {
            const unsigned k = k0 + ((rem0 + e >= plane) ? 1u : 0u);
            const unsigned c = q0 + e + st * k;
            acc[e] = acc[e] + nih * x[c];
            acc[e] = acc[e] + ih * x[c + st];
          }
          return;
        }
      } else {   // a == 0
        const unsigned w = op.n[0] - 1u;
        if (w >= (unsigned)W) {
          const unsigned jk0 = q0 / w;
          const unsigned i0 = q0 - jk0 * w;
#pragma unroll
          for (int e = 0; e < W; ++e) {
            const unsigned jk = jk0 + ((i0 + e >= w) ? 1u : 0u);
            const unsigned c = q0 + e + jk;
            acc[e] = acc[e] + nih * x[c];
            acc[e] = acc[e] + ih * x[c + 1u];
          }
          return;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < W; ++e) acc[e] = op_fwd_terms1<T>(op, r0 + e, x, acc[e]);
}

template <typename T, int W>
__device__ __forceinline__ void op_forward_n(const OpDev& op, unsigned r0, const T* __restrict__ x, T (&s)[W]) {
#pragma unroll
  for (int e = 0; e < W; ++e) s[e] = (T)0;
  switch (op.mode) {
    case SIPB_BLOCK_RIGHT:
      op_fwd_terms<T, W>(op, r0, x + op.npts, s);
      break;
    case SIPB_BLOCK_BOTH:
      op_fwd_terms<T, W>(op, r0, x, s);
      op_fwd_terms<T, W>(op, r0, x + op.npts, s);
      break;
    default:
      op_fwd_terms<T, W>(op, r0, x, s);
  }
}

template <typename T>
__device__ __forceinline__ T op_forward(const OpDev& op, i64 r, const T* __restrict__ x) {
  T s[1];
  op_forward_n<T, 1>(op, (unsigned)r, x, s);
  return s[0];
}

// ---- grid-point coordinates with carry ---------------------------------------------------------
struct GridIdx {
  unsigned cc;       // index inside one N-block
  unsigned i, j, k;
  bool upper;        // second Minkowski half
};
__device__ __forceinline__ GridIdx grid_decode(i64 c, i64 npts, const unsigned (&n)[3]) {
  GridIdx g;
  g.upper = c >= npts;
  g.cc = (unsigned)(g.upper ? c - npts : c);
  const unsigned t = g.cc / n[0];
  g.i = g.cc - t * n[0];
  g.k = t / n[1];
  g.j = t - g.k * n[1];
  return g;
}
__device__ __forceinline__ void grid_next(GridIdx& g, unsigned npts, const unsigned (&n)[3]) {
  g.cc += 1u;
  if (g.cc == npts) {          // crossed into the second Minkowski half
    g.cc = 0u; g.i = 0u; g.j = 0u; g.k = 0u; g.upper = true;
    return;
  }
  if (++g.i == n[0]) {
    g.i = 0u;
    if (++g.j == n[1]) { g.j = 0u; ++g.k; }
  }
}

__device__ __forceinline__ bool op_touches_half(int mode, bool upper) {
  return mode == SIPB_BLOCK_PLAIN || mode == SIPB_BLOCK_BOTH || (mode == SIPB_BLOCK_LEFT && !upper) ||
         (mode == SIPB_BLOCK_RIGHT && upper);
}

// ---- value fetchers for the adjoint gathers: NV values per row, W consecutive rows, signed start index ----
template <typename T>
struct FetchPlain {          // v[row]
  static constexpr int NV = 1;
  const T* __restrict__ v;
  template <int W> __device__ __forceinline__ void get(int row0, T (&out)[1][W]) const { load_any<T, W>(v + row0, out[0]); }
};
template <typename T>
struct FetchAxpy {           // rho*y[row] + l[row]      (rhs_compose.jl:28)
  static constexpr int NV = 1;
  T rho;
  const T* __restrict__ y;
  const T* __restrict__ l;
  template <int W> __device__ __forceinline__ void get(int row0, T (&out)[1][W]) const {
    T a[W], b[W];
    load_any<T, W>(y + row0, a);
    load_any<T, W>(l + row0, b);
#pragma unroll
    for (int e = 0; e < W; ++e) out[0][e] = rho * a[e] + b[e];
  }
};
template <typename T>
struct FetchDiff {           // y[row] - y_old[row]      (update_y_l.jl:82)
  static constexpr int NV = 1;
  const T* __restrict__ y;
  const T* __restrict__ yo;
  template <int W> __device__ __forceinline__ void get(int row0, T (&out)[1][W]) const {
    T a[W], b[W];
    load_any<T, W>(y + row0, a);
    load_any<T, W>(yo + row0, b);
#pragma unroll
    for (int e = 0; e < W; ++e) out[0][e] = a[e] - b[e];
  }
};
template <typename T>
struct FetchAxpyDiff {       // both of the above from ONE load of y (fused rhs + dual residual gather)
  static constexpr int NV = 2;
  T rho;
  const T* __restrict__ y;
  const T* __restrict__ l;
  const T* __restrict__ yo;
  template <int W> __device__ __forceinline__ void get(int row0, T (&out)[2][W]) const {
    T a[W], b[W], c[W];
    load_any<T, W>(y + row0, a);
    load_any<T, W>(l + row0, b);
    load_any<T, W>(yo + row0, c);
#pragma unroll
    for (int e = 0; e < W; ++e) {
      out[0][e] = rho * a[e] + b[e];
      out[1][e] = a[e] - c[e];
    }
  }
};

// ---- adjoint for a group of W consecutive grid points that stay inside one grid line (caller checks
//      g0.i + W <= n0 and a single Minkowski half): the rows the group needs are consecutive too, so they
//      are fetched with (at most) two wide loads per block instead of 2W scalar gathers.  F::NV independent
//      row vectors are pushed through the same index arithmetic at once (t[q][e] = (A' v_q)[g0 + e]). ------
template <typename T, int W, typename F>
__device__ __forceinline__ void op_adjoint_line(const OpDev& op, int mode, const GridIdx& g0, const F& f,
                                                T (&t)[F::NV][W]) {
  constexpr int NV = F::NV;
  auto scalar_val = [&](int row, T (&v1)[NV]) {
    T v[NV][1];
    f.template get<1>(row, v);
#pragma unroll
    for (int q = 0; q < NV; ++q) v1[q] = v[q][0];
  };
#pragma unroll
  for (int q = 0; q < NV; ++q)
#pragma unroll
    for (int e = 0; e < W; ++e) t[q][e] = (T)0;
  if (!op_touches_half(mode, g0.upper)) return;
  const unsigned cc0 = g0.cc, i0 = g0.i, j = g0.j, k = g0.k;
  if (op.kind == SIPB_OP_IDENTITY) {
    T v[NV][W];
    f.template get<W>((int)cc0, v);
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
      for (int e = 0; e < W; ++e) t[q][e] = t[q][e] + v[q][e];
    return;
  }
  if (op.kind == SIPB_OP_DXZ) {
    const unsigned w = op.n[0] - 1u;
    const T a = (T)op.a_xz;
    const bool jl = j >= 1u, jh = j < op.n[1] - 1u;
#pragma unroll
    for (int e = 0; e < W; ++e) {
      const unsigned i = i0 + e;
      const bool il = i >= 1u, ih_ = i < op.n[0] - 1u;
      const int r = (int)(i + w * j);
      T acc[NV], v1[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) acc[q] = (T)0;
      if (il && jl) { scalar_val(r - 1 - (int)w, v1);
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = acc[q] + a * v1[q]; }
      if (ih_ && jl) { scalar_val(r - (int)w, v1);
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = acc[q] + (-a) * v1[q]; }
      if (il && jh) { scalar_val(r - 1, v1);
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = acc[q] + (-a) * v1[q]; }
      if (ih_ && jh) { scalar_val(r, v1);
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = acc[q] + a * v1[q]; }
#pragma unroll
      for (int q = 0; q < NV; ++q) t[q][e] = acc[q];
    }
    return;
  }
  for (int b = 0; b < op.nblk; ++b) {
    const int a = op.axis[b];
    const T ih = (T)op.ih[a];
    const T nih = -ih;
    const int base = (int)op.rs[b];
    if (a == 0) {
      const int q0 = base + (int)(cc0 - (j + op.n[1] * k));       // row of (i0, j, k)
      if (i0 + (unsigned)W == op.n[0]) {                          // group ends the line: the row of i = n0-1 does
#pragma unroll                                                    // not exist -> element-wise (one group per line)
        for (int e = 0; e < W; ++e) {
          T v1[NV];
          if (i0 + e >= 1u) { scalar_val(q0 + e - 1, v1);
#pragma unroll
            for (int q = 0; q < NV; ++q) t[q][e] = t[q][e] + ih * v1[q]; }
          if (i0 + e < op.n[0] - 1u) { scalar_val(q0 + e, v1);
#pragma unroll
            for (int q = 0; q < NV; ++q) t[q][e] = t[q][e] + nih * v1[q]; }
        }
      } else {
        T v[NV][W], prev[NV];
        f.template get<W>(q0, v);
        if (i0 >= 1u) scalar_val(q0 - 1, prev);
        else {
#pragma unroll
          for (int q = 0; q < NV; ++q) prev[q] = (T)0;
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          T pr = prev[q];
#pragma unroll
          for (int e = 0; e < W; ++e) {
            if (i0 + e >= 1u) t[q][e] = t[q][e] + ih * pr;
            t[q][e] = t[q][e] + nih * v[q][e];            // i < n0-1 holds for the whole group here
            pr = v[q][e];
          }
        }
      }
    } else {
      // axes 1 and 2: the two rows of every grid point are W consecutive rows each
      int st, q0;
      bool has_lo, has_hi;
      if (a == 1) {
        st = (int)op.n[0];
        q0 = base + (int)(cc0 - op.n[0] * k);                     // row of (i0, j, k)
        has_lo = j >= 1u;
        has_hi = j < op.n[1] - 1u;
      } else {
        st = (int)(op.n[0] * op.n[1]);
        q0 = base + (int)cc0;
        const unsigned kg = k + op.kofs;                          // slabs: global plane number, halo row at q0 - st
        has_lo = kg >= 1u;
        has_hi = kg < op.nlast - 1u;
      }
      if (has_lo) {
        T v[NV][W];
        f.template get<W>(q0 - st, v);
#pragma unroll
        for (int q = 0; q < NV; ++q)
#pragma unroll
          for (int e = 0; e < W; ++e) t[q][e] = t[q][e] + ih * v[q][e];
      }
      if (has_hi) {
        T v[NV][W];
        f.template get<W>(q0, v);
#pragma unroll
        for (int q = 0; q < NV; ++q)
#pragma unroll
          for (int e = 0; e < W; ++e) t[q][e] = t[q][e] + nih * v[q][e];
      }
    }
  }
}

// single grid point (line wraps, tails)
template <typename T, typename F>
__device__ __forceinline__ void op_adjoint_pt(const OpDev& op, int mode, const GridIdx& g, const F& f, T (&t)[F::NV]) {
  T tt[F::NV][1];
  op_adjoint_line<T, 1>(op, mode, g, f, tt);
#pragma unroll
  for (int q = 0; q < F::NV; ++q) t[q] = tt[q][0];
}

}  // namespace sipb
