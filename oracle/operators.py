"""Oracle restatement of the linear-operator / CDS index work (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/get_discrete_Grad.jl, get_TD_operator.jl, mat2CDS.jl, CDS_MVp.jl,
CDS_MVp_MT.jl, CDS_scaled_add!.jl, Q_update!.jl (CDS branch) and the Q assembly of
PARSDMM_initialize.jl:216-230.  SciPy CSC matrices stand in for Julia's SparseMatrixCSC; SciPy's
csc_matvec / csr_matvec accumulate in the same (ascending stored index) order as Julia's
spmatmul / adjoint mul!, without FMA.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------------------------
# get_discrete_Grad.jl
# ----------------------------------------------------------------------------------------------
def _D1(n: int, h, TF) -> sp.csc_matrix:
    """1-D forward difference, (n-1) x n, values (-1)/h and (1)/h computed in TF.

    get_discrete_Grad.jl:22-23 / :58-60:
      Dx = spdiagm(0 => ones(TF,n-1)*-1, 1 => ones(TF,n-1)*1); Dx = Dx[1:end-1,:] ./ h
    """
    h = TF(h)
    neg = (np.ones(n - 1, dtype=TF) * TF(-1)) / h
    pos = (np.ones(n - 1, dtype=TF) * TF(1)) / h
    D = sp.diags([neg, pos], [0, 1], shape=(n - 1, n), format="csc", dtype=TF)
    return D


def _I(n: int, TF) -> sp.csc_matrix:
    return sp.identity(n, dtype=TF, format="csc")


def _kron(*mats) -> sp.csc_matrix:
    out = mats[0]
    for m in mats[1:]:
        out = sp.kron(out, m, format="csc")
    out = sp.csc_matrix(out)
    out.sort_indices()
    return out


def get_discrete_Grad(*args):
    """2-D: get_discrete_Grad(n1,n2,h1,h2,TD_type)   (get_discrete_Grad.jl:16-37)
       3-D: get_discrete_Grad(n1,n2,n3,h1,h2,h3,TD_type) (get_discrete_Grad.jl:51-76)
    h1.. must already be TF scalars (np.float32/np.float64)."""
    if len(args) == 5:
        n1, n2, h1, h2, TD_type = args
        TF = type(h1)
        Ix, Iz = _I(n1, TF), _I(n2, TF)
        Dx, Dz = _D1(n1, h1, TF), _D1(n2, h2, TF)
        if TD_type == "D_z":
            return _kron(Dz, Ix)
        if TD_type == "D_x":
            return _kron(Iz, Dx)
        if TD_type in ("TV", "D2D"):
            D = sp.vstack([_kron(Dz, Ix), _kron(Iz, Dx)], format="csc", dtype=TF)
            D.sort_indices()
            return D
        raise ValueError("unknown 2D TD_type " + TD_type)
    n1, n2, n3, h1, h2, h3, TD_type = args
    TF = type(h1)
    Ix, Iy, Iz = _I(n1, TF), _I(n2, TF), _I(n3, TF)
    Dx, Dy, Dz = _D1(n1, h1, TF), _D1(n2, h2, TF), _D1(n3, h3, TF)
    if TD_type == "D_z":
        return _kron(Dz, Iy, Ix)
    if TD_type == "D_y":
        return _kron(Iz, Dy, Ix)
    if TD_type == "D_x":
        return _kron(Iz, Iy, Dx)
    if TD_type in ("TV", "D3D"):
        D = sp.vstack([_kron(Dz, Iy, Ix), _kron(Iz, Dy, Ix), _kron(Iz, Iy, Dx)], format="csc", dtype=TF)
        D.sort_indices()
        return D
    raise ValueError("unknown 3D TD_type " + TD_type)


# ----------------------------------------------------------------------------------------------
# get_TD_operator.jl:12-95 (sparse banded operators only; JOLI transforms are out of scope)
# ----------------------------------------------------------------------------------------------
def get_TD_operator(comp_grid, TD_type: str, TF):
    """Returns (TD_OP, AtA_diag, dense, TD_n, banded)."""
    h1 = TF(comp_grid.d[0])
    h2 = TF(comp_grid.d[1])
    n1 = int(comp_grid.n[0])
    n2 = int(comp_grid.n[1])
    if len(comp_grid.n) == 3 and comp_grid.n[2] > 1:  # get_TD_operator.jl:26
        h3 = TF(comp_grid.d[2])
        n3 = int(comp_grid.n[2])
        if TD_type in ("TV", "D3D"):
            return (get_discrete_Grad(n1, n2, n3, h1, h2, h3, TD_type), False, False,
                    (n1 - 1 + n1 + n1, n2 - 1 + n2 + n2, n3 - 1 + n3 + n3), True)
        if TD_type == "D_z":
            return get_discrete_Grad(n1, n2, n3, h1, h2, h3, TD_type), False, False, (n1, n2, n3 - 1), True
        if TD_type == "D_x":
            return get_discrete_Grad(n1, n2, n3, h1, h2, h3, TD_type), False, False, (n1 - 1, n2, n3), True
        if TD_type == "D_y":
            return get_discrete_Grad(n1, n2, n3, h1, h2, h3, TD_type), False, False, (n1, n2 - 1, n3), True
        if TD_type == "identity":
            return _I(n1 * n2 * n3, TF), True, False, (n1, n2, n3), True
        raise ValueError("oracle: transform-domain operator %r is outside the CDS hot path" % TD_type)
    # 2-D
    if TD_type in ("TV", "D2D"):
        return get_discrete_Grad(n1, n2, h1, h2, TD_type), False, False, ((n1 - 1) + n1, n2 + (n2 - 1)), True
    if TD_type == "D_z":
        return get_discrete_Grad(n1, n2, h1, h2, TD_type), False, False, (n1, n2 - 1), True
    if TD_type == "D_x":
        return get_discrete_Grad(n1, n2, h1, h2, TD_type), False, False, (n1 - 1, n2), True
    if TD_type == "D_xz":  # get_TD_operator.jl:69-73
        D_x = get_discrete_Grad(n1, n2, h1, h2, "D_x")
        D_z = get_discrete_Grad(n1 - 1, n2, h1, h2, "D_z")
        A = sp.csc_matrix(D_z @ D_x)
        A.sort_indices()
        return A.astype(TF), False, False, (n1 - 1, n2 - 1), True
    if TD_type == "identity":
        return _I(n1 * n2, TF), True, False, (n1, n2), True
    raise ValueError("oracle: transform-domain operator %r is outside the CDS hot path" % TD_type)


# ----------------------------------------------------------------------------------------------
# mat2CDS.jl:7-32
# ----------------------------------------------------------------------------------------------
def mat2CDS(A: sp.spmatrix):
    """Sparse -> compressed diagonal storage.  Returns (R [m x ndiag, Fortran order], offset int64[])."""
    A = sp.csc_matrix(A)
    TF = A.dtype.type
    coo = A.tocoo()                       # findnz(A): every STORED entry (mat2CDS.jl:10)
    d = np.sort(coo.col.astype(np.int64) - coo.row.astype(np.int64))
    if d.size:
        keep = np.concatenate(([True], np.diff(d) != 0))  # d[findall(!iszero, diff([-Inf;d]))]
        offset = d[keep]
    else:
        offset = np.zeros(0, dtype=np.int64)
    m, n = A.shape
    R = np.zeros((m, offset.size), dtype=TF, order="F")
    for i, off in enumerate(offset):
        dA = A.diagonal(int(off))
        if off >= 0:
            R[: dA.size, i] = dA           # mat2CDS.jl:24
        else:
            R[m - dA.size:, i] = dA        # mat2CDS.jl:26
    return R, offset.astype(np.int64)


# ----------------------------------------------------------------------------------------------
# CDS_MVp.jl:9-28 / CDS_MVp_MT.jl:9-25 (+ Ax_CDS_MT zero fill, argmin_x.jl:72-78)
# ----------------------------------------------------------------------------------------------
def CDS_MVp(N: int, ndiags: int, R: np.ndarray, offset: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """y += A*x, one pass per diagonal in the order of `offset` (accumulation order matters)."""
    for i in range(ndiags):
        d = int(offset[i])
        r0 = max(0, -d)           # 1-based max(1,1-d)
        r1 = min(N, N - d)        # exclusive upper bound of 1-based min(N,N-d)
        c0 = max(0, d)
        if r1 > r0:
            y[r0:r1] = y[r0:r1] + R[r0:r1, i] * x[c0:c0 + (r1 - r0)]
    return y


def Ax_CDS(x: np.ndarray, Q: np.ndarray, Q_offsets: np.ndarray) -> np.ndarray:
    """Ax_CDS_MT (argmin_x.jl:72-78): zero-fill then CDS_MVp_MT."""
    out = np.zeros_like(x)
    return CDS_MVp(Q.shape[0], Q.shape[1], Q, Q_offsets, x, out)


# ----------------------------------------------------------------------------------------------
# CDS_scaled_add!.jl:8-26 and Q_update!.jl:45-49
# ----------------------------------------------------------------------------------------------
def CDS_scaled_add(A: np.ndarray, B: np.ndarray, A_offsets, B_offsets, alpha) -> None:
    A_offsets = np.asarray(A_offsets)
    for k in range(len(B_offsets)):
        cols = np.nonzero(A_offsets == B_offsets[k])[0]
        if cols.size == 0:
            raise RuntimeError("attempted to update a diagonal in A in CDS storage that does not exist. "
                               "A and B need to have the same nonzero diagonals")
        for c in cols:
            A[:, c] = A[:, c] + alpha * B[:, k]


def Q_update(Q, AtA, set_Prop, rho, ind_updated, log_rho_row, Q_offsets):
    """Q_update!.jl:45-49: Q += (rho_new - rho_logged) * AtA_i for every changed i (ascending i)."""
    TF = Q.dtype.type
    for ii in ind_updated:
        CDS_scaled_add(Q, AtA[ii], Q_offsets, set_Prop.AtA_offsets[ii], TF(rho[ii]) - TF(log_rho_row[ii]))
    return Q


def assemble_Q(AtA, AtA_offsets, rho):
    """PARSDMM_initialize.jl:216-230.

    all_offsets is a 999x99 ZERO-padded table scanned column-major by `unique`, so the order is:
    offsets of AtA[1] (ascending), then 0 (from the padding, if not seen yet), then the unseen
    offsets of AtA[2], ...   (limits: <=999 diagonals per operator, <=99 operators)."""
    TF = AtA[0].dtype.type
    assert len(AtA) <= 99 and all(len(o) <= 999 for o in AtA_offsets)
    all_offsets = np.zeros((999, 99), dtype=np.int64)
    for i in range(len(AtA)):
        all_offsets[: len(AtA_offsets[i]), i] = AtA_offsets[i]
    flat = all_offsets.flatten(order="F")
    _, first = np.unique(flat, return_index=True)
    Q_offsets = flat[np.sort(first)].astype(np.int64)   # unique() keeps first-appearance order
    Q = np.zeros((AtA[0].shape[0], Q_offsets.size), dtype=TF, order="F")
    for i in range(len(AtA)):
        for j in range(len(AtA_offsets[i])):
            col = np.nonzero(Q_offsets == AtA_offsets[i][j])[0]
            for c in col:
                Q[:, c] = Q[:, c] + TF(rho[i]) * AtA[i][:, j]
    return Q, Q_offsets


# ----------------------------------------------------------------------------------------------
# sparse mat-vec helpers with Julia's accumulation order
# ----------------------------------------------------------------------------------------------
def spmv(A: sp.csc_matrix, x: np.ndarray) -> np.ndarray:
    """mul!(s, A, x) for SparseMatrixCSC (update_y_l.jl:43)."""
    return np.asarray(A @ x).ravel().astype(A.dtype, copy=False)


def spmv_t(A: sp.csc_matrix, v: np.ndarray) -> np.ndarray:
    """A' * v (rhs_compose.jl:28, update_y_l.jl:84): per output column a left fold over stored rows."""
    return np.asarray(A.T @ v).ravel().astype(A.dtype, copy=False)


def AtA_sparse(A: sp.csc_matrix) -> sp.csc_matrix:
    """TD_OP' * TD_OP (PARSDMM_precompute_distribute.jl:47)."""
    C = sp.csc_matrix(sp.csc_matrix(A.T) @ A)
    C.sort_indices()
    return C.astype(A.dtype)
