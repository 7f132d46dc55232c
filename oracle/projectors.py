"""Oracle restatement of the projectors / proximal maps on the hot path (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/projectors/project_bounds!.jl:3-25, project_l1_Duchi!.jl:21-52,
project_l2!.jl:3-16, project_annulus!.jl:3-21, project_cardinality!.jl:3-21, prox_l2s!.jl:3-6,
prox_l1!.jl:8-10.  All functions mutate their first argument in place and return it, like the
Julia `!` functions.
"""
from __future__ import annotations

import numpy as np

from .sip_types import norm1, norm2


def project_bounds(x: np.ndarray, LB, UB) -> np.ndarray:
    """Scalar bounds: x[j] = max(LB, min(x[j], UB)) (project_bounds!.jl:3-12);
    vector bounds: min then max (project_bounds!.jl:14-25)."""
    TF = x.dtype.type
    if np.ndim(LB) == 0:
        np.minimum(x, TF(UB), out=x)
        np.maximum(x, TF(LB), out=x)
    else:
        np.minimum(x, np.asarray(UB, dtype=TF), out=x)
        np.maximum(np.asarray(LB, dtype=TF), x, out=x)
    return x


def _accumulate_pairwise(c: np.ndarray, v: np.ndarray, s, i1: int, n: int):
    """Julia Base._accumulate_pairwise!(+, c, v, s, i1, n) (base/accumulate.jl; Julia >= 1.0):
    pairwise cumulative sum with leaves of fewer than 128 elements.  `cumsum!` on a Float vector
    dispatches here (project_l1_Duchi!.jl:38).  Third-party (Julia Base) arithmetic restated."""
    TF = c.dtype.type
    if n < 128:
        loc = np.cumsum(v[i1:i1 + n], dtype=TF)     # s_ = op(s_, v[i]) sequentially in TF
        c[i1:i1 + n] = s + loc                      # c[i] = op(s, s_)
        return loc[-1]
    n2 = n >> 1
    s_ = _accumulate_pairwise(c, v, s, i1, n2)
    s_ = s_ + _accumulate_pairwise(c, v, s + s_, i1 + n2, n - n2)
    return s_


def julia_cumsum(v: np.ndarray) -> np.ndarray:
    """cumsum!(sv, u) for a Float vector: c[1]=v[1], then pairwise accumulate of the rest."""
    TF = v.dtype.type
    c = np.empty_like(v)
    n = v.size
    if n == 0:
        return c
    # Base._accumulate1!: c[1] = v[1]; _accumulate_pairwise!(op, c, v, v1, 2, n-1)
    c[0] = v[0]
    if n > 1:
        with np.errstate(over="ignore"):
            _accumulate_pairwise(c, v, TF(v[0]), 1, n - 1)
    return c


def project_l1_Duchi(v: np.ndarray, b) -> np.ndarray:
    """project_l1_Duchi!.jl:21-52 (real vectors)."""
    TF = v.dtype.type
    b = TF(b)
    if b <= TF(0):
        raise ValueError("Radius of L1 ball is negative")          # :22
    if norm1(v) <= b:                                               # :23
        return v
    lv = v.size
    u = np.sort(np.abs(v))[::-1].copy()                             # :30-37 (sort order is alg-independent)
    sv = julia_cumsum(u)                                            # :38
    # :41-44  while u[rho+1] > ((sv[rho+1]-b)/(rho+1)) && (rho+1) < lv; rho += 1
    j = np.arange(1, lv + 1).astype(TF)
    cond = u > (sv - b) / j
    rho = int(np.argmin(cond)) if not cond.all() else lv            # number of leading trues
    rho = min(rho, lv - 1)
    rho = max(1, rho)                                               # :45
    theta = max(TF(0), (sv[rho - 1] - b) / TF(rho))                 # :46
    theta = TF(theta)
    v[:] = np.sign(v) * np.maximum(np.abs(v) - theta, TF(0))        # :49
    return v


def project_l2(x: np.ndarray, sigma) -> np.ndarray:
    """project_l2!.jl:3-16."""
    TF = x.dtype.type
    sigma = TF(sigma)
    nl2 = norm2(x)
    if nl2 <= sigma:
        return x
    x *= TF(sigma / nl2)
    return x


def project_annulus(x: np.ndarray, sigma_min, sigma_max) -> np.ndarray:
    """project_annulus!.jl:3-21."""
    TF = x.dtype.type
    sigma_min, sigma_max = TF(sigma_min), TF(sigma_max)
    nl2 = norm2(x)
    if sigma_min <= nl2 <= sigma_max:
        return x
    if nl2 > sigma_max:
        x *= TF(sigma_max / nl2)
    elif nl2 < sigma_min and nl2 > 0:
        x *= TF(sigma_min / nl2)
    elif nl2 < sigma_min and nl2 == 0:
        # ones(TF,n) .* (sigma_min ./ sqrt(length(x))): sqrt(Int) is Float64, rounded to TF by copy!
        x[:] = TF(np.float64(sigma_min) / np.sqrt(np.float64(x.size)))
    return x


def _isless_key(x: np.ndarray) -> np.ndarray:
    """Unsigned integer key whose order is Julia's `isless` on floats: -0.0 < 0.0, every NaN after +Inf."""
    u = x.view(np.uint32 if x.dtype == np.float32 else np.uint64).copy()
    top = np.array(1, dtype=u.dtype) << np.array(8 * u.dtype.itemsize - 1, dtype=u.dtype)
    neg = (u & top) != 0
    key = np.where(neg, ~u, u | top)
    key[np.isnan(x)] = np.iinfo(u.dtype).max
    return key


def project_histogram_relaxed(x: np.ndarray, LB: np.ndarray, UB: np.ndarray) -> np.ndarray:
    """project_histogram_relaxed.jl:9-26: sort_ind = sortperm(x) (stable, `isless` order); the sorted values are clamped
    element by element — min(., UB[j]) first, then max(LB[j], .) — by the (already sorted) bounds and moved back."""
    TF = x.dtype.type
    LB = np.asarray(LB, dtype=TF)
    UB = np.asarray(UB, dtype=TF)
    sort_ind = np.argsort(_isless_key(x), kind="stable")             # :11
    xs = x[sort_ind]                                                 # :12
    xs = np.where(np.isnan(xs) | (xs < UB), xs, UB)                  # :15 min(x[j], UB[j])  (Julia min propagates NaN)
    xs = np.where(np.isnan(xs) | (xs > LB), xs, LB)                  # :16 max(LB[j], x[j])
    xs = np.where(np.isnan(LB) | np.isnan(UB), TF(np.nan), xs)
    x[sort_ind] = xs                                                 # :18-19 (inverse permutation)
    return x


def project_cardinality(x: np.ndarray, k: int) -> np.ndarray:
    """project_cardinality!.jl:3-21: sort_ind = sortperm(x, by=abs, rev=true) (stable => among equal
    magnitudes the lower index ranks first and is kept); x[sort_ind[k+1:end]] .= 0."""
    TF = x.dtype.type
    k = int(k)
    sort_ind = np.argsort(-np.abs(x), kind="stable")
    x[sort_ind[k:]] = TF(0.0)
    return x


def prox_l2s(x: np.ndarray, rho, m: np.ndarray) -> np.ndarray:
    """prox_l2s!.jl:3-6:  x .= (x .* rho .+ m) ./ (rho .+ 1.0)  — the literal 1.0 is Float64, so the
    division runs in Float64 and the result is rounded to TF on assignment."""
    TF = x.dtype.type
    num = x * TF(rho) + m                                       # TF
    den = np.float64(TF(rho)) + np.float64(1.0)                 # Float64
    x[:] = (num.astype(np.float64) / den).astype(TF)
    return x


def prox_l1(x: np.ndarray, rho) -> np.ndarray:
    """prox_l1!.jl:8-10: x .= sign.(x) .* max.(0, abs.(x) .- (1 ./ rho))."""
    TF = x.dtype.type
    thr = TF(TF(1) / TF(rho))          # 1 ./ rho : Int / TF -> TF
    x[:] = np.sign(x) * np.maximum(TF(0.0), np.abs(x) - thr)
    return x


def _fiber_axis(ndim: int, direction: str) -> int:
    """2-D: "x" = columns x[:,i] (axis 0), "z" = rows x[i,:] (axis 1); 3-D: "x","y","z" = axes 0,1,2."""
    if ndim == 2:
        return {"x": 0, "z": 1}[direction]
    return {"x": 0, "y": 1, "z": 2}[direction]


def project_bounds_fiber(x: np.ndarray, LB, UB, TD_n, mode) -> np.ndarray:
    """project_bounds!.jl:38-88 (matrix :38-55, tensor fiber modes :57-88): every fiber along `mode[2]`
    gets x .= min.(max.(x, LB), UB) — max first — with LB/UB indexed along the fiber."""
    TF = x.dtype.type
    if mode[0] != "fiber":
        raise NotImplementedError("bound constraints per slice of a tensor currently not implemented, yet...")   # :83
    X = x.reshape(tuple(TD_n), order="F")
    ax = _fiber_axis(X.ndim, mode[1])
    shp = [1] * X.ndim
    shp[ax] = X.shape[ax]
    lb = np.asarray(LB, dtype=TF).reshape(shp)
    ub = np.asarray(UB, dtype=TF).reshape(shp)
    X[...] = np.minimum(np.maximum(X, lb), ub)
    x[:] = X.ravel(order="F")
    return x


def project_cardinality_fiber(x: np.ndarray, k: int, TD_n, mode) -> np.ndarray:
    """project_cardinality!.jl:23-62 (matrix) and :64-113 (tensor, fiber modes): in every fiber keep the k
    largest magnitudes (stable sortperm => lower index inside the fiber wins ties), zero the rest."""
    TF = x.dtype.type
    if mode[0] != "fiber":
        raise NotImplementedError("oracle: slice modes are outside the hot path")
    X = x.reshape(tuple(TD_n), order="F")
    ax = _fiber_axis(X.ndim, mode[1])
    Xm = np.moveaxis(X, ax, 0)                       # fibers are now columns Xm[:, ...]
    order = np.argsort(-np.abs(Xm), axis=0, kind="stable")
    drop = order[int(k):]
    np.put_along_axis(Xm, drop, TF(0.0), axis=0)
    x[:] = X.ravel(order="F")
    return x


def project_cardinality_slice(x: np.ndarray, k: int, TD_n, mode) -> np.ndarray:
    """project_cardinality!.jl:115-146 (3-D tensor, slice modes): every 2-D slice orthogonal to `mode[2]` keeps its k
    largest magnitudes; ties are broken by the position inside the permuted / reshaped slice (stable sortperm),
    which is the column-major order of the remaining two axes.  As in the reference the "x" and "y" modes work on
    a permuted COPY — the input is NOT mutated and the projected vector is only returned — while "z" (a plain
    reshape) mutates the input."""
    TF = x.dtype.type
    n1, n2, n3 = (int(v) for v in TD_n)
    X = x.reshape((n1, n2, n3), order="F")
    if mode[1] == "x":
        Xp = np.transpose(X, (1, 2, 0)).reshape((n2 * n3, n1), order="F").copy(order="F")       # :122-124
    elif mode[1] == "y":
        Xp = np.transpose(X, (0, 2, 1)).reshape((n1 * n3, n2), order="F").copy(order="F")       # :125-127
    elif mode[1] == "z":
        Xp = X.reshape((n1 * n2, n3), order="F")                                                 # :128 (shares memory)
    else:
        raise ValueError("unknown slice direction")
    order = np.argsort(-np.abs(Xp), axis=0, kind="stable")                                       # :133-136
    np.put_along_axis(Xp, order[int(k):], TF(0.0), axis=0)
    if mode[1] == "x":
        out = np.transpose(Xp.reshape((n2, n3, n1), order="F"), (2, 0, 1))                       # :139-141
    elif mode[1] == "y":
        out = np.transpose(Xp.reshape((n1, n3, n2), order="F"), (0, 2, 1))                       # :142-144
    else:
        x[:] = Xp.reshape(-1, order="F")
        return x
    return np.ascontiguousarray(out.reshape(-1, order="F"))
