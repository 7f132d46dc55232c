"""Oracle restatement of the PARSDMM solver loop (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/PARSDMM.jl:25-278, PARSDMM_initialize.jl:6-318, rhs_compose.jl:6-40,
argmin_x.jl:6-78 (CDS branch), cg.jl:44-128, update_y_l.jl:6-109, adapt_rho_gamma.jl:8-132,
stop_PARSDMM.jl:7-54.  Serial mode only (`options.parallel == false`).

Scalar types: Julia keeps TF scalars (norm/dot return TF); Float64 literals promote (argmin_x.jl:34
`0.1*...`).  NumPy >= 2 treats Python floats as weak scalars, so every Float64 literal of the Julia
code is written as np.float64 here and every TF literal as TF(...).
"""
from __future__ import annotations

import time

import numpy as np

from . import operators as ops
from . import projectors as proj
from .sip_types import convert_options, dot, eps, log_type_PARSDMM, norm2


# ----------------------------------------------------------------------------------------------
# rhs_compose.jl:24-36
# ----------------------------------------------------------------------------------------------
def rhs_compose(rhs, l, y, rho, TD_OP, p):
    TF = rhs.dtype.type
    rhs[:] = TF(0)
    for ii in range(p):
        temp = ops.spmv_t(TD_OP[ii], TF(rho[ii]) * y[ii] + l[ii])
        rhs += temp                      # BLAS.axpy!(1, temp, rhs)
    return rhs


# ----------------------------------------------------------------------------------------------
# cg.jl:44-128 (M = identity, so z is an alias of r)
# ----------------------------------------------------------------------------------------------
def cg(A, b, tol, maxIter, x):
    """Returns (x, flag, relres, iter).  `x` is the starting guess and is updated in place."""
    TF = b.dtype.type
    n = b.size
    if norm2(b) == 0:                                              # :47
        return np.zeros(n, dtype=TF), -9, TF(0), 0
    r = b - A(x)                                                   # :52
    p = r.copy()                                                   # :55 (z = M(r) = r)
    nr0 = norm2(b)                                                 # :61
    resvec = np.zeros(maxIter, dtype=TF)
    flag = -1
    with np.errstate(all="ignore"):
        if norm2(r) / nr0 <= tol:                                  # :73-76
            return x, 0, resvec[0], 1
        lastIter = 0
        for it in range(1, maxIter + 1):
            lastIter = it
            Ap = A(p)                                              # :85
            gamma = dot(r, r)                                      # :86 dot(r,z)
            alpha = gamma / dot(p, Ap)                             # :88
            if alpha == np.inf or alpha < 0:                       # :91-93
                flag = -2
                break
            x += alpha * p                                         # :95
            r -= alpha * Ap                                        # :97
            resvec[it - 1] = norm2(r) / nr0                        # :100
            if resvec[it - 1] <= tol:                              # :104-106
                flag = 0
                break
            beta = dot(r, r) / gamma                               # :110
            p = r + beta * p                                       # :114 axpby!(1, z, beta, p)
    return x, flag, resvec[lastIter - 1], lastIter


# ----------------------------------------------------------------------------------------------
# argmin_x.jl:6-70, CDS branch :23-39
# ----------------------------------------------------------------------------------------------
def argmin_x(Q, rhs, x, x_solve_tol_ref, i, Q_offsets):
    TF = rhs.dtype.type

    def Af1(v):
        return ops.Ax_CDS(v, Q, Q_offsets)

    with np.errstate(all="ignore"):
        ratio = np.float64(0.1) * np.float64(norm2(Af1(x) - rhs)) / np.float64(norm2(rhs))   # 0.1 is Float64
        cur = _jl_max(ratio, np.float64(TF(10) * eps(TF)))
        if i < 3:                                                  # :33-34
            x_solve_tol_ref = TF(cur)
        else:                                                      # :36
            x_solve_tol_ref = TF(_jl_min(cur, np.float64(x_solve_tol_ref)))
    x, flag, relres, it = cg(Af1, rhs, x_solve_tol_ref, 1000, x)   # :39
    return x, it, relres, TF(x_solve_tol_ref)


def _jl_max(a, b):
    """Julia max: NaN-propagating."""
    if np.isnan(a) or np.isnan(b):
        return np.float64(np.nan)
    return a if a > b else b


def _jl_min(a, b):
    if np.isnan(a) or np.isnan(b):
        return np.float64(np.nan)
    return a if a < b else b


# ----------------------------------------------------------------------------------------------
# update_y_l.jl:6-109 (the loop-fusion formulas :65-78; the BLAS branch is algebraically identical)
# ----------------------------------------------------------------------------------------------
def update_y_l(x, p, i, y, y_old, l, l_old, rho, gamma, prox, TD_OP, log, P_sub, counter, x_hat, r_pri, s,
               feasibility_only):
    TF = x.dtype.type
    rho1 = (TF(1.0) / rho).astype(TF)                              # :34
    for ii in range(p):
        y_old[ii][:] = y[ii]                                       # :39-40
        l_old[ii][:] = l[ii]
        s[ii][:] = ops.spmv(TD_OP[ii], x)                          # :43
        if gamma[ii] == 1:                                         # :65-70
            y[ii][:] = s[ii] - l[ii] * rho1[ii]
            y[ii] = prox[ii](y[ii])
            r_pri[ii][:] = -s[ii] + y[ii]
            l[ii][:] = l[ii] + rho[ii] * r_pri[ii]
        else:                                                      # :71-78
            x_hat[ii][:] = gamma[ii] * s[ii] + (TF(1.0) - gamma[ii]) * y[ii]
            y[ii][:] = x_hat[ii] - l[ii] * rho1[ii]
            y[ii] = prox[ii](y[ii])
            r_pri[ii][:] = -s[ii] + y[ii]
            l[ii][:] = l[ii] + rho[ii] * (-x_hat[ii] + y[ii])
        log.r_pri[i - 1, ii] = norm2(r_pri[ii])                    # :81
        x_hat[ii][:] = y[ii] - y_old[ii]                           # :82
        log.r_dual[i - 1, ii] = rho[ii] * norm2(ops.spmv_t(TD_OP[ii], x_hat[ii]))   # :84
        if i % 10 == 0 and ((not feasibility_only and ii < p - 1) or feasibility_only):     # :90-94
            x_hat[ii][:] = s[ii]
            P_sub[ii](x_hat[ii])
            with np.errstate(all="ignore"):
                log.set_feasibility[counter - 1, ii] = norm2(x_hat[ii] - s[ii]) / (norm2(s[ii]) + TF(100) * eps(TF))
    if i % 10 == 0:                                                # :103-105
        counter += 1
    return counter


def adapt_decide(TF, d_dHh_dlh, n_d_H_hat, n_d_l_hat, n_d_l, n_d_G_hat, d_dGh_dl, rho_ii, gamma_ii, adjust_rho,
                 adjust_gamma):
    """Scalar part of adapt_rho_gamma.jl:31-37,55-126 for one set: from the six reductions to the new
    (rho_i, gamma_i).  All arguments are TF scalars."""
    safeguard = TF(1e-10) if TF == np.float64 else TF(1e-6)        # :31-35
    eps_correlation = TF(0.3)                                      # :37
    with np.errstate(all="ignore"):
        alpha_reliable = False                                 # :55-59
        alpha_correlation = TF(0)
        if (n_d_H_hat * n_d_l_hat) > safeguard and (n_d_H_hat ** 2) > safeguard and d_dHh_dlh > safeguard:
            alpha_reliable = True
            alpha_correlation = d_dHh_dlh / (n_d_H_hat * n_d_l_hat)
        beta_reliable = False                                  # :61-65
        beta_correlation = TF(0)
        if (n_d_G_hat * n_d_l) > safeguard and (n_d_G_hat ** 2) > safeguard and d_dGh_dl > safeguard:
            beta_reliable = True
            beta_correlation = d_dGh_dl / (n_d_G_hat * n_d_l)

        alpha_comp = False                                     # :67-77
        alpha_hat = TF(0)
        if alpha_reliable and alpha_correlation > eps_correlation:
            alpha_comp = True
            alpha_hat_MG = d_dHh_dlh / (n_d_H_hat ** 2)
            alpha_hat_SD = (n_d_l_hat ** 2) / d_dHh_dlh
            if (TF(2.0) * alpha_hat_MG) > alpha_hat_SD:
                alpha_hat = alpha_hat_MG
            else:
                alpha_hat = alpha_hat_SD - alpha_hat_MG / TF(2.0)
        beta_comp = False                                      # :79-89
        beta_hat = TF(0)
        if beta_reliable and beta_correlation > eps_correlation:
            beta_comp = True
            beta_hat_MG = d_dGh_dl / (n_d_G_hat ** 2)
            beta_hat_SD = (n_d_l ** 2) / d_dGh_dl
            if (TF(2.0) * beta_hat_MG) > beta_hat_SD:
                beta_hat = beta_hat_MG
            else:
                beta_hat = beta_hat_SD - beta_hat_MG / TF(2.0)

        if adjust_rho and not adjust_gamma:                    # :92-101
            if alpha_comp and beta_comp:
                rho_ii = np.sqrt(alpha_hat * beta_hat)
            elif alpha_comp and not beta_comp:
                rho_ii = alpha_hat
            elif not alpha_comp and beta_comp:
                rho_ii = beta_hat
        elif adjust_rho and adjust_gamma:                      # :102-115
            if alpha_comp and beta_comp:
                rho_ii = np.sqrt(alpha_hat * beta_hat)
                gamma_ii = TF(1.0) + ((TF(2.0) * np.sqrt(alpha_hat * beta_hat)) / (alpha_hat + beta_hat))
            elif alpha_comp and not beta_comp:
                rho_ii = alpha_hat
                gamma_ii = TF(1.9)
            elif not alpha_comp and beta_comp:
                rho_ii = beta_hat
                gamma_ii = TF(1.1)
            else:
                gamma_ii = TF(1.5)
        elif not adjust_rho and adjust_gamma:                  # :116-126
            if alpha_comp and beta_comp:
                gamma_ii = TF(1.0) + ((TF(2.0) * np.sqrt(alpha_hat * beta_hat)) / (alpha_hat + beta_hat))
            elif alpha_comp and not beta_comp:
                gamma_ii = TF(1.9)
            elif not alpha_comp and beta_comp:
                gamma_ii = TF(1.1)
            else:
                gamma_ii = TF(1.5)
    return rho_ii, gamma_ii


# ----------------------------------------------------------------------------------------------
# adapt_rho_gamma.jl:8-132
# ----------------------------------------------------------------------------------------------
def adapt_rho_gamma(gamma, rho, adjust_gamma, adjust_rho, y, y_old, s, s_0, l, l_hat_0, l_0, l_old, y_0, p,
                    l_hat):
    TF = rho.dtype.type
    with np.errstate(all="ignore"):
        for ii in range(p):
            l_hat[ii][:] = l_old[ii] + rho[ii] * (-s[ii] + y_old[ii])      # :41
            d_l_hat = l_hat[ii] - l_hat_0[ii]
            d_H_hat = s[ii] - s_0[ii]
            d_dHh_dlh = dot(d_H_hat, d_l_hat)                      # :46
            n_d_H_hat = norm2(d_H_hat)
            n_d_l_hat = norm2(d_l_hat)
            d_l = l[ii] - l_0[ii]
            n_d_l = norm2(d_l)
            d_G_hat = -(y[ii] - y_0[ii])
            n_d_G_hat = norm2(d_G_hat)
            d_dGh_dl = dot(d_G_hat, d_l)                           # :53

            rho[ii], gamma[ii] = adapt_decide(TF, d_dHh_dlh, n_d_H_hat, n_d_l_hat, n_d_l, n_d_G_hat, d_dGh_dl, rho[ii],
                                              gamma[ii], adjust_rho, adjust_gamma)
    return rho, gamma, l_hat


# ----------------------------------------------------------------------------------------------
# stop_PARSDMM.jl:7-54  (i and counter are 1-based like the reference)
# ----------------------------------------------------------------------------------------------
def _jl_maximum(a):
    a = np.asarray(a)
    if np.isnan(a).any():
        return np.nan
    return a.max()


def stop_PARSDMM(log, i, evol_rel_tol, feas_tol, obj_tol, adjust_rho, adjust_gamma, adjust_feasibility_rho,
                 ind_ref, counter, TF):
    stop = False
    with np.errstate(all="ignore"):
        if i > 6:                                                  # :23
            ob = log.obj.astype(TF)
            rel = np.abs((ob[i - 6:i] - ob[i - 7:i - 1]) / ob[i - 7:i - 1])
            if _jl_maximum(log.set_feasibility[counter - 2, :]) < feas_tol and _jl_maximum(rel) < obj_tol:
                stop = True
        if i > 5 and _jl_maximum(log.evol_x[i - 6:i]) < evol_rel_tol:   # :29
            stop = True
        lo = max(i - 50, 1)
        if i > 20 and adjust_rho and log.r_pri_total[i - 1] > _jl_maximum(log.r_pri_total[lo - 1:i - 1]):  # :35
            adjust_rho = False
            adjust_feasibility_rho = False
            adjust_gamma = False
            ind_ref = i
        lo2 = max(ind_ref, max(i - 50, 1))
        if (not adjust_rho) and i > (ind_ref + 25) and \
                log.r_pri_total[i - 1] > _jl_maximum(log.r_pri_total[lo2 - 1:i - 1]):                       # :49
            stop = True
    return stop, adjust_rho, adjust_gamma, adjust_feasibility_rho, ind_ref


# ----------------------------------------------------------------------------------------------
# PARSDMM_initialize.jl:6-318 (serial) + PARSDMM.jl:25-258
# ----------------------------------------------------------------------------------------------
def PARSDMM(m, AtA, TD_OP, set_Prop, P_sub, comp_grid, options, x=None, l=None, y=None):
    """Returns (x, log_PARSDMM, l, y).  `x`, `l`, `y` (if given) are updated in place."""
    TF = m.dtype.type
    t0 = time.perf_counter()
    if getattr(options, "parallel", False):
        raise NotImplementedError("oracle restates the serial path only")
    convert_options(options, TF)                                   # PARSDMM.jl:43
    maxit = int(options.maxit)
    evol_rel_tol, feas_tol, obj_tol = options.evol_rel_tol, options.feas_tol, options.obj_tol
    rho_ini = options.rho_ini
    rho_update_frequency = int(options.rho_update_frequency)
    gamma_ini = options.gamma_ini
    adjust_rho, adjust_gamma = bool(options.adjust_rho), bool(options.adjust_gamma)
    adjust_feasibility_rho = bool(options.adjust_feasibility_rho)
    feasibility_only = bool(options.feasibility_only)
    zero_ini_guess = bool(options.zero_ini_guess)
    if x is None:
        x = np.zeros(m.size, dtype=TF)
    l = [] if l is None else l
    y = [] if y is None else y

    # ---- PARSDMM_initialize ------------------------------------------------------------------
    ind_ref = maxit                                                # init.jl:30
    if not options.Minkowski:                                      # :31-38
        N = x.size
    elif zero_ini_guess:
        N = x.size * 2
    else:
        assert x.size == 2 * m.size
        N = x.size
    p = len(TD_OP)
    pp = p if feasibility_only else p - 1                          # :54-56
    rho = np.empty(p, dtype=TF)                                    # :58-63
    if len(rho_ini) == 1:
        rho[:] = rho_ini[0]
    else:
        rho[:] = np.asarray(rho_ini, dtype=TF)
    m_orig = m.copy()
    prox = list(P_sub)                                             # :64-71
    if not feasibility_only:
        # prox for the distance term always sees the CURRENT rho[p] (closure semantics, PARSDMM.jl:234-241)
        prox.append(lambda inp: proj.prox_l2s(inp, rho_box[0][-1], m_orig))
    rho_box = [rho]

    stop = False                                                   # :83-104
    feasibility_initial = np.zeros(len(P_sub), dtype=TF)
    m_ext = np.concatenate([m, np.zeros(m.size, dtype=TF)]) if options.Minkowski else m
    with np.errstate(all="ignore"):
        for ii in range(len(P_sub)):
            Am = ops.spmv(TD_OP[ii], m_ext)
            PAm = P_sub[ii](Am.copy())
            feasibility_initial[ii] = norm2(PAm - Am) / (norm2(Am) + TF(100) * eps(TF))
    if _jl_maximum(feasibility_initial) < feas_tol:
        stop = True
    for ii in range(pp):                                           # :107-114
        if set_Prop.ncvx[ii]:
            rho_update_frequency = 3
            adjust_gamma = False
            gamma_ini = TF(0.75)

    if len(l) == 0:                                                # :120-127
        l = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(p)]
    if len(y) == 0:
        y = [np.zeros(TD_OP[i].shape[0], dtype=TF) for i in range(p)]
    gamma = np.full(p, gamma_ini, dtype=TF)                        # :159
    ly = [TD_OP[i].shape[0] for i in range(p)]
    zl = lambda: [np.zeros(n, dtype=TF) for n in ly]               # noqa: E731
    y_0, y_old, l_0, l_old, l_hat_0, l_hat = zl(), zl(), zl(), zl(), zl(), zl()
    x_hat, s_0, s, r_pri = zl(), zl(), zl(), zl()
    x_old = np.zeros(N, dtype=TF)
    rhs = np.zeros(N, dtype=TF)

    Q, Q_offsets = ops.assemble_Q(AtA, set_Prop.AtA_offsets, rho)  # :216-230

    log = log_type_PARSDMM(np.zeros((maxit, pp)), np.zeros((maxit, p)), np.zeros((maxit, p)), np.zeros(maxit),
                           np.zeros(maxit), np.zeros(maxit), np.zeros(maxit), np.zeros((maxit, p)),
                           np.zeros((maxit, p)), np.zeros(maxit, dtype=np.int64), np.zeros(maxit), {})   # :233-236
    log.set_feasibility[0, :] = feasibility_initial
    if zero_ini_guess:                                             # :304-313
        for v in l:
            v[:] = TF(0)
        for v in y:
            v[:] = TF(0)
        x[:] = TF(0)

    # ---- PARSDMM.jl:63-82 : feasible input ----------------------------------------------------
    if stop:
        x = m.copy()                                               # copy!(x,m) (resizes x if needed)
        if options.Minkowski:
            x = np.concatenate([x, np.zeros(x.size, dtype=TF)])
        _trim(log, 1, 1)
        log.timing = {"total": time.perf_counter() - t0}
        return x, log, l, y
    if options.Minkowski and x.size == m.size:                     # :85-89
        x = np.concatenate([x, np.zeros(m.size, dtype=TF)])

    counter = 2                                                    # :91
    x_solve_tol_ref = TF(1.0)                                      # :93
    timing = {k: 0.0 for k in ("initialization", "form rhs for linear system", "argmin x",
                               "argmin y and l update", "stopping conditions check",
                               "adjust rho and gamma", "Q-update")}
    timing["initialization"] = time.perf_counter() - t0

    for i in range(1, maxit + 1):                                  # :97
        t = time.perf_counter()
        rhs = rhs_compose(rhs, l, y, rho, TD_OP, p)                # :101
        timing["form rhs for linear system"] += time.perf_counter() - t
        t = time.perf_counter()
        x_old[:] = x                                               # :106
        x, it, relres, x_solve_tol_ref = argmin_x(Q, rhs, x, x_solve_tol_ref, i, Q_offsets)   # :107
        log.cg_it[i - 1] = it
        log.cg_relres[i - 1] = relres
        timing["argmin x"] += time.perf_counter() - t
        t = time.perf_counter()
        counter = update_y_l(x, p, i, y, y_old, l, l_old, rho, gamma, prox, TD_OP, log, P_sub, counter,
                             x_hat, r_pri, s, feasibility_only)    # :133
        log.r_dual_total[i - 1] = _tf_sum(log.r_dual[i - 1, :], TF)   # :134
        log.r_pri_total[i - 1] = _tf_sum(log.r_pri[i - 1, :], TF)     # :138
        with np.errstate(all="ignore"):
            if not options.Minkowski:                              # :139-143
                log.obj[i - 1] = TF(0.5) * norm2(x - m) ** 2
            else:
                log.obj[i - 1] = TF(0.5) * norm2(ops.spmv(TD_OP[-1], x) - m) ** 2
            log.evol_x[i - 1] = norm2(x_old - x) / norm2(x)        # :145
        log.rho[i - 1, :] = rho
        log.gamma[i - 1, :] = gamma
        timing["argmin y and l update"] += time.perf_counter() - t

        t = time.perf_counter()
        stop, adjust_rho, adjust_gamma, adjust_feasibility_rho, ind_ref = stop_PARSDMM(
            log, i, evol_rel_tol, feas_tol, obj_tol, adjust_rho, adjust_gamma, adjust_feasibility_rho, ind_ref,
            counter, TF)                                           # :153
        timing["stopping conditions check"] += time.perf_counter() - t
        if stop:                                                   # :154-158
            _trim(log, i, counter)
            log.timing = timing
            return x, log, l, y

        t = time.perf_counter()
        if i == 1:                                                 # :164-180
            for ii in range(p):
                l_hat[ii][:] = l_old[ii] + rho[ii] * (-s[ii] + y_old[ii])
                l_hat_0[ii][:] = l_hat[ii]
                y_0[ii][:] = y[ii]
                s_0[ii][:] = s[ii]
                l_0[ii][:] = l[ii]
        if (adjust_rho or adjust_gamma) and i % rho_update_frequency == 0:    # :182-207
            rho, gamma, l_hat = adapt_rho_gamma(gamma, rho, adjust_gamma, adjust_rho, y, y_old, s, s_0, l, l_hat_0,
                                                l_0, l_old, y_0, p, l_hat)
            if i > 1:
                for ii in range(p):
                    l_hat_0[ii][:] = l_hat[ii]
                    y_0[ii][:] = y[ii]
                    s_0[ii][:] = s[ii]
                    l_0[ii][:] = l[ii]
        if adjust_feasibility_rho and i % 10 == 0:                 # :213-223
            row = log.set_feasibility[counter - 2, :]
            if i > 10 and row.size:
                idx = _jl_findmax_index(row)
                rho[idx] = TF(2.0) * rho[idx]
        rho = np.maximum(np.minimum(rho, TF(1e4)), TF(1e-2)).astype(TF)      # :226 (new array)
        rho_box[0] = rho
        timing["adjust rho and gamma"] += time.perf_counter() - t

        t = time.perf_counter()
        ind_updated = np.nonzero(rho != log.rho[i - 1, :].astype(TF))[0]     # :230
        Q = ops.Q_update(Q, AtA, set_Prop, rho, ind_updated, log.rho[i - 1, :], Q_offsets)    # :243
        timing["Q-update"] += time.perf_counter() - t
        if i == maxit:                                             # :249-252
            _trim(log, i, counter)
    log.timing = timing
    return x, log, l, y


def _tf_sum(row, TF):
    """sum() of a row of TF-valued `Real`s: sequential left fold in TF."""
    acc = TF(0)
    with np.errstate(all="ignore"):
        for v in row:
            acc = TF(acc + TF(v))
    return acc


def _jl_findmax_index(row):
    """findmax: first index of the maximum; NaN wins (Julia isless semantics)."""
    row = np.asarray(row)
    nan = np.isnan(row)
    if nan.any():
        return int(np.argmax(nan))
    return int(np.argmax(row))


def _trim(log, i, counter):
    """output_check_PARSDMM (PARSDMM.jl:261-278): keep rows 1:i, set_feasibility rows 1:counter."""
    log.obj = log.obj[:i]
    log.evol_x = log.evol_x[:i]
    log.r_pri_total = log.r_pri_total[:i]
    log.r_dual_total = log.r_dual_total[:i]
    log.r_pri = log.r_pri[:i, :]
    log.r_dual = log.r_dual[:i, :]
    log.cg_it = log.cg_it[:i]
    log.cg_relres = log.cg_relres[:i]
    log.set_feasibility = log.set_feasibility[:counter, :]
    log.gamma = log.gamma[:i, :]
    log.rho = log.rho[:i, :]
