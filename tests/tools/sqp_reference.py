"""tests/tools/sqp_reference.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the batched SQP solver `ntgb_solve_sqp` (ntg_b200/csrc/ntg_sqp.cuh), one problem
at a time, driven by the CPU oracle's evaluations.  It exists so that the algorithm (dual active-set
QP of Goldfarb and Idnani on the reduced space, elastic retry, L1 merit, damped BFGS) can be checked
here without a GPU, and so that the GPU solver's answers have an independent implementation of the
SAME algorithm to be compared with (tests/test_gpu_sqp.py).  The role NPSOL plays for the reference
(/root/reference/src/ntg.c:250-253).
"""
from __future__ import annotations

import numpy as np

INF = 1e300
BIG = 1e19  # |bound| >= BIG: no bound (NPSOL's "infinite bound size", ntg.c:248 sets 1e20... the expanded bounds use it)


def gi_qp(G, g0, Arows, bl, bu, max_iter=None):
    """min 1/2 x'Gx + g0'x  s.t.  bl <= Arows x <= bu (rows with bl == bu are equalities).
    Goldfarb-Idnani dual active set.  Returns x, lam (signed: >0 lower bound active, <0 upper),
    istate (0 free, 1 lower, 2 upper, 3 equality), status (0 ok, 1 infeasible, 2 iteration limit), iterations."""
    n = G.shape[0]
    m = Arows.shape[0]
    L = np.linalg.cholesky(G)
    J = np.linalg.inv(L).T.copy()          # J J' = G^-1
    R = np.zeros((n, n))
    x = -J @ (J.T @ g0)
    q = 0
    act = []            # (row, sign, droppable)
    u = np.zeros(n + 1)
    state = np.zeros(m, dtype=int)
    iseq = (bl == bu)
    if max_iter is None:
        max_iter = 10 * (n + m) + 20
    scale = np.maximum(1.0, np.abs(Arows).max(axis=1)) if m else np.ones(0)
    it = 0
    status = 0
    while True:
        it += 1
        if it > max_iter:
            status = 2
            break
        # most violated constraint, equalities first
        ax = Arows @ x
        vlo = np.where(bl > -BIG, bl - ax, -INF)
        vhi = np.where(bu < BIG, ax - bu, -INF)
        viol = np.maximum(vlo, vhi) / scale
        viol[state != 0] = -INF
        tol = 1e-10
        cand = -1
        if iseq.any():
            ve = np.where(iseq, viol, -INF)
            if ve.max() > tol:
                cand = int(np.argmax(ve))
        if cand < 0:
            if m == 0 or viol.max() <= tol:
                break
            cand = int(np.argmax(viol))
        sgn = 1.0 if vlo[cand] >= vhi[cand] else -1.0
        npv = sgn * Arows[cand]
        b = bl[cand] if sgn > 0 else -bu[cand]
        s = npv @ x - b                     # < 0
        u[q] = 0.0
        inner = 0
        while True:
            inner += 1
            if inner > 4 * (n + m) + 10:
                status = 2
                break
            d = J.T @ npv
            z = J[:, q:] @ d[q:]
            r = np.linalg.solve(R[:q, :q], d[:q]) if q > 0 else np.zeros(0)
            zn = z @ npv
            t2 = -s / zn if zn > 1e-13 * max(1.0, npv @ npv) else INF
            t1, l = INF, -1
            for k in range(q):
                if act[k][2] and r[k] > 0.0:
                    tk = u[k] / r[k]
                    if tk < t1:
                        t1, l = tk, k
            t = min(t1, t2)
            if t >= INF:
                status = 1
                break
            if t2 >= INF:
                u[:q] -= t * r
                u[q] += t
                q = _drop(J, R, u, act, state, l, q)
                continue
            x = x + t * z
            u[:q] -= t * r
            u[q] += t
            if t == t2:
                # add the constraint: Givens rotations fold d[q+1:] into d[q]
                dd = d.copy()
                for j in range(n - 1, q, -1):
                    a, bb = dd[j - 1], dd[j]
                    h = np.hypot(a, bb)
                    if h == 0.0:
                        continue
                    c_, s_ = a / h, bb / h
                    dd[j - 1], dd[j] = h, 0.0
                    cj1, cj = J[:, j - 1].copy(), J[:, j].copy()
                    J[:, j - 1] = c_ * cj1 + s_ * cj
                    J[:, j] = -s_ * cj1 + c_ * cj
                R[:q + 1, q] = dd[:q + 1]
                act.append((cand, sgn, not iseq[cand]))
                state[cand] = 3 if iseq[cand] else (1 if sgn > 0 else 2)
                q += 1
                break
            q = _drop(J, R, u, act, state, l, q)
            s = npv @ x - b
        if status:
            break
    lam = np.zeros(m)
    for k in range(q):
        lam[act[k][0]] = act[k][1] * u[k]
    return x, lam, state, status, it, (cand if status == 1 else -1)


def soft_qp(B, g, A, bl, bu, nu_cap, rho, max_pass=100):
    """QP with rows that may turn SOFT: a row that cannot be satisfied together with the others (the
    dual active-set method finds no step when it tries to add it), or whose multiplier exceeds nu_cap,
    leaves the constraint set and enters the objective as  nu_cap*|shortfall| + rho/2*shortfall^2
    (linearised at the violated side).  With no soft rows this is the plain SQP subproblem."""
    m = A.shape[0]
    soft = np.zeros(m, dtype=int)    # 0 hard, +1 soft at its lower bound, -1 soft at its upper bound
    for npass in range(max_pass):
        S = soft != 0
        r = np.where(soft > 0, bl, np.where(soft < 0, bu, 0.0))
        sg = soft.astype(float)
        AS = A[S]
        B2 = B + rho * AS.T @ AS
        g2 = g - AS.T @ (sg[S] * nu_cap + rho * r[S])
        bl2 = np.where(soft > 0, -1e20, bl)
        bu2 = np.where(soft < 0, 1e20, bu)
        x, lam, state, st, it, failed = gi_qp(B2, g2, A, bl2, bu2)
        if st == 2:
            return x, lam, state, 2, soft
        if st == 1:
            ax = A[failed] @ x
            soft[failed] = 1 if (bl[failed] > -BIG and bl[failed] - ax >= ax - bu[failed]) or bu[failed] >= BIG else -1
            continue
        over = np.abs(lam) > nu_cap
        over &= ~S
        over &= bl != bu if False else True
        if over.any():
            soft[over] = np.where(lam[over] > 0, 1, -1)
            continue
        # multipliers of the soft rows
        res = A @ x - r
        lam = lam + np.where(soft > 0, nu_cap + rho * (-res), 0.0) - np.where(soft < 0, nu_cap + rho * res, 0.0)
        soft_qp.passes = npass + 1
        return x, lam, state, 0, soft
    return x, lam, state, 3, soft


def _drop(J, R, u, act, state, l, q):
    """remove active constraint at position l; u[q] (the multiplier of the constraint being added) moves to u[q-1]"""
    state[act[l][0]] = 0
    n = J.shape[0]
    for k in range(l, q - 1):
        R[:, k] = R[:, k + 1]
        u[k] = u[k + 1]
        act[k] = act[k + 1]
    u[q - 1] = u[q]
    u[q] = 0.0
    R[:, q - 1] = 0.0
    act.pop()
    q -= 1
    for j in range(l, q):
        a, bb = R[j, j], R[j + 1, j]
        h = np.hypot(a, bb)
        if h == 0.0:
            continue
        c_, s_ = a / h, bb / h
        rj, rj1 = R[j, :].copy(), R[j + 1, :].copy()
        R[j, :] = c_ * rj + s_ * rj1
        R[j + 1, :] = -s_ * rj + c_ * rj1
        R[j + 1, j] = 0.0
        cj, cj1 = J[:, j].copy(), J[:, j + 1].copy()
        J[:, j] = c_ * cj + s_ * cj1
        J[:, j + 1] = -s_ * cj + c_ * cj1
    return q


class ReducedNLP:
    """min f(C) s.t. A_eq C = b_eq (eliminated: C = Cpart + N y), hl <= [A_in C; c(C)] <= hu."""

    def __init__(self, port, spec):
        self.port, self.spec = port, spec
        nC = spec.nC
        o = port.eval(spec, np.zeros((1, nC)), mode_obj=0, mode_con=0, dense=False, band=False, linear=True)
        A, bl, bu = o["A"], o["bl"], o["bu"]
        nclin = spec.nclin
        lbl, ubl = bl[nC:nC + nclin], bu[nC:nC + nclin]
        eq = lbl == ubl
        Ae, be = A[eq], lbl[eq]
        self.Ai, self.hl_li, self.hu_li = A[~eq], lbl[~eq], ubl[~eq]
        if Ae.shape[0]:
            U, s, Vt = np.linalg.svd(Ae)
            rank = int((s > 1e-10 * s[0]).sum())
            self.N = Vt[rank:].T.copy()
            self.Cpart = np.linalg.lstsq(Ae, be, rcond=None)[0]
        else:
            self.N = np.eye(nC)
            self.Cpart = np.zeros(nC)
        self.hl = np.concatenate([self.hl_li, bl[nC + nclin:]])
        self.hu = np.concatenate([self.hu_li, bu[nC + nclin:]])
        self.nr = self.N.shape[1]
        self.m = self.hl.size
        self.nevals = 0

    def eval(self, y, deriv=True):
        C = self.Cpart + self.N @ y
        self.nevals += 1
        spec, nC = self.spec, self.spec.nC
        if deriv:
            e = self.port.eval(spec, C[None, :], mode_obj=2, mode_con=2, dense=True, band=False)
            h = np.concatenate([self.Ai @ C, e["c"][0]])
            if spec.ncnln if hasattr(spec, "ncnln") else e["c"].shape[1]:
                Jd = np.nan_to_num(e["Jdense"][0], nan=0.0)
                Jd = Jd.T if Jd.shape[0] == nC else Jd
            else:
                Jd = np.zeros((0, nC))
            Jall = np.vstack([self.Ai, Jd])
            return float(e["f"][0]), self.N.T @ e["g"][0], h, Jall @ self.N
        e = self.port.eval(spec, C[None, :], mode_obj=0, mode_con=0, dense=False, band=False)
        return float(e["f"][0]), None, np.concatenate([self.Ai @ C, e["c"][0]]), None


def row_viol(h, hl, hu):
    return np.maximum(np.maximum(np.where(hl > -BIG, hl - h, 0.0), np.where(hu < BIG, h - hu, 0.0)), 0.0)


def sqp(nlp, y0, gtol=1e-6, ctol=1e-8, max_iter=60, verbose=False, rho_pen=1e4, init_scale=False, restor_hard=False, scalar_nu=True, restor_scale=True):
    """One problem.  Returns dict(y, f, viol, iters, evals, status, lam, istate).

    Per iteration: QP on the reduced space (Goldfarb-Idnani).  If the linearised constraints are
    inconsistent, the rows violated at the current point are taken OUT of the constraint set and
    into the objective as rho/2 * (a_i d - r_i)^2 (rows satisfied now stay hard constraints, so d = 0
    is feasible and the QP cannot fail): a regularised Gauss-Newton step on the violation."""
    nr, m = nlp.nr, nlp.m
    hl, hu = nlp.hl, nlp.hu
    y = y0.copy()
    B = np.eye(nr)
    nu = np.zeros(m)
    lam = np.zeros(m)
    ev0 = nlp.nevals
    f, gr, h, Jr = nlp.eval(y)
    status = 0
    istate = np.zeros(m, dtype=int)
    scaled = not init_scale
    bscale = np.maximum(1.0, np.maximum(np.where(np.abs(hl) < BIG, np.abs(hl), 0), np.where(np.abs(hu) < BIG, np.abs(hu), 0)))
    for it in range(max_iter):
        try:
            np.linalg.cholesky(B)
        except np.linalg.LinAlgError:
            B = np.eye(nr)
        bl, bu = hl - h, hu - h
        x, lam_new, istate, st, _, _ = gi_qp(B, gr, Jr, bl, bu)
        restor = False
        soft = np.zeros(m, dtype=int)
        npass_used = 1
        if st != 0:
            restor = True
            vl = (hl > -BIG) & (bl > 0)       # violated below: want a.d >= bl
            vu = (hu < BIG) & (bu < 0)        # violated above: want a.d <= bu
            V = vl | vu
            soft = V.astype(int)
            r = np.where(vl, bl, np.where(vu, bu, 0.0))
            rho = rho_pen * max(1.0, np.trace(B) / nr) / max(1e-300, (Jr[V] ** 2).sum(axis=1).max())
            AV = Jr[V]
            B2 = B + rho * AV.T @ AV
            g2 = gr - rho * AV.T @ r[V]
            if restor_hard:
                bl2 = np.where(V, -1e20, bl)
                bu2 = np.where(V, 1e20, bu)
                x, lam_new, istate, st2, _, _ = gi_qp(B2, g2, Jr, bl2, bu2)
            else:
                x = -np.linalg.solve(B2, g2)
                lam_new = np.zeros(m)
                istate = np.zeros(m, dtype=int)
            res = Jr @ x - r
            lam_new = lam_new + np.where(vl, rho * (-res), 0.0) + np.where(vu, rho * (-res), 0.0)
        d = x
        viol = row_viol(h, hl, hu)
        vmax = (viol / bscale).max() if m else 0.0
        grL = gr - Jr.T @ lam_new
        kkt = np.abs(grL).max()
        if verbose:
            print(f"passes {npass_used} it {it:3d} f {f:.10g} viol {vmax:.3e} |grL| {kkt:.3e} |d| {np.abs(d).max():.3e} soft {int((soft != 0).sum())} nact {(istate != 0).sum()}")
        if vmax <= ctol and kkt <= gtol * max(1.0, abs(f)) and not restor:
            status = 1
            lam = lam_new
            break
        if restor and np.abs(d).max() < 1e-12:
            status = 4  # stationary point of the violation
            break
        al = np.abs(lam_new)
        nu = np.maximum(al, 0.5 * (nu + al))
        if scalar_nu:
            nu = np.full(m, nu.max() if m else 0.0)
        phi0 = f + nu @ viol
        # directional derivative of the L1 merit along d
        Jd = Jr @ d
        dv = np.where((hl > -BIG) & (bl > 0), -Jd, 0.0) + np.where((hu < BIG) & (bu < 0), Jd, 0.0)
        D = gr @ d + nu @ dv
        if D > -1e-14:
            D = -abs(d @ B @ d)
        alpha, ok = 1.0, False
        for _ in range(16):
            ft, _, ht, _ = nlp.eval(y + alpha * d, deriv=False)
            phit = ft + nu @ row_viol(ht, hl, hu)
            if phit <= phi0 + 1e-4 * alpha * D:
                ok = True
                break
            alpha *= 0.5
        if not ok:
            if np.allclose(B, np.eye(nr)):
                status = 2
                break
            B = np.eye(nr)
            nu = np.zeros(m)
            scaled = not init_scale
            continue
        s = alpha * d
        y = y + s
        grL_old = gr - Jr.T @ lam_new
        gr_old = gr
        f, gr, h, Jr = nlp.eval(y)
        uu = (gr - Jr.T @ lam_new) - grL_old
        Bs = B @ s
        sBs = s @ Bs
        su = s @ uu
        if not restor:
            if not scaled and su > 0:
                B = B * (su / sBs if init_scale == 2 else (uu @ uu) / su * nr / np.trace(B))
                Bs = B @ s
                sBs = s @ Bs
                scaled = True
            if su < 0.2 * sBs:       # Powell damping
                th = 0.8 * sBs / (sBs - su)
                uu = th * uu + (1 - th) * Bs
                su = s @ uu
            if sBs > 0 and su > 0:
                B = B - np.outer(Bs, Bs) / sBs + np.outer(uu, uu) / su
        elif restor_scale:
            suf = s @ (gr - gr_old)
            if sBs > 0 and suf > sBs:
                B = B * min(suf / sBs, 1e8)
        lam = lam_new
    viol = row_viol(h, hl, hu)
    return dict(y=y, C=nlp.Cpart + nlp.N @ y, f=f, viol=viol.max() if m else 0.0, iters=it + 1, evals=nlp.nevals - ev0,
                status=status, lam=lam, istate=istate)


def sqp_via_core(lib, nlp, y0, gtol=1e-6, ctol=1e-8, max_iter=60, rho_pen=1e4, nalpha=16):
    """The same iteration with the direction computed by the C++ core of ntg_sqp.cuh compiled for the
    host (tests/tools/sqp_host.cpp); line search and bookkeeping restate k_sqp_update."""
    import ctypes as C
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    nr, m = nlp.nr, nlp.m
    hl, hu = np.ascontiguousarray(nlp.hl), np.ascontiguousarray(nlp.hu)
    y = y0.copy()
    B = np.eye(nr).ravel().copy()
    lam, sprev, grLold, grold, d = np.zeros(max(m, 1)), np.zeros(nr), np.zeros(nr), np.zeros(nr), np.zeros(nr)
    scal = np.zeros(8)
    flag = np.zeros(8, dtype=np.int32)
    flag[3] = 1
    istate = np.zeros(max(m, 1), dtype=np.int32)
    ev0 = nlp.nevals
    status = 0
    lib.sqp_host_step.argtypes = [C.c_int, C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 12 + [C.POINTER(C.c_int)] * 2 + [C.c_double] * 3
    for it in range(max_iter):
        f, gr, h, Jr = nlp.eval(y)
        gr, h, Jr = np.ascontiguousarray(gr), np.ascontiguousarray(h), np.ascontiguousarray(Jr)
        lib.sqp_host_step(nr, m, f, dp(gr), dp(h), dp(hl), dp(hu), dp(Jr), dp(B), dp(lam), dp(sprev), dp(grLold), dp(grold), dp(d), dp(scal),
                          ip(flag), ip(istate), gtol, ctol, rho_pen)
        if flag[5] != 0:
            status = int(flag[5])
            break
        nu, phi0, D = scal[0], scal[1], scal[2]
        alpha, ok, best, abest = 1.0, False, np.inf, 0.0
        for _ in range(nalpha):
            ft, _, ht, _ = nlp.eval(y + alpha * d, deriv=False)
            phit = ft + nu * row_viol(ht, hl, hu).sum()
            if phit < best:
                best, abest = phit, alpha
            if phit <= phi0 + 1e-4 * alpha * D:
                ok = True
                best, abest = phit, alpha
                break
            alpha *= 0.5
        if not (best < phi0):
            if flag[3]:
                status = 2
                break
            flag[2] = 1      # reset B, nu; no previous step
            flag[0] = 0
            continue
        sprev[:] = abest * d
        y = y + sprev
        flag[0] = 1
        flag[1] = flag[4]
    f, _, h, _ = nlp.eval(y, deriv=False)
    return dict(y=y, C=nlp.Cpart + nlp.N @ y, f=f, viol=row_viol(h, hl, hu).max() if m else 0.0, iters=it + 1,
                evals=nlp.nevals - ev0, status=status, lam=lam[:m].copy(), istate=istate[:m].copy())
