"""TEST INFRASTRUCTURE ONLY -- numpy/scipy prototype of the interior-point loop ("chd-ipm") that both
the C++ oracle (oracle/ipm_oracle.cpp) and the CUDA product implement.  It exists to pin the algorithm
down in ~200 readable lines; it evaluates functions through the C++ oracle and solves the condensed
KKT system with scipy's sparse LU.

chd-ipm = IPOPT's published primal-dual barrier framework (Waechter & Biegler 2006: slack
reformulation, fraction-to-the-boundary rule, monotone barrier update, IPOPT's scaled optimality error
and termination test, gradient-based NLP scaling) with two documented substitutions that make it
batchable on a GPU: the Lagrangian Hessian is the Gauss-Newton Hessian of the (least-squares) cost
instead of L-BFGS, and globalisation is an l1-merit Armijo backtracking instead of the filter +
restoration phase.  See DESIGN.md section "IPM".
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

INF = 1e19


class Opts:
    tol = 1e-3                 # phys_optim.cpp:578
    constr_viol_tol = 1e-4     # IPOPT defaults
    dual_inf_tol = 1.0
    compl_inf_tol = 1e-4
    mu_init = 0.1
    kappa_eps = 10.0
    kappa_mu = 0.2
    theta_mu = 1.5
    tau_min = 0.99
    kappa1 = 1e-2
    kappa2 = 1e-2
    kappa_sigma = 1e10
    s_max = 100.0
    scal_max_grad = 100.0
    bound_relax = 1e-8
    delta_w = 1e-6
    delta_c = 1e-8
    dw_min = 1e-8
    dw_max = 1e4
    dw_inc = 4.0
    dw_dec = 3.0
    eta = 1e-4
    rho = 0.1
    max_backtrack = 25
    gamma_theta = 1e-5
    gamma_phi = 1e-5
    s_phi = 2.3
    s_theta = 1.1
    delta_sw = 1.0
    eta_phi = 1e-8
    use_curvature = True


def solve_stage(o, stage, max_iter, opts=Opts, verbose=False, log=None):
    o.set_stage(stage)
    n, m = o.n, o.m
    x = o.get_x()
    xlo, xhi = o.var_bounds()
    fixed = xlo == xhi
    x[fixed] = xlo[fixed]
    bnd = (~fixed) & ((xlo > -INF) | (xhi < INF))   # bounded free variables -> extra identity rows
    bidx = np.nonzero(bnd)[0]
    nb = len(bidx)
    Jb = sp.csr_matrix((np.ones(nb), (np.arange(nb), bidx)), shape=(nb, n))
    cl, cu = o.con_bounds()
    cl = np.concatenate([cl, xlo[bidx]])
    cu = np.concatenate([cu, xhi[bidx]])
    M = m + nb
    eq = cl == cu
    iq = ~eq
    free = ~fixed

    def evaluate(xx, jac=True):
        o.set_x(xx)
        f = o.cost()
        c = np.concatenate([o.cons(), xx[bidx]])
        if not jac:
            return f, c
        g = o.grad()
        J = sp.vstack([o.jac(), Jb]).tocsr()
        return f, c, g, J

    f, c, g, J = evaluate(x)
    # gradient-based scaling (IPOPT nlp_scaling_method=gradient-based, max gradient 100)
    gmax = np.abs(g[free]).max() if free.any() else 0.0
    sf = min(1.0, opts.scal_max_grad / gmax) if gmax > opts.scal_max_grad else 1.0
    rowmax = np.asarray(abs(J[:, free]).max(axis=1).todense()).ravel()
    sc = np.where(rowmax > opts.scal_max_grad, opts.scal_max_grad / np.maximum(rowmax, 1e-300), 1.0)
    sc = np.maximum(sc, 1e-8)
    cls, cus = cl * sc, cu * sc
    hasL = iq & (cl > -INF)
    hasU = iq & (cu < INF)
    # relaxed inequality bounds
    dL = np.where(hasL, cls - opts.bound_relax * np.maximum(1, np.abs(cls)), -np.inf)
    dU = np.where(hasU, cus + opts.bound_relax * np.maximum(1, np.abs(cus)), np.inf)
    Sc = sp.diags(sc)

    d = sc * c
    # slack initialisation pushed into the interior
    s = d.copy()
    pL = np.minimum(opts.kappa1 * np.maximum(1, np.abs(dL)), opts.kappa2 * (dU - dL))
    pU = np.minimum(opts.kappa1 * np.maximum(1, np.abs(dU)), opts.kappa2 * (dU - dL))
    pL = np.where(np.isfinite(pL), pL, opts.kappa1 * np.maximum(1, np.abs(dL)))
    pU = np.where(np.isfinite(pU), pU, opts.kappa1 * np.maximum(1, np.abs(dU)))
    s = np.where(hasL, np.maximum(s, dL + pL), s)
    s = np.where(hasU, np.minimum(s, dU - pU), s)
    y = np.zeros(M)
    zL = np.where(hasL, 1.0, 0.0)
    zU = np.where(hasU, 1.0, 0.0)
    mu = opts.mu_init
    nu = 1.0
    mu_min = min(opts.tol, opts.compl_inf_tol) / (opts.kappa_eps + 1.0)  # IPOPT monotone-mu floor
    nfree = int(free.sum())
    eqi = np.nonzero(eq)[0]
    iqi = np.nonzero(iq)[0]
    nE = len(eqi)
    status = -1
    it = 0
    hist = []
    n_bounds = int(hasL.sum() + hasU.sum())
    filt, mu_filter, n_ls_fail = [], None, 0
    n_refactor_total = 0
    delta_w = opts.delta_w
    theta_max = theta_min = 0.0

    for it in range(max_iter + 1):
        gs = sf * g
        Js = Sc @ J
        d = sc * c
        W = sf * o.cost_hessian()
        r_eq = d[eq] - cls[eq]
        r_iq = d[iq] - s[iq]
        gapL = np.where(hasL, s - dL, 1.0)
        gapU = np.where(hasU, dU - s, 1.0)
        rx = (gs + Js.T @ y)[free]
        rs = (-y - zL + zU)[iq]
        compL = np.where(hasL, gapL * zL, 0.0)
        compU = np.where(hasU, gapU * zU, 0.0)
        # IPOPT optimality error
        ysum = np.abs(y).sum() + zL.sum() + zU.sum()
        s_d = max(opts.s_max, ysum / max(M + n_bounds, 1)) / opts.s_max
        s_c = max(opts.s_max, (zL.sum() + zU.sum()) / max(n_bounds, 1)) / opts.s_max
        dual_inf = max(np.abs(rx).max() if nfree else 0.0, np.abs(rs).max() if len(iqi) else 0.0)
        cviol = max(np.abs(r_eq).max() if nE else 0.0, np.abs(r_iq).max() if len(iqi) else 0.0)

        def compl_err(mm):
            e = 0.0
            if hasL.any():
                e = max(e, np.abs(compL[hasL] - mm).max())
            if hasU.any():
                e = max(e, np.abs(compU[hasU] - mm).max())
            return e

        E0 = max(dual_inf / s_d, cviol, compl_err(0.0) / s_c)
        # unscaled measures
        viol_unscaled = 0.0
        cc = c
        viol_unscaled = max(np.maximum(cl - cc, 0).max(), np.maximum(cc - cu, 0).max()) if M else 0.0
        dual_unscaled = dual_inf / sf
        compl_unscaled = compl_err(0.0) / sf
        hist.append((it, f, E0, cviol, dual_inf, mu))
        if verbose:
            print("it %3d f %.6e E0 %.2e viol %.2e (unsc %.2e) dual %.2e compl %.2e mu %.1e nu %.1e" %
                  (it, f, E0, cviol, viol_unscaled, dual_inf, compl_err(0.0), mu, nu))
        if (E0 <= opts.tol and viol_unscaled <= opts.constr_viol_tol and dual_unscaled <= opts.dual_inf_tol
                and compl_unscaled <= opts.compl_inf_tol):
            status = 0
            break
        if it == max_iter:
            status = -1
            break
        while True:
            Emu = max(dual_inf / s_d, cviol, compl_err(mu) / s_c)
            if Emu <= opts.kappa_eps * mu and mu > mu_min:
                mu = max(mu_min, min(opts.kappa_mu * mu, mu ** opts.theta_mu))
            else:
                break
        tau = max(opts.tau_min, 1 - mu)
        # condensed KKT
        sigL = np.where(hasL, zL / gapL, 0.0)
        sigU = np.where(hasU, zU / gapU, 0.0)
        Sig = (sigL + sigU)[iq]
        bvec = (np.where(hasL, mu / gapL, 0.0) - np.where(hasU, mu / gapU, 0.0))[iq]
        JE = Js[eqi][:, free]
        JI = Js[iqi][:, free]
        yc = (sc * y)[:m]   # multipliers of the unscaled rows of the real constraints
        if opts.use_curvature == 2:
            # PSD-by-construction curvature: y+ * Jd^T Jd of the squared-distance rows only
            ypos = np.zeros(m); off_ = 0
            for nm_, rows_ in o.constraint_sets():
                if nm_.startswith('leg-length') or nm_.startswith('ee-dist'):
                    ypos[off_:off_ + rows_] = np.maximum(yc[off_:off_ + rows_], 0.0)
                off_ += rows_
            Wc = o.lag_hessian(ypos)
        else:
            Wc = o.lag_hessian(yc) if opts.use_curvature else 0.0 * W
        Wf = (W + Wc)[free][:, free]
        rhs = np.concatenate([-(gs[free] + JE.T @ y[eq]) - JI.T @ (Sig * r_iq - bvec), -r_eq])
        n_refactor = 0
        while True:
            H = (Wf + JI.T @ sp.diags(Sig) @ JI + delta_w * sp.eye(nfree)).tocsc()
            K = sp.bmat([[H, JE.T], [JE, -opts.delta_c * sp.eye(nE)]]).tocsc()
            if opts.use_curvature == 1:
                # quasi-definite test: the condensed primal block must be positive definite (then the
                # pivot signs of an unpivoted LDL^T are + on variables, - on equality rows in any order)
                try:
                    np.linalg.cholesky(H.toarray())
                    okH = True
                except np.linalg.LinAlgError:
                    okH = False
                if not okH and n_refactor < 15:
                    delta_w = max(delta_w * 8.0, 1e-4)
                    n_refactor += 1
                    continue
            break
        n_refactor_total += n_refactor
        lu = spla.splu(K)
        sol = lu.solve(rhs)
        sol += lu.solve(rhs - K @ sol)
        dxf = sol[:nfree]
        dyE = sol[nfree:]
        dx = np.zeros(n)
        dx[free] = dxf
        ds_i = JI @ dxf + r_iq
        dyI = Sig * ds_i - y[iq] - bvec
        ds = np.zeros(M)
        ds[iq] = ds_i
        dy = np.zeros(M)
        dy[eq] = dyE
        dy[iq] = dyI
        dzL = np.where(hasL, mu / gapL - zL - sigL * ds, 0.0)
        dzU = np.where(hasU, mu / gapU - zU + sigU * ds, 0.0)
        # fraction to the boundary
        a_pr = 1.0
        mL = hasL & (ds < 0)
        if mL.any():
            a_pr = min(a_pr, (-tau * gapL[mL] / ds[mL]).min())
        mU = hasU & (ds > 0)
        if mU.any():
            a_pr = min(a_pr, (tau * gapU[mU] / ds[mU]).min())
        a_du = 1.0
        mz = hasL & (dzL < 0)
        if mz.any():
            a_du = min(a_du, (-tau * zL[mz] / dzL[mz]).min())
        mz = hasU & (dzU < 0)
        if mz.any():
            a_du = min(a_du, (-tau * zU[mz] / dzU[mz]).min())
        # filter line search (Waechter & Biegler 2006, Algorithm A) without the restoration phase
        theta = np.abs(r_eq).sum() + np.abs(r_iq).sum()

        def barrier(ss):
            return -mu * (np.log((ss - dL)[hasL]).sum() + np.log((dU - ss)[hasU]).sum())

        if it == 0:
            theta_max = 1e4 * max(1.0, theta)
            theta_min = 1e-4 * max(1.0, theta)
        if mu != mu_filter:
            filt = []
            mu_filter = mu
        phi0 = sf * f + barrier(s)
        dphi = gs @ dx - mu * ((ds / gapL)[hasL].sum() - (ds / gapU)[hasU].sum())
        alpha = a_pr
        accepted = False
        ftype = False
        for ls in range(opts.max_backtrack):
            xt = x + alpha * dx
            st = s + alpha * ds
            ft, ct = evaluate(xt, jac=False)
            dt_ = sc * ct
            theta_t = np.abs(dt_[eq] - cls[eq]).sum() + np.abs(dt_[iq] - st[iq]).sum()
            phit = sf * ft + barrier(st)
            ok = np.isfinite(phit) and theta_t <= theta_max
            if ok:
                for (tf, pf) in filt:
                    if theta_t >= tf and phit >= pf:
                        ok = False
                        break
            if ok:
                switching = dphi < 0 and alpha * (-dphi) ** opts.s_phi > opts.delta_sw * theta ** opts.s_theta
                if theta <= theta_min and switching:
                    if phit <= phi0 + opts.eta_phi * alpha * dphi:
                        accepted, ftype = True, True
                else:
                    if theta_t <= (1 - opts.gamma_theta) * theta or phit <= phi0 - opts.gamma_phi * theta:
                        accepted = True
                        ftype = switching and phit <= phi0 + opts.eta_phi * alpha * dphi
            if accepted:
                break
            alpha *= 0.5
        if accepted and not ftype:
            filt.append(((1 - opts.gamma_theta) * theta, phi0 - opts.gamma_phi * theta))
        if not accepted:
            n_ls_fail += 1
            if verbose:
                print("   line search failed; taking the smallest step")
        # Levenberg-Marquardt style adaptation of the primal regularisation
        if ls == 0:
            delta_w = max(delta_w / opts.dw_dec, opts.dw_min)
        else:
            delta_w = min(delta_w * opts.dw_inc ** min(ls, 3), opts.dw_max)
        x = x + alpha * dx
        s = s + alpha * ds
        y = y + alpha * dy
        zL = zL + a_du * dzL
        zU = zU + a_du * dzU
        # keep z within kappa_sigma of mu/gap
        gapL = np.where(hasL, s - dL, 1.0)
        gapU = np.where(hasU, dU - s, 1.0)
        zL = np.where(hasL, np.clip(zL, mu / (opts.kappa_sigma * gapL), opts.kappa_sigma * mu / gapL), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (opts.kappa_sigma * gapU), opts.kappa_sigma * mu / gapU), 0.0)
        f, c, g, J = evaluate(x)
        if log is not None:
            log.append(dict(it=it, alpha=alpha, a_pr=a_pr, a_du=a_du, mu=mu, nu=nu, dw=delta_w))
    o.set_x(x)
    return dict(status=status, iters=it, n_ls_fail=n_ls_fail, n_refactor=n_refactor_total, f=f, x=x, hist=hist, viol=viol_unscaled, E0=E0)
