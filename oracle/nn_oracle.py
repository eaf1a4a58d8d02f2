"""CPU oracle for the NONLINEAR path (Flux fnn / resnet dynamics + NMPC) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (same reasons as oracle/mpc_oracle.py: the reference solves this NLP with Ipopt through JuMP; neither
exists here, and the reference's own tests only assert |x_linear - x_nl| <= 0.5, test/computation_mpc_test.jl:152-169).
This file RESTATES
  * the network layout the reference's NL modelers read out of `Flux.params`
    (fnn/mpc_modeler_implementation_fnn.jl:88-143, resnet/mpc_modeler_implementation_resnet.jl:87-142):
    W_in (no bias) -> n_hidden x [W_j, b_j] with activation -> W_out (no bias); resnet adds the skip y_{j-1};
  * the NLP those modelers + design_mpc.jl:405-465 define (absolute-coordinate dynamics, input box, deviation cost);
  * `nmpc_sqp`: the algorithmic twin of the CUDA SQP kernel (Gauss-Newton SQP, per-problem condensed QP solved by the same
    OSQP-style ADMM as the linear path, Armijo backtracking on the true cost);
  * `nmpc_local_opt`: an independent solve (scipy L-BFGS-B on the single-shooting NLP with the adjoint gradient) plus a
    first-order KKT certificate, standing in for Ipopt.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
from __future__ import annotations

import dataclasses
import numpy as np

from . import mpc_oracle as mo

STATUS_STALLED = 2       # OSQP's "solved inaccurate" code, used for a line-search stall
ACTIVATIONS = ("relu", "tanh", "sigmoid", "swish", "identity")     # ids shared with include/mpcb200.h (MPCB_ACT_*)


def act(name, h):
    if name == "relu": return np.maximum(h, 0.0)
    if name == "tanh": return np.tanh(h)
    if name == "sigmoid": return 1.0 / (1.0 + np.exp(-h))
    if name == "swish": return h / (1.0 + np.exp(-h))
    if name == "identity": return h
    raise ValueError(name)


def dact(name, h):
    if name == "relu": return (h > 0.0).astype(float)          # NNlib.relu': 0 at 0
    if name == "tanh": return 1.0 - np.tanh(h) ** 2
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-h)); return s * (1.0 - s)
    if name == "swish":
        s = 1.0 / (1.0 + np.exp(-h)); return s + h * s * (1.0 - s)
    if name == "identity": return np.ones_like(h)
    raise ValueError(name)


@dataclasses.dataclass
class NeuralModel:
    """`Flux.params(system.f)` as the reference parses it: params[1] = W_in, then (W_j, b_j) pairs, last = W_out."""
    arch: str                 # "fnn" | "resnet" | "polynet" | "densenet"
    activation: str
    W_in: np.ndarray          # (n_neur, nx + nu)
    W_h: list                 # n_hidden x (n_neur, n_neur)
    b_h: list                 # n_hidden x (n_neur,)
    W_out: np.ndarray         # (nx, n_neur)

    @property
    def nx(self): return self.W_out.shape[0]
    @property
    def nu(self): return self.W_in.shape[1] - self.W_out.shape[0]
    @property
    def n_neur(self): return self.W_in.shape[0]
    @property
    def n_hid(self): return len(self.W_h)


def hidden_states(m: NeuralModel, x, u):
    """y[:, j] of the reference (fnn.jl:126-141 / resnet.jl:125-140), j = 1..n_hid+1, for a batch: list of (B, n_neur)
    plus the pre-activations."""
    xu = np.concatenate([np.atleast_2d(x), np.atleast_2d(u)], axis=1)
    ys = [xu @ m.W_in.T]; pre = []
    for W, b in zip(m.W_h, m.b_h):
        h = ys[-1] @ W.T + b
        if m.arch == "densenet":       # densenet.jl:139-155: the new block is PREPENDED to the previous layer's vector
            pre.append(h)
            ys.append(np.concatenate([act(m.activation, h), ys[-1]], axis=1))
            continue
        if m.arch == "polynet":        # polynet.jl:132-149: branch = act(W y + b); y+ = y + branch + act(W branch + b)  (same W, b)
            br = act(m.activation, h)
            h2 = br @ W.T + b
            pre.append((h, h2))
            ys.append(ys[-1] + br + act(m.activation, h2))
        else:
            pre.append(h)
            ys.append(act(m.activation, h) + (ys[-1] if m.arch == "resnet" else 0.0))
    return ys, pre


def forward(m: NeuralModel, x, u):
    """x_{k+1} = W_out y[:, end]  (fnn.jl:143)."""
    ys, _ = hidden_states(m, x, u)
    return ys[-1] @ m.W_out.T


def jacobian(m: NeuralModel, x, u):
    """f, A = df/dx (B, nx, nx), Bm = df/du (B, nx, nu): forward-mode through the chain."""
    ys, pre = hidden_states(m, x, u)
    Bn = ys[0].shape[0]
    Jm = np.broadcast_to(m.W_in, (Bn,) + m.W_in.shape).copy()          # d y_1 / d [x;u]
    for W, h in zip(m.W_h, pre):
        if m.arch == "densenet":
            Jm = np.concatenate([dact(m.activation, h)[:, :, None] * (W @ Jm), Jm], axis=1)
            continue
        if m.arch == "polynet":
            Jb = dact(m.activation, h[0])[:, :, None] * (W @ Jm)
            Jm = Jm + Jb + dact(m.activation, h[1])[:, :, None] * (W @ Jb)
            continue
        D = dact(m.activation, h)[:, :, None]
        Jn = D * (W @ Jm)
        Jm = Jn + Jm if m.arch == "resnet" else Jn
    Jf = m.W_out @ Jm
    return ys[-1] @ m.W_out.T, Jf[:, :, :m.nx], Jf[:, :, m.nx:]


def rollout(m: NeuralModel, x0, u):
    """x (B, H+1, nx) from x0 (B, nx) and u (B, H, nu)."""
    x0 = np.atleast_2d(x0); Bn, H = u.shape[0], u.shape[1]
    x = np.zeros((Bn, H + 1, m.nx)); x[:, 0] = x0
    for k in range(H):
        x[:, k + 1] = forward(m, x[:, k], u[:, k])
    return x


def reference_nl_residual(m: NeuralModel, x, u):
    """Max violation of the reference's NL-modeler equality constraints by a trajectory (B, H+1, nx), (B, H, nu), with the
    hidden variables y set by their defining equations (fnn.jl:122-144): what Ipopt drives to zero."""
    r = 0.0
    for k in range(u.shape[1]):
        r = max(r, np.abs(x[:, k + 1] - forward(m, x[:, k], u[:, k])).max())
    return r


def reference_nl_variable_count(m: NeuralModel, H):
    """x, e_x, x_reference (nx x (H+1)), u, e_u, u_reference (nu x H), y (n_neur x (n_hid+1) x H)  (fnn.jl:111-119)."""
    return 3 * m.nx * (H + 1) + 3 * m.nu * H + m.n_neur * (m.n_hid + 1) * H


# ----------------------------------------------------------------------------------------------------------------------
# The NLP in single-shooting form (equivalent: x and y are uniquely determined by x0 and u through the equalities)
#   J(u) = e_H' P e_H + sum_{k<H} (e_k' Q e_k + eps_k' R eps_k) [+ sum_{k<H-1} (u_k - u_{k+1})' S (u_k - u_{k+1})]
# ----------------------------------------------------------------------------------------------------------------------
def constant_hessian(nu, H, R, S):
    """Hc = 2 (I (x) R) + 2 D'(I (x) S) D with the reference's switches (design_mpc.jl:436-447: R-term only if R[1,1] != 0,
    S-term only inside that branch and only if S[1,1] != 0)."""
    nz = nu * H
    Hc = np.zeros((nz, nz))
    if R[0, 0] != 0.0:
        Hc += 2.0 * np.kron(np.eye(H), R)
        if S[0, 0] != 0.0:
            D = np.zeros((nu * (H - 1), nz))
            for k in range(H - 1):
                D[k * nu:(k + 1) * nu, k * nu:(k + 1) * nu] = np.eye(nu)
                D[k * nu:(k + 1) * nu, (k + 1) * nu:(k + 2) * nu] = -np.eye(nu)
            Hc += 2.0 * D.T @ np.kron(np.eye(H - 1), S) @ D
    return Hc


def objective(m, Q, P, Hc, u, x0, xref, uref):
    Bn, H, nu = u.shape
    x = rollout(m, x0, u)
    e = x - xref[:, None, :]
    du = (u - uref[:, None, :]).reshape(Bn, -1)
    J = 0.5 * np.einsum("bi,ij,bj->b", du, Hc, du)
    J += np.einsum("bki,ij,bkj->b", e[:, :H], Q, e[:, :H]) + np.einsum("bi,ij,bj->b", e[:, H], P, e[:, H])
    return J, x


def linearize_trajectory(m, Q, P, Hc, u, x0, xref, uref):
    """One Gauss-Newton linearisation: J, gradient g (B, nz), GN Hessian Pc (B, nz, nz) (incl. Hc), the trajectory and the
    sensitivities Gamma_k = d x_k / d u for k = 1..H, shape (B, H, nx, nz)."""
    Bn, H, nu = u.shape; nx = m.nx; nz = nu * H
    x = np.zeros((Bn, H + 1, nx)); x[:, 0] = x0
    Gam = np.zeros((Bn, nx, nz))
    du = (u - uref[:, None, :]).reshape(Bn, nz)
    Pc = np.broadcast_to(Hc, (Bn, nz, nz)).copy()
    Gall = np.zeros((Bn, H, nx, nz))                       # Gamma_1 .. Gamma_H
    g = du @ Hc.T
    e0 = x[:, 0] - xref
    J = 0.5 * np.einsum("bi,bi->b", du, g) + np.einsum("bi,ij,bj->b", e0, Q, e0)
    for k in range(H):
        f, A, Bm = jacobian(m, x[:, k], u[:, k])
        x[:, k + 1] = f
        Gam = A @ Gam
        Gam[:, :, k * nu:(k + 1) * nu] = Bm
        Gall[:, k] = Gam
        W = P if k + 1 == H else Q
        e = f - xref
        WG = W @ Gam                                           # (B, nx, nz)
        Pc += 2.0 * np.einsum("bia,bic->bac", Gam, WG)
        g += 2.0 * np.einsum("bia,ij,bj->ba", Gam, W, e)
        J += np.einsum("bi,ij,bj->b", e, W, e)
    return J, g, Pc, x, Gall


@dataclasses.dataclass
class SqpSettings:
    qp: mo.AdmmSettings = dataclasses.field(default_factory=lambda: mo.AdmmSettings(eps_abs=1e-9, eps_rel=0.0, sigma=0.0, check_every=5, max_iter=1000))
    sqp_max_iter: int = 20
    sqp_tol: float = 1e-6         # ||step||_inf
    ls_max: int = 12              # Armijo halvings
    ls_c1: float = 1e-4
    ls_noise: float = 1e-10       # round-off floor of the cost evaluation, relative to max(1, |J|)


def admm_box_per_problem(Kinv, q, lb, ub, s: mo.AdmmSettings, rho, x0, y0):
    """The box-only ADMM of mpc_oracle.admm_condensed with a PER-PROBLEM cached inverse Kinv (B, nz, nz); same update
    order, same termination at (x~, z+, y+).  Warm start as OSQP: x = x0, z = x0, y = y0."""
    Bn, nz = q.shape
    x = x0.copy(); z = x0.copy(); ys = y0 / rho
    max_iter = -(-s.max_iter // s.check_every) * s.check_every
    iters = np.zeros(Bn, np.int32); status = np.full(Bn, mo.STATUS_MAX_ITER, np.int32)
    xo = np.zeros((Bn, nz)); yo = np.zeros((Bn, nz)); pres = np.zeros(Bn); dres = np.zeros(Bn)
    active = np.ones(Bn, bool); qn = np.abs(q).max(1)
    for it in range(1, max_iter + 1):
        r = rho * (z - ys) + s.sigma * x - q
        t = np.einsum("bij,bj->bi", Kinv, r)
        w = s.alpha * t + (1 - s.alpha) * z + ys
        zn = np.minimum(np.maximum(w, lb), ub)
        ysn = w - zn
        x = s.alpha * t + (1 - s.alpha) * x
        z, ys = zn, ysn
        if it % s.check_every == 0:
            y = rho * ys
            pc = r - (s.sigma + rho) * t
            rp = np.abs(t - z).max(1); rd = np.abs(pc + q + y).max(1)
            ep = s.eps_abs + s.eps_rel * np.maximum(np.abs(t).max(1), np.abs(z).max(1))
            ed = s.eps_abs + s.eps_rel * np.maximum(np.maximum(np.abs(pc).max(1), np.abs(y).max(1)), qn)
            conv = (rp <= ep) & (rd <= ed)
            fin = active & (conv | (it >= max_iter))
            status[active & conv] = mo.STATUS_SOLVED
            iters[fin] = it; xo[fin] = t[fin]; yo[fin] = y[fin]; pres[fin] = rp[fin]; dres[fin] = rd[fin]
            active &= ~fin
            if not active.any(): break
    return xo, yo, iters, status, pres, dres


def admm_box_gen_per_problem(Kinv, q, lb, ub, G, lo_g, hi_g, rho_g, s: mo.AdmmSettings, rho, x0, y0, yg0, nball=0, rad=None):
    """As admm_box_per_problem plus GENERAL rows  lo_g <= G v <= hi_g  (linearised state box and / or terminal equality,
    lo = hi), per-row step sizes rho_g; Kinv is the inverse of  Kgn + (sigma + rho) I + G' diag(rho_g) G.  With nball > 0 the LAST
    nball rows form one ball |G v - lo_g|_2 <= rad (contractive terminal set): projected onto the ball instead of a box.
    Returns the box and row multipliers."""
    Bn, nz = q.shape
    x = x0.copy(); z = x0.copy(); ys = y0 / rho
    zg = np.einsum("bij,bj->bi", G, x0); ysg = yg0 / rho_g
    max_iter = -(-s.max_iter // s.check_every) * s.check_every
    iters = np.zeros(Bn, np.int32); status = np.full(Bn, mo.STATUS_MAX_ITER, np.int32)
    xo = np.zeros((Bn, nz)); yo = np.zeros((Bn, nz)); ygo = np.zeros_like(lo_g); pres = np.zeros(Bn); dres = np.zeros(Bn)
    active = np.ones(Bn, bool); qn = np.abs(q).max(1)
    for it in range(1, max_iter + 1):
        r = rho * (z - ys) + s.sigma * x - q + np.einsum("bij,bi->bj", G, rho_g * (zg - ysg))
        t = np.einsum("bij,bj->bi", Kinv, r)
        tg = np.einsum("bij,bj->bi", G, t)
        w = s.alpha * t + (1 - s.alpha) * z + ys
        zn = np.minimum(np.maximum(w, lb), ub)
        ysn = w - zn
        wg = s.alpha * tg + (1 - s.alpha) * zg + ysg
        zgn = np.minimum(np.maximum(wg, lo_g), hi_g)
        if nball:
            dv = wg[:, -nball:] - lo_g[:, -nball:]
            nd = np.sqrt((dv ** 2).sum(1))
            sc = np.where(nd > rad, rad / np.maximum(nd, 1e-300), 1.0)
            zgn[:, -nball:] = lo_g[:, -nball:] + sc[:, None] * dv
        ysgn = wg - zgn
        x = s.alpha * t + (1 - s.alpha) * x
        z, ys, zg, ysg = zn, ysn, zgn, ysgn
        if it % s.check_every == 0:
            y = rho * ys; yg = rho_g * ysg
            kt = r - (s.sigma + rho) * t - np.einsum("bij,bi->bj", G, rho_g * tg)      # Kgn x~ from the cached factor
            gs = kt + np.einsum("bij,bi->bj", G, yg)
            rp = np.maximum(np.abs(t - z).max(1), np.abs(tg - zg).max(1))
            rd = np.abs(gs + q + y).max(1)
            ep = s.eps_abs + s.eps_rel * np.maximum(np.maximum(np.abs(t).max(1), np.abs(z).max(1)), np.maximum(np.abs(tg).max(1), np.abs(zg).max(1)))
            ed = s.eps_abs + s.eps_rel * np.maximum(np.maximum(np.abs(gs).max(1), np.abs(y).max(1)), qn)
            conv = (rp <= ep) & (rd <= ed)
            fin = active & (conv | (it >= max_iter))
            status[active & conv] = mo.STATUS_SOLVED
            iters[fin] = it; xo[fin] = t[fin]; yo[fin] = y[fin]; ygo[fin] = yg[fin]; pres[fin] = rp[fin]; dres[fin] = rd[fin]
            active &= ~fin
            if not active.any(): break
    return xo, yo, ygo, iters, status, pres, dres


def nmpc_sqp(m: NeuralModel, Q, R, S, P, H, umin, umax, x0, xref, uref, rho, s: SqpSettings = None, u_init=None, y_init=None,
             terminal="none", rho_eq_scale=1e3, xmin=None, xmax=None, state_constraint=False):
    """Twin of the CUDA kernel `nmpc_sqp_kernel`: per problem, repeat { rollout + Jacobians -> GN condensed QP in absolute
    inputs v (box umin <= v <= umax) -> ADMM warm-started at (u, y) -> step d = v - u -> Armijo backtracking on J } until
    ||d||_inf <= sqp_tol (status 1) or sqp_max_iter (status -2); a failed line search ends with status 2.
    terminal="equality" (design_mpc.jl:330-331, e_x[:,end] == 0): the QP carries the linearised rows
    Gamma_H v = Gamma_H u - e_H(u) with row-equilibrated step sizes, the line search runs on the l1 merit
    J + mu |e_H|_1 with mu = max(mu, 1.1 |multipliers|_inf), and a QP that does not converge within the inner cap is
    reported as primal infeasible (-3).
    state_constraint (fnn.jl:146-154, `mpc_state_constraint` present): rows xmin <= x_k(u) + Gamma_k (v - u) <= xmax for
    k = 1..H (the fixed column k = 0 is constant), row-equilibrated, violations enter the same merit.
    General rows are ordered [state rows (k, i) ..., terminal rows]."""
    s = s or SqpSettings()
    x0 = np.atleast_2d(np.asarray(x0, float)); Bn = x0.shape[0]; nx, nu = m.nx, m.nu; nz = nu * H
    xref = np.broadcast_to(np.atleast_2d(np.asarray(xref, float)), (Bn, nx)); uref = np.broadcast_to(np.atleast_2d(np.asarray(uref, float)), (Bn, nu))
    lb = np.tile(np.asarray(umin, float), H); ub = np.tile(np.asarray(umax, float), H)
    Hc = constant_hessian(nu, H, np.asarray(R, float), np.asarray(S, float))
    u = np.tile(np.clip(uref, umin, umax), (1, H)) if u_init is None else np.array(u_init, float).reshape(Bn, nz)
    y = np.zeros((Bn, nz)) if y_init is None else np.array(y_init, float).reshape(Bn, nz)
    status = np.full(Bn, mo.STATUS_MAX_ITER, np.int32); sqp_iters = np.zeros(Bn, np.int32); inner = np.zeros(Bn, np.int64)
    step = np.zeros(Bn); qp_dres = np.zeros(Bn)
    ball = terminal == "contractive"       # e_H' e_H <= 0.9 e_0' e_0 (design_mpc.jl:333-340): the terminal rows are projected onto a ball
    eq = terminal == "equality" or ball; sb = bool(state_constraint)
    rad = np.sqrt(0.9) * np.sqrt(((x0 - xref) ** 2).sum(1))
    xmin_ = np.asarray(xmin, float) if sb else None; xmax_ = np.asarray(xmax, float) if sb else None
    mg = (nx * H if sb else 0) + (nx if eq else 0)
    yg = np.zeros((Bn, mg)); mu = np.zeros(Bn)

    def violation(xt, sel=None):
        """l1 measure of the nonlinear constraints along a trajectory (B', H+1, nx)."""
        xr_ = xref[idx] if sel is None else xref[idx][sel]
        c = np.zeros(xt.shape[0])
        if sb: c += (np.maximum(xt[:, 1:] - xmax_, 0.0) + np.maximum(xmin_ - xt[:, 1:], 0.0)).sum((1, 2))
        if ball: c += np.maximum(np.sqrt(((xt[:, H] - xr_) ** 2).sum(1)) - (rad[idx] if sel is None else rad[idx][sel]), 0.0)
        elif eq: c += np.abs(xt[:, H] - xr_).sum(1)
        return c

    act_ = np.ones(Bn, bool)
    for it in range(1, s.sqp_max_iter + 1):
        idx = np.flatnonzero(act_)
        if idx.size == 0: break
        ua = u[idx]
        J0, g, Pc, xa, GH = linearize_trajectory(m, Q, P, Hc, ua.reshape(-1, H, nu), x0[idx], xref[idx], uref[idx])
        q = g - np.einsum("bij,bj->bi", Pc, ua)
        qp_failed = np.zeros(idx.size, bool)
        if eq or sb:
            Gs, los, his, rhos = [], [], [], []
            if sb:
                Gk = GH.reshape(idx.size, H * nx, nz)
                gu = np.einsum("bij,bj->bi", Gk, ua)
                xk = xa[:, 1:].reshape(idx.size, H * nx)
                Gs.append(Gk); los.append(np.tile(xmin_, H) - xk + gu); his.append(np.tile(xmax_, H) - xk + gu)
                rhos.append(rho / np.maximum((Gk ** 2).sum(2), 1e-12))
            if eq:
                GT = GH[:, H - 1]
                eH = xa[:, H] - xref[idx]
                bq = np.einsum("bij,bj->bi", GT, ua) - eH
                Gs.append(GT); los.append(bq); his.append(bq)
                n2 = (GT ** 2).sum(2)
                rhos.append(np.repeat((rho * nx / np.maximum(n2.sum(1), 1e-12))[:, None], nx, 1) if ball else rho_eq_scale * rho / np.maximum(n2, 1e-12))
            Gg = np.concatenate(Gs, 1); lo_g = np.concatenate(los, 1); hi_g = np.concatenate(his, 1); rho_g = np.concatenate(rhos, 1)
            K = Pc + (s.qp.sigma + rho) * np.eye(nz) + np.einsum("bia,bi,bic->bac", Gg, rho_g, Gg)
            Kinv = np.linalg.inv(K); Kinv = 0.5 * (Kinv + Kinv.transpose(0, 2, 1))
            v, yn, ygn, its, st_qp, _pr, dr = admm_box_gen_per_problem(Kinv, q, lb, ub, Gg, lo_g, hi_g, rho_g, s.qp, rho, ua, y[idx], yg[idx],
                                                                       nball=nx if ball else 0, rad=rad[idx])
            qp_failed = st_qp != mo.STATUS_SOLVED
            yg[idx] = ygn
            # exact-penalty weight: dual norm of the violation measure (l1 measure -> max norm; the ball's l2 distance -> l2 norm; state rows l1)
            if ball: mu[idx] = np.maximum(mu[idx], 1.1 * np.maximum(np.sqrt((ygn[:, -nx:] ** 2).sum(1)), np.abs(ygn[:, :-nx]).max(1, initial=0.0)))
            else: mu[idx] = np.maximum(mu[idx], 1.1 * np.abs(ygn).max(1))
            c0 = violation(xa)
        else:
            K = Pc + (s.qp.sigma + rho) * np.eye(nz)
            Kinv = np.linalg.inv(K); Kinv = 0.5 * (Kinv + Kinv.transpose(0, 2, 1))
            v, yn, its, _st, _pr, dr = admm_box_per_problem(Kinv, q, lb, ub, s.qp, rho, ua, y[idx])
            c0 = np.zeros(idx.size)
        d = v - ua
        gd = np.einsum("bi,bi->b", g, d) - mu[idx] * c0          # directional derivative of the l1 merit (the step zeroes the linearised rows)
        J0 = J0 + mu[idx] * c0
        stp = np.abs(d).max(1)
        small = stp <= s.sqp_tol                      # converged: take the full step, no line search
        small &= ~qp_failed
        t = np.ones(idx.size); ok = small | qp_failed; un = np.where(small[:, None], v, ua)
        for _ in range(s.ls_max + 1):
            todo = ~ok
            if not todo.any(): break
            cand = ua[todo] + t[todo, None] * d[todo]
            Jc, xc = objective(m, Q, P, Hc, cand.reshape(-1, H, nu), x0[idx][todo], xref[idx][todo], uref[idx][todo])
            if eq or sb: Jc = Jc + mu[idx][todo] * violation(xc, todo)
            good = Jc <= J0[todo] + s.ls_c1 * t[todo] * gd[todo] + s.ls_noise * np.maximum(1.0, np.abs(J0[todo]))
            tt = np.flatnonzero(todo)
            un[tt[good]] = cand[good]; ok[tt[good]] = True
            t[tt[~good]] *= 0.5
        u[idx] = np.where(ok[:, None], un, ua); y[idx] = yn
        sqp_iters[idx] = it; inner[idx] += its; step[idx] = stp; qp_dres[idx] = dr
        status[idx[small]] = mo.STATUS_SOLVED
        status[idx[~ok]] = STATUS_STALLED             # no Armijo decrease along the SQP direction (kink of a relu network)
        status[idx[qp_failed]] = mo.STATUS_PRIMAL_INF  # linearised terminal rows + input box not solvable within the inner cap
        act_[idx[small | ~ok | qp_failed]] = False
    uu = u.reshape(Bn, H, nu)
    J, x = objective(m, Q, P, Hc, uu, x0, xref, uref)
    return {"u": uu, "x": x, "e_x": x - xref[:, None, :], "e_u": uu - uref[:, None, :], "objective": J, "y": y, "y_terminal": yg, "status": status,
            "iters": sqp_iters, "inner_iters": inner, "step": step, "qp_dual_res": qp_dres}


def grad_adjoint(m, Q, P, Hc, u, x0, xref, uref):
    """J and dJ/du by the adjoint recursion (independent of the forward-mode Gamma accumulation used by the twin)."""
    Bn, H, nu = u.shape; nx = m.nx
    x = np.zeros((Bn, H + 1, nx)); x[:, 0] = x0
    As, Bs = [], []
    for k in range(H):
        f, A, Bm = jacobian(m, x[:, k], u[:, k]); x[:, k + 1] = f; As.append(A); Bs.append(Bm)
    e = x - xref[:, None, :]
    du = (u - uref[:, None, :]).reshape(Bn, -1)
    g = (du @ Hc.T).reshape(Bn, H, nu)
    lam = 2.0 * e[:, H] @ P.T
    for k in range(H - 1, -1, -1):
        g[:, k] += np.einsum("bij,bi->bj", Bs[k], lam)
        lam = np.einsum("bij,bi->bj", As[k], lam) + (2.0 * e[:, k] @ Q.T)
    J = 0.5 * np.einsum("bi,bi->b", du, du @ Hc.T) + np.einsum("bki,ij,bkj->b", e[:, :H], Q, e[:, :H]) + np.einsum("bi,ij,bj->b", e[:, H], P, e[:, H])
    return J, g.reshape(Bn, -1)


def kkt_residual(g, u, lb, ub, tol=1e-9):
    """First-order optimality of a box-constrained problem: the projected gradient (inf-norm)."""
    at_l = u <= lb + tol; at_u = u >= ub - tol
    pg = np.where(at_l & ~at_u, np.minimum(g, 0.0), np.where(at_u & ~at_l, np.maximum(g, 0.0), np.where(at_l & at_u, 0.0, g)))
    return np.abs(pg).max(-1)


def nmpc_local_opt(m, Q, R, S, P, H, umin, umax, x0, xref, uref, u_init=None):
    """Independent solve of ONE problem (Ipopt stand-in): scipy L-BFGS-B on the single-shooting NLP with the adjoint
    gradient, started from `u_init` (default: the reference input).  Returns u (H, nu), J and the KKT residual."""
    from scipy.optimize import minimize
    nu = m.nu; nz = nu * H
    x0 = np.asarray(x0, float)[None]; xref = np.asarray(xref, float)[None]; uref = np.asarray(uref, float)[None]
    Hc = constant_hessian(nu, H, np.asarray(R, float), np.asarray(S, float))
    lb = np.tile(np.asarray(umin, float), H); ub = np.tile(np.asarray(umax, float), H)

    def fg(v):
        J, g = grad_adjoint(m, Q, P, Hc, v.reshape(1, H, nu), x0, xref, uref)
        return float(J[0]), g[0]

    v0 = np.tile(np.clip(uref[0], umin, umax), H) if u_init is None else np.asarray(u_init, float).ravel()
    r = minimize(fg, v0, jac=True, method="L-BFGS-B", bounds=list(zip(lb, ub)), options={"maxiter": 5000, "ftol": 1e-16, "gtol": 1e-12, "maxcor": 40})
    J, g = fg(r.x)
    return r.x.reshape(H, nu), J, float(kkt_residual(g, r.x, lb, ub))


# --------------------------------------------------------------------------------------------------
# The reference's LINEAR method on a black-box model, one design per problem
# --------------------------------------------------------------------------------------------------
def relinearized_linear_mpc(m: NeuralModel, Q, R, S, H, umin, umax, x0, xref, uref, P=None, xmin=None, xmax=None,
                            state_constraint=False, terminal="none", eps=1e-10):
    """For every problem i: linearise the network at (xref_i, uref_i) (proceed_system_linearization, design_mpc.jl:319-323),
    P_i = are(A_i, B_i, Q, R) (design_mpc.jl:327) unless P is given, build the linear modeler's QP on that system
    (linear.jl:45-93 via mpc_oracle.condense) and solve it: exactly (active-set KKT) when the general rows are
    equalities or absent, with the ADMM twin at tight tolerance otherwise.  Checker for mpcb_solve_relinearized_batch.
    Returns dict of u, x, e_x, objective, P, A, B and `solved` (False where the problem is infeasible)."""
    x0 = np.atleast_2d(np.asarray(x0, float)); n = x0.shape[0]
    xref = np.broadcast_to(np.atleast_2d(np.asarray(xref, float)), x0.shape)
    uref = np.broadcast_to(np.atleast_2d(np.asarray(uref, float)), (n, m.nu))
    _, A, B = jacobian(m, xref, uref)
    out = {k: [] for k in ("u", "x", "e_x", "objective", "P", "solved")}
    for i in range(n):
        Pi = mo.dare(A[i], B[i], Q, R) if P is None else np.asarray(P, float)
        c = mo.condense(A[i], B[i], Q, R, S, Pi, H, umin, umax, xmin, xmax, state_constraint=state_constraint, terminal=terminal)
        p = mo.pack_params(x0[i], xref[i], uref[i])
        v = None
        if c.mg == 0 or c.eq_mask.all():
            try: v, _ = mo.qp_exact(c, p[0])
            except RuntimeError: v = None          # no KKT certificate: the equality rows are not reachable inside the input box
        if v is None:
            r = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=eps, eps_rel=eps, check_every=10, max_iter=100000))
            v = r["v"][0]; out["solved"].append(r["status"][0] == 1)
        else:
            out["solved"].append(True)
        rec = mo.recover(c, v, p)
        for k in ("u", "x", "e_x", "objective"): out[k].append(rec[k][0])
        out["P"].append(Pi)
    res = {k: np.array(v) for k, v in out.items()}
    res["A"] = A; res["B"] = B
    return res
