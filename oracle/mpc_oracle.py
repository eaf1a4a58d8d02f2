"""CPU oracle for the MPC solve loop -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (AutomationLabsModelPredictiveControl.jl v0.1.4) is Julia, and the arithmetic of
its hot path lives in un-vendored third-party solvers (OSQP.jl "0.8" -> libosqp 0.6.x, Ipopt.jl "1",
ControlSystems.jl "1" `are`).  Neither Julia nor those solvers exist in the build container or on the GPU
box, and no reference test pins a numeric solution tighter than atol=0.5
(test/computation_mpc_test.jl:1053-1054).  This file therefore RESTATES
  * the reference's optimisation problem (variables / constraints / cost), and
  * the published OSQP ADMM algorithm (Stellato et al. 2020, OSQP 0.6 defaults),
and is pinned only by (i) the reference's structural known-answers (constraint counts 74/75/78,
test/terminal_ingredient_test.jl:160,237,317), (ii) the decoded quadruple-tank fixture (tests/golden), and
(iii) an independent exact KKT solve (`qp_exact`) that certifies optimality to 1e-9.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
package.  The product path (automationlabsmodelpredictivecontrol.jl_b200) never does.

All matrices here are numpy row-major [row, col]; "column k" of the reference's nx x (H+1) Julia matrices is
index [:, k] here as well (0-based: reference column 1 == index 0).
"""
from __future__ import annotations

import dataclasses
import numpy as np
import scipy.sparse as sp

OSQP_INFTY = 1e30


# --------------------------------------------------------------------------------------------------
# Terminal cost: P = are(Discrete, A, B, Q, R)                     (src/sub/design_mpc.jl:327)
# --------------------------------------------------------------------------------------------------
def dare(A, B, Q, R, tol=1e-13, max_iter=100000):
    """Discrete algebraic Riccati equation by plain value iteration (independent of the product's SDA
    doubling solver and of scipy; tests compare all three)."""
    A = np.asarray(A, float); B = np.asarray(B, float); Q = np.asarray(Q, float); R = np.asarray(R, float)
    P = Q.copy()
    for _ in range(max_iter):
        BtP = B.T @ P
        K = np.linalg.solve(R + BtP @ B, BtP @ A)
        Pn = Q + A.T @ P @ (A - B @ K)
        Pn = 0.5 * (Pn + Pn.T)
        if np.abs(Pn - P).max() <= tol * max(1.0, np.abs(Pn).max()):
            return Pn
        P = Pn
    raise RuntimeError("dare: value iteration did not converge")


# --------------------------------------------------------------------------------------------------
# The reference's QP in its own (sparse, redundant) encoding
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SparseQP:
    """min 1/2 v' P v + q' v  s.t.  l <= A v <= u   (OSQP form; P is the FULL symmetric matrix here)."""
    P: sp.csc_matrix
    q: np.ndarray
    A: sp.csc_matrix
    l: np.ndarray
    u: np.ndarray
    idx: dict            # variable name -> index array shaped like the reference's JuMP container
    rows: dict           # constraint group name -> row slice
    n_jump_constraints: int   # what JuMP.num_constraints-style counting of the tests sees (74/75/78)
    x0_rows: np.ndarray  # rows of the `JuMP.fix(x[i,1], x0[i])` equalities (computation_mpc.jl:23-27)
    xref_rows: np.ndarray
    uref_rows: np.ndarray


def build_reference_qp(A, B, Q, R, S, P, H, xref, uref, x0, umin, umax, xmin=None, xmax=None,
                       state_constraint=False, terminal="none") -> SparseQP:
    """Restates linear/mpc_modeler_implementation_linear.jl:34-100 (variables, constraints),
    design_mpc.jl:330-331 (terminal equality) and design_mpc.jl:405-465 (cost; NO 1/2 factor, stage
    index 1..H includes the fixed initial deviation, S-term only if S[1,1] != 0).

    Variable order = JuMP creation order (linear.jl:48-55): x, e_x, x_reference (nx x (H+1) each,
    column-major), u, e_u, u_reference (nu x H each) [, delta_u (nu x H) if S != 0, design_mpc.jl:423-427].
    Rows: AffExpr-EqualTo group first (dynamics, e_x def, e_u def, terminal, delta_u def), then the
    one-sided AffExpr-LessThan bounds, then the VariableRef-EqualTo `fix`es bridged to single-entry
    equality rows (OSQP's MOI wrapper accepts affine rows only) -- this matches the constraint-type
    list the tests assert (test/modeler_implementation_test.jl:86-105).
    """
    A = np.asarray(A, float); B = np.asarray(B, float)
    nx, nu = B.shape
    xref = np.asarray(xref, float).reshape(nx, -1); uref = np.asarray(uref, float).reshape(nu, -1)
    if xref.shape[1] == 1: xref = np.repeat(xref, H + 1, axis=1)       # main_mpc.jl:111
    if uref.shape[1] == 1: uref = np.repeat(uref, H, axis=1)           # main_mpc.jl:112
    use_S = float(np.asarray(S)[0, 0]) != 0.0
    use_R = float(np.asarray(R)[0, 0]) != 0.0

    off = 0
    idx = {}
    for name, r, c in (("x", nx, H + 1), ("e_x", nx, H + 1), ("x_reference", nx, H + 1),
                       ("u", nu, H), ("e_u", nu, H), ("u_reference", nu, H)):
        idx[name] = off + np.arange(r * c).reshape(c, r).T        # [i, k] -> column-major offset
        off += r * c
    if use_S:
        idx["delta_u"] = off + np.arange(nu * H).reshape(H, nu).T
        off += nu * H
    n = off

    ri, ci, vv, lo, up = [], [], [], [], []
    rows = {}
    r = 0

    def add_row(cols, vals, l_, u_):
        nonlocal r
        ri.extend([r] * len(cols)); ci.extend(cols); vv.extend(vals); lo.append(l_); up.append(u_)
        r += 1

    n_jump = 0
    # --- dynamics in deviation coordinates, NO affine term (linear.jl:58-60) ---
    s = r
    for k in range(H):
        for i in range(nx):
            cols = [idx["e_x"][i, k + 1]] + list(idx["e_x"][:, k]) + list(idx["e_u"][:, k])
            vals = [1.0] + list(-A[i, :]) + list(-B[i, :])
            add_row(cols, vals, 0.0, 0.0)
    rows["dynamics"] = slice(s, r); n_jump += r - s
    # --- deviation definitions (linear.jl:81-87) ---
    s = r
    for k in range(H + 1):
        for i in range(nx):
            add_row([idx["e_x"][i, k], idx["x"][i, k], idx["x_reference"][i, k]], [1.0, -1.0, 1.0], 0.0, 0.0)
    rows["e_x_def"] = slice(s, r); n_jump += r - s
    s = r
    for k in range(H):
        for i in range(nu):
            add_row([idx["e_u"][i, k], idx["u"][i, k], idx["u_reference"][i, k]], [1.0, -1.0, 1.0], 0.0, 0.0)
    rows["e_u_def"] = slice(s, r); n_jump += r - s
    # --- terminal ingredient (design_mpc.jl:330-340) ---
    s = r
    if terminal == "equality":
        for i in range(nx):
            add_row([idx["e_x"][i, H]], [1.0], 0.0, 0.0)
        n_jump += nx
    elif terminal == "contractive":
        n_jump += 1      # one quadratic constraint; not representable in a QP -> counted only
    rows["terminal"] = slice(s, r)
    # --- input-rate definition (design_mpc.jl:429-432): only i < H ---
    s = r
    if use_S:
        for k in range(H - 1):
            for i in range(nu):
                add_row([idx["delta_u"][i, k], idx["u"][i, k], idx["u"][i, k + 1]], [1.0, -1.0, 1.0], 0.0, 0.0)
        n_jump += r - s
    rows["delta_u_def"] = slice(s, r)
    # --- one-sided state bounds incl. the fixed initial column (linear.jl:62-70) ---
    s = r
    if state_constraint:
        for k in range(H + 1):
            for i in range(nx):
                add_row([idx["x"][i, k]], [1.0], -OSQP_INFTY, float(xmax[i]))
                add_row([idx["x"][i, k]], [-1.0], -OSQP_INFTY, -float(xmin[i]))   # lb - x <= 0
        n_jump += r - s
    rows["state_bounds"] = slice(s, r)
    # --- one-sided input bounds (linear.jl:73-78) ---
    s = r
    for k in range(H):
        for i in range(nu):
            add_row([idx["u"][i, k]], [1.0], -OSQP_INFTY, float(umax[i]))
            add_row([idx["u"][i, k]], [-1.0], -OSQP_INFTY, -float(umin[i]))
    rows["input_bounds"] = slice(s, r); n_jump += r - s
    # --- references fixed (linear.jl:90-100) and x[:,1] fixed (computation_mpc.jl:23-27) ---
    s = r
    for k in range(H + 1):
        for i in range(nx):
            add_row([idx["x_reference"][i, k]], [1.0], xref[i, k], xref[i, k])
    xref_rows = np.arange(s, r); s = r
    for k in range(H):
        for i in range(nu):
            add_row([idx["u_reference"][i, k]], [1.0], uref[i, k], uref[i, k])
    uref_rows = np.arange(s, r); s = r
    for i in range(nx):
        add_row([idx["x"][i, 0]], [1.0], float(x0[i]), float(x0[i]))
    x0_rows = np.arange(s, r)
    rows["fixes"] = slice(xref_rows[0], r)
    m = r

    Amat = sp.csc_matrix((vv, (ri, ci)), shape=(m, n))
    # --- cost (design_mpc.jl:436-465), OSQP's P = 2 * blkdiag(...) because JuMP has no 1/2 ---
    Pm = sp.lil_matrix((n, n))
    for k in range(H):     # stage cost over reference columns 1..H  (index 0..H-1)
        ix = idx["e_x"][:, k]; Pm[np.ix_(ix, ix)] = 2.0 * np.asarray(Q, float)
        if use_R:
            iu = idx["e_u"][:, k]; Pm[np.ix_(iu, iu)] = 2.0 * np.asarray(R, float)
    ix = idx["e_x"][:, H]; Pm[np.ix_(ix, ix)] = 2.0 * np.asarray(P, float)
    if use_S and use_R:    # the reference only adds the S term in the branch where R != 0 too (:436-447)
        for k in range(H - 1):
            idu = idx["delta_u"][:, k]; Pm[np.ix_(idu, idu)] = 2.0 * np.asarray(S, float)
    return SparseQP(sp.csc_matrix(Pm), np.zeros(n), Amat, np.array(lo, float), np.array(up, float), idx, rows,
                    n_jump, x0_rows, xref_rows, uref_rows)


# --------------------------------------------------------------------------------------------------
# Condensed restatement (what the GPU solves) -- SURVEY.md section 8(a) "condensed restatement"
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class CondensedQP:
    """Decision variable v = vec(u) in ABSOLUTE input coordinates (nz = nu*H, stage-major).

        min 1/2 v' Pc v + q(p)' v     s.t.  umin <= v <= umax (box rows, identity)
                                            lg + bg(p) <= G v <= ug + bg(p)   (general rows)
    with per-problem parameter p = [x0; xref; uref] (constant references over the horizon, as
    proceed_controller builds them, main_mpc.jl:105-117):   q(p) = Lq p,   bg(p) = Lb p.
    `cost_const(p)` restores the reference's objective value (terms that do not depend on v).
    """
    nx: int; nu: int; H: int
    Pc: np.ndarray; Lq: np.ndarray
    lb: np.ndarray; ub: np.ndarray
    G: np.ndarray; Lb: np.ndarray; lg: np.ndarray; ug: np.ndarray; eq_mask: np.ndarray
    Phi: np.ndarray      # (H+1, nx, nx)    e_k = Phi[k] e0 + Gam[k] (v - 1 (x) uref)
    Gam: np.ndarray      # (H+1, nx, nz)
    Qs: np.ndarray; Ps: np.ndarray; Rs: np.ndarray; Ss: np.ndarray
    A: np.ndarray; B: np.ndarray
    nball: int = 0       # terminal "contractive": the first nball general rows form ONE ball  |G v - b(p)|_2 <= sqrt(0.9) |x0 - xref|_2

    @property
    def nz(self): return self.nu * self.H
    @property
    def mg(self): return self.G.shape[0]


def condense(A, B, Q, R, S, P, H, umin, umax, xmin=None, xmax=None, state_constraint=False,
             terminal="none") -> CondensedQP:
    A = np.asarray(A, float); B = np.asarray(B, float)
    nx, nu = B.shape; nz = nu * H
    Q = np.asarray(Q, float); R = np.asarray(R, float); S = np.asarray(S, float); P = np.asarray(P, float)
    use_S = S[0, 0] != 0.0; use_R = R[0, 0] != 0.0
    Phi = np.zeros((H + 1, nx, nx)); Gam = np.zeros((H + 1, nx, nz))
    Phi[0] = np.eye(nx)
    for k in range(H):
        Phi[k + 1] = A @ Phi[k]
        Gam[k + 1] = A @ Gam[k]
        Gam[k + 1][:, k * nu:(k + 1) * nu] += B
    # J = e_H' P e_H + sum_{k<H} e_k' Q e_k + eps_k' R eps_k  (+ S-term),  eps = v - ubar
    Pc = np.zeros((nz, nz)); Fe = np.zeros((nz, nx))
    for k in range(H):
        Pc += Gam[k].T @ Q @ Gam[k]; Fe += Gam[k].T @ Q @ Phi[k]
    Pc += Gam[H].T @ P @ Gam[H]; Fe += Gam[H].T @ P @ Phi[H]
    if use_R:
        Pc += np.kron(np.eye(H), R)
    Pc_dev = 2.0 * Pc                      # Hessian w.r.t. eps (deviation inputs)
    Psum = Pc_dev.copy()
    if use_S and use_R:
        D = np.zeros((nu * (H - 1), nz))
        for k in range(H - 1):
            D[k * nu:(k + 1) * nu, k * nu:(k + 1) * nu] = np.eye(nu)
            D[k * nu:(k + 1) * nu, (k + 1) * nu:(k + 2) * nu] = -np.eye(nu)
        Psum = Psum + 2.0 * D.T @ np.kron(np.eye(H - 1), S) @ D      # delta_u = u_k - u_{k+1}: absolute == deviation
    Psum = 0.5 * (Psum + Psum.T)
    E1 = np.kron(np.ones((H, 1)), np.eye(nu))      # ubar = E1 uref
    # gradient wrt v at v: Psum v + 2 Fe e0 - Pc_dev ubar   (S-term is invariant to the constant shift)
    # p = [x0; xref; uref],  e0 = x0 - xref
    Lq = np.hstack([2.0 * Fe, -2.0 * Fe, -Pc_dev @ E1])
    lb = np.tile(np.asarray(umin, float), H); ub = np.tile(np.asarray(umax, float), H)
    Gs, Lbs, lgs, ugs, eqs = [], [], [], [], []
    nball = 0
    if terminal == "equality":      # e_H = 0  <=>  Gam_H v = Gam_H ubar - Phi_H e0
        Gs.append(Gam[H]); Lbs.append(np.hstack([-Phi[H], Phi[H], Gam[H] @ E1]))
        lgs.append(np.zeros(nx)); ugs.append(np.zeros(nx)); eqs.append(np.ones(nx, bool))
    elif terminal == "contractive":  # e_H' e_H <= 0.9 e_0' e_0  (design_mpc.jl:333-340, P_contract = I): e_H = Gam_H v - b(p), a ball
        Gs.append(Gam[H]); Lbs.append(np.hstack([-Phi[H], Phi[H], Gam[H] @ E1]))
        lgs.append(np.zeros(nx)); ugs.append(np.zeros(nx)); eqs.append(np.zeros(nx, bool)); nball = nx
    if state_constraint:            # xmin <= xref + e_k <= xmax for k = 1..H (k = 0 is the fixed x0: constant)
        for k in range(1, H + 1):
            Gs.append(Gam[k]); Lbs.append(np.hstack([-Phi[k], Phi[k] - np.eye(nx), Gam[k] @ E1]))
            lgs.append(np.asarray(xmin, float)); ugs.append(np.asarray(xmax, float)); eqs.append(np.zeros(nx, bool))
    if Gs:
        G = np.vstack(Gs); Lb = np.vstack(Lbs); lg = np.concatenate(lgs); ug = np.concatenate(ugs); eq = np.concatenate(eqs)
    else:
        G = np.zeros((0, nz)); Lb = np.zeros((0, 2 * nx + nu)); lg = np.zeros(0); ug = np.zeros(0); eq = np.zeros(0, bool)
    return CondensedQP(nx, nu, H, Psum, Lq, lb, ub, G, Lb, lg, ug, eq, Phi, Gam, Q, P, R, S, A, B, nball)


def pack_params(x0, xref, uref):
    """p = [x0; xref; uref] per problem, shape (Bn, 2nx+nu); 1-D references broadcast."""
    x0 = np.atleast_2d(np.asarray(x0, float)); Bn = x0.shape[0]
    xref = np.broadcast_to(np.atleast_2d(np.asarray(xref, float)), (Bn, x0.shape[1]))
    uref = np.atleast_2d(np.asarray(uref, float)); uref = np.broadcast_to(uref, (Bn, uref.shape[1]))
    return np.hstack([x0, xref, uref])


def recover(c: CondensedQP, v, p):
    """From absolute inputs v (Bn, nz) rebuild the reference's result matrices by rolling the deviation
    dynamics e_{k+1} = A e_k + B eps_k (linear.jl:59).  Returns dict of x, e_x (Bn, H+1, nx), u, e_u
    (Bn, H, nu) and the reference objective J (design_mpc.jl:449-456)."""
    nx, nu, H = c.nx, c.nu, c.H
    v = np.atleast_2d(v); p = np.atleast_2d(p); Bn = v.shape[0]
    x0, xref, uref = p[:, :nx], p[:, nx:2 * nx], p[:, 2 * nx:]
    u = v.reshape(Bn, H, nu); e_u = u - uref[:, None, :]
    e_x = np.zeros((Bn, H + 1, nx)); e_x[:, 0] = x0 - xref
    for k in range(H):
        e_x[:, k + 1] = e_x[:, k] @ c.A.T + e_u[:, k] @ c.B.T
    x = e_x + xref[:, None, :]
    J = np.einsum("bi,ij,bj->b", e_x[:, H], c.Ps, e_x[:, H])
    J += np.einsum("bki,ij,bkj->b", e_x[:, :H], c.Qs, e_x[:, :H])
    if c.Rs[0, 0] != 0.0:
        J += np.einsum("bki,ij,bkj->b", e_u, c.Rs, e_u)
        if c.Ss[0, 0] != 0.0:
            du = u[:, :-1] - u[:, 1:]
            J += np.einsum("bki,ij,bkj->b", du, c.Ss, du)
    return {"x": x, "e_x": e_x, "u": u, "e_u": e_u, "objective": J}


# --------------------------------------------------------------------------------------------------
# Condensed OSQP-style ADMM: the algorithmic twin of the CUDA kernels (same update order, same checks)
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class AdmmSettings:
    rho: float = 0.0          # <= 0: automatic  sqrt(lmin(Pc) lmax(Pc))
    rho_eq_scale: float = 1e3  # OSQP RHO_EQ_OVER_RHO_INEQ
    sigma: float = 1e-6
    alpha: float = 1.6
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    eps_prim_inf: float = 1e-4
    max_iter: int = 4000
    check_every: int = 25
    ineq_scale: float = 1.0   # multiplies the step size of the inequality general rows (second rung of the rho ladder)
    cold_init: int = 0        # cold start: 0 = OSQP's zeros (default), 1 = the clipped unconstrained optimum with a dual guess (cold_start_point)
    init_kappa: float = 2.0


STATUS_SOLVED, STATUS_MAX_ITER, STATUS_PRIMAL_INF = 1, -2, -3


def auto_rho(Pc):
    ev = np.linalg.eigvalsh(Pc)
    return float(np.sqrt(max(ev[0], 1e-12) * ev[-1]))


def admm_matrices(c: CondensedQP, s: AdmmSettings):
    """Stacked operator T = [I;G] K^-1 [I,G'] with K = Pc + sigma I + rho I + G' diag(rho_g) G and the check
    operator C = [[Pc, G'],[G, 0]]  (DESIGN.md section 3).  rho_g,i = rho (x rho_eq_scale on equality rows) / |G_i|^2: what
    OSQP's row equilibration of the constraint matrix amounts to for a per-row step size."""
    nz, mg = c.nz, c.mg
    rho = s.rho if s.rho > 0 else auto_rho(c.Pc)
    rho_g = np.where(c.eq_mask, s.rho_eq_scale * rho, s.ineq_scale * rho) / np.maximum((c.G ** 2).sum(1), 1e-12)   # row-equilibrated step sizes
    if c.nball:      # the projection onto a ball is closed-form only for one common step size on its rows
        rho_g[:c.nball] = rho / np.maximum((c.G[:c.nball] ** 2).sum(1).mean(), 1e-12)
    K = c.Pc + (s.sigma + rho) * np.eye(nz) + c.G.T @ (rho_g[:, None] * c.G)
    Kinv = np.linalg.inv(K); Kinv = 0.5 * (Kinv + Kinv.T)
    Ac = np.vstack([np.eye(nz), c.G])
    T = Ac @ Kinv @ Ac.T
    C = np.zeros((nz + mg, nz + mg)); C[:nz, :nz] = c.Pc; C[:nz, nz:] = c.G.T; C[nz:, :nz] = c.G
    rho_vec = np.concatenate([np.full(nz, rho), rho_g])
    return T, C, rho_vec, rho


def cold_start_point(c: CondensedQP, p, rho, kappa):
    """Cold-start iterate of settings.cold_init = 1 (csrc: the refill / init code of every ADMM kernel; the map Lv = -Pc^-1 Lq is a
    per-system constant from host_design.cpp).  x = clip(v_unc, lb, ub) with v_unc = Lv p the unconstrained optimum; dual guess on the
    box rows y = -kappa rho (x - v_unc): zero where the box is inactive, the sign of the true multiplier where it clips (the exact
    value would be -(Pc (x - v_unc)), a full operator pass; kappa rho is a one-number stand-in for Pc); general rows start at
    z = G x, y = 0 like an OSQP warm start.  It is a starting point only: the iteration, its fixed point and the termination test are
    unchanged."""
    p = np.atleast_2d(p)
    Lv = -np.linalg.solve(c.Pc, c.Lq)
    vunc = p @ Lv.T
    x = np.minimum(np.maximum(vunc, c.lb), c.ub)
    y = np.concatenate([-kappa * rho * (x - vunc), np.zeros((p.shape[0], c.mg))], axis=1)
    return x, y


def admm_condensed(c: CondensedQP, p, s: AdmmSettings, v0=None, y0=None):
    """Batched over problems (rows of p).  Mirrors OSQP's iteration (update_xz_tilde / update_x / update_z /
    update_y, osqp 0.6 `osqp_solve`) on the reduced KKT system with a cached operator.

    Termination (every `check_every` iterations; max_iter is rounded up to a multiple of it) applies OSQP's
    criteria  r <= eps_abs + eps_rel * max(...)  to the triple (x~, z+, y+) -- the exact minimiser x~ of the
    x-subproblem instead of OSQP's relaxed x -- because its dual residual is available from the cached factor
    without a second operator pass; x~ is also the returned solution.  Both sequences have
    the same limit.  Residuals:
        prim = |[x~; G x~] - z+|_inf                      norm: max(|[x~; G x~]|, |z+|)
        dual = |Pc x~ + G' y+_g + q + y+_box|_inf         norm: max(|Pc x~ + G' y+_g|, |y+_box|, |q|)
    """
    nz, mg = c.nz, c.mg; nt = nz + mg
    T, C, rho_vec, rho = admm_matrices(c, s)
    rinv = 1.0 / rho_vec
    p = np.atleast_2d(p); Bn = p.shape[0]
    q = p @ c.Lq.T
    b = p @ c.Lb.T if mg else np.zeros((Bn, 0))
    lo = np.concatenate([np.broadcast_to(c.lb, (Bn, nz)), c.lg + b], axis=1)
    hi = np.concatenate([np.broadcast_to(c.ub, (Bn, nz)), c.ug + b], axis=1)
    nb_ = c.nball
    if nb_:          # ball rows: centre b(p), radius sqrt(0.9) |x0 - xref|_2; no box on them
        rad = np.sqrt(0.9) * np.linalg.norm(p[:, :c.nx] - p[:, c.nx:2 * c.nx], axis=1)
        cen = b[:, :nb_]
        lo[:, nz:nz + nb_] = -OSQP_INFTY; hi[:, nz:nz + nb_] = OSQP_INFTY
    Ac = np.vstack([np.eye(nz), c.G])
    if v0 is None and s.cold_init:
        v0, y0 = cold_start_point(c, p, rho, s.init_kappa)
    if v0 is None:
        x = np.zeros((Bn, nz)); z = np.zeros((Bn, nt)); ys = np.zeros((Bn, nt))
    else:                                   # OSQP warm start: x = x0, z = A x0, y = y0
        x = np.array(v0, float).reshape(Bn, nz); z = x @ Ac.T; ys = np.array(y0, float).reshape(Bn, nt) * rinv
    max_iter = -(-s.max_iter // s.check_every) * s.check_every
    iters = np.zeros(Bn, np.int32); status = np.full(Bn, STATUS_MAX_ITER, np.int32)
    pres = np.zeros(Bn); dres = np.zeros(Bn)
    xo = np.zeros((Bn, nz)); yo = np.zeros((Bn, nt))
    active = np.ones(Bn, bool)
    qn = np.abs(q).max(1)
    for it in range(1, max_iter + 1):
        r = rho_vec * (z - ys)
        r[:, :nz] += s.sigma * x - q
        t = r @ T                      # T symmetric: [x~; z~_g] with z~_g = G x~
        w = s.alpha * t + (1 - s.alpha) * z + ys
        zn = np.minimum(np.maximum(w, lo), hi)
        if nb_:
            dev = w[:, nz:nz + nb_] - cen
            nrm = np.linalg.norm(dev, axis=1)
            scale = np.where(nrm > rad, rad / np.maximum(nrm, 1e-300), 1.0)
            zn[:, nz:nz + nb_] = cen + dev * scale[:, None]
        ysn = w - zn
        x = s.alpha * t[:, :nz] + (1 - s.alpha) * x
        dy = rho_vec * (ysn - ys)
        z, ys = zn, ysn
        if it % s.check_every == 0:
            y = rho_vec * ys
            g = t[:, :nz] @ c.Pc + y[:, nz:] @ c.G
            rp = np.abs(t - z).max(1)
            rd = np.abs(g + q + y[:, :nz]).max(1)
            ep = s.eps_abs + s.eps_rel * np.maximum(np.abs(t).max(1), np.abs(z).max(1))
            ed = s.eps_abs + s.eps_rel * np.maximum(np.maximum(np.abs(g).max(1), np.abs(y[:, :nz]).max(1)), qn)
            conv = (rp <= ep) & (rd <= ed)
            # OSQP primal infeasibility certificate on delta_y (osqp 0.6 is_primal_infeasible)
            ndy = np.abs(dy).max(1)
            supp = (hi * np.maximum(dy, 0)).sum(1) + (lo * np.minimum(dy, 0)).sum(1)
            atdy = np.abs(dy @ Ac).max(1)
            pinf = (ndy > s.eps_prim_inf) & (supp < -s.eps_prim_inf * ndy) & (atdy <= s.eps_prim_inf * ndy) & ~conv
            if mg == 0 or nb_: pinf[:] = False     # a non-empty box is always feasible; no certificate is evaluated for ball rows
            fin = active & (conv | pinf | (it >= max_iter))
            status[active & conv] = STATUS_SOLVED
            status[active & pinf] = STATUS_PRIMAL_INF
            pres[fin] = rp[fin]; dres[fin] = rd[fin]
            iters[fin] = it; xo[fin] = t[fin, :nz]; yo[fin] = y[fin]
            active = active & ~fin
            if not active.any():
                break
    return {"v": xo, "y": yo, "iters": iters, "status": status, "prim_res": pres, "dual_res": dres, "rho": rho}


# --------------------------------------------------------------------------------------------------
# Stage-wise (Riccati) form of the x-update: twin of csrc/admm_riccati.cu / host_design.cpp::riccati_factors
# --------------------------------------------------------------------------------------------------
def riccati_factors(c: CondensedQP, sigma, rho):
    """K x~ = r with K = Pc + (sigma + rho) I (box-only, no S term) is the optimality system of the LQ problem
        min sum_{k=1..H} 1/2 e_k' W_k e_k + sum_{k<H} (1/2 u_k' Rh u_k - r_k' u_k),  e_0 = 0,  e_{k+1} = A e_k + B u_k
    with W_k = 2Q (k < H), 2P (k = H), Rh = 2R + (sigma + rho) I -- the reference's own stage-wise structure
    (linear.jl:48-60) with the ADMM penalty folded into the input weight.  Returns per-stage (K_k, Lam_k^-1, Acl_k)."""
    assert c.mg == 0 and not (c.Ss[0, 0] != 0.0 and c.Rs[0, 0] != 0.0), "stage-wise form: box-only problems without the S term"
    A, B, H, nx, nu = c.A, c.B, c.H, c.nx, c.nu
    Rh = (2.0 * c.Rs if c.Rs[0, 0] != 0.0 else np.zeros((nu, nu))) + (sigma + rho) * np.eye(nu)
    Pi = 2.0 * c.Ps
    K = np.zeros((H, nu, nx)); Li = np.zeros((H, nu, nu)); Acl = np.zeros((H, nx, nx))
    for k in range(H - 1, -1, -1):
        Li[k] = np.linalg.inv(Rh + B.T @ Pi @ B)
        K[k] = Li[k] @ B.T @ Pi @ A
        Acl[k] = A - B @ K[k]
        Pi = (2.0 * c.Qs if k > 0 else 0.0) + A.T @ Pi @ Acl[k]
        Pi = 0.5 * (Pi + Pi.T)
    return K, Li, Acl


def riccati_apply(c: CondensedQP, fac, r):
    """x~ = K^-1 r for a batch of right-hand sides r (Bn, nz) by one backward and one forward sweep (the kernel's two sweeps)."""
    K, Li, Acl = fac
    H, nx, nu, B = c.H, c.nx, c.nu, c.B
    r = np.atleast_2d(r); Bn = r.shape[0]
    h = r.reshape(Bn, H, nu)
    pi = np.zeros((Bn, nx)); d = np.zeros((Bn, H, nu))
    for k in range(H - 1, -1, -1):
        d[:, k] = (h[:, k] + pi @ B) @ Li[k].T
        pi = pi @ Acl[k] - h[:, k] @ K[k]
    e = np.zeros((Bn, nx)); u = np.zeros((Bn, H, nu))
    for k in range(H):
        u[:, k] = d[:, k] - e @ K[k].T
        e = e @ Acl[k].T + d[:, k] @ B.T
    return u.reshape(Bn, H * nu)


# --------------------------------------------------------------------------------------------------
# Exact solve (ground truth): primal-dual active set + KKT certificate
# --------------------------------------------------------------------------------------------------
def admm_condensed_ladder(c: CondensedQP, p, s: AdmmSettings, ladder_iter, kappa=10.0):
    """Twin of the rho ladder (settings.ladder_iter / ladder_kappa, mpcb_api.cu): a first pass capped at ladder_iter iterations;
    the problems it leaves unsolved continue from their iterate (x, y) with the step size of the inequality general rows
    multiplied by kappa, for the remaining max_iter - ladder_iter iterations; their iteration counts continue from the cap."""
    ladder_iter = -(-ladder_iter // s.check_every) * s.check_every
    rho = s.rho if s.rho > 0 else auto_rho(c.Pc)
    r = admm_condensed(c, p, dataclasses.replace(s, rho=rho, max_iter=ladder_iter))
    idx = np.flatnonzero(r["status"] == STATUS_MAX_ITER)
    if idx.size:
        r2 = admm_condensed(c, np.atleast_2d(p)[idx], dataclasses.replace(s, rho=rho, max_iter=s.max_iter - ladder_iter, ineq_scale=float(kappa)),
                            v0=r["v"][idx], y0=r["y"][idx])
        for k in r:
            if isinstance(r[k], np.ndarray) and r[k].shape[:1] == r["status"].shape and k in r2: r[k][idx] = r2[k]
        r["iters"][idx] = r2["iters"] + ladder_iter
    r["second_rung"] = idx
    return r


def qp_exact(c: CondensedQP, p, v_init=None, tol=1e-9, max_pdas=60):
    """Exact optimum of ONE condensed problem (strictly convex => the KKT point is unique).  Handles box
    rows and EQUALITY general rows (terminal constraint); inequality general rows are not supported here.
    Returns (v, info) and raises if the KKT certificate cannot be established to `tol`."""
    nz, mg = c.nz, c.mg
    assert mg == 0 or c.eq_mask.all(), "qp_exact: only equality general rows supported"
    p = np.asarray(p, float).ravel()
    q = c.Lq @ p
    bg = (c.Lb @ p + c.lg) if mg else np.zeros(0)
    lb, ub, Pc, G = c.lb, c.ub, c.Pc, c.G
    if v_init is None:
        v_init = admm_condensed(c, p[None], AdmmSettings(eps_abs=1e-8, eps_rel=1e-8, check_every=10, max_iter=20000))["v"][0]
    v = np.array(v_init, float)

    def kkt_fixed(AU, AL):
        I = ~(AU | AL); fixed = ~I
        vv = np.where(AU, ub, np.where(AL, lb, 0.0))
        ni = int(I.sum())
        KK = np.zeros((ni + mg, ni + mg)); KK[:ni, :ni] = Pc[np.ix_(I, I)]
        rhs = np.zeros(ni + mg); rhs[:ni] = -(q[I] + Pc[np.ix_(I, fixed)] @ vv[fixed])
        if mg:
            KK[:ni, ni:] = G[:, I].T; KK[ni:, :ni] = G[:, I]
            rhs[ni:] = bg - G[:, fixed] @ vv[fixed]
        sol = np.linalg.lstsq(KK, rhs, rcond=None)[0] if mg else np.linalg.solve(KK, rhs)
        vv[I] = sol[:ni]; nu_ = sol[ni:]
        mu = -(Pc @ vv + q + (G.T @ nu_ if mg else 0.0)); mu[I] = 0.0     # multiplier of the box rows
        return vv, mu, nu_

    nu_ = np.zeros(mg)
    if mg:
        nu_ = np.linalg.lstsq(G.T, -(Pc @ v + q), rcond=None)[0] * 0.0
    mu = -(Pc @ v + q)
    seen = set()
    for it in range(max_pdas):
        AU = (mu + (v - ub)) > 0
        AL = (mu + (v - lb)) < 0
        key = (AU.tobytes(), AL.tobytes())
        if key in seen:
            break
        seen.add(key)
        v, mu, nu_ = kkt_fixed(AU, AL)
    # certificate
    g = Pc @ v + q + (G.T @ nu_ if mg else 0.0)
    feas = max(np.max(lb - v, initial=0.0), np.max(v - ub, initial=0.0))
    eqres = np.abs(G @ v - bg).max(initial=0.0) if mg else 0.0
    at_l = v <= lb + 1e-12; at_u = v >= ub - 1e-12; free = ~(at_l | at_u)
    stat = max(np.abs(g[free]).max(initial=0.0), np.max(-g[at_l & ~at_u], initial=0.0), np.max(g[at_u & ~at_l], initial=0.0))
    scale = max(1.0, np.abs(q).max())
    ok = feas <= tol and eqres <= tol * max(1.0, np.abs(bg).max(initial=0.0)) and stat <= tol * scale
    info = {"pdas_iters": it + 1, "feas": feas, "eqres": eqres, "stationarity": stat, "ok": bool(ok),
            "n_active": int((at_l | at_u).sum()), "nu": nu_, "mu": -g}
    if not ok:
        raise RuntimeError(f"qp_exact: KKT certificate failed {info}")
    return v, info


def u0_metric(u0, u0_star, umin, umax):
    """||u0 - u0*||_inf / max(||u0*||_inf, ||umax - umin||_inf)   (plain relative error is
    ill-defined because many optimal u0 sit on the bound 0)."""
    scale = np.maximum(np.abs(u0_star).max(-1), np.abs(np.asarray(umax) - np.asarray(umin)).max())
    return np.abs(u0 - u0_star).max(-1) / scale
