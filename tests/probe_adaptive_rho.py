"""Experiment (not a test; run by hand: python tests/probe_adaptive_rho.py): would OSQP-style per-problem adaptive rho shorten the state-box tail?
Part 1: the stragglers of a capped first pass continued with (a) the fixed batch-wide rho, (b) OSQP's adaptive rule (rho <- rho sqrt(rp_rel / rd_rel) when the
ratio leaves [1/tol, tol]) with a per-problem refactorisation.  Part 2: the OSQP port itself (scaling + adaptive rho, reference formulation) on the same problems.
Output of the round-2 run: profiles/r02/experiment_adaptive_rho.txt."""
import json, numpy as np, sys, time
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from oracle import mpc_oracle as mo, osqp_ref as orf
g = json.load(open(str(__import__('pathlib').Path(__file__).resolve().parent / 'golden' / 'qt_linear_model.json'))); sc = g['scenario']
A, B = np.array(g['A']), np.array(g['B'])
Q = 100*np.eye(4); R = 0.1*np.eye(2); S = np.zeros((2,2)); P = mo.dare(A,B,Q,R)
xmin, xmax = np.full(4,0.55), np.full(4,0.75)
for H in (10, 20):
    c = mo.condense(A,B,Q,R,S,P,H,sc['umin'],sc['umax'],xmin,xmax,state_constraint=True)
    n = 4000
    rng = np.random.default_rng(7)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    p = mo.pack_params(x0, xref, np.array(sc['u_ref']))
    s = mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5, sigma=0.0, max_iter=300)
    t = time.time(); r1 = mo.admm_condensed(c, p, s); print("H", H, "first pass", time.time()-t, "unsolved", (r1["status"]!=1).sum(), "of", n)
    idx = np.flatnonzero(r1["status"] != 1)
    # per-problem adaptive rho on the stragglers (dense, one problem at a time), warm-started from the first pass
    nz, mg = c.nz, c.mg; nt = nz+mg
    rho0 = mo.auto_rho(c.Pc)
    base = np.concatenate([np.full(nz, rho0), rho0/np.maximum((c.G**2).sum(1),1e-12)])
    Ac = np.vstack([np.eye(nz), c.G])
    res = []
    for i in idx[:40]:
        q = c.Lq @ p[i]; b = c.Lb @ p[i]
        lo = np.concatenate([c.lb, c.lg + b]); hi = np.concatenate([c.ub, c.ug + b])
        x = r1["v"][i].copy(); z = Ac @ x; y = r1["y"][i].copy()
        scale = 1.0; alpha = 1.6
        def factor(scale):
            rv = base*scale
            K = c.Pc + np.diag(rv[:nz]) + c.G.T @ (rv[nz:,None]*c.G)
            return rv, np.linalg.inv(K)
        rv, Kinv = factor(scale); nfac = 1; last_adapt = 0
        for it in range(1, 20001):
            rhs = rv*z - y; rhs_x = rhs[:nz] - q + c.G.T @ rhs[nz:]
            xt = Kinv @ rhs_x; t_ = Ac @ xt
            w = alpha*t_ + (1-alpha)*z + y/rv
            zn = np.clip(w, lo, hi); y = rv*(w - zn); z = zn
            if it % 5 == 0:
                gvec = c.Pc @ xt + c.G.T @ y[nz:]
                rp = np.abs(t_ - z).max(); rd = np.abs(gvec + q + y[:nz]).max()
                nA = max(np.abs(t_).max(), np.abs(z).max()); nD = max(np.abs(gvec).max(), np.abs(y[:nz]).max(), np.abs(q).max())
                if rp <= 1e-7 + 1e-7*nA and rd <= 1e-7 + 1e-7*nD: break
                if it - last_adapt >= 50:
                    est = np.sqrt((rp/max(nA,1e-30)) / max(rd/max(nD,1e-30), 1e-30))
                    if est > 5 or est < 0.2:
                        scale *= est; scale = min(max(scale, 1e-4), 1e6)
                        rv, Kinv = factor(scale); nfac += 1; last_adapt = it
        res.append((it, nfac, scale))
    res = np.array(res)
    print(" adaptive on", len(res), "stragglers: iterations mean", res[:,0].mean(), "max", res[:,0].max(), "refactors mean", res[:,1].mean(), "final scale range", res[:,2].min(), res[:,2].max())
    s2 = mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5, sigma=0.0, max_iter=20000)
    r2 = mo.admm_condensed(c, p[idx[:40]], s2, v0=r1["v"][idx[:40]], y0=r1["y"][idx[:40]])
    print(" fixed rho continuing:        iterations mean", r2["iters"].mean(), "max", r2["iters"].max())

# ---- part 2
g = json.load(open(str(__import__('pathlib').Path(__file__).resolve().parent / 'golden' / 'qt_linear_model.json'))); sc = g['scenario']
A, B = np.array(g['A']), np.array(g['B'])
Q = 100*np.eye(4); R = 0.1*np.eye(2); S = np.zeros((2,2)); P = mo.dare(A,B,Q,R)
xmin, xmax = np.full(4,0.55), np.full(4,0.75)
H = 10
c = mo.condense(A,B,Q,R,S,P,H,sc['umin'],sc['umax'],xmin,xmax,state_constraint=True)
n = 4000
rng = np.random.default_rng(7)
xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4)); uref = np.array(sc['u_ref'])
p = mo.pack_params(x0, xref, uref)
s = mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5, sigma=0.0, max_iter=20000)
r1 = mo.admm_condensed(c, p, s)
order = np.argsort(-r1["iters"])[:30]
print("ours: worst iters", r1["iters"][order][:10], "mean all", r1["iters"].mean())
qp = mo.build_reference_qp(A,B,Q,R,S,P,H,xref[0],uref,x0[0],sc['umin'],sc['umax'],xmin,xmax,state_constraint=True)
prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u)
rows = np.concatenate([qp.x0_rows, qp.xref_rows, qp.uref_rows])
sel = np.concatenate([qp.idx["u"].T.ravel(), qp.idx["x"].T.ravel()])
vals = np.hstack([x0, np.tile(xref,(1,H+1)), np.tile(np.broadcast_to(uref,(n,2)),(1,H))])
for adaptive in (1, 0):
    st = orf.default_settings(eps_abs=1e-7, eps_rel=1e-7, max_iter=100000, adaptive_rho=adaptive)
    r = orf.solve_batch(prob, st, rows, vals, sel, cold_start=True, nthreads=8)
    print("OSQP adaptive", adaptive, ": mean", r["iters"].mean(), "max", r["iters"].max(), "on our worst 10:", r["iters"][order][:10], "its own worst:", np.sort(r["iters"])[-5:])
    if adaptive:
        w = orf.Workspace(prob, st)
        for i in order[:5]:
            w.update_bounds(rows, vals[i]); o = w.solve(cold_start=True); print("  problem", i, "iters", o["iters"], "rho updates", o["rho_updates"], "rho", o["rho"])
