"""Experiment (not a test; run by hand): would a SOFTENED relu derivative in the Gauss-Newton linearisation (slope clip(0.5 + h / 2 delta, 0, 1) inside a band
|h| < delta, exact relu in the rollout and the line search) reduce the kink stalls of the SQP on the reference relu FNN fixture?  Round-2 answer: no --
profiles/r02/experiment_relu_softened_derivative.txt (status 1: 0.914 exact, 0.914 / 0.902 / 0.812 for delta = 1e-3 / 1e-2 / 5e-2)."""
import numpy as np, sys, json, time
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1])); sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent))
from oracle import mpc_oracle as mo, nn_oracle as no
from conftest import load_nn_fixture
g = json.load(open(str(__import__('pathlib').Path(__file__).resolve().parent / 'golden' / 'qt_linear_model.json'))); sc = g['scenario']
qt = {"Q": sc["Q"]*np.eye(4), "R": sc["R"]*np.eye(2), "S": np.zeros((2,2)), "xmin": np.array(sc["xmin"]), "xmax": np.array(sc["xmax"]), "umin": np.array(sc["umin"]), "umax": np.array(sc["umax"]),
      "x_ref": np.array(sc["x_ref"]), "u_ref": np.array(sc["u_ref"])}
m = load_nn_fixture("qt_fnn_model.json")
H, n = 20, 256
rng = np.random.default_rng(0)
x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4)); xref = rng.uniform(0.4, 1.0, (n, 4)); uref = qt["u_ref"].copy()
_f, Jx, Ju = no.jacobian(m, qt["x_ref"][None], qt["u_ref"][None])
A, B = Jx[0], Ju[0]
P = mo.dare(A, B, qt["Q"], qt["R"])
c = mo.condense(A, B, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
rho = mo.auto_rho(c.Pc)
print("rho", rho)
orig_dact = no.dact
def run(label):
    t = time.time()
    r = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), rho)
    st = r["status"]
    print(label, "status1 %.3f stalled %.3f cap %.3f | sqp iters mean %.1f | %.0fs" % ((st==1).mean(), (st==2).mean(), (st==-2).mean(), r["iters"].mean(), time.time()-t))
    return r
r0 = run("exact derivative")
res = {"exact": r0}
for delta in (1e-3, 1e-2, 5e-2):
    def dact(name, h, delta=delta):
        if name == "relu": return np.clip(0.5 + h / (2*delta), 0.0, 1.0)
        return orig_dact(name, h)
    no.dact = dact
    res[delta] = run("band %g" % delta)
no.dact = orig_dact
J0 = r0["objective"]
for k, r in res.items():
    if k == "exact": continue
    rel = (r["objective"] - J0) / np.abs(J0)
    print("band", k, ": objective vs exact-derivative run: better by >1e-6 on %.3f, worse by >1e-6 on %.3f, median rel %.2e, worst %.2e" % ((rel < -1e-6).mean(), (rel > 1e-6).mean(), np.median(rel), rel.max()))
