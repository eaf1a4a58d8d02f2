import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a box without a GPU skips the gpu-marked tests instead of failing them.  On a box WITH a
    GPU nothing is skipped: a missing libmpcb200.so then fails loudly (there is no CPU fallback to hide behind)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device visible)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def qt():
    """Quadruple-tank linear model decoded from the reference fixture + the scenario of
    test/computation_mpc_test.jl:981-1040 (tests/golden/make_golden_qt_model.py)."""
    g = json.loads((ROOT / "tests" / "golden" / "qt_linear_model.json").read_text())
    sc = g["scenario"]
    return {"A": np.array(g["A"]), "B": np.array(g["B"]), "Q": sc["Q"] * np.eye(4), "R": sc["R"] * np.eye(2), "S": np.zeros((2, 2)),
            "xmin": np.array(sc["xmin"]), "xmax": np.array(sc["xmax"]), "umin": np.array(sc["umin"]), "umax": np.array(sc["umax"]),
            "x_ref": np.array(sc["x_ref"]), "u_ref": np.array(sc["u_ref"]), "x0": np.array(sc["x0"])}


def qt_batch(qt, n, seed=0):
    """BASELINE.md section 4, config 2 inputs."""
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4))
    xref = rng.uniform(0.4, 1.0, (n, 4))
    return x0, xref, qt["u_ref"].copy()


@pytest.fixture(scope="session")
def mpc():
    import almpc_b200
    return almpc_b200


def load_nn_fixture(name):
    """tests/golden/qt_fnn_model.json (decoded reference fixture) / qt_resnet_model.json (trained here): oracle model."""
    from oracle import nn_oracle as no
    g = json.loads((ROOT / "tests" / "golden" / name).read_text())
    return no.NeuralModel(g["arch"], g["activation"], np.array(g["W_in"]), [np.array(w) for w in g["W_h"]], [np.array(b) for b in g["b_h"]],
                          np.array(g["W_out"]))


@pytest.fixture(scope="session")
def fnn_model():
    return load_nn_fixture("qt_fnn_model.json")


@pytest.fixture(scope="session")
def resnet_model():
    return load_nn_fixture("qt_resnet_model.json")


def assert_matches_twin(res, tw, tight=1e-9, loose=1e-6, min_same=0.99, status_frac=1.0, check=5):
    """CUDA result vs the oracle twin on EVERY problem of the batch (not only on the subset whose iteration counts agree):
    same statuses (on at least `status_frac` of the batch), solutions to round-off where the iteration counts agree (`tight`), and
    to the accuracy both solves guarantee (`loose`) where a borderline termination check flipped on a 1e-16 difference -- those may
    differ by whole check periods, never by more than a few."""
    import numpy as np
    n = res["iters"].shape[0]
    v = res["u"].reshape(n, -1)
    st_eq = res["status"] == tw["status"]
    assert st_eq.mean() >= status_frac, ("statuses", st_eq.mean())
    same = (res["iters"] == tw["iters"]) & st_eq
    assert same.mean() >= min_same, ("same iteration count", same.mean())
    assert np.abs(v[same] - tw["v"][same]).max() < tight
    both = st_eq & (res["status"] == 1)
    if both.any():
        assert np.abs(v[both] - tw["v"][both]).max() < loose, np.abs(v[both] - tw["v"][both]).max()
        assert np.abs(res["iters"][both] - tw["iters"][both]).max() <= 4 * check
    return v, same


def exact_sample(n, frac=0.01, at_least=48, seed=0):
    """Random sample (>= 1 % of the batch) for the comparisons with the exact optimum (a per-problem active-set solve on the CPU)."""
    import numpy as np
    k = min(n, max(at_least, int(np.ceil(frac * n))))
    return np.sort(np.random.default_rng(seed).choice(n, k, replace=False))
