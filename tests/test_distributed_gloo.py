"""World-size-2 gloo test of the N > 1 host logic of the one-process-per-GPU mode: shard ranges, the flat gather payload the
kernels write directly, the gather on rank 0 (`almpc_b200.parallel`, the same functions bench.py calls over NCCL).  The per-shard
"solve" is stood in by the oracle twin (there is no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import qt_batch


def test_shard_ranges(mpc):
    from almpc_b200.parallel import shard_range
    for batch in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, qtd, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import almpc_b200  # noqa: F401
    from almpc_b200.parallel import gather_payloads, payload_doubles, payload_views, shard_range
    from oracle import mpc_oracle as mo
    lo, hi = shard_range(n, rank, world)
    x0, xref, uref = qtd["x0"][lo:hi], qtd["xref"][lo:hi], qtd["uref"]
    P = mo.dare(qtd["A"], qtd["B"], qtd["Q"], qtd["R"])
    c = mo.condense(qtd["A"], qtd["B"], qtd["Q"], qtd["R"], np.zeros((2, 2)), P, 10, qtd["umin"], qtd["umax"])
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-5, eps_rel=1e-5, check_every=5))
    J = mo.recover(c, tw["v"], p)["objective"]
    # every rank's payload has the size of the LARGEST shard (equal-size gather); a rank fills the first hi - lo rows of each block
    nmax = max(b - a for a, b in (shard_range(n, r, world) for r in range(world)))
    payload = torch.zeros(payload_doubles(nmax, 2), dtype=torch.float64)
    pv = payload_views(payload, nmax, 2)
    m = hi - lo
    pv["u0"][:m] = torch.from_numpy(tw["v"][:, :2].copy()); pv["objective"][:m] = torch.from_numpy(J)
    pv["prim_res"][:m] = torch.from_numpy(tw["prim_res"]); pv["dual_res"][:m] = torch.from_numpy(tw["dual_res"])
    pv["status"][:m] = torch.from_numpy(tw["status"]); pv["iters"][:m] = torch.from_numpy(tw["iters"])
    gathered = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
    gather_payloads(payload, gathered, dst=0)
    if rank == 0:
        out = {}
        for r in range(world):
            a, b = shard_range(n, r, world)
            for k, v in payload_views(gathered[r], nmax, 2).items(): out.setdefault(k, []).append(v[: b - a].numpy().copy())
        ret.put({k: np.concatenate(v) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(qt, mpc):
    from oracle import mpc_oracle as mo
    n = 101                                  # odd: unequal shards
    x0, xref, uref = qt_batch(qt, n, seed=3)
    qtd = dict(qt, x0=x0, xref=xref, uref=uref)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, qtd, n, ret)) for r in range(2)]
    for p_ in procs: p_.start()
    got = ret.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], np.zeros((2, 2)), P, 10, qt["umin"], qt["umax"])
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-5, eps_rel=1e-5, check_every=5))
    assert np.array_equal(got["iters"], tw["iters"]) and np.array_equal(got["status"], tw["status"])
    assert np.abs(got["u0"] - tw["v"][:, :2]).max() < 1e-12        # problem order preserved across unequal shards
