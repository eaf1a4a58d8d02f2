"""World-size-2 gloo test of the N > 1 host logic: shard ranges, payload packing, ordered gather on rank 0.
The per-shard "solve" is stood in by the oracle twin (there is no GPU here); on the GPU box bench.py runs the same
plumbing over NCCL with libmpcb200 doing the solves."""
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
    from almpc_b200.parallel import gather_to_rank0, pack_payload, shard_range, unpack_payload
    from oracle import mpc_oracle as mo
    lo, hi = shard_range(n, rank, world)
    x0, xref, uref = qtd["x0"][lo:hi], qtd["xref"][lo:hi], qtd["uref"]
    P = mo.dare(qtd["A"], qtd["B"], qtd["Q"], qtd["R"])
    c = mo.condense(qtd["A"], qtd["B"], qtd["Q"], qtd["R"], np.zeros((2, 2)), P, 10, qtd["umin"], qtd["umax"])
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-5, eps_rel=1e-5, check_every=5))
    J = mo.recover(c, tw["v"], p)["objective"]
    t = torch.from_numpy
    payload = pack_payload(t(tw["v"][:, :2].copy()), t(tw["iters"]), t(tw["status"]), t(tw["prim_res"]), t(tw["dual_res"]), t(J))
    full = gather_to_rank0(payload, n)
    if rank == 0:
        out = unpack_payload(full, 2)
        ret.put({k: v.numpy() for k, v in out.items()})
    else:
        assert full is None
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
