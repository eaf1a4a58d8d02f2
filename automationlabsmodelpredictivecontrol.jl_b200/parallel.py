"""Multi-GPU plumbing: the batch shards trivially (every (x0, reference) pair is an independent QP sharing read-only
per-system constants), so there is NO data-path collective during the solve; one process per
GPU solves its contiguous shard and a single gather of solutions + convergence statistics ends the step
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of problems owned by `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world <= 0 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_payload(u0, iters, status, prim_res, dual_res, objective):
    """One float64 row per problem: [u0..., iters, status, prim_res, dual_res, objective] -- a single contiguous buffer
    so the final gather is one collective."""
    import torch
    cols = [u0, iters.to(torch.float64).unsqueeze(1), status.to(torch.float64).unsqueeze(1), prim_res.unsqueeze(1),
            dual_res.unsqueeze(1), objective.unsqueeze(1)]
    return torch.cat(cols, dim=1).contiguous()


def unpack_payload(payload, nu: int):
    import torch
    return {"u0": payload[:, :nu], "iters": payload[:, nu].to(torch.int32), "status": payload[:, nu + 1].to(torch.int32),
            "prim_res": payload[:, nu + 2], "dual_res": payload[:, nu + 3], "objective": payload[:, nu + 4]}


def gather_to_rank0(payload, batch: int, group=None):
    """Gather per-rank payload shards (possibly of unequal length) on rank 0, in problem order.  Returns the full
    (batch, ncol) tensor on rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    ncol = payload.shape[1]
    buf = torch.zeros((nmax, ncol), dtype=payload.dtype, device=payload.device)
    buf[: payload.shape[0]] = payload
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
