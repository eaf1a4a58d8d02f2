"""Multi-GPU plumbing of the one-process-per-GPU mode: the batch shards trivially (every (x0, reference) pair is an independent QP
sharing read-only per-system constants), so there is NO data-path collective during the solve; every rank
solves its contiguous shard and a single gather of solutions + convergence statistics ends the step (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  bench.py and tests/test_distributed_gloo.py both go through these functions.  (The other mode --
one process driving several GPUs through one multi-device handle -- lives in the C ABI, csrc/multi_device.hpp.)"""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of problems owned by `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world <= 0 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def payload_doubles(n: int, nu: int) -> int:
    """Float64 slots of the flat per-shard gather payload: u0 (n x nu) | objective n | prim_res n | dual_res n | status n, iters n as int32."""
    return (nu + 4) * n


def payload_views(payload, n: int, nu: int):
    """Views into the flat float64 payload that the kernels write DIRECTLY (the io pointers of the device entry point at them), so the
    final gather needs no packing pass: 8 nu + 32 bytes per problem."""
    import torch
    u0 = payload[: nu * n].view(n, nu)
    obj = payload[nu * n:(nu + 1) * n]; pres = payload[(nu + 1) * n:(nu + 2) * n]; dres = payload[(nu + 2) * n:(nu + 3) * n]
    ints = payload[(nu + 3) * n:(nu + 4) * n].view(torch.int32)
    return {"u0": u0, "objective": obj, "prim_res": pres, "dual_res": dres, "status": ints[:n], "iters": ints[n:]}


def gather_payloads(payload, gathered, dst: int = 0, group=None, async_op: bool = False):
    """The one collective of the path: every rank's payload to rank `dst` (`gathered`: list of world tensors there, None elsewhere).
    async_op: returns the work handle -- bench.py double-buffers the payload so that step k's gather overlaps step k+1's solve."""
    import torch.distributed as dist
    return dist.gather(payload, gathered, dst=dst, group=group, async_op=async_op)
