#!/usr/bin/env python
"""bench.py -- the driver contract for this repo.

    python bench.py --gpus N --steps K --warmup W            # our arm (libmpcb200 on B200)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port of OSQP, all host threads)

Workload = BASELINE.json configs[1]: quadruple-tank linear tracking MPC (nx=4, nu=2, H=20), batch of 65,536 random initial
states / references per GPU, cold start.  One "step" = one pass of the hot path (update_initialization! + calculate! for the
whole batch: condensed ADMM solve + result recovery).  Solver settings are the PARITY settings (eps_abs = eps_rel = 1e-7,
check every 5 iterations), under which tests/test_gpu_linear.py proves u0 within 1e-4, objective within 1e-6 and residuals
<= 1e-5 of the exact optimum.

  value : solves/s, inputs already resident in HBM, device-pointer C-ABI entry, CUDA events on the launching stream
  e2e   : solves/s through the host-array C-ABI entry with pinned HOST buffers (H2D + solve + recover + D2H of u, e_u, x,
          e_x, u0, objective, status, iters, residuals -- everything calculate! hands back) inside the timed region
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H = 20
BATCH = 65536
EPS = 1e-7
CHECK = 5
SIGMA = 0.0    # OSQP's sigma only regularises a semidefinite P; the condensed K = Pc + rho I is positive definite without it
NCU_DRAM_BYTES_PER_LAUNCH = 4257024   # profiles/r01/onchip_qt_h20_ncu_full.txt
FP64_PEAK_TFLOPS = 37.1   # DMMA.8x8x4 peak measured on this pool's B200 (profiles/micro/fp64_peak_r01.jsonl);
                          # MEASURED_PEAKS.json has no FP64 entry (cuBLAS DGEMM 8192^3 measured 35.5 in the same run)


def qt_model():
    g = json.loads((ROOT / "tests" / "golden" / "qt_linear_model.json").read_text())
    sc = g["scenario"]
    return (np.array(g["A"]), np.array(g["B"]), np.array(sc["xmin"]), np.array(sc["xmax"]), np.array(sc["umin"]), np.array(sc["umax"]),
            np.array(sc["x_ref"]), np.array(sc["u_ref"]), np.array(sc["x0"]))


def make_batch(n, seed):
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = qt_model()
    rng = np.random.default_rng(seed)
    return rng.uniform(xmin, xmax, (n, 4)), rng.uniform(0.4, 1.0, (n, 4)), u_ref.copy()


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7: continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"): reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def algorithmic_flops(info, iters, check_every):
    """SURVEY 8(d) / DESIGN.md section 5: per solve  it*(2 nt^2) [T r]  +  (it/check)*(2 nt^2) [termination pass]  +  2 nt np [q = Lq p]."""
    nt, npar = info.nt, 2 * info.nx + info.nu
    it = iters.astype(np.float64)
    return float((it * 2 * nt * nt + (it / check_every) * 2 * nt * nt + 2 * nt * npar).sum())


def executed_flops(info, iters, check_every):
    """What the condensed kernels actually execute: box-only problems get the dual residual in closed form (no termination pass);
    problems with general rows run a pass with C = [[Pc, G'], [G, 0]] only at the checks where some problem of the warp / row block can
    terminate (primal residual converged) -- counted here as ONE pass per problem, its last check: a lower bound of what ran."""
    nt, npar = info.nt, 2 * info.nx + info.nu
    it = iters.astype(np.float64)
    chk = 2.0 * nt * nt if info.mg > 0 else 0.0
    return float((it * 2 * nt * nt + chk + 2 * nt * npar).sum())


def kernel_name(info, sigma):
    """The dominant kernel of a controller, as ncu names it."""
    k = info.kernel
    if k == 1:
        has_g = info.mg > 0
        sig = has_g or sigma != 0.0
        minb = {8: 4, 16: 4, 24: 3 if has_g else 4, 32: 2 if has_g else 4, 40: 2 if has_g else 3, 48: 2 if has_g else 3}.get(info.nt_pad, 2 if (has_g or sig) else 3)
        return f"admm_onchip_kernel<{info.nt_pad},{str(has_g).lower()},{str(sig).lower()},{minb}>"
    if k == 3: return f"admm_smem{'g' if info.mg > 0 else ''}_kernel<{info.nt_pad},{str(sigma != 0.0).lower()},*>"
    if k == 4: return f"admm_riccati_kernel<{info.nx},{info.nu},{str(sigma != 0.0).lower()}>"
    return "stream_iter_kernel"


def run_ours(args):
    import torch
    import torch.distributed as dist
    import almpc_b200 as mpc
    from almpc_b200 import _lib

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gloo = dist.new_group(backend="gloo")          # host-side barrier for the phases in which only rank 0 drives the GPUs

    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = qt_model()
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    Cn = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200",
                                mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS, mpc_b200_check_every=CHECK, mpc_b200_sigma=SIGMA, mpc_b200_device=local)
    m = Cn.tuning.modeler
    info = m.info
    n = args.batch
    x0_h, xref_h, uref_h = make_batch(n, seed=rank)       # rng(0) on rank 0 == BASELINE.md config 2; other ranks: other shards

    # ---------------- device-resident leg ----------------
    f64 = dict(dtype=torch.float64, device=dev)
    x0 = torch.from_numpy(x0_h).to(dev); xref = torch.from_numpy(xref_h).to(dev); uref = torch.from_numpy(uref_h).to(dev)
    u = torch.empty((n, H, 2), **f64); e_u = torch.empty_like(u); x = torch.empty((n, H + 1, 4), **f64); e_x = torch.empty_like(x)
    # what the final gather carries lives in ONE flat buffer the kernels write directly (no packing pass):
    # [u0 n x 2 | objective n | prim_res n | dual_res n | status n (i32) iters n (i32)] = 6 n doubles
    from almpc_b200 import parallel
    payload = torch.empty(parallel.payload_doubles(n, 2), **f64)
    pv = parallel.payload_views(payload, n, 2)
    u0, obj, pres, dres, status, iters = pv["u0"], pv["objective"], pv["prim_res"], pv["dual_res"], pv["status"], pv["iters"]
    io = _lib.BatchIO()
    io.batch = n; io.x0 = x0.data_ptr(); io.xref = xref.data_ptr(); io.uref = uref.data_ptr(); io.xref_broadcast = 0; io.uref_broadcast = 1
    io.u = u.data_ptr(); io.e_u = e_u.data_ptr(); io.x = x.data_ptr(); io.e_x = e_x.data_ptr(); io.u0 = u0.data_ptr()
    io.objective = obj.data_ptr(); io.prim_res = pres.data_ptr(); io.dual_res = dres.data_ptr(); io.status = status.data_ptr(); io.iters = iters.data_ptr()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    # The one collective of the path: the gather of u0 + convergence stats on rank 0 (NCCL over NVLink).  It is double-buffered and
    # asynchronous: step k's gather runs on NCCL's stream while step k+1 solves into the other payload buffer, and step k+1's timed
    # interval ends only after the launching stream has waited for gather k -- so an interval is max(solve, what is left of the previous
    # gather), and the last gather is waited for inside the timed region too.
    payloads = [payload]; ios = [io]; gathered = [None, None]; works = [None, None]
    if world > 1:
        payload_b = torch.empty_like(payload)
        pvb = parallel.payload_views(payload_b, n, 2)
        io_b = _lib.BatchIO()
        for f, _t in io._fields_: setattr(io_b, f, getattr(io, f))
        io_b.u0 = pvb["u0"].data_ptr(); io_b.objective = pvb["objective"].data_ptr(); io_b.prim_res = pvb["prim_res"].data_ptr()
        io_b.dual_res = pvb["dual_res"].data_ptr(); io_b.status = pvb["status"].data_ptr(); io_b.iters = pvb["iters"].data_ptr()
        payloads.append(payload_b); ios.append(io_b)
        gathered = [[torch.empty_like(payload) for _ in range(world)] if rank == 0 else None for _ in range(2)]
    step_no = [0]

    def step():
        stream = torch.cuda.current_stream().cuda_stream
        b = step_no[0] & 1 if world > 1 else 0
        m.solve_batch_device(ios[b], stream)
        if world > 1:
            if works[b ^ 1] is not None: works[b ^ 1].wait()          # the launching stream waits for the previous step's gather
            works[b] = parallel.gather_payloads(payloads[b], gathered[b], dst=0, async_op=True)
        step_no[0] += 1

    def drain():
        for w in works:
            if w is not None: w.wait()

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the per-step events)
        evs[k][0].record()
        step()
        evs[k][1].record()
    tail[0].record(); drain(); tail[1].record()         # what is left of the last gather
    barrier()
    t_wall = time.perf_counter() - t_wall
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms_total = torch.tensor([sum(ms_steps) + tail[0].elapsed_time(tail[1])], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    clocks = sampler.stop() if sampler else None

    if world > 1:         # results of the last step live in the buffer it wrote
        pvl = parallel.payload_views(payloads[(step_no[0] - 1) & 1], n, 2)
        it_np = pvl["iters"].cpu().numpy(); st_np = pvl["status"].cpu().numpy()
    else:
        it_np = iters.cpu().numpy(); st_np = status.cpu().numpy()
    launches_full_step = int(m.timing()["kernel_launches"])     # kernels of libmpcb200 the last full step launched (counted by the library)
    # dominant kernel alone (solve kernel, no recover): time it live with events for the roofline
    io_solve = _lib.BatchIO()
    for f, _t in io._fields_: setattr(io_solve, f, getattr(io, f))
    for f in ("u", "e_u", "x", "e_x", "u0", "objective"): setattr(io_solve, f, None)
    torch.cuda.synchronize()
    kms = []
    for k in range(args.steps):
        flush.zero_()
        kev[k][0].record(); m.solve_batch_device(io_solve, torch.cuda.current_stream().cuda_stream); kev[k][1].record()
    torch.cuda.synchronize()
    kms = [a.elapsed_time(b) for a, b in kev]
    k_ms = sum(kms) / len(kms)
    flops = algorithmic_flops(info, it_np, CHECK)
    achieved_alg = flops / (k_ms * 1e-3) / 1e12
    flops_exec = executed_flops(info, it_np, CHECK)
    achieved = flops_exec / (k_ms * 1e-3) / 1e12

    # ---------------- end-to-end leg: host-array C ABI with pinned host buffers, every rank on its own shard concurrently ----------------
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
    hx0 = pin((n, 4)); hxr = pin((n, 4)); hx0.copy_(torch.from_numpy(x0_h)); hxr.copy_(torch.from_numpy(xref_h))
    out = {"u": pin((n, H, 2)).numpy(), "e_u": pin((n, H, 2)).numpy(), "x": pin((n, H + 1, 4)).numpy(), "e_x": pin((n, H + 1, 4)).numpy(),
           "u0": pin((n, 2)).numpy(), "objective": pin((n,)).numpy(), "prim_res": pin((n,)).numpy(), "dual_res": pin((n,)).numpy(),
           "status": pin((n,), torch.int32).numpy(), "iters": pin((n,), torch.int32).numpy()}
    h2d = hx0.numel() * 8 + hxr.numel() * 8 + uref_h.size * 8
    d2h = sum(v.nbytes for v in out.values())
    for _ in range(3):
        m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, out=out)
    e2e_sum = 0.0
    for _ in range(args.steps):
        flush.zero_(); barrier()
        t0 = time.perf_counter()
        m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, out=out)      # synchronous: returns with results on the host
        e2e_sum += time.perf_counter() - t0
    tim = m.timing()
    e2e_max = torch.tensor([e2e_sum], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    e2e_val = world * n * args.steps / float(e2e_max.item())          # whole job: all ranks' problems over the slowest rank's time
    assert np.array_equal(out["iters"], it_np)

    # ---------------- the same end-to-end call returning only what a closed loop applies (u0 + convergence stats, 48 B per problem) ----------------
    out0 = {k: out[k] for k in ("u0", "objective", "prim_res", "dual_res", "status", "iters")}
    for _ in range(3): m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, want=("u0", "objective"), out=dict(out0))
    e2e0_sum = 0.0
    for _ in range(args.steps):
        flush.zero_(); barrier()
        t0 = time.perf_counter()
        m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, want=("u0", "objective"), out=dict(out0))
        e2e0_sum += time.perf_counter() - t0
    e2e0_max = torch.tensor([e2e0_sum], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(e2e0_max, op=dist.ReduceOp.MAX)
    e2e_u0_val = world * n * args.steps / float(e2e0_max.item())

    # ---------------- pure device->host copy of one step's results (same bytes, same page-locked arrays, every rank at once): the PCIe ceiling of e2e ----------------
    d2h_src = torch.empty(d2h // 8, dtype=torch.float64, device=dev); d2h_dst = torch.empty(d2h // 8, dtype=torch.float64).pin_memory()
    for _ in range(2): d2h_dst.copy_(d2h_src, non_blocking=True); torch.cuda.synchronize()
    cp_sum = 0.0
    for _ in range(5):
        barrier()
        t0 = time.perf_counter(); d2h_dst.copy_(d2h_src, non_blocking=True); torch.cuda.synchronize(); cp_sum += time.perf_counter() - t0
    cp_max = torch.tensor([cp_sum / 5], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(cp_max, op=dist.ReduceOp.MAX)
    d2h_copy = {"bytes_per_rank": int(d2h), "ms_max_over_ranks": float(cp_max.item()) * 1e3, "GBps_per_rank": d2h / float(cp_max.item()) / 1e9,
                "GBps_aggregate": world * d2h / float(cp_max.item()) / 1e9,
                "e2e_ceiling_solves_per_s": world * n / float(cp_max.item()),
                "what": "torch copy_ of one step's result bytes device -> page-locked host, all ranks concurrently after a barrier, mean of 5, max over ranks"}
    del d2h_src, d2h_dst

    # ---------------- a stream of independent batches: two handles on two streams, consecutive steps overlap (the next batch ramps up while the
    # previous one drains and recovers); K steps timed as ONE interval, no flush: two alternating 137 MB working sets exceed the 126 MB L2 ----------------
    pipelined = None
    try:
        C2 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200",
                                    mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS, mpc_b200_check_every=CHECK, mpc_b200_sigma=SIGMA, mpc_b200_device=local)
        m2 = C2.tuning.modeler
        io2, t2 = _device_io(_lib, dev, n, 4, 2, H, x0_h, xref_h, uref_h)
        pstreams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        pair = [(m, io), (m2, io2)]

        def pstep(k):
            mm, ii = pair[k & 1]
            mm.solve_batch_device(ii, pstreams[k & 1].cuda_stream)
        torch.cuda.synchronize()
        for k in range(4): pstep(k)
        torch.cuda.synchronize(); barrier()
        pa = torch.cuda.Event(enable_timing=True); pb = torch.cuda.Event(enable_timing=True)
        pa.record()
        for s_ in pstreams: s_.wait_event(pa)
        for k in range(args.steps): pstep(k)
        for s_ in pstreams: torch.cuda.current_stream().wait_stream(s_)
        pb.record(); torch.cuda.synchronize()
        pms = torch.tensor([pa.elapsed_time(pb)], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        assert np.array_equal(t2["iters"].cpu().numpy(), it_np)
        pipelined = {"value": world * n * args.steps / (float(pms.item()) * 1e-3), "unit": "solves/s", "ms_per_step": float(pms.item()) / args.steps,
                     "what": "the same K steps issued alternately on two handles / two streams and timed as one interval (a stream of independent batches: the next "
                             "batch's solve ramps up while the previous one drains and its results are recovered); no gather; no L2 flush -- two alternating "
                             "137 MB working sets exceed the L2"}
        m2.close()
    except Exception as e:
        pipelined = {"error": repr(e)[:200]}

    # ---------------- strong scaling: configs[1]'s fixed 65 536 problems split over the ranks (device-timed, with the gather) ----------------
    strong = None
    if world > 1:
        ns = min(hi - lo for lo, hi in (parallel.shard_range(BATCH, r, world) for r in range(world)))
        io_s = _lib.BatchIO()
        for f, _t in io._fields_: setattr(io_s, f, getattr(io, f))
        io_s.batch = ns
        pay_s = payload[:6 * ns]        # (the first ns rows of each block are what this shard writes; the gather below moves the same 48 B per problem)
        gathered_s = [torch.empty_like(pay_s) for _ in range(world)] if rank == 0 else None

        def step_s():
            m.solve_batch_device(io_s, torch.cuda.current_stream().cuda_stream)
            parallel.gather_payloads(pay_s, gathered_s, dst=0)
        for _ in range(3): step_s()
        barrier()
        se = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush.zero_(); se[k][0].record(); step_s(); se[k][1].record()
        barrier()
        ms_s = torch.tensor([sum(a.elapsed_time(b) for a, b in se)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
        strong = {"total_batch": ns * world, "batch_per_gpu": ns, "ms_per_step": float(ms_s.item()) / args.steps, "value": ns * world * args.steps / (float(ms_s.item()) * 1e-3),
                  "unit": "solves/s", "what": "fixed total batch split over the ranks (strong scaling), device-timed, max over ranks"}

    # ---------------- ONE process driving all N GPUs through one multi-device handle (mpcb_settings.n_devices): rank 0 alone, the others wait on the host ----------------
    single = None
    if world > 1:
        dist.barrier(group=gloo)
        if rank == 0:
            try:
                single = single_process_multi_device(mpc, _lib, world, n, args.steps, flush)
            except Exception as e:      # the contract line must survive a failure of this extra leg
                single = {"error": repr(e)[:300]}
        dist.barrier(group=gloo)
    line = None
    if rank == 0:
        # ---------------- config 1: closed-loop single-solve latency (B = 1, warm start), p50 ----------------
        lat = closed_loop_latency(mpc, Cn)
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(sample=args.cpu_sample)
            lat["cpu_p50_us"] = cpu.pop("latency_p50_us")
        extra = None
        if world == 1 and not args.no_extra:
            extra = other_configs(mpc, dev, cpu=not args.no_cpu)
        line = {
            "metric": "MPC QP solves/sec", "value": world * n * args.steps / (ms_total * 1e-3), "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: quadruple-tank linear tracking MPC nx=4 nu=2 H=20, batch 65536 random x0/x_ref per GPU, cold start",
                       "batch_per_gpu": n, "eps_abs": EPS, "eps_rel": EPS, "check_every": CHECK, "sigma": SIGMA, "alpha": 1.6, "rho": info.rho, "kernel": "onchip-dmma",
                       "l2": "flushed between timed steps (256 MiB memset)", "parallelism": f"batch-shard x{world}, one NCCL gather of u0+objective+residuals+status+iters (48 B/problem) per step, double-buffered: it overlaps the next step's solve" if world > 1 else "single GPU",
                       "outputs": "u,e_u,x,e_x,u0,objective,status,iters,residuals"},
            "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "what": "mpcb_solve_linear_batch on page-locked host arrays, every rank on its shard concurrently, max over ranks",
                    "phases_ms": {k: round(v, 4) for k, v in tim.items() if k.endswith("_ms")},
                    "u0_only": {"value": e2e_u0_val, "unit": "solves/s", "d2h_bytes_per_step": int(sum(v.nbytes for v in out0.values())) * world,
                                "what": "same call returning u0 + objective + residuals + status + iters only (what a closed loop consumes)"},
                    "d2h_copy_ceiling": d2h_copy},
            "gpu_launches": int(launches_full_step * args.steps),
            "roofline": {"bound": "tensor", "kernel": kernel_name(info, SIGMA), "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (n == BATCH and info.kernel == 1 and info.nt_pad == 40) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, ncu --set full capture "
                                           "profiles/r01/onchip_qt_h20_ncu_full.txt (4.26 MB read, 0 written: outputs stay in L2 until recover reads them)",
                         "peak_source": "FP64 DMMA peak measured by profiles/micro/fp64_peak.cu on this pool (MEASURED_PEAKS.json has no FP64 number)",
                         "kernel_ms": k_ms, "executed_flops_per_launch": flops_exec, "algorithmic_flops_per_launch_8d": flops, "mean_iters": float(it_np.mean()),
                         "frac_algorithmic_8d": achieved_alg / FP64_PEAK_TFLOPS,
                         "note": "achieved / frac count the flops the kernel EXECUTES (it*2*nt^2 + 2*nt*np per solve); SURVEY 8(d)'s algorithmic count adds one "
                                 "termination pass per check that box-only problems do not run (closed-form dual residual): frac_algorithmic_8d"},
            "solver": {"mean_iters": float(it_np.mean()), "max_iters": int(it_np.max()), "solved_frac": float((st_np == 1).mean())},
            "latency": lat, "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["configs"] = extra
        if pipelined is not None:
            line["pipelined"] = pipelined
        if strong is not None:
            line["strong_scaling"] = strong
        if single is not None:
            line["single_process_multi_device"] = single
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def single_process_multi_device(mpc, _lib, ndev, n_per_dev, steps, flush):
    """a single host process (the Julia host of the north_star) drives all GPUs through ONE handle.  Device entry:
    buffers on device 0, shards fanned out and gathered back over NVLink peer copies inside the call.  Host entry: page-locked host
    arrays in, results out, one host thread + stream pair per device."""
    import torch
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = qt_model()
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    Cn = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS,
                                mpc_b200_check_every=CHECK, mpc_b200_sigma=SIGMA, mpc_b200_devices=list(range(ndev)))
    m = Cn.tuning.modeler
    n = ndev * n_per_dev
    x0_h = np.concatenate([make_batch(n_per_dev, seed=r)[0] for r in range(ndev)]); xref_h = np.concatenate([make_batch(n_per_dev, seed=r)[1] for r in range(ndev)])
    uref_h = make_batch(1, 0)[2]
    dev = torch.device("cuda", 0)
    io, t = _device_io(_lib, dev, n, 4, 2, H, x0_h, xref_h, uref_h)
    step = lambda: m.solve_batch_device(io, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_(); ev[k][0].record(); step(); ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    u0_only = _lib.BatchIO()
    for f, _t in io._fields_: setattr(u0_only, f, getattr(io, f))
    for f in ("u", "e_u", "x", "e_x"): setattr(u0_only, f, None)
    step0 = lambda: m.solve_batch_device(u0_only, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): step0()
    torch.cuda.synchronize()
    for k in range(steps):
        flush.zero_(); ev[k][0].record(); step0(); ev[k][1].record()
    torch.cuda.synchronize()
    ms0 = sum(a.elapsed_time(b) for a, b in ev) / steps
    st = t["status"].cpu().numpy()
    # host entry
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
    hx0 = pin((n, 4)); hxr = pin((n, 4)); hx0.copy_(torch.from_numpy(x0_h)); hxr.copy_(torch.from_numpy(xref_h))
    out = {"u": pin((n, H, 2)).numpy(), "e_u": pin((n, H, 2)).numpy(), "x": pin((n, H + 1, 4)).numpy(), "e_x": pin((n, H + 1, 4)).numpy(),
           "u0": pin((n, 2)).numpy(), "objective": pin((n,)).numpy(), "prim_res": pin((n,)).numpy(), "dual_res": pin((n,)).numpy(),
           "status": pin((n,), torch.int32).numpy(), "iters": pin((n,), torch.int32).numpy()}
    for _ in range(3): m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, out=out)
    t0 = time.perf_counter()
    for _ in range(steps): m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, out=out)
    e2e = n * steps / (time.perf_counter() - t0)
    out0 = {k: out[k] for k in ("u0", "objective", "prim_res", "dual_res", "status", "iters")}
    for _ in range(3): m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, want=("u0", "objective"), out=dict(out0))
    t0 = time.perf_counter()
    for _ in range(steps): m.solve_batch(hx0.numpy(), hxr.numpy(), uref_h, want=("u0", "objective"), out=dict(out0))
    e2e0 = n * steps / (time.perf_counter() - t0)
    res = {"devices": ndev, "batch": n, "device_entry": {"ms_per_step": ms, "value": n / ms * 1e3, "unit": "solves/s",
                                                          "what": "mpcb_solve_linear_batch_device on ONE multi-device handle: buffers on device 0, full outputs (2 KB/problem) gathered back "
                                                                  "over NVLink peer copies inside the call; CUDA events on device 0's stream"},
           "device_entry_u0_only": {"ms_per_step": ms0, "value": n / ms0 * 1e3, "unit": "solves/s", "what": "same, gathering u0 + stats only (48 B/problem)"},
           "host_entry_e2e": {"value": e2e, "unit": "solves/s", "what": "mpcb_solve_linear_batch on page-locked host arrays, one host thread + stream pair per device, full outputs"},
           "host_entry_e2e_u0_only": {"value": e2e0, "unit": "solves/s"}, "solved_frac": float((st == 1).mean())}
    m.close()
    return res


# ----------------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2], [3], [4] as driver-run records next to the configs[1] headline (N = 1): each at the PARITY settings its
# GPU tests prove (tests/test_gpu_at_size.py), device-timed with CUDA events, with its own roofline entry.
# ----------------------------------------------------------------------------------------------------------------------------
def _time_device(solve, reps, flush):
    import torch
    for _ in range(2): solve()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); solve(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def _device_io(_lib, dev, n, nx, nu, H, x0_h, xref_h, uref_h, full=True):
    import torch
    f64 = dict(dtype=torch.float64, device=dev)
    t = {"x0": torch.from_numpy(np.ascontiguousarray(x0_h)).to(dev), "xref": torch.from_numpy(np.ascontiguousarray(xref_h)).to(dev),
         "uref": torch.from_numpy(np.ascontiguousarray(uref_h)).to(dev), "status": torch.empty(n, dtype=torch.int32, device=dev),
         "iters": torch.empty(n, dtype=torch.int32, device=dev), "u0": torch.empty((n, nu), **f64), "objective": torch.empty(n, **f64),
         "prim_res": torch.empty(n, **f64), "dual_res": torch.empty(n, **f64)}
    if full:
        t.update({"u": torch.empty((n, H, nu), **f64), "e_u": torch.empty((n, H, nu), **f64), "x": torch.empty((n, H + 1, nx), **f64), "e_x": torch.empty((n, H + 1, nx), **f64)})
    io = _lib.BatchIO(); io.batch = n
    io.xref_broadcast = int(xref_h.ndim == 1); io.uref_broadcast = int(uref_h.ndim == 1)
    for k, v in t.items(): setattr(io, k, v.data_ptr())
    return io, t


def other_configs(mpc, dev, cpu=True):
    import torch
    from almpc_b200 import _lib
    out = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = lambda: torch.cuda.current_stream().cuda_stream
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = qt_model()

    # ---- configs[3]: horizon sweep, quadruple tank, batch 16 384, automatic kernel choice (on-chip -> shared-memory -> stage-wise)
    sweep = []
    n = 16384
    x0_h, xref_h, uref_h = make_batch(n, seed=0)
    for Hh in (10, 20, 30, 50, 75, 100, 150, 200):
        sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
        kw4 = dict(mpc_solver="b200", mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS, mpc_b200_check_every=CHECK, mpc_b200_sigma=SIGMA)
        io, t = _device_io(_lib, dev, n, 4, 2, Hh, x0_h, xref_h, uref_h)
        io_s = _lib.BatchIO()
        for f, _t in io._fields_: setattr(io_s, f, getattr(io, f))
        for f in ("u", "e_u", "x", "e_x", "u0", "objective"): setattr(io_s, f, None)
        # automatic step size sqrt(lmin lmax) first (what a caller gets without tuning), then the step size tuned on a 4096-problem sample
        Ca = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 5, list(x_ref), list(u_ref), **kw4)
        ma = Ca.tuning.modeler
        k_ms_auto = _time_device(lambda: ma.solve_batch_device(io_s, stream()), 2, flush)
        it_auto = float(t["iters"].cpu().numpy().mean()); rho_auto = ma.info.rho
        ma.close()
        Cn = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 5, list(x_ref), list(u_ref), mpc_b200_rho_tune=(x0_h[:4096], xref_h[:4096], uref_h, 7), **kw4)
        m = Cn.tuning.modeler; info = m.info
        ms = _time_device(lambda: m.solve_batch_device(io, stream()), 3, flush)
        k_ms = _time_device(lambda: m.solve_batch_device(io_s, stream()), 3, flush)
        it = t["iters"].cpu().numpy().astype(np.float64); st = t["status"].cpu().numpy()
        rec = {"H": Hh, "nz": info.nz, "kernel_id": info.kernel, "kernel": kernel_name(info, SIGMA), "ms": ms, "solves_per_s": n / ms * 1e3, "kernel_ms": k_ms,
               "mean_iters": float(it.mean()), "max_iters": int(it.max()), "solved_frac": float((st == 1).mean()),
               "step_size": {"rho": info.rho, "how": "mpcb_tune_rho, 4096-problem sample, 7 candidates rho_auto * 2^j (design time)", "automatic_rho": rho_auto,
                             "kernel_ms_with_automatic_rho": k_ms_auto, "mean_iters_with_automatic_rho": it_auto}}
        if info.kernel == 4:      # stage-wise kernel: bound by streaming its per-problem state (6 row passes of nz doubles per iteration)
            by = float(it.sum()) * 6 * info.nz * 8; fl = float(it.sum()) * Hh * 2 * (2 * 16 + 4 * 8 + 4)
            rec["roofline"] = {"bound": "hbm", "achieved": by / (k_ms * 1e-3) / 1e9, "peak": _hbm_peak(), "unit": "GB/s", "frac": by / (k_ms * 1e-3) / 1e9 / _hbm_peak(),
                               "fp64_tflops": fl / (k_ms * 1e-3) / 1e12, "fp64_frac": fl / (k_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                               "note": "algorithmic state bytes = iterations x 6 passes x nz x 8 B (w read twice + written, q read, d written + read); stage flops = iterations x H x 2 x 68"}
        else:
            fl = executed_flops(info, it, CHECK)
            rec["roofline"] = {"bound": "tensor", "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fl / (k_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS}
        rho_t = info.rho
        m.close()
        # the other kernel family at the same H and step size: the crossover is a driver-run number, not only a profiles/ file
        other = 4 if info.kernel != 4 else 2
        try:
            Co = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 5, list(x_ref), list(u_ref), mpc_b200_kernel=other, mpc_b200_rho=rho_t, **kw4)
            mo_ = Co.tuning.modeler
            ko_ms = _time_device(lambda: mo_.solve_batch_device(io_s, stream()), 2, flush)
            rec["other_family"] = {"kernel_id": other, "kernel": kernel_name(mo_.info, SIGMA), "kernel_ms": ko_ms,
                                   "what": "stage-wise (Riccati) kernel forced" if other == 4 else "condensed streamed DMMA GEMM kernel forced"}
            mo_.close()
        except Exception as e:
            rec["other_family"] = {"error": repr(e)[:200]}
        sweep.append(rec)
    out.append({"config": "configs[3]: quadruple-tank tracking MPC, horizon sweep H = 10..200, batch 16384 (rng(0) inputs of configs[1]), cold start",
                "eps_abs": EPS, "eps_rel": EPS, "check_every": CHECK, "sigma": SIGMA, "outputs": "u,e_u,x,e_x,u0,objective,status,iters,residuals (device resident)",
                "crossover": "condensed DMMA kernels (on chip) up to nz = 120 (H = 60), stage-wise (Riccati) kernel beyond; every H also carries the OTHER kernel family's time "
                             "(other_family): the curves cross between H = 50 and H = 75",
                "sweep": sweep})

    # ---- not a BASELINE config: the general-row workloads the shared-memory resident general-row kernel (admm_smemg.cuh) took over from the streamed path
    try:
        gen = []
        n = 16384
        rng7 = np.random.default_rng(7)
        for label, Hh, extra, x0g, xrg in (
                ("state box [0.55, 0.75] on x_1..x_H, H=20 (nt=120), rho ladder 300/x10", 20, dict(mpc_state_constraint=True, mpc_b200_ladder_iter=300, mpc_b200_max_iter=20000),
                 rng7.uniform(0.62, 0.72, (n, 4)), rng7.uniform(0.70, 0.82, (n, 4))),
                ("terminal equality, H=40 (nt=84), x0 within 0.0015 of the reference (all feasible)", 40, dict(mpc_terminal_ingredient="equality"),
                 np.tile(x_ref, (n, 1)) + 0.0015 * rng7.standard_normal((n, 4)), np.tile(x_ref, (n, 1)))):
            sb = "mpc_state_constraint" in extra
            sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(np.full(4, 0.55) if sb else xmin, np.full(4, 0.75) if sb else xmax), mpc.Hyperrectangle(umin, umax))
            Cg = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS,
                                        mpc_b200_check_every=CHECK, mpc_b200_sigma=SIGMA, **extra)
            mg_ = Cg.tuning.modeler
            io, t = _device_io(_lib, dev, n, 4, 2, Hh, x0g, xrg, np.asarray(u_ref))
            ms = _time_device(lambda: mg_.solve_batch_device(io, stream()), 3, flush)
            it = t["iters"].cpu().numpy().astype(np.float64); st = t["status"].cpu().numpy()
            fl = executed_flops(mg_.info, it, CHECK)
            # one problem through the host entry (the closed-loop use with a terminal / state constraint): the CTA-cooperative kernel takes small batches
            ts1 = []
            for _ in range(33):
                t0 = time.perf_counter(); mg_.solve_batch(x0g[:1], xrg[:1], np.asarray(u_ref), want=("u0",)); ts1.append((time.perf_counter() - t0) * 1e6)
            gen.append({"workload": label, "batch": n, "nt": mg_.info.nt, "kernel_id": mg_.info.kernel, "kernel": kernel_name(mg_.info, SIGMA), "ms": ms, "solves_per_s": n / ms * 1e3,
                        "one_problem_cold_start_p50_us": float(np.median(ts1[3:])), "one_problem_kernel": "admm_coop_kernel",
                        "mean_iters": float(it.mean()), "max_iters": int(it.max()), "solved_frac": float((st == 1).mean()),
                        "frac_fp64": fl / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS})
            mg_.close()
        out.append({"config": "general rows (not a BASELINE config): quadruple tank with state-box / terminal-equality rows at 64 < nt <= 120 -- on the streamed path until round 2 "
                              "(252 ms and 3.6 ms, profiles/r02/smemg_sbox20_v1.jsonl), now on the shared-memory resident general-row kernel; the state-box time is the iteration "
                              "tail of a few problems (one warp's latency), not throughput",
                    "eps_abs": EPS, "eps_rel": EPS, "check_every": CHECK, "sigma": SIGMA, "workloads": gen})
    except Exception as e:      # an extra record must never take the bench line down
        out.append({"config": "general rows", "error": repr(e)})

    # ---- configs[2]: random stable LTI nx = 64, nu = 16, H = 50, terminal LQR cost + terminal equality, batch 8 192
    rng = np.random.default_rng(1); nx, nu, Hh, n = 64, 16, 50, 8192
    G = rng.standard_normal((nx, nx)); A3 = 0.95 * G / np.abs(np.linalg.eigvals(G)).max(); B3 = rng.standard_normal((nx, nu)) / 8
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A3, B3, mpc.Hyperrectangle(-1e3 * np.ones(nx), 1e3 * np.ones(nx)), mpc.Hyperrectangle(-np.ones(nu), np.ones(nu)))
    x0_h = np.random.default_rng(3).standard_normal((n, nx)); xr = np.zeros(nx); ur = np.zeros(nu)
    kw3 = dict(mpc_solver="b200", mpc_terminal_ingredient="equality", mpc_b200_eps_abs=EPS, mpc_b200_eps_rel=EPS, mpc_b200_check_every=10, mpc_b200_sigma=0.0,
               mpc_b200_max_iter=4000)
    t0 = time.perf_counter()
    Ca = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 1, [0.0] * nx, [0.0] * nu, **kw3)        # automatic step size sqrt(lmin lmax)
    design_s = time.perf_counter() - t0
    io, t = _device_io(_lib, dev, n, nx, nu, Hh, x0_h, xr, ur)
    ma = Ca.tuning.modeler
    ms_auto = _time_device(lambda: ma.solve_batch_device(io, stream()), 2, flush)
    it_auto = float(t["iters"].cpu().numpy().mean()); rho_auto = ma.info.rho
    ma.close()
    t0 = time.perf_counter()
    Cn = mpc.proceed_controller(sys_, "model_predictive_control", Hh, 1, [0.0] * nx, [0.0] * nu, mpc_b200_rho_tune=(x0_h[:512], xr, ur), **kw3)   # tuned on a 512-problem sample
    tune_s = time.perf_counter() - t0
    m = Cn.tuning.modeler; info = m.info
    ms = _time_device(lambda: m.solve_batch_device(io, stream()), 3, flush)
    launches = int(m.timing()["kernel_launches"])
    it = t["iters"].cpu().numpy().astype(np.float64); st = t["status"].cpu().numpy()
    fl = executed_flops(info, it, 10)
    rec = {"config": "configs[2]: random stable LTI (rng(1)) nx=64 nu=16 H=50, input box [-1,1]^16, terminal LQR cost + terminal equality, batch 8192, x0 = N(0,I) (rng(3)), cold start",
           "eps_abs": EPS, "eps_rel": EPS, "check_every": 10, "sigma": 0.0, "nz": info.nz, "mg": info.mg, "nt_pad": info.nt_pad, "kernel_id": info.kernel, "kernel": kernel_name(info, 0.0),
           "design_s": design_s, "ms": ms, "solves_per_s": n / ms * 1e3, "gpu_launches": launches,
           "step_size": {"rho": info.rho, "how": "mpcb_tune_rho on a 512-problem sample of the batch, 7 candidates rho_auto * 2^j (design time, outside the timed region: "
                                               "the batch-wide counterpart of OSQP's per-problem adaptive rho)", "tune_s": tune_s, "candidates": m.rho_tuning,
                         "automatic_rho": rho_auto, "ms_with_automatic_rho": ms_auto, "mean_iters_with_automatic_rho": it_auto}, "mean_iters": float(it.mean()), "max_iters": int(it.max()),
           "solved_frac": float((st == 1).mean()), "infeasible": int((st == -3).sum()), "iteration_cap": int((st == -2).sum()),
           "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                        "note": "whole solve (all launches incl. checks, compaction, recovery); flops = iterations x 2 nt^2 + ONE pass with [[Pc,G'],[G,0]] per problem (check passes are skipped while no row of a block can terminate), nt = 864 (unpadded), "
                                "counted per problem up to ITS termination"}}
    if cpu:
        rec["cpu_baseline"] = _cpu_lti(A3, B3, nx, nu, Hh, Cn.tuning.terminal_ingredient.P, x0_h)
    out.append(rec)
    m.close()

    # ---- configs[4]: NMPC, ResNet surrogate of the quadruple tank, H = 20, batch 4 096 (SQP kernel in place of Ipopt)
    g = json.loads((ROOT / "tests" / "golden" / "qt_resnet_model.json").read_text())
    f = mpc.ResNet(np.array(g["W_in"]), [(np.array(w), np.array(b)) for w, b in zip(g["W_h"], g["b_h"])], np.array(g["W_out"]), activation=g["activation"])
    sys_ = mpc.ConstrainedBlackBoxControlDiscreteSystem(f, 4, 2, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    Cn = mpc.proceed_controller(sys_, "model_predictive_control", 20, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_programming_type="non_linear")
    mod = Cn.tuning.modeler
    n = 4096
    x0_h, xref_h, uref_h = make_batch(n, seed=0)
    io, t = _device_io(_lib, dev, n, 4, 2, 20, x0_h, xref_h, uref_h)
    inner = torch.empty(n, dtype=torch.int32, device=dev); io.inner_iters = inner.data_ptr()
    ms = _time_device(lambda: mod.solve_batch_device(io, stream()), 3, flush)
    it = t["iters"].cpu().numpy().astype(np.float64); inn = inner.cpu().numpy().astype(np.float64); st = t["status"].cpu().numpy()
    nz = 40
    fl = float((inn * 2 * nz * nz + it * (2 * nz * nz * 4 * 21 + 2 * nz ** 3 / 3 + 20 * 2 * (13 * 6 + 13 * 13 + 4 * 13) * (1 + 6))).sum())
    rec = {"config": "configs[4]: NMPC with the ResNet surrogate (6 -> 13 -> [13x13 residual relu] -> 4, tests/golden/qt_resnet_model.json), H=20, batch 4096, inputs of configs[1], cold start",
           "sqp_tol": 1e-6, "inner_eps_abs": 1e-9, "kernel": "nmpc_sqp_kernel", "ms": ms, "solves_per_s": n / ms * 1e3, "sqp_iters_mean": float(it.mean()), "sqp_iters_max": int(it.max()),
           "inner_admm_iters_mean": float(inn.mean()), "solved_frac": float((st == 1).mean()),
           "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                        "note": "useful flops = inner ADMM iterations x 2 nz^2 + SQP iterations x (K = Hc + 2 sum Gamma' W Gamma: 2 nz^2 nx (H+1); Cholesky-equivalent of the inverse nz^3 2/3; "
                                "network + forward-mode Jacobian per stage); the kernel is latency/issue bound, not FP64 bound (DESIGN.md section 5.4)"}}
    if cpu:
        rec["cpu_baseline"] = _cpu_nmpc(g, x0_h, xref_h, uref_h)
    out.append(rec)
    mod.close()
    return out


def _hbm_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        return 6650.0      # B200_PROFILING.md fallback


def _cpu_lti(A3, B3, nx, nu, Hh, P, x0_h, sample=None):
    """The reference's path for configs[2] restated: OSQP port on the reference's sparse encoding (12 192 variables), library defaults, one
    workspace per thread, bounded sample."""
    from oracle import mpc_oracle as mo, osqp_ref as orf
    ncores = os.cpu_count() or 1
    sample = sample or max(ncores, 16)
    qp = mo.build_reference_qp(A3, B3, 100 * np.eye(nx), 0.1 * np.eye(nu), np.zeros((nu, nu)), P, Hh, np.zeros(nx), np.zeros(nu), x0_h[0], -np.ones(nu), np.ones(nu),
                               terminal="equality")
    prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u)
    rows = np.concatenate([qp.x0_rows, qp.xref_rows, qp.uref_rows])
    sel = np.concatenate([qp.idx["u"].T.ravel(), qp.idx["x"].T.ravel()])
    vals = np.hstack([x0_h[:sample], np.zeros((sample, nx * (Hh + 1) + nu * Hh))])
    stg = orf.default_settings()
    orf.solve_batch(prob, stg, rows, vals[:ncores], sel, nthreads=ncores)
    t0 = time.perf_counter()
    r = orf.solve_batch(prob, stg, rows, vals, sel, cold_start=True, nthreads=ncores)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "solves/s", "cores": ncores, "kind": "port",
            "sample": f"{sample} problems, OSQP 0.6 defaults (eps 1e-3) on the reference's sparse model (n=12192), cold start; mean iters {float(r['iters'].mean()):.0f}; "
                      f"solved {float((r['status'] == 1).mean()):.2f}"}


def _cpu_nmpc(g, x0_h, xref_h, uref_h, sample=24):
    """Ipopt stand-in for configs[4]: the oracle's independent L-BFGS-B solve of the same NLP, one problem at a time on one core (no Ipopt / JuMP here)."""
    from oracle import mpc_oracle as mo, nn_oracle as no
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = qt_model()
    mdl = no.NeuralModel(g["arch"], g["activation"], np.array(g["W_in"]), [np.array(w) for w in g["W_h"]], [np.array(b) for b in g["b_h"]], np.array(g["W_out"]))
    Q = 100.0 * np.eye(4); R = 0.1 * np.eye(2); S = np.zeros((2, 2))
    _, Aj, Bj = no.jacobian(mdl, x_ref[None], u_ref[None]); P = mo.dare(Aj[0], Bj[0], Q, R)
    t0 = time.perf_counter()
    for i in range(sample):
        no.nmpc_local_opt(mdl, Q, R, S, P, 20, umin, umax, x0_h[i], xref_h[i], uref_h)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "solves/s", "cores": 1, "kind": "port", "sample": f"{sample} problems, scipy L-BFGS-B on the single-shooting NLP (Ipopt stand-in), one core"}


def closed_loop_latency(mpc, Cn, steps=200):
    """x0 = 0.6, 200 closed-loop steps of x+ = A x + B u0 in deviation form, warm start."""
    A, B, *_rest, x_ref, u_ref, x0 = qt_model()
    x = x0.copy(); ts = []
    for k in range(steps + 20):
        t0 = time.perf_counter()
        mpc.update_initialization(Cn, x)
        mpc.calculate(Cn, warm_start=True, want=("u", "e_u", "x", "e_x"))
        dt = time.perf_counter() - t0
        if k >= 20: ts.append(dt)
        u0 = Cn.computation_results.u[:, 0]
        x = x_ref + A @ (x - x_ref) + B @ (u0 - u_ref)
    out = {"p50_us": statistics.median(ts) * 1e6, "p99_us": float(np.percentile(ts, 99)) * 1e6, "steps": steps,
           "what": "update_initialization!+calculate! B=1 warm start through the Python host mirror, host wall clock incl. transfers"}
    # the same closed loop straight through the C ABI with a prebuilt mpcb_batch_io (what a Julia ccall pays: no Python
    # object churn between the measurement points except one ctypes call)
    from almpc_b200 import _lib
    m = Cn.tuning.modeler; i = m.info; L = _lib.lib()
    xb = np.array(x0, np.float64)[None].copy(); xr = np.array(x_ref, np.float64); ur = np.array(u_ref, np.float64)
    bufs = {"u": np.empty((1, i.horizon, i.nu)), "e_u": np.empty((1, i.horizon, i.nu)), "x": np.empty((1, i.horizon + 1, i.nx)),
            "e_x": np.empty((1, i.horizon + 1, i.nx)), "y": np.empty((1, i.nt)), "prim_res": np.empty(1), "dual_res": np.empty(1)}
    wu = np.zeros((1, i.nz)); wy = np.zeros((1, i.nt)); st = np.empty(1, np.int32); it = np.empty(1, np.int32)
    io = _lib.BatchIO(); io.batch = 1; io.x0 = xb.ctypes.data; io.xref = xr.ctypes.data; io.uref = ur.ctypes.data; io.xref_broadcast = 1; io.uref_broadcast = 1
    for k, v in bufs.items(): setattr(io, k, v.ctypes.data)
    io.status = st.ctypes.data; io.iters = it.ctypes.data
    ref = C.byref(io); h = m._h; f = L.mpcb_solve_linear_batch
    tc = []
    for k in range(steps + 20):
        if k == 1: io.warm_u = wu.ctypes.data; io.warm_y = wy.ctypes.data
        t0 = time.perf_counter()
        rc = f(h, ref)
        dt = time.perf_counter() - t0
        assert rc == 0 and st[0] == 1
        if k >= 20: tc.append(dt)
        wu[:] = bufs["u"].reshape(1, -1); wy[:] = bufs["y"]
        xb[0] = x_ref + A @ (xb[0] - x_ref) + B @ (bufs["u"][0, 0] - u_ref)
    out["c_abi_p50_us"] = statistics.median(tc) * 1e6; out["c_abi_p99_us"] = float(np.percentile(tc, 99)) * 1e6
    return out


def _reference_problem():
    from oracle import mpc_oracle as mo, osqp_ref as orf
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, x0 = qt_model()
    Q = 100.0 * np.eye(4); R = 0.1 * np.eye(2); S = np.zeros((2, 2)); P = mo.dare(A, B, Q, R)
    qp = mo.build_reference_qp(A, B, Q, R, S, P, H, x_ref, u_ref, x0, umin, umax)
    prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u)
    rows = np.concatenate([qp.x0_rows, qp.xref_rows, qp.uref_rows])
    sel = np.concatenate([qp.idx["u"].T.ravel(), qp.idx["x"].T.ravel()])
    return mo, orf, qp, prob, rows, sel


def _reference_vals(x0, xref, uref):
    n = x0.shape[0]
    return np.hstack([x0, np.tile(xref, (1, H + 1)), np.tile(np.broadcast_to(uref, (n, 2)), (1, H))])


def cpu_baseline(sample):
    """The reference's CPU path restated (oracle/osqp_ref.c: OSQP 0.6 defaults on the reference's sparse formulation, one
    persistent workspace per thread, cold start), timed on this box's host cores on a bounded sample of the same workload."""
    mo, orf, qp, prob, rows, sel = _reference_problem()
    x0, xref, uref = make_batch(sample, seed=0)
    st = orf.default_settings()
    vals = _reference_vals(x0, xref, uref)
    ncores = os.cpu_count() or 1                               # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    orf.solve_batch(prob, st, rows, vals[:256], sel, nthreads=ncores)          # warm the threads / page in
    t0 = time.perf_counter()
    r = orf.solve_batch(prob, st, rows, vals, sel, cold_start=True, nthreads=ncores)
    dt = time.perf_counter() - t0
    # single-solve latency: closed loop, warm start, persistent workspace (what one JuMP model + OSQP does)
    A, B, *_rest, x_ref, u_ref, xx0 = qt_model()
    w = orf.Workspace(prob, st); x = xx0.copy(); ts = []
    ucols = qp.idx["u"][:, 0]
    for k in range(220):
        t1 = time.perf_counter()
        w.update_bounds(qp.x0_rows, x); s = w.solve(cold_start=False)
        d = time.perf_counter() - t1
        if k >= 20: ts.append(d)
        x = x_ref + A @ (x - x_ref) + B @ (s["x"][ucols] - u_ref)
    return {"value": sample / dt, "unit": "solves/s", "cores": ncores, "kind": "port",
            "sample": f"{sample} problems of configs[1] (rng(0)), OSQP 0.6 defaults eps=1e-3 (the reference sets no solver attribute), "
                      f"cold start, OpenMP over problems; mean iters {float(r['iters'].mean()):.1f}; solved {float((r['status'] == 1).mean()):.3f}",
            "latency_p50_us": statistics.median(ts) * 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    mo, orf, qp, prob, rows, sel = _reference_problem()
    st = orf.default_settings()
    n = args.ref_sample
    x0, xref, uref = make_batch(n, seed=0)
    vals = _reference_vals(x0, xref, uref)
    ncores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        orf.solve_batch(prob, st, rows, vals[: max(256, n // 8)], sel, nthreads=ncores)
    t0 = time.perf_counter()
    its = []
    for _ in range(args.steps):
        r = orf.solve_batch(prob, st, rows, vals, sel, cold_start=True, nthreads=ncores); its.append(float(r["iters"].mean()))
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = (f"{n} problems of configs[1] per step (rng(0)); oracle port of OSQP 0.6 (oracle/osqp_ref.c) on the reference's sparse "
              f"formulation (n=372, m=412), library defaults eps=1e-3, cold start, one workspace per thread; mean iters {np.mean(its):.1f}")
    print(json.dumps({
        "impl": "reference", "metric": "MPC QP solves/sec", "value": val, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: quadruple-tank linear tracking MPC nx=4 nu=2 H=20, bounded sample of the 65536 batch", "batch_per_step": n},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--cpu-sample", type=int, default=32768)
    ap.add_argument("--ref-sample", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[0],[2],[3],[4] records (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
