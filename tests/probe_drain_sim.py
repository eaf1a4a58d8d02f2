"""Experiment (not a test; run by hand): slot-level model of the on-chip kernel's end-of-launch drain on one B200 (148 SMs x 3 CTAs x 4 warps x 8 slots, one warp of
every CTA per scheduler), with the measured per-scheduler speeds (1 / 2 / 3 active warps: 0.57 / 0.76 / 0.80 of the pipe) and the iteration
histogram of the headline batch: would compacting the active slots of a CTA into fewer warps after the queue is empty shorten the drain?
Round-2 answer (profiles/r02/experiment_drain_compaction_sim.txt): launch 972 -> 936 time units (-3.8 %), against 833 for a drain-free launch -- not worth
moving register fragments between warps."""
import numpy as np, sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
rng = np.random.default_rng(0)
# iteration histogram of the headline batch (cold_init = 0): from the twin on 4096 problems
from oracle import mpc_oracle as mo
import bench
A, B, xmin, xmax, umin, umax, x_ref, u_ref, x0_ = bench.qt_model()
Q = 100*np.eye(4); R = 0.1*np.eye(2); S = np.zeros((2,2)); P = mo.dare(A,B,Q,R)
c = mo.condense(A,B,Q,R,S,P,20,umin,umax)
x0, xref, uref = bench.make_batch(8192, 0); p = mo.pack_params(x0, xref, uref)
its = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5, sigma=0.0, cold_init=0))["iters"] // 5   # in check periods
print("mean periods", its.mean(), "max", its.max())
# per-iteration time of a warp as a function of the number of ACTIVE warps on its scheduler (units of P = pipe time per warp-iteration)
RT = {0: 0.0, 1: 1.75, 2: 2.63/1.0, 3: 3.75}     # round time: all warps on the scheduler complete one iteration in RT[k]
def simulate(policy, nprob=65536, nsm=148, ctas=3):
    # global queue shared by all SMs; event-driven per scheduler is complex -> approximate: time advances per SM in "periods"; each scheduler's
    # period takes RT[k] * 5; warps on different schedulers drift, which we ignore by using the SM's mean round time (pipe is per scheduler, so use per-scheduler clocks)
    q = list(rng.choice(its, nprob))
    qi = 0
    # state: per SM, per CTA, per warp(=scheduler id), 8 slots remaining periods
    rem = np.zeros((nsm, ctas, 4, 8), int)
    clock = np.zeros((nsm, 4))          # per-scheduler time
    # event loop: pick the scheduler with the smallest clock, advance it one period
    import heapq
    heap = [(0.0, s, k) for s in range(nsm) for k in range(4)]
    heapq.heapify(heap)
    tend = 0.0
    qarr = np.array(q); n = len(qarr)
    while heap:
        t, s, k = heapq.heappop(heap)
        # refill empty slots of the warps on this scheduler
        for cta in range(ctas):
            for sl in range(8):
                if rem[s, cta, k, sl] == 0 and qi < n:
                    rem[s, cta, k, sl] = qarr[qi]; qi += 1
        if policy == "compact" and qi >= n:
            # within each CTA, pack active slots into the fewest warps; CTA c vacates schedulers starting from (c) rotation
            for cta in range(ctas):
                act = rem[s, cta][rem[s, cta] > 0]
                order = [(cta + j) % 4 for j in range(4)]      # warp fill order rotated per CTA so that CTAs vacate different schedulers
                new = np.zeros((4, 8), int)
                for j, v in enumerate(act): new[order[j // 8], j % 8] = v
                rem[s, cta] = new
        nact = sum(1 for cta in range(ctas) if (rem[s, cta, k] > 0).any())
        if nact == 0:
            if policy == "compact" and qi >= n and (rem[s] > 0).any():
                heapq.heappush(heap, (t + 0.5, s, k))      # idle scheduler waits (may receive slots at the next compaction)
            continue
        dt = RT[nact] * 5
        for cta in range(ctas):
            m = rem[s, cta, k] > 0
            rem[s, cta, k][m] -= 1
        tend = max(tend, t + dt)
        heapq.heappush(heap, (t + dt, s, k))
    return tend
base = simulate("none"); comp = simulate("compact")
ideal = its.mean() * 5 * 65536 / (148 * 12 * 8) * 3.75 / 1.0
print("none", base, "compact", comp, "ratio", comp / base, "ideal steady", ideal)
