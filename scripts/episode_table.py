"""Episode-reward table of DESIGN.md §5 (north_star criterion 3): discounted reward on RockSample(7,8) with a
SIR belief of 2^16 particles, 40-step episodes, the same environment draws for every arm.

    python scripts/episode_table.py --arm oracle --queries 1000000 --episodes 24 --procs 6 --out profiles/x.json
    python scripts/episode_table.py --arm gpu --queries 1000000 --episodes 24 [--inflight-max 256] --out ...

The oracle arm is pure CPU (one episode per process) and deterministic, so it can run anywhere; the GPU arm
needs the GPU box.  Episode e uses planner seed 300 + e and environment seed 4242 + e in both arms."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

STEPS, NPART = 40, 1 << 16


def one_episode(job):
    arm, Q, ep, extra = job
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from oracle import oracle as O
    from test_gpu_parity import _run_episodes
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    if arm == "gpu":
        def mk(e):
            return pj.solve(pj.POMCPSolver(c=20.0, rng=300 + e, tree_queries=Q, search_mode="batched", **extra), pomdp,
                            belief_mode=abi.BELIEF_EXTERNAL)
    else:
        def mk(e):
            return O.OraclePlanner(pomdp, O.default_params(c=20.0, seed=300 + e, tree_queries=Q,
                                                           belief_mode=abi.BELIEF_EXTERNAL, **extra))
    t = time.time()
    r = _run_episodes(O, pomdp, mk, 1, STEPS, 4242, arm == "gpu", NPART, first_ep=ep)
    return ep, float(r[0]), time.time() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["gpu", "oracle"], required=True)
    ap.add_argument("--queries", type=int, required=True)
    ap.add_argument("--episodes", type=int, default=24)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--first", type=int, default=0, help="index of the first episode (seeds 300 + e, 4242 + e)")
    ap.add_argument("--inflight-max", type=int, default=0)
    ap.add_argument("--rollouts-per-leaf", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    extra = {}
    if a.inflight_max:
        extra["inflight_max"] = a.inflight_max
    if a.rollouts_per_leaf:
        extra["rollouts_per_leaf"] = a.rollouts_per_leaf
    jobs = [(a.arm, a.queries, ep, extra) for ep in range(a.first, a.first + a.episodes)]
    t0 = time.time()
    if a.procs > 1 and a.arm == "oracle":
        with mp.get_context("spawn").Pool(a.procs) as pool:
            res = pool.map(one_episode, jobs, chunksize=1)
    else:
        res = [one_episode(j) for j in jobs]
    r = np.array([x[1] for x in sorted(res)])
    out = {"arm": a.arm, "queries": a.queries, "episodes": len(r), "first_episode": a.first, "steps": STEPS, "particles": NPART, "extra": extra,
           "rewards": r.tolist(), "mean": float(r.mean()), "se": float(r.std(ddof=1) / np.sqrt(len(r))),
           "wall_s": time.time() - t0, "episode_s": [x[2] for x in sorted(res)]}
    print(json.dumps({k: v for k, v in out.items() if k not in ("rewards", "episode_s")}), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
