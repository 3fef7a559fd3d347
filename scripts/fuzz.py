"""Developer tool (GPU box): random solver / updater configurations against the invariants every search and update
must keep - a wider, longer sweep than tests/test_gpu_fuzz.py (which keeps a seeded slice of it in the suite).

    python scripts/fuzz.py [--cases 300] [--seed 0]

Per case: a problem, POMCPSolver or POMCPDPWSolver with random knobs, then a short episode - search, environment
step, belief update (unweighted re-root with or without reinvigoration, or SIR on an external particle set), search
again.  Checked: the counters add up (simulations + terminal skips = queries x trees of every search), the root holds
its visits, values of visited actions are finite, the only exceptions are the reference's own (depletion, all samples
terminal).  Prints one line per failure and a summary; exit code 1 on any failure."""
import argparse
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pomcp_jl_b200 as pj  # noqa: E402
from pomcp_jl_b200 import _abi as abi  # noqa: E402


def one_case(rng, idx):
    problems = [("tiger", pj.TigerPOMDP(), 100.0), ("baby", pj.BabyPOMDP(-5.0, -10.0), 10.0),
                ("rs44", pj.RockSamplePOMDP(4, 4, seed=44), 10.0), ("rs78", pj.RockSamplePOMDP(7, 8, seed=7008), 20.0),
                ("lightdark", pj.LightDark1D(), 1.0)]
    name, pomdp, c = problems[int(rng.integers(0, len(problems)))]
    mode = str(rng.choice(["external", "unweighted", "unweighted", "sir"]))
    if name == "lightdark" and mode == "unweighted":
        mode = "sir"  # continuous observations never re-visit a child: the reference pairs them with an external updater
    T = int(rng.choice([1, 1, 1, 2, 4])) if mode in ("external", "sir") else 1
    Q = int(rng.choice([1, 5, 64, 300, 3000, 30000]))
    kw = dict(c=c, rng=int(rng.integers(1, 1 << 30)), tree_queries=Q, n_trees=T,
              eps=float(rng.choice([0.01, 0.2, 0.9])), max_depth=int(rng.choice([1, 2, 5, 2 ** 63 - 1])),
              init_N=int(rng.choice([0, 0, 3])), init_V=float(rng.choice([0.0, -5.0, 7.5])),
              inflight_max=int(rng.choice([0, 0, 1, 3, 64, 1000])), inflight_min=int(rng.choice([0, 0, 1, 50])),
              inflight_ramp_div=int(rng.choice([0, 0, 1, 64])),
              rollouts_per_leaf=int(rng.choice([0, 0, 1, 31, 64, 128])))
    if name == "baby" and rng.random() < 0.3:
        kw["estimate_value"] = pj.RolloutEstimator(pj.FeedWhenCrying())
    elif rng.random() < 0.15:
        kw["estimate_value"] = float(rng.normal())
    dpw = rng.random() < 0.35
    if dpw:
        solver = pj.POMCPDPWSolver(enable_action_pw=bool(rng.random() < 0.6), k_action=float(rng.choice([0.1, 1.0, 10.0])),
                                   alpha_action=float(rng.choice([0.2, 0.5])), k_observation=float(rng.choice([1.0, 2.0, 10.0])),
                                   alpha_observation=float(rng.choice([0.25, 0.5])), **kw)
    else:
        solver = pj.POMCPSolver(**kw)
    bm = abi.BELIEF_UNWEIGHTED if mode == "unweighted" else abi.BELIEF_EXTERNAL
    ctx = dict(case=idx, problem=name, mode=mode, dpw=dpw, **{k: v for k, v in kw.items() if k != "estimate_value"})
    gp = pj.solve(solver, pomdp, belief_mode=bm)
    try:
        n_part = int(rng.choice([1, 100, 5000, 70000]))
        if mode == "sir":  # the SIR update needs a particle belief
            if name == "lightdark":
                parts = np.zeros(n_part, dtype=pomdp.state_dtype)
                parts["y"] = rng.normal(2.0, 3.0, n_part)
            elif name in ("tiger", "baby"):
                parts = (rng.random(n_part) < 0.5).astype(np.uint32)
            else:
                rocks = rng.integers(0, 1 << pomdp.k, n_part).astype(np.uint32)
                parts = (np.uint32(pomdp.start[0]) | np.uint32(pomdp.start[1] << 4) | (rocks << 8)).astype(np.uint32)
            gp.set_belief_particles(parts)
        else:
            gp.set_belief_initial()
        s = gp.initial_state(seed=int(rng.integers(0, 1000)))
        root_n = 0
        for step in range(int(rng.integers(1, 5))):
            gp.reset_counters()
            try:
                a, N, V = gp.search(Q)
            except pj.AllSamplesTerminal:
                break
            cs = gp.counters()
            assert cs["simulations"] + cs["terminal_skips"] == T * Q, ("counters", ctx, step, cs)
            assert 0 <= a < pomdp.n_actions, ("action", ctx, step, a)
            assert np.isfinite(V[N > kw["init_N"]]).all(), ("values", ctx, step, N, V)
            if T == 1:
                rn = gp.root_stats()[0]
                assert rn == root_n + Q - cs["terminal_skips"], ("root visits", ctx, step, rn, root_n, cs)
            sp, o, r, term = gp.generate_sor(s, a, 77 + idx, step)
            if term:
                break
            s = sp
            try:
                if mode == "unweighted":
                    if rng.random() < 0.5:
                        gp.update_unweighted(a, o)
                    else:
                        gp.update_reinvigorated(a, o, seed=step, n_min=int(rng.choice([10, 500])))
                    root_n = gp.root_stats()[0]
                elif mode == "sir":
                    n_out = int(rng.choice([1, 100, 5000, 70000]))
                    gp.pf_update_sir(a, o, seed=step, n_out=n_out)
                    bp = gp.belief_particles()
                    assert len(bp) == n_out, ("sir size", ctx, step, len(bp), n_out)
                    root_n = 0
                else:
                    gp.set_belief_initial()
                    root_n = 0
            except (pj.ParticleDepletion, pj.AllSamplesTerminal):
                break
    finally:
        gp.close()
    return ctx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    bad = 0
    for i in range(a.cases):
        try:
            one_case(rng, i)
        except (NotImplementedError, ValueError) as ex:  # a combination the library rejects by design
            print(f"case {i}: rejected: {type(ex).__name__}: {str(ex)[:120]}", flush=True)
        except Exception:  # noqa: BLE001
            bad += 1
            print(f"case {i}: FAILED\n{traceback.format_exc()[-1500:]}", flush=True)
    print(f"{a.cases} cases, {bad} failures")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
