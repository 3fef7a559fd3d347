"""GPU test of K10 (needs >= 2 GPUs, e.g. `gpurun --gpus 2`): root-parallel search through the
C ABI — one tree per rank, one ncclAllReduce of [N_a, N_a*V_a] inside pomcp_search_rootparallel —
must give every rank the same merged decision, equal to the host-side merge of the per-rank
statistics, and the per-rank trees must be distinct complete searches."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, uid_path, out_dir, queries):
    sys.path.insert(0, ROOT)
    import time
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200.api import nccl_unique_id

    if rank == 0:
        uid = nccl_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.05)
        uid = open(uid_path, "rb").read()
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    solver = pj.POMCPSolver(c=20.0, rng=99, tree_queries=queries, search_mode="batched", device=rank, rank=rank)
    pl = pj.solve(solver, pomdp)
    pl.comm_init(uid, rank, world)
    pl.set_belief_initial()
    a, N, V = pl.search_rootparallel(queries)
    rootN, ownN, ownV = pl.root_stats()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[a, rootN], N, V, ownN, ownV]))
    # belief update on every rank's own tree: the merged action and the real observation are
    # the same everywhere; a rank whose tree lacks the child rebuilds its belief with the device-side
    # reinvigorator instead of depleting, so the second decision merges complete searches again.
    o = 0 if a >= 5 else 2
    outcome = pl.update_reinvigorated(int(a), o, seed=500 + rank, n_min=4096)
    n_bel = len(pl.belief_particles())
    a2, N2, V2 = pl.search_rootparallel(queries)
    rootN2, ownN2, _ = pl.root_stats()
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.concatenate([[a2, outcome, n_bel, rootN2], N2, V2, ownN2]))
    pl.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_rootparallel_nccl_merge(tmp_path):
    import torch.multiprocessing as mp

    world, queries = 2, 50000
    mp.spawn(_worker, args=(world, str(tmp_path / "uid"), str(tmp_path), queries), nprocs=world, join=True)
    nA = 13
    r = [np.load(tmp_path / f"r{k}.npy") for k in range(world)]
    merged = [x[2:2 + 2 * nA] for x in r]
    own_N = [x[2 + 2 * nA:2 + 3 * nA] for x in r]
    own_V = [x[2 + 3 * nA:2 + 4 * nA] for x in r]
    assert r[0][0] == r[1][0] and np.array_equal(merged[0], merged[1])  # same decision on every rank
    assert all(x[1] == queries for x in r) and all(n.sum() == queries - 1 for n in own_N)
    assert not np.array_equal(own_N[0], own_N[1])  # distinct Philox streams per rank
    n = own_N[0] + own_N[1]
    v = (own_N[0] * own_V[0] + own_N[1] * own_V[1]) / np.maximum(n, 1)
    assert np.array_equal(merged[0][:nA], n)
    assert np.allclose(merged[0][nA:], v, rtol=0, atol=1e-9)
    best = 0
    for i in range(1, nA):
        if v[i] >= v[best]:
            best = i
    assert int(r[0][0]) == best
    # second decision after the per-rank belief update
    s2 = [np.load(tmp_path / f"s{k}.npy") for k in range(world)]
    assert s2[0][0] == s2[1][0] and np.array_equal(s2[0][4:4 + 2 * nA], s2[1][4:4 + 2 * nA])
    for x in s2:
        assert x[1] in (0, 1, 2) and x[2] >= 4096 or x[1] == 0
        own = x[4 + 2 * nA:4 + 3 * nA]
        assert own.sum() >= queries - 1  # tree reuse: the re-rooted child keeps its counts
    assert np.array_equal(s2[0][4:4 + nA], s2[0][4 + 2 * nA:] + s2[1][4 + 2 * nA:])
