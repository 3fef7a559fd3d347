"""CPU test of the N>1 path (world_size 2, gloo): root-parallel search = one independent tree per
rank (rank mixed into the Philox counter), root [N_a, N_a*V_a] summed by one all-reduce, then
V_a = sum(N V)/sum(N) and the reference's arg-max rule (src/solver.jl:60-70).  The per-rank trees
come from the CPU oracle here (no GPU in this container); the same packing / merge rule is what
pomcp_search_rootparallel does around ncclAllReduce on the device (tests/test_gpu_multirank.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, queries, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import pomcp_jl_b200 as pj
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pomdp = pj.RockSamplePOMDP(4, 4, seed=44)
    sp = O.default_params(c=20.0, seed=99, rank=rank)
    op = O.OraclePlanner(pomdp, sp)
    op.set_belief_initial()
    rc, a, N, V = op.search(queries)
    assert rc == 0
    nA = pomdp.n_actions
    packed = torch.zeros(2 * nA, dtype=torch.float64)
    packed[:nA] = torch.from_numpy(N.astype(np.float64))
    packed[nA:] = torch.from_numpy(N.astype(np.float64) * V)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    n = packed[:nA].numpy()
    v = np.where(n > 0, packed[nA:].numpy() / np.maximum(n, 1), 0.0)
    best = 0
    for i in range(1, nA):
        if v[i] >= v[best]:
            best = i
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[best], n, v, N.astype(np.float64), V]))
    dist.barrier()
    dist.destroy_process_group()


def test_rootparallel_merge_world2(oracle, tmp_path):
    import torch.multiprocessing as mp
    import pomcp_jl_b200 as pj

    world, queries = 2, 3000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, queries, str(tmp_path)), nprocs=world, join=True)
    pomdp = pj.RockSamplePOMDP(4, 4, seed=44)
    nA = pomdp.n_actions
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    # both ranks agree on the merged decision and statistics
    assert np.array_equal(r0[:1 + 2 * nA], r1[:1 + 2 * nA])
    # the two trees are different (distinct Philox streams) but both are complete searches
    N0, N1 = r0[1 + 2 * nA:1 + 3 * nA], r1[1 + 2 * nA:1 + 3 * nA]
    assert N0.sum() == N1.sum() == queries - 1 and not np.array_equal(N0, N1)
    # same answer as the oracle's in-process root-parallel search (threads instead of ranks)
    rc, a, N, V, cs = oracle.search_rootparallel(pomdp, oracle.default_params(c=20.0, seed=99), None, 2, 2, queries)
    assert rc == 0
    assert int(r0[0]) == a
    assert np.array_equal(r0[1:1 + nA], N.astype(np.float64))
    assert np.allclose(r0[1 + nA:1 + 2 * nA], V, rtol=0, atol=1e-12)
    assert cs["simulations"] == 2 * queries
