"""world_size-2 gloo tests of the multi-GPU host logic on CPU: partition plan, owner routing, the two all-to-alls and
reassembly (blight_b200/dist.py), with the oracle standing in for the per-rank lookup kernel, plus replica sharding."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blight_b200 import api
from blight_b200 import dist as bdist
from tests import common


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, workdir):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ub, uo, rb, ro, _ = common.small_case()
        flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=3, b=6, threads=1)
        info = flat.info()
        lb = 2 * info["m"] - 1 - info["n_log2"]
        plan = bdist.PartitionPlan.balanced(flat.group_sizes(), world, lb)
        a, b = plan.group_range(rank)
        part_blob = os.path.join(workdir, f"part{rank}.blflat")
        flat.slice(a, b).save(part_blob)
        whole_blob = os.path.join(workdir, f"whole{rank}.blflat")
        flat.save(whole_blob)
        local = oracle.CPort(part_blob)     # this rank answers only its own MPHF groups
        whole = oracle.CPort(whole_blob)    # ground truth

        # this rank's shard of the reads
        n_reads = 600
        lo, hi = bdist.shard_range(n_reads, rank, world)
        canon_l, mini_l, want_l = [], [], []
        for r in range(lo, hi):
            seq = rb[int(ro[r]):int(ro[r + 1])]
            ids, c, mn = whole.query_sequence(seq, with_kmers=True)
            canon_l.append(c); mini_l.append(mn); want_l.append(ids)
        canon = torch.from_numpy(np.concatenate(canon_l).view(np.int64))
        mini = torch.from_numpy(np.concatenate(mini_l).view(np.int32))
        want = np.concatenate(want_l)

        def lookup_local(c, mn):
            cu = c.numpy().view(np.uint64)
            mu = mn.numpy().view(np.uint32)
            # every pair routed here must belong to this rank's group range
            g = (mu.astype(np.int64) >> lb)
            assert ((g >= a) & (g < b)).all()
            return torch.tensor([local.query_get_hash(int(x), int(y)) for x, y in zip(cu, mu)], dtype=torch.int64)

        got = bdist.exchange_lookup(canon, mini, plan, lookup_local).numpy()
        assert np.array_equal(got, want), f"rank {rank}: partitioned ids differ from the whole index"

        # replica-mode bookkeeping: shard sizes and the counter all-reduce
        ctr = torch.tensor([int((want >= 0).sum()), int((want < 0).sum()), len(want), 0], dtype=torch.int64)
        tot = bdist.all_reduce_counters(ctr.clone())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(want)], dtype=torch.int64))
        assert int(tot[2]) == int(sum(int(s) for s in sizes))
        assert int(tot[0]) + int(tot[1]) == int(tot[2])
        open(os.path.join(workdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_partitioned_exchange_gloo_world2():
    with tempfile.TemporaryDirectory() as wd:
        mp.spawn(_worker, args=(2, _free_port(), wd), nprocs=2, join=True)
        assert os.path.exists(os.path.join(wd, "ok0")) and os.path.exists(os.path.join(wd, "ok1"))


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            got = [bdist.shard_range(n, r, w) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in got) - min(b - a for a, b in got) <= 1


def test_partition_plan_balanced():
    rng = np.random.default_rng(1)
    sizes = rng.integers(0, 100000, 1024)
    for w in (2, 4, 8):
        plan = bdist.PartitionPlan.balanced(sizes, w, lb=4)
        assert plan.cuts[0] == 0 and plan.cuts[-1] == 1024 and all(a < b for a, b in zip(plan.cuts, plan.cuts[1:]))
        per = [int(sizes[a:b].sum()) for a, b in zip(plan.cuts, plan.cuts[1:])]
        assert max(per) < 1.1 * sum(per) / w
        mini = torch.from_numpy(rng.integers(0, 1024 << 4, 10000).astype(np.int32))
        own = plan.owner_of(mini).numpy()
        g = mini.numpy().astype(np.int64) >> 4
        for r in range(w):
            assert ((g[own == r] >= plan.cuts[r]) & (g[own == r] < plan.cuts[r + 1])).all()
    with pytest.raises(ValueError):
        bdist.PartitionPlan.balanced(np.ones(4), 8, lb=0)
    # degenerate: all mass in one group still gives every rank a non-empty range
    skew = np.zeros(16); skew[3] = 100
    plan = bdist.PartitionPlan.balanced(skew, 4, lb=0)
    assert all(a < b for a, b in zip(plan.cuts, plan.cuts[1:]))
