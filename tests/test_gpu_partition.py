"""The fused bucket-partitioned path (csrc/part_kernels.cu: super-k-mer records dispatched to the owner's inbox, ids
stored back into the source's buffer) against the oracle, on ONE GPU: a loop-back partition of one rank, and a simulated
partition of three ranks whose inboxes and id buffers all live on GPU 0 (the peer pointers are then local pointers, the
kernels are exactly the ones the NVLink path runs)."""
import numpy as np
import pytest

from blight_b200 import api, synth
from blight_b200 import dist as bdist
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    torch.cuda.set_device(0)
    return torch


def _dev_batch(torch, rb, ro, k=31):
    koff = synth.kmer_offsets(ro, k)
    return (torch.from_numpy(rb).cuda(), torch.from_numpy(ro.astype(np.int64)).cuda(),
            torch.from_numpy(koff.astype(np.int64)).cuda(), int(koff[-1]))


@pytest.mark.parametrize("mode", ["stream", "pull", "direct", "legacy"])
@pytest.mark.parametrize("shape", [(9, 6, 6), (7, 5, 0), (11, 8, 8)])
def test_loopback_partition_matches_oracle(shape, mode, tmp_path, torch_cuda, monkeypatch):
    """stream / pull / direct: the session of part_session.cu (ordering by device-side flags) with its return paths — 32-bit id
    streams pushed into the source's return region or left in the owner's memory for the source to fetch, + scatter pass at
    the source; or int64 ids stored by the owner straight into the source's id array;
    legacy: round 1's Python pipeline (a counter all-to-all per sub-batch)."""
    torch = torch_cuda
    if mode == "legacy":
        monkeypatch.setenv("BLIGHT_PART_PIPELINE", "legacy")
    else:
        monkeypatch.setenv("BLIGHT_PART_RETURN", mode)
    m, n, b = shape
    g, ub, uo, rb, ro = common.synthetic(600_000, 6000, seed=11 + m, sub_rate=0.03)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    port = common.cport_of(flat, tmp_path)
    want, wctr = port.query_reads(rb, ro)
    info = flat.info()
    plan = bdist.PartitionPlan([0, info["n_mphf"]], 2 * m - 1 - n)
    ps = bdist.PartitionedSet(plan, flat, 0, 31, m)
    d_b, d_o, d_k, total = _dev_batch(torch, rb, ro)
    ps.enable_fused(sub_positions=1 << 17)  # several sub-batches, both inbox buffers
    n0 = api.launch_count()
    ids, ctr = ps.query_reads_fused(d_b, d_o, d_k, total)
    torch.cuda.synchronize()
    assert api.launch_count() >= n0 + 2
    assert np.array_equal(ids.cpu().numpy(), want)
    assert [int(c) for c in ctr.cpu()[:3]] == [int(wctr[0]), int(wctr[1]), int(wctr[2])]
    _, ctr2 = ps.query_reads_fused(d_b, d_o, want_ids=False)
    torch.cuda.synchronize()
    assert torch.equal(ctr2.cpu(), ctr.cpu())
    # the experimental kernel order (dispatch of the next sub-batch ahead of the lookup of this one): same answers
    # the other kernel orders (dispatch of the next sub-batch ahead of / beside the lookup of this one): same answers
    for order in ("ahead", "overlap"):
        monkeypatch.setenv("BLIGHT_PART_ORDER", order)
        ps.enable_fused(sub_positions=1 << 17, order=order)  # the order is fixed when the buffers are set up
        ids3, ctr3 = ps.query_reads_fused(d_b, d_o, d_k, total)
        torch.cuda.synchronize()
        assert np.array_equal(ids3.cpu().numpy(), want) and torch.equal(ctr3.cpu(), ctr.cpu()), order
        _, ctr4 = ps.query_reads_fused(d_b, d_o, want_ids=False)
        torch.cuda.synchronize()
        assert torch.equal(ctr4.cpu(), ctr.cpu()), order


def test_loopback_ragged_and_tiny_reads(tmp_path, torch_cuda):
    """Read lengths 1..4000 and a block of reads of 31-36 bases (more than 64 super-k-mers per 256-base strip: the run
    table overflows and the surplus k-mers travel as single-k-mer records); invalid bases raise like the reference."""
    torch = torch_cuda
    rng = np.random.default_rng(3)
    g, ub, uo, _, _ = common.synthetic(400_000, 10, seed=19)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=0, b=5, threads=0)
    port = common.cport_of(flat, tmp_path)
    lens = np.concatenate([rng.integers(1, 80, 400), rng.integers(31, 37, 3000), rng.integers(100, 4000, 200), [31, 30, 32, 2048, 4096 + 30]])
    rng.shuffle(lens)
    starts = rng.integers(0, len(g) - 4200, len(lens))
    rb = np.concatenate([g[s:s + l] for s, l in zip(starts, lens)])
    ro = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    want, wctr = port.query_reads(rb, ro)
    plan = bdist.PartitionPlan([0, flat.info()["n_mphf"]], 2 * 9 - 1 - 6)
    ps = bdist.PartitionedSet(plan, flat, 0, 31, 9)
    d_b, d_o, d_k, total = _dev_batch(torch, rb, ro)
    ps.enable_fused(sub_positions=1 << 16, records_per_position=1.0)
    ids, ctr = ps.query_reads_fused(d_b, d_o, d_k, total)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy(), want)
    assert (int(ctr[0]), int(ctr[1])) == (int(wctr[0]), int(wctr[1]))
    # an inbox too small for the batch: detected, answered through the plain path, same ids
    ps2 = bdist.PartitionedSet(plan, flat, 0, 31, 9)
    ps2.enable_fused(sub_positions=1 << 16, records_per_position=0.001)
    ids2, ctr3 = ps2.query_reads_fused(d_b, d_o, d_k, total)
    torch.cuda.synchronize()
    assert np.array_equal(ids2.cpu().numpy()[:total], want)
    bad = rb.copy()
    bad[int(ro[5]) + 3] = ord("N")
    lens5 = int(ro[6] - ro[5])
    d_bad = torch.from_numpy(bad).cuda()
    if lens5 >= 31:
        with pytest.raises(api.InvalidBase):
            ps.query_reads_fused(d_bad, d_o, d_k, total)


def test_three_owners_two_sources_on_one_gpu(tmp_path, torch_cuda):
    """Partition over 3 ranks simulated on GPU 0: ranks 0 and 1 hold reads, every rank owns a slice of the index."""
    torch = torch_cuda
    m, n, b, world = 9, 8, 6, 3
    g, ub, uo, rb, ro = common.synthetic(800_000, 8000, seed=23, sub_rate=0.02)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    port = common.cport_of(flat, tmp_path)
    want, wctr = port.query_reads(rb, ro)
    plan = bdist.PartitionPlan.balanced(flat.group_sizes(), world, 2 * m - 1 - n)
    owners = [flat.slice(*plan.group_range(r)).upload(0) for r in range(world)]
    # two sources: first and second half of the reads
    half = (len(ro) - 1) // 2
    parts = []
    for lo, hi in ((0, half), (half, len(ro) - 1)):
        o = ro[lo:hi + 1] - ro[lo]
        parts.append((rb[int(ro[lo]):int(ro[hi])], o, want[int(synth.kmer_offsets(ro, 31)[lo]):int(synth.kmer_offsets(ro, 31)[hi])]))
    cap, kcap = 1 << 16, 1 << 21
    RB = api.RUN_RECORD_BYTES
    inbox = torch.zeros(world * world * cap * RB, dtype=torch.uint8, device="cuda")   # [owner][source][cap]
    ret = torch.zeros(2, world, kcap, dtype=torch.int32, device="cuda")               # [source][owner][kcap]
    side = torch.zeros(2, world * cap * 16, dtype=torch.uint8, device="cuda")         # [source][owner][cap]
    counts = torch.zeros(world, world, dtype=torch.int64, device="cuda")              # [source][owner], packed
    ctr = torch.zeros(api.N_CTR, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ids_bufs = []
    for s, (pb, po, _) in enumerate(parts):
        d_b, d_o, d_k, total = _dev_batch(torch, pb, po)
        ids_bufs.append(torch.full((total,), -7, dtype=torch.int64, device="cuda"))
        rt = api.PartRoute()
        rt.world, rt.rank, rt.lb, rt.cap, rt.kcap = world, s, plan.lb, cap, kcap
        rt.side = side[s].data_ptr()
        for i, c in enumerate(plan.cuts):
            rt.cuts[i] = c
        for d in range(world):
            rt.inbox[d] = inbox.data_ptr() + (d * world + s) * cap * RB
        api.part_dispatch(31, m, d_b, d_o, d_k, rt, counts[s], ctr, err)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    for d in range(world):
        regions = [inbox.data_ptr() + (d * world + s) * cap * RB for s in range(world)]
        cnt_d = counts[:, d].contiguous()
        assert int(cnt_d[2]) == 0 and int(cnt_d[0]) > 0
        ret_ptrs = [ret[s, d].data_ptr() for s in range(2)] + [0]
        api.part_lookup(owners[d], regions, cnt_d, ret_ptrs, cap, kcap, ctr)
    bases = np.asarray([o.info["id_base"] for o in owners], dtype=np.uint64)
    assert bases[0] == 0 and bases[1] > 0 and bases[2] > bases[1]  # owners answer with slice-local 32-bit ids
    for s in range(2):
        api.part_scatter(side[s].data_ptr(), cap, counts[s], ret[s].data_ptr(), kcap, world, world * cap, ids_bufs[s], bases)
    torch.cuda.synchronize()
    for s in range(2):
        assert np.array_equal(ids_bufs[s].cpu().numpy(), parts[s][2]), s
    assert (int(ctr[0]), int(ctr[1]), int(ctr[2])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]))


def test_three_sessions_on_one_gpu(tmp_path, torch_cuda):
    """Three ranks of the partitioned path as three sessions of ONE process on GPU 0 (connect_local, one stream each): the
    ordering between ranks is nothing but the device-side flags, exactly as between the GPUs of a box. Ranks hold unequal
    shares of the reads (one holds none); ids land in each rank's own id array; both kernel orders."""
    torch = torch_cuda
    m, n, b, world = 9, 8, 6, 3
    g, ub, uo, rb, ro = common.synthetic(800_000, 9000, seed=29, sub_rate=0.02)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    port = common.cport_of(flat, tmp_path)
    want, wctr = port.query_reads(rb, ro)
    plan = bdist.PartitionPlan.balanced(flat.group_sizes(), world, 2 * m - 1 - n)
    owners = [flat.slice(*plan.group_range(r)).upload(0) for r in range(world)]
    koff_all = synth.kmer_offsets(ro, 31)
    cutsr = [0, 6000, 9000, 9000]  # reads per rank: 6000, 3000, 0
    for order, ret in (("serial", "stream"), ("ahead", "stream"), ("overlap", "stream"), ("serial", "pull"), ("ahead", "pull"), ("overlap", "pull"),
                       ("serial", "direct"), ("overlap", "direct")):
        sub = 1 << 17
        sess, batches = [], []
        for r in range(world):
            lo, hi = cutsr[r], cutsr[r + 1]
            pb = rb[int(ro[lo]):int(ro[hi])] if hi > lo else np.zeros(0, dtype=np.uint8)
            po = (ro[lo:hi + 1] - ro[lo]) if hi > lo else np.zeros(1, dtype=np.uint64)
            d_b, d_o, d_k, total = _dev_batch(torch, pb, po)
            batches.append((d_b, d_o, d_k, total, want[int(koff_all[lo]):int(koff_all[hi])]))
            sess.append(api.PartSession(owners[r], world, r, plan.lb, plan.cuts, sub, 1 << 15, max(total, 1), order=order, return_path=ret))
        for r in range(world):
            for q in range(world):
                if q != r:
                    sess[r].connect_local(q, sess[q])
        n_sub = max((bt[0].numel() + sub - 1) // sub for bt in batches)
        streams = [torch.cuda.Stream() for _ in range(world)]
        ctrs = [torch.zeros(api.N_CTR, dtype=torch.int64, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for rep in range(2):  # twice: the sequence numbers carry over between batches
            for r in range(world):
                ctrs[r].zero_()
            torch.cuda.synchronize()
            for r in range(world):
                d_b, d_o, d_k, total, _ = batches[r]
                sess[r].query(d_b, d_o, d_k, n_sub, ctrs[r], stream=streams[r])
            torch.cuda.synchronize()
            for r in range(world):
                assert sess[r].status(stream=streams[r]) == 0
                total, w = batches[r][3], batches[r][4]
                got = sess[r].ids_tensor("cuda")[:total].cpu().numpy()
                assert np.array_equal(got, w), (order, ret, rep, r)
            tot = sum(c.cpu().numpy() for c in ctrs)
            assert (int(tot[0]), int(tot[1]), int(tot[2])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]))
        # counting mode
        for r in range(world):
            ctrs[r].zero_()
        torch.cuda.synchronize()
        for r in range(world):
            d_b, d_o, d_k, total, _ = batches[r]
            sess[r].query(d_b, d_o, None, n_sub, ctrs[r], stream=streams[r])
        torch.cuda.synchronize()
        tot = sum(c.cpu().numpy() for c in ctrs)
        assert (int(tot[0]), int(tot[1]), int(tot[2])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]))
        for s_ in sess:
            s_.close()
