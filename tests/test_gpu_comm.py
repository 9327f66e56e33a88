"""Several GPUs behind the C ABI from ONE process (blight_comm, csrc/comm.cu): replica and bucket-partitioned modes must
return what a single device holding the whole index returns — which is what the oracle returns. On a one-GPU box the
ranks are GPU 0 listed several times (separate indices, sessions and streams; the ordering between ranks is still only
the device-side flags); on a multi-GPU box they are the real devices with peer access over NVLink."""
import numpy as np
import pytest

from blight_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


def _devices(n_sim):
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n > 1 else [0] * n_sim


@pytest.mark.parametrize("mode", [api.COMM_REPLICA, api.COMM_PARTITION])
@pytest.mark.parametrize("shape", [(9, 8, 6), (7, 5, 0), (11, 6, 8)])
def test_comm_matches_oracle(shape, mode, tmp_path):
    m, n, b = shape
    g, ub, uo, rb, ro = common.synthetic(700_000, 7000, seed=41 + m, sub_rate=0.03)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    port = common.cport_of(flat, tmp_path)
    want, wctr = port.query_reads(rb, ro)
    comm = api.Comm(flat, _devices(3), mode)
    if mode == api.COMM_PARTITION:
        assert comm.cuts[0] == 0 and comm.cuts[-1] == 1 << n and sum(comm.kmers) == flat.info()["number_kmer"]
    n0 = api.launch_count()
    ids, ctr = comm.query_reads_host(rb, ro)
    assert api.launch_count() > n0
    assert np.array_equal(ids, want)
    assert [int(c) for c in ctr[:3]] == [int(c) for c in wctr[:3]]
    _, ctr2 = comm.query_reads_host(rb, ro, want_ids=False)
    assert np.array_equal(ctr, ctr2)
    # twice: buffers and sequence numbers carry over
    ids3, _ = comm.query_reads_host(rb, ro)
    assert np.array_equal(ids3, want)
    # one sequence, and fewer reads than devices
    one = comm.query_sequence_host(rb[:150].tobytes())
    assert np.array_equal(one, want[:120])
    assert len(comm.query_sequence_host(rb[:30].tobytes())) == 0
    # FASTA text with the reference's pairing quirks, and a file
    text = synth.fasta_bytes(rb, ro) + b"hdr-without-gt\n" + rb[:300].tobytes() + b"\n\n" + b"SWALLOWED\n" + b">x\n\n" + b">y\n" + rb[300:450].tobytes()
    single = flat.upload(0).query_fasta_host(text)
    c4 = comm.query_fasta_host(text)
    assert [int(c) for c in c4[:3]] == [int(c) for c in single[:3]]
    path = tmp_path / "reads.fa"
    path.write_bytes(text)
    c5 = comm.query_file_host(str(path))
    assert [int(c) for c in c5[:3]] == [int(c) for c in single[:3]]
    # invalid base under a queried k-mer: std::domain_error in the reference (kmer.h:68)
    bad = rb[:3000].copy()
    bad[1000] = ord("N")
    with pytest.raises(api.InvalidBase):
        comm.query_reads_host(bad, ro[:21])
    comm.close()


def test_comm_partition_ragged_and_overflow(tmp_path, monkeypatch):
    """Read lengths 1..4000 and a block of reads of 31-36 bases (far more super-k-mers per base than the inboxes are sized
    for: the overflow is detected and the batch answered again with one record slot per position)."""
    rng = np.random.default_rng(7)
    g, ub, uo, _, _ = common.synthetic(400_000, 10, seed=19)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=0, b=5, threads=0)
    port = common.cport_of(flat, tmp_path)
    lens = np.concatenate([rng.integers(1, 80, 400), rng.integers(31, 37, 60000), rng.integers(100, 4000, 200), [31, 30, 32, 2048, 4096 + 30]])
    rng.shuffle(lens)
    starts = rng.integers(0, len(g) - 4200, len(lens))
    rb = np.concatenate([g[s:s + l] for s, l in zip(starts, lens)])
    ro = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    want, wctr = port.query_reads(rb, ro)
    monkeypatch.setenv("BLIGHT_PART_CAP", "3000")  # inbox regions of 3000 records: this batch needs far more
    comm = api.Comm(flat, _devices(2), api.COMM_PARTITION)
    ids, ctr = comm.query_reads_host(rb, ro)
    assert np.array_equal(ids, want)
    assert (int(ctr[0]), int(ctr[1])) == (int(wctr[0]), int(wctr[1]))
    comm.close()
