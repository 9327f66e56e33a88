"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Bit-exact: identifiers are integers; -1 for absent; the reference's own false positives at b > 0 included."""
import os

import numpy as np
import pytest

from blight_b200 import api, synth
from tests import common
from tests.golden import fixtures

pytestmark = pytest.mark.gpu
ANS = fixtures.answers()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    torch.cuda.set_device(0)
    return torch


@pytest.mark.parametrize("shape", common.LAMBDA_SHAPES)
def test_lambda_self_query(shape, tmp_path, torch_cuda):
    """BASELINE config 1 (and its other shapes): ids identical to the reference's, a bijection on [0, N)."""
    m, n, s, b = shape
    flat = common.build_lambda(m, n, s, b)
    idx = flat.upload(0)
    bases, offs = fixtures.lambda_unitigs()
    n0 = api.launch_count()
    ids, ctr = idx.query_reads_host(bases, offs)
    assert api.launch_count() > n0
    g = ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]
    assert [int(x) for x in ids[:8]] == g["first8"]
    assert fixtures.digest(ids) == g["sha256"]
    assert np.array_equal(np.sort(ids), np.arange(48462))
    assert [int(c) for c in ctr] == [48462, 0, 48462, 0]
    # file_query on FASTA text: Good == N, Erroneous == 0 (SURVEY §4 self-query smoke)
    ctr2 = idx.query_fasta_host(fixtures.lambda_fasta())
    assert [int(c) for c in ctr2] == [48462, 0, 48462, 0]
    # absent random k-mers: all -1, same as the reference
    absent = synth.random_canonical_kmers(100000, 31, seed=12345)
    a = idx.query_kmers_host(absent)
    assert fixtures.digest(a) == ANS["absent"][f"m{m}_n{n}_s{s}_b{b}"]["sha256"]
    assert (a == -1).all()


@pytest.mark.parametrize("shape", common.SMALL_SHAPES)
def test_error_reads_match_reference_golden(shape, torch_cuda):
    """Error-bearing + ragged reads against ids produced by the reference itself (incl. its b>0 false positives)."""
    m, n, b = shape
    ub, uo, rb, ro, z = common.small_case()
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=min(n, 3), b=b, threads=0)
    idx = flat.upload(0)
    ids, ctr = idx.query_reads_host(rb, ro)
    key = f"m{m}_n{n}_b{b}"
    assert np.array_equal(ids, z["ids_" + key].astype(np.int64))
    assert int(ctr[0]) == ANS["small"][key]["found"] and int(ctr[1]) == ANS["small"][key]["not_found"]
    # bool mode (counters only) agrees
    _, ctr2 = idx.query_reads_host(rb, ro, want_ids=False)
    assert np.array_equal(ctr, ctr2)


@pytest.mark.parametrize("shape", [(7, 5, 6), (9, 10, 0), (11, 6, 8), (7, 13, 3), (15, 20, 5), (5, 0, 2)])
def test_fresh_reads_vs_oracle(shape, tmp_path, torch_cuda):
    """Seeded synthetic genome + reads with 3 % substitutions, both strands: GPU == oracle id for id."""
    torch = torch_cuda
    m, n, b = shape
    g, ub, uo, rb, ro = common.synthetic(600_000, 6000, seed=31 + m + b, sub_rate=0.03)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    want, wctr = port.query_reads(rb, ro)
    # device-buffer entry point
    d_b = torch.from_numpy(rb).cuda()
    d_o = torch.from_numpy(ro.astype(np.int64)).cuda()
    koff = synth.kmer_offsets(ro, 31)
    d_k = torch.from_numpy(koff.astype(np.int64)).cuda()
    ids, ctr = idx.query_reads(d_b, d_o, d_k, int(koff[-1]))
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy()[:int(koff[-1])], want)
    c = ctr.cpu().numpy()
    assert (int(c[0]), int(c[1]), int(c[2]), int(c[3])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]), 0)
    # front end alone: canonical k-mers and minimizers
    canon, mini, fctr = api.reads_to_kmers(31, m, d_b, d_o, d_k, int(koff[-1]))
    torch.cuda.synchronize()
    r0 = rb[:150 * 50]
    for i in range(50):
        _, wc, wm = port.query_sequence(r0[150 * i:150 * (i + 1)], with_kmers=True)
        lo = int(koff[i])
        assert np.array_equal(canon[lo:lo + 120].cpu().numpy().view(np.uint64), wc)
        assert np.array_equal(mini[lo:lo + 120].cpu().numpy().view(np.uint32), wm)
    # k-mer entry points: with the minimizer recomputed in the kernel, and with the caller's minimizer
    ids2 = idx.query_kmers(canon.contiguous())
    ids3 = idx.query_kmers(canon.contiguous(), mini=mini.contiguous())
    torch.cuda.synchronize()
    assert np.array_equal(ids2.cpu().numpy(), want)
    assert np.array_equal(ids3.cpu().numpy(), want)
    # absent
    absent = synth.random_canonical_kmers(300000, 31, seed=99)
    assert np.array_equal(idx.query_kmers_host(absent), port.query_kmers(absent))


def test_both_read_kernels_agree(tmp_path, torch_cuda):
    """The per-k-mer kernel and the super-k-mer kernel (anchor + one-window prediction through the valid bitmap / the
    position->id table, negative filter in front of the residual lookups) forced either way and with each optional
    table switched off (fresh processes, the knobs are read once): all must reproduce the oracle, ids and counters,
    including on reads with many errors and on an index with b=8."""
    import subprocess, sys, json as _json
    code = r"""
import sys, os, json, numpy as np
sys.path.insert(0, os.getcwd())
from blight_b200 import api
from tests import common
import oracle, tempfile
out = {}
for (m, n, b, sub) in [(7, 5, 6, 0.02), (9, 9, 8, 0.05), (11, 6, 3, 0.0)]:
    g, ub, uo, rb, ro = common.synthetic(500_000, 5000, seed=5 + m, sub_rate=sub)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=0, b=b, threads=0)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "x.blflat"); flat.save(p); port = oracle.CPort(p)
        want, wctr = port.query_reads(rb, ro)
    idx = flat.upload(0)
    ids, ctr = idx.query_reads_host(rb, ro)
    _, ctr2 = idx.query_reads_host(rb, ro, want_ids=False)
    out[f"{m}_{n}_{b}"] = [bool(np.array_equal(ids, want)), [int(c) for c in ctr[:3]] == [int(c) for c in wctr[:3]],
                           [int(c) for c in ctr2[:3]] == [int(c) for c in wctr[:3]]]
print(json.dumps(out))
"""
    variants = [
        {"BLIGHT_READS_KERNEL": "plain"},
        {"BLIGHT_READS_KERNEL": "sk"},                                                   # + position->id table + filter
        {"BLIGHT_READS_KERNEL": "sk", "BLIGHT_POS_ID": "0", "BLIGHT_FILTER_BITS": "0"},  # valid bitmap only
        {"BLIGHT_READS_KERNEL": "sk", "BLIGHT_FILTER_BITS": "3"},                        # many filter false positives
        {"BLIGHT_READS_KERNEL": "sk", "BLIGHT_FILTER_ANCHORS": "0", "BLIGHT_POS_ID": "0"},
        {"BLIGHT_READS_KERNEL": "plain", "BLIGHT_EXACT_POS": "0"},                       # the reference's truncated positions + 2^b scan only
        {"BLIGHT_READS_KERNEL": "sk", "BLIGHT_EXACT_POS": "0"},
    ]
    for kern in variants:
        env = dict(os.environ, **kern)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stderr[-2000:]
        res = _json.loads(r.stdout.strip().splitlines()[-1])
        assert all(all(v) for v in res.values()), (kern, res)


def test_other_k(tmp_path, torch_cuda):
    """k != 31 (k=21 m=7, k=16 m=5, k=25 m=11)."""
    for k, m, n, b in [(21, 7, 4, 4), (16, 5, 3, 2), (25, 11, 8, 6)]:
        g = synth.random_genome(200_000, seed=k)
        st, ln = synth.cut_unitigs(g, k, 700, seed=k + 1)
        ub, uo = synth.concat_sequences(g, st, ln)
        rb, ro = synth.simulate_reads(g, 1500, 100, 0.02, 0.5, seed=k + 2)
        flat = api.FlatIndex.build_seqs(ub, uo, k=k, m=m, n=n, s=0, b=b, threads=0)
        port = common.cport_of(flat, tmp_path, f"k{k}.blflat")
        idx = flat.upload(0)
        want, wctr = port.query_reads(rb, ro)
        ids, ctr = idx.query_reads_host(rb, ro)
        assert np.array_equal(ids, want), (k, m)
        assert int(ctr[0]) == int(wctr[0])


def test_edge_shapes(tmp_path, torch_cuda):
    """Empty batch, reads shorter than k, exactly k, invalid bases (only when a queried k-mer covers them),
    lower case, FASTA pairing quirks."""
    flat = common.build_lambda(7, 5, 3, 6)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    bases, offs = fixtures.lambda_unitigs()
    ids, ctr = idx.query_reads_host(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert len(ids) == 0 and int(ctr.sum()) == 0
    assert len(idx.query_sequence_host(bases[:30].tobytes())) == 0   # < k: nothing (blight.cpp:577-579)
    one = idx.query_sequence_host(bases[:31].tobytes())
    assert len(one) == 1 and one[0] == port.query_sequence(bases[:31])[0]
    low = bases[100:400].tobytes().lower()
    assert np.array_equal(idx.query_sequence_host(low), port.query_sequence(bases[100:400]))
    bad = bases[100:400].copy()
    bad[150] = ord("N")
    with pytest.raises(api.InvalidBase):   # std::domain_error in the reference (kmer.h:68)
        idx.query_sequence_host(bad.tobytes())
    # an invalid byte inside a read shorter than k is never looked at by the reference (blight.cpp:782)
    rb = np.concatenate([np.frombuffer(b"ACGTNNNN", dtype=np.uint8), bases[1000:1200]])
    ro = np.array([0, 8, 208], dtype=np.uint64)
    ids, ctr = idx.query_reads_host(rb, ro)
    assert np.array_equal(ids, port.query_sequence(bases[1000:1200]))
    # FASTA text with the reference's pairing quirks
    s1, s2 = bases[2000:2300].tobytes(), bases[30000:30100].tobytes()
    text = b"hdr-without-gt\n" + s1 + b"\n\n" + b"SWALLOWED\n" + b">x\n\n" + b">y\n" + s2  # no trailing newline
    ctr = idx.query_fasta_host(text)
    assert [int(c) for c in ctr] == [270 + 70, 0, 340, 0]


def test_ragged_read_lengths(tmp_path, torch_cuda):
    """Read lengths from 1 to 5000, many tiles, reads crossing tile boundaries."""
    rng = np.random.default_rng(5)
    g, ub, uo, _, _ = common.synthetic(500_000, 10, seed=9)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=0, b=5, threads=0)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    lens = np.concatenate([rng.integers(1, 80, 500), rng.integers(100, 5000, 300), [31, 30, 32, 2048, 2049, 4096 + 30]])
    rng.shuffle(lens)
    starts = rng.integers(0, len(g) - 5001, len(lens))
    rb = np.concatenate([g[s:s + l] for s, l in zip(starts, lens)])
    ro = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    want, wctr = port.query_reads(rb, ro)
    ids, ctr = idx.query_reads_host(rb, ro)
    assert np.array_equal(ids, want)
    assert int(ctr[0]) == int(wctr[0]) and int(ctr[1]) == int(wctr[1])


def test_host_batch_in_many_chunks(tmp_path, torch_cuda, monkeypatch):
    """The host entry points copy and query a batch chunk by chunk (copy/compute overlap); tiny chunks must give the
    same ids and counters as one chunk, including reads that straddle chunk boundaries."""
    g, ub, uo, rb, ro = common.synthetic(400_000, 5000, seed=77, sub_rate=0.02)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=7, s=0, b=6, threads=0)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    want, wctr = port.query_reads(rb, ro)
    for kb in ("1", "7", "100000"):
        monkeypatch.setenv("BLIGHT_HOST_CHUNK_KB", kb)
        ids, ctr = idx.query_reads_host(rb, ro)
        assert np.array_equal(ids, want), kb
        assert (int(ctr[0]), int(ctr[1]), int(ctr[2])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]))
        _, ctr2 = idx.query_reads_host(rb, ro, want_ids=False)
        assert np.array_equal(ctr, ctr2)
    text = synth.fasta_bytes(rb, ro)
    monkeypatch.setenv("BLIGHT_HOST_CHUNK_KB", "3")
    c3 = idx.query_fasta_host(text)
    assert (int(c3[0]), int(c3[1])) == (int(wctr[0]), int(wctr[1]))


def test_file_query_streams_the_file(tmp_path, torch_cuda, monkeypatch):
    """file_query(path): the streaming reader (chunks, carried partial records, plain and gzip, last line without a
    newline, the reference's pairing quirks) gives the counters of the in-memory path and of the oracle."""
    import gzip
    g, ub, uo, rb, ro = common.synthetic(400_000, 6000, seed=55, sub_rate=0.02)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=7, s=0, b=6, threads=0)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    want, wctr = port.query_reads(rb, ro)
    text = synth.fasta_bytes(rb, ro)
    quirks = b"hdr-without-gt\n" + rb[:300].tobytes() + b"\n\n" + b"SWALLOWED\n" + b">x\n\n" + b">y\n" + rb[300:450].tobytes()  # no trailing newline
    ref_q = idx.query_fasta_host(quirks)
    plain, gz = tmp_path / "reads.fa", tmp_path / "reads.fa.gz"
    plain.write_bytes(text + quirks)
    with gzip.open(gz, "wb") as f:
        f.write(text + quirks)
    expect = [int(wctr[0]) + int(ref_q[0]), int(wctr[1]) + int(ref_q[1])]
    for kb in ("1", "5", "300", "0"):
        monkeypatch.setenv("BLIGHT_STREAM_CHUNK_KB", kb)
        for path in (plain, gz):
            c = idx.query_file_host(str(path))
            assert [int(c[0]), int(c[1])] == expect, (kb, path)
    monkeypatch.setenv("BLIGHT_FILE_QUERY", "whole")
    c = idx.query_file_host(str(plain))
    assert [int(c[0]), int(c[1])] == expect
    monkeypatch.delenv("BLIGHT_FILE_QUERY")
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    assert int(idx.query_file_host(str(empty)).sum()) == 0
    with pytest.raises(api.BlightError):
        idx.query_file_host(str(tmp_path / "missing.fa"))


def test_large_scale_properties(torch_cuda):
    """5 Mbp index / 1.2 M reads (beyond what the oracle checks in seconds): size-independent properties —
    self-query ids are a bijection on [0, N); a k-mer and its reverse complement get the same id; queries are
    idempotent; bool counts equal the number of non-negative ids; every found id of an error-free read run is
    consistent between the read-level and k-mer-level entry points."""
    torch = torch_cuda
    g = synth.random_genome(5_000_000, seed=42)
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    ub, uo = synth.concat_sequences(g, st, ln)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=7, n=5, s=3, b=6, threads=0)
    N = flat.info()["number_kmer"]
    idx = flat.upload(0)
    d_b = torch.from_numpy(ub).cuda()
    d_o = torch.from_numpy(uo.astype(np.int64)).cuda()
    koff = synth.kmer_offsets(uo, 31)
    d_k = torch.from_numpy(koff.astype(np.int64)).cuda()
    ids, ctr = idx.query_reads(d_b, d_o, d_k, int(koff[-1]))
    torch.cuda.synchronize()
    ids = ids[:int(koff[-1])]
    assert int(koff[-1]) == N
    assert int(ctr[0]) == N and int(ctr[1]) == 0
    srt = torch.sort(ids).values
    assert torch.equal(srt, torch.arange(N, device=srt.device))
    # reads, both strands, with errors
    rb = synth.torch_simulate_reads(torch.from_numpy(g).cuda(), 1_200_000, 150, 0.01, 0.5, seed=44)
    ro = torch.arange(0, 1_200_001, device="cuda", dtype=torch.int64) * 150
    ko = torch.arange(0, 1_200_001, device="cuda", dtype=torch.int64) * 120
    a, ca = idx.query_reads(rb, ro, ko, 1_200_000 * 120)
    b, cb = idx.query_reads(rb, ro, ko, 1_200_000 * 120)
    _, cc = idx.query_reads(rb, ro, want_ids=False)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(ca, cb) and torch.equal(ca, cc)
    assert int(ca[0]) == int((a >= 0).sum()) and int(ca[1]) == int((a < 0).sum())
    assert int(a.max()) < N and int(a.min()) >= -1
    frac = int(ca[0]) / (1_200_000 * 120)
    assert 0.70 < frac < 0.85  # ~79 % of read k-mers are error free
    canon, mini, _ = api.reads_to_kmers(31, 7, rb, ro, ko, 1_200_000 * 120)
    c = idx.query_kmers(canon.contiguous(), mini=mini.contiguous())
    torch.cuda.synchronize()
    assert torch.equal(a, c)


def test_packed_reads_on_the_device(tmp_path, torch_cuda):
    """Reads held as 2-bit codes (blight_query_reads_packed: 16 bases per word, first base in the high bits, nuc2int's codes):
    same ids and counters as the ASCII entry point and the oracle, ragged lengths included; and the host packer's output
    (what blight_query_reads_host sends for its packed chunks) is that very layout."""
    torch = torch_cuda
    rng = np.random.default_rng(11)
    g, ub, uo, _, _ = common.synthetic(500_000, 10, seed=21)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=0, b=5, threads=0)
    port = common.cport_of(flat, tmp_path)
    idx = flat.upload(0)
    lens = np.concatenate([rng.integers(1, 80, 300), rng.integers(100, 3000, 400), [31, 30, 32, 2048, 4096 + 30]])
    rng.shuffle(lens)
    starts = rng.integers(0, len(g) - 5001, len(lens))
    rb = np.concatenate([g[s:s + l] for s, l in zip(starts, lens)])
    sub = rng.random(len(rb)) < 0.02
    rb = np.where(sub, synth.ACGT[rng.integers(0, 4, len(rb))], rb).astype(np.uint8)
    ro = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    want, wctr = port.query_reads(rb, ro)
    code = ((rb >> 1) & 3).astype(np.uint64)
    pad = (-len(code)) % 16
    code = np.concatenate([code, np.zeros(pad + 64, dtype=np.uint64)])
    sh = np.arange(30, -2, -2, dtype=np.uint64)
    packed = (code.reshape(-1, 16) << sh).sum(1).astype(np.uint32)
    d_p = torch.from_numpy(packed.view(np.int32)).cuda()
    d_o = torch.from_numpy(ro.astype(np.int64)).cuda()
    koff = synth.kmer_offsets(ro, 31)
    d_k = torch.from_numpy(koff.astype(np.int64)).cuda()
    ids, ctr = idx.query_reads_packed(d_p, d_o, len(rb), d_k, int(koff[-1]))
    _, ctr2 = idx.query_reads_packed(d_p, d_o, len(rb), want_ids=False)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy()[:int(koff[-1])], want)
    assert [int(c) for c in ctr.cpu()[:3]] == [int(c) for c in wctr[:3]] == [int(c) for c in ctr2.cpu()[:3]]
