"""construct_index on the GPU (csrc/gpu_builder.cu): the flat image must equal, word for word, the one the
reference's own construct_index leaves (golden SHA-256 of the reference's export; the live reference where oracle/_ref
travelled with the repo) — which also makes the identifiers a bijection on [0, N) and the indexed k-mer set the reference's.
On top of the image: queries on the GPU-built index answer like the oracle, and a 100 M-k-mer graph builds faster on the
GPU than with the host builder."""
import hashlib
import os
import time

import numpy as np
import pytest

import oracle
from blight_b200 import api, synth
from tests import common
from tests.golden import fixtures

pytestmark = pytest.mark.gpu
ANS = fixtures.answers()


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def _spans(offsets):
    offsets = np.asarray(offsets, dtype=np.uint64)
    return offsets[:-1].copy(), np.diff(offsets.astype(np.int64)).astype(np.uint64)


@pytest.mark.parametrize("shape", common.LAMBDA_SHAPES)
def test_lambda_gpu_build_equals_reference_export(shape, tmp_path):
    m, n, s, b = shape
    bases, offs = fixtures.lambda_unitigs()
    st, ln = _spans(offs)
    n0 = api.launch_count()
    flat = api.FlatIndex.build_gpu(bases, st, ln, k=31, m=m, n=n, s=s, b=b)
    assert api.launch_count() > n0 + 10
    p = os.path.join(str(tmp_path), "g.blflat")
    flat.save(p)
    assert _sha(p) == ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]["blob_sha256"]
    i = flat.info()
    assert i["number_kmer"] == 48462 and i["number_super_kmer"] == ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]["number_super_kmer"]
    # queried: a bijection on [0, N), the reference's ids
    idx = flat.upload(0)
    ids, ctr = idx.query_reads_host(bases, offs)
    assert fixtures.digest(ids) == ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]["sha256"]
    assert np.array_equal(np.sort(ids), np.arange(48462)) and int(ctr[0]) == 48462


@pytest.mark.parametrize("shape", common.SMALL_SHAPES)
def test_small_gpu_build_equals_reference_export(shape, tmp_path):
    m, n, b = shape
    ub, uo, _, _, _ = common.small_case()
    st, ln = _spans(uo)
    flat = api.FlatIndex.build_gpu(ub, st, ln, k=31, m=m, n=n, s=min(n, 3), b=b)
    p = os.path.join(str(tmp_path), "g.blflat")
    flat.save(p)
    assert _sha(p) == ANS["small"][f"m{m}_n{n}_b{b}"]["blob_sha256"]


@pytest.mark.parametrize("shape", [(31, 7, 5, 6), (31, 9, 12, 2), (31, 11, 8, 8), (31, 13, 16, 5), (31, 5, 0, 3), (21, 7, 4, 4), (16, 5, 3, 2), (25, 11, 8, 6),
                                   (31, 3, 2, 0)])
def test_gpu_build_equals_host_build(shape, tmp_path):
    """Overlapping spans of one genome (how the benchmark feeds unitigs), tiny and huge bucket counts, other k, sequences
    shorter than k in between: the two builders must produce the same image; absent / near-miss k-mers then answer alike
    by construction, which the oracle confirms on the GPU-built blob."""
    k, m, n, b = shape
    g = synth.random_genome(700_000, seed=k + m)
    st, ln = synth.cut_unitigs(g, k, 900, seed=k + m + 1)
    st = np.concatenate([st, [5, 100]]).astype(np.uint64)   # two sequences shorter than k: skipped by both
    ln = np.concatenate([ln, [k - 1, 3]]).astype(np.uint64)
    host = api.FlatIndex.build_spans(g, st, ln, k, m, n, min(n, 3), b, threads=0)
    dev = api.FlatIndex.build_gpu(g, st, ln, k, m, n, min(n, 3), b)
    assert dev.equals(host), dev.difference(host)
    port = common.cport_of(dev, tmp_path)
    idx = dev.upload(0)
    rb, ro = synth.simulate_reads(g, 3000, 100, 0.03, 0.5, seed=9)  # 3 % substitutions: present, absent and near-miss k-mers
    want, wctr = port.query_reads(rb, ro)
    ids, ctr = idx.query_reads_host(rb, ro)
    assert np.array_equal(ids, want) and int(ctr[0]) == int(wctr[0])
    N = dev.info()["number_kmer"]
    found = ids[ids >= 0]
    assert found.max() < N


def test_gpu_build_file_and_errors(tmp_path):
    fa = os.path.join(str(tmp_path), "l.fa")
    open(fa, "wb").write(fixtures.lambda_fasta())
    a = api.FlatIndex.build_file_gpu(fa, 31, 7, 5, 3, 6)
    assert a.equals(common.build_lambda(7, 5, 3, 6))
    bases, offs = fixtures.lambda_unitigs()
    st, ln = _spans(offs)
    bad = bases.copy()
    bad[1000] = ord("N")
    with pytest.raises(api.InvalidBase):
        api.FlatIndex.build_gpu(bad, st, ln, 31, 7, 5, 3, 6)
    with pytest.raises(api.BlightError):
        api.FlatIndex.build_file_gpu(os.path.join(str(tmp_path), "nope.fa"), 31, 7, 5, 3, 6)
    empty = api.FlatIndex.build_gpu(np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64), 31, 7, 5, 3, 6)
    assert empty.info()["number_kmer"] == 0


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_gpu_build_equals_live_reference(tmp_path):
    g, ub, uo, _, _ = common.synthetic(1_000_000, 10, seed=77)
    fa = os.path.join(str(tmp_path), "u.fa")
    open(fa, "wb").write(synth.fasta_bytes(ub, uo))
    for (m, n, s, b) in [(7, 5, 3, 6), (9, 12, 4, 2)]:
        ref = oracle.Reference(31, m, n, s, 1, b)
        ref.construct_index(fa)
        rp = os.path.join(str(tmp_path), "ref.blflat")
        ref.export(rp)
        ours = api.FlatIndex.build_file_gpu(fa, 31, m, n, s, b)
        theirs = api.FlatIndex.load(rp)
        assert ours.equals(theirs), ours.difference(theirs)


def test_100m_kmers_build_faster_than_host_builder():
    """BASELINE configs[2] graph (100 M 31-mers, k31 m7 n5 b6; and m9 n10): same image as the host builder, in less time."""
    g = synth.random_genome(100_000_000, seed=42)
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    for m, n in ((7, 5), (9, 10)):
        t0 = time.perf_counter()
        host = api.FlatIndex.build_spans(g, st, ln, 31, m, n, 3, 6, threads=os.cpu_count() or 1)
        t_host = time.perf_counter() - t0
        api.FlatIndex.build_gpu(g[:2_000_000], st[:500], ln[:500], 31, m, n, 3, 6)  # warm-up: context, kernels
        t0 = time.perf_counter()
        dev = api.FlatIndex.build_gpu(g, st, ln, 31, m, n, 3, 6)
        t_gpu = time.perf_counter() - t0
        assert dev.equals(host), dev.difference(host)
        assert dev.info()["number_kmer"] == 100_000_000 - 30
        print(f"m={m} n={n}: host builder {t_host:.2f} s ({os.cpu_count()} cores), GPU builder {t_gpu:.2f} s wall, {dev.gpu_build_seconds:.3f} s on the device")
        assert t_gpu < t_host
