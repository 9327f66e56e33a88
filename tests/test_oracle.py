"""The oracle (plain-C restatement) against the golden vectors made from the reference itself, and — where the
compiled reference is available (this container) — against the reference directly on fresh random inputs."""
import os

import numpy as np
import pytest

import oracle
from blight_b200 import synth
from tests import common
from tests.golden import fixtures

ANS = fixtures.answers()


@pytest.mark.parametrize("shape", common.LAMBDA_SHAPES)
def test_cport_lambda_matches_reference_golden(shape, tmp_path):
    m, n, s, b = shape
    flat = common.build_lambda(m, n, s, b)
    port = common.cport_of(flat, tmp_path)
    bases, offs = fixtures.lambda_unitigs()
    ids, ctr = port.query_reads(bases, offs)
    g = ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]
    assert port.number_kmer == g["number_kmer"] == 48462
    assert [int(x) for x in ids[:8]] == g["first8"]
    assert fixtures.digest(ids) == g["sha256"]
    assert np.array_equal(np.sort(ids), np.arange(len(ids)))  # bijection on [0, N): ids are 0-based
    assert int(ctr[0]) == 48462 and int(ctr[1]) == 0 and int(ctr[2]) == 48462
    absent = synth.random_canonical_kmers(100000, 31, seed=12345)
    a = port.query_kmers(absent)
    ga = ANS["absent"][f"m{m}_n{n}_s{s}_b{b}"]
    assert int((a >= 0).sum()) == ga["n_found"] == 0
    assert fixtures.digest(a) == ga["sha256"]


@pytest.mark.parametrize("shape", common.SMALL_SHAPES)
def test_cport_error_reads_match_reference_golden(shape, tmp_path):
    """Error-bearing reads incl. ragged lengths; includes the reference's own false positives at b > 0."""
    m, n, b = shape
    ub, uo, rb, ro, z = common.small_case()
    from blight_b200 import api
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=min(n, 3), b=b, threads=2)
    port = common.cport_of(flat, tmp_path)
    ids, ctr = port.query_reads(rb, ro)
    key = f"m{m}_n{n}_b{b}"
    want = z["ids_" + key].astype(np.int64)
    assert np.array_equal(ids, want)
    assert int(ctr[0]) == ANS["small"][key]["found"] and int(ctr[1]) == ANS["small"][key]["not_found"]


def test_golden_has_reference_false_positives():
    """The golden set must actually exercise F7: more k-mers are 'found' at b=8 than at b=0 on the same reads."""
    assert ANS["small"]["m7_n5_b8"]["found"] > ANS["small"]["m7_n5_b0"]["found"]


def test_cport_rejects_invalid_base(tmp_path):
    flat = common.build_lambda(7, 5, 3, 6)
    port = common.cport_of(flat, tmp_path)
    bases, offs = fixtures.lambda_unitigs()
    q = bases[:100].copy()
    q[50] = ord("N")
    with pytest.raises(ValueError):
        port.query_sequence(q)
    assert len(port.query_sequence(bases[:30])) == 0  # shorter than k: empty (blight.cpp:577-579)
    assert len(port.query_sequence(bases[:31])) == 1


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", [(7, 5, 6), (9, 10, 3), (11, 3, 8), (7, 0, 0)])
def test_cport_equals_reference_on_fresh_inputs(shape, tmp_path):
    m, n, b = shape
    g, ub, uo, rb, ro = common.synthetic(400_000, 4000, seed=100 + m + b, sub_rate=0.03)
    fa = os.path.join(str(tmp_path), "u.fa")
    open(fa, "wb").write(synth.fasta_bytes(ub, uo))
    ref = oracle.Reference(31, m, n, min(n, 3), 1, b)
    ref.construct_index(fa)
    blob = os.path.join(str(tmp_path), "ref.blflat")
    ref.export(blob)
    port = oracle.CPort(blob)
    rids, f, nf, _ = ref.query_reads(rb, ro, threads=2)
    pids, ctr = port.query_reads(rb, ro)
    assert np.array_equal(rids, pids)
    assert (f, nf) == (int(ctr[0]), int(ctr[1]))
    absent = synth.random_canonical_kmers(200000, 31, seed=5)
    assert np.array_equal(ref.query_kmers(absent, threads=2), port.query_kmers(absent))
    # minimizer restatement == patched minimizer_naive
    for x in absent[:2000]:
        assert port.minimizer(int(x)) == ref.minimizer(int(x))


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_import_roundtrip(tmp_path):
    """A blob imported back into a reference object answers like the object it was exported from."""
    bases, offs = fixtures.lambda_unitigs()
    flat = common.build_lambda(7, 5, 3, 6)
    blob = os.path.join(str(tmp_path), "x.blflat")
    flat.save(blob)
    ref = oracle.Reference.from_blob(blob, 31, 7)
    ids, f, nf, _ = ref.query_reads(bases, offs, threads=1)
    assert fixtures.digest(ids) == ANS["lambda"]["m7_n5_s3_b6"]["sha256"]


def test_bucket_end_shapes_hold_bucket_end_keys(tmp_path):
    """The adversarial shapes of tests/test_gpu_index_text.py do contain keys the reference answers only through a window
    no own-bucket lookup answers; the C port returns the known identifiers for the first shape."""
    for i, shape in enumerate(common.BUCKET_END_SHAPES):
        flat, text, port, holes = common.bucket_end_case(shape, str(tmp_path))
        assert len(holes) > 0, shape
        if i == 0:
            assert {hex(k): v for k, v in holes} == common.KNOWN_BUCKET_END_IDS


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape", common.BUCKET_END_SHAPES[:2])
def test_bucket_end_keys_against_live_reference(shape, tmp_path):
    """The index text as one read, through the reference's own construct_index + query_sequence_hash: same ids as the C
    port on our builder's blob, including the keys found through a window past their bucket's end (blight.cpp:729-740)."""
    G, gs, um, us, k, m, n, b = shape
    flat, text, port, holes = common.bucket_end_case(shape, str(tmp_path))
    g = synth.random_genome(G, seed=gs)
    st, ln = synth.cut_unitigs(g, k, um, seed=us)
    ub, uo = synth.concat_sequences(g, st, ln)
    fa = os.path.join(str(tmp_path), "u.fa")
    open(fa, "wb").write(synth.fasta_bytes(ub, uo))
    ref = oracle.Reference(k, m, n, 0, 1, b)
    ref.construct_index(fa)
    off = np.array([0, len(text)], dtype=np.uint64)
    rids, f, nf, _ = ref.query_reads(text, off, threads=1)
    pids, ctr = port.query_reads(text, off)
    assert np.array_equal(rids, pids) and f == int(ctr[0])
    hk = np.array([h[0] for h in holes], dtype=np.uint64)
    assert [int(v) for v in ref.query_kmers(hk)] == [h[1] for h in holes]
