"""The host index builder (blight_b200/csrc/builder.cpp) must reproduce the reference's construct_index bit for bit."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from blight_b200 import api, synth
from tests import common
from tests.golden import fixtures

ANS = fixtures.answers()


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("shape", common.LAMBDA_SHAPES)
def test_lambda_blob_equals_reference_export(shape, tmp_path):
    m, n, s, b = shape
    flat = common.build_lambda(m, n, s, b)
    p = os.path.join(str(tmp_path), "a.blflat")
    flat.save(p)
    assert _sha(p) == ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]["blob_sha256"]
    i = flat.info()
    assert i["number_kmer"] == 48462
    assert i["number_super_kmer"] == ANS["lambda"][f"m{m}_n{n}_s{s}_b{b}"]["number_super_kmer"]


@pytest.mark.parametrize("shape", common.SMALL_SHAPES)
def test_small_blob_equals_reference_export(shape, tmp_path):
    m, n, b = shape
    ub, uo, _, _, _ = common.small_case()
    for threads in (1, 3):
        flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=m, n=n, s=min(n, 3), b=b, threads=threads)
        p = os.path.join(str(tmp_path), f"a{threads}.blflat")
        flat.save(p)
        assert _sha(p) == ANS["small"][f"m{m}_n{n}_b{b}"]["blob_sha256"]


def test_lambda_known_answers():
    """C1 index recap."""
    flat = common.build_lambda(7, 5, 3, 6)
    i = flat.info()
    assert (i["number_kmer"], i["number_super_kmer"], i["largest_mphf"], i["largest_bucket"]) == (48462, 3708, 3036, 1337)


def test_file_and_memory_builds_agree(tmp_path):
    fa = os.path.join(str(tmp_path), "l.fa")
    open(fa, "wb").write(fixtures.lambda_fasta())
    a = api.FlatIndex.build_file(fa, 31, 7, 5, 3, 6, 2)
    b = common.build_lambda(7, 5, 3, 6)
    assert a.equals(b), a.difference(b)
    import gzip
    with gzip.open(fa + ".gz", "wb") as f:
        f.write(fixtures.lambda_fasta())
    c = api.FlatIndex.build_file(fa + ".gz", 31, 7, 5, 3, 6, 2)  # gzip-transparent like zstr::ifstream
    assert c.equals(b)


def test_save_load_roundtrip(tmp_path):
    a = common.build_lambda(9, 5, 3, 6)
    p = os.path.join(str(tmp_path), "r.blflat")
    a.save(p)
    b = api.FlatIndex.load(p)
    assert a.equals(b)
    with open(p, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(api.BlightError):
        api.FlatIndex.load(p)


def test_constructor_validation():
    """kmer_Set_Light constructor checks (blight.h:75-92) -> BLIGHT_ERR_INVALID_ARG."""
    L = api.lib()
    assert L.blight_check_params(31, 7, 5, 3, 6) == api.OK
    assert L.blight_check_params(31, 8, 5, 3, 6) == api.ERR_INVALID_ARG   # even m
    assert L.blight_check_params(33, 7, 5, 3, 6) == api.ERR_INVALID_ARG   # k too large
    assert L.blight_check_params(31, 17, 5, 3, 6) == api.ERR_INVALID_ARG  # m too large
    assert L.blight_check_params(31, 7, 14, 3, 6) == api.ERR_INVALID_ARG  # n > 2m-1
    assert L.blight_check_params(31, 7, 5, 6, 6) == api.ERR_INVALID_ARG   # s > n
    with pytest.raises(ValueError):
        api.KmerSetLight(31, 8, 5, 3, 1, 6)


def test_invalid_base_and_missing_file(tmp_path):
    bases, offs = fixtures.lambda_unitigs()
    bad = bases.copy()
    bad[1000] = ord("N")
    with pytest.raises(api.InvalidBase):
        api.FlatIndex.build_seqs(bad, offs, 31, 7, 5, 3, 6, 1)
    with pytest.raises(api.BlightError) as e:
        api.FlatIndex.build_file(os.path.join(str(tmp_path), "nope.fa"), 31, 7, 5, 3, 6, 1)
    assert e.value.code == api.ERR_IO


def test_fasta_record_pairing_quirks(tmp_path):
    """blight.cpp:212-229: the header line is skipped whatever it holds, an empty header swallows the next line, an
    empty sequence drops the record, a missing final newline is fine, lower case is accepted."""
    bases, offs = fixtures.lambda_unitigs()
    s = [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(4)]
    plain = b">a\n" + s[1] + b"\n>b\n" + s[2] + b"\n"
    quirky = b"no-gt-header\n" + s[1] + b"\n\n" + b"SWALLOWED\n" + b">x\n\n" + b">b\n" + s[2].lower()
    pa, pb = os.path.join(str(tmp_path), "a.fa"), os.path.join(str(tmp_path), "b.fa")
    open(pa, "wb").write(plain)
    open(pb, "wb").write(quirky)
    a = api.FlatIndex.build_file(pa, 31, 7, 5, 3, 6, 1)
    b = api.FlatIndex.build_file(pb, 31, 7, 5, 3, 6, 1)
    assert a.equals(b), a.difference(b)
    if oracle.reference_available():
        ref = oracle.Reference(31, 7, 5, 3, 1, 6)
        ref.construct_index(pb)
        rp = os.path.join(str(tmp_path), "ref.blflat")
        ref.export(rp)
        assert api.FlatIndex.load(rp).equals(b)


def test_slice_keeps_global_ids(tmp_path):
    """Partition slices: groups outside the range are empty, ids inside stay global, the union answers like the whole."""
    ub, uo, rb, ro, _ = common.small_case()
    flat = api.FlatIndex.build_seqs(ub, uo, 31, 9, 6, 3, 6, 2)
    whole = common.cport_of(flat, tmp_path, "w.blflat")
    want, _ = whole.query_reads(rb[:150 * 400], ro[:401])
    sizes = flat.group_sizes()
    assert int(sizes.sum()) == flat.info()["number_kmer"]
    cuts = [0, 20, 41, 64]
    got = np.full_like(want, -1)
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = flat.slice(a, b)
        assert part.info()["number_kmer"] == int(sizes[a:b].sum())
        port = common.cport_of(part, tmp_path, f"p{a}.blflat")
        ids, _ = port.query_reads(rb[:150 * 400], ro[:401])
        assert not np.any((ids >= 0) & (got >= 0))
        got = np.where(ids >= 0, ids, got)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_fresh_synthetic_equals_reference(tmp_path):
    g, ub, uo, _, _ = common.synthetic(1_000_000, 10, seed=77)
    fa = os.path.join(str(tmp_path), "u.fa")
    open(fa, "wb").write(synth.fasta_bytes(ub, uo))
    for (m, n, s, b) in [(7, 5, 3, 6), (9, 12, 4, 2)]:
        ref = oracle.Reference(31, m, n, s, 1, b)
        ref.construct_index(fa)
        rp = os.path.join(str(tmp_path), "ref.blflat")
        ref.export(rp)
        ours = api.FlatIndex.build_file(fa, 31, m, n, s, b, 4)
        theirs = api.FlatIndex.load(rp)
        assert ours.equals(theirs), ours.difference(theirs)


def test_corrupted_blobs_are_refused(tmp_path):
    """flat_validate: what the device code later indexes with must be in range — bucket starts that are not the running
    sum of the lengths, level bits that would rank past the group's keys, b beyond the constructor's limit."""
    a = common.build_lambda(7, 5, 3, 6)
    p = os.path.join(str(tmp_path), "ok.blflat")
    a.save(p)
    raw = np.fromfile(p, dtype=np.uint8)
    text, bstart, bnuc, h = common.read_blob_text(p)
    first = int(np.nonzero(bnuc)[0][1])  # a non-empty bucket that is not the first
    for what in ("bucket_start", "level_bits", "b"):
        bad = raw.copy()
        if what == "bucket_start":
            bad[128 + 8 * first:128 + 8 * first + 8].view(np.uint64)[0] += 1
        elif what == "b":
            bad[8 + 16:8 + 20].view(np.uint32)[0] = 25
        else:
            off = 128 + 8 * h["n_buckets"] + (4 * h["n_buckets"] + 7) // 8 * 8 + h["n_mphf"] * 208 + 8 * h["seq_words"]
            pos_words = int(raw[40:128].view(np.uint64)[7])
            bits0 = off + 8 * pos_words
            bad[bits0:bits0 + 4096] = 0xFF  # far more ones than the first group has keys
        q = os.path.join(str(tmp_path), what + ".blflat")
        bad.tofile(q)
        with pytest.raises(api.BlightError):
            api.FlatIndex.load(q)
