"""Shared helpers of the test suite: golden cases -> flat index -> blob -> oracle."""
from __future__ import annotations

import os

import numpy as np

import oracle
from blight_b200 import api, synth
from tests.golden import fixtures

LAMBDA_SHAPES = [(7, 5, 3, 6), (9, 5, 3, 6), (11, 5, 3, 8), (7, 13, 3, 6), (7, 0, 0, 6), (7, 5, 3, 0)]
SMALL_SHAPES = [(7, 5, 0), (7, 5, 3), (7, 5, 6), (7, 5, 8), (9, 8, 6), (11, 4, 6), (5, 9, 4)]


def unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    return fixtures._unpack(packed, n)


def small_case():
    z = fixtures.small_reads()
    uo, ro = z["unitig_offsets"], z["read_offsets"]
    ub = unpack2(z["unitig_bases_packed"], int(uo[-1]))
    rb = unpack2(z["read_bases_packed"], int(ro[-1]))
    return ub, uo, rb, ro, z


def build_lambda(m, n, s, b, threads=2) -> api.FlatIndex:
    bases, offs = fixtures.lambda_unitigs()
    return api.FlatIndex.build_seqs(bases, offs, k=31, m=m, n=n, s=s, b=b, threads=threads)


def cport_of(flat: api.FlatIndex, tmp_path, name="idx.blflat") -> oracle.CPort:
    p = os.path.join(str(tmp_path), name)
    flat.save(p)
    return oracle.CPort(p)


def synthetic(genome_len=300_000, n_reads=2000, seed=1, sub_rate=0.01, unitig_mean=1000):
    g = synth.random_genome(genome_len, seed=seed)
    st, ln = synth.cut_unitigs(g, 31, unitig_mean, seed=seed + 1)
    ub, uo = synth.concat_sequences(g, st, ln)
    rb, ro = synth.simulate_reads(g, n_reads, 150, sub_rate, 0.5, seed=seed + 2)
    return g, ub, uo, rb, ro


def read_blob_text(path: str):
    """Index text of a BLFLAT01 blob (flat_index.hpp): (ASCII bases of bucketSeq, bucket_start, bucket_nuc, header dict).
    Nucleotide p sits at bits 2p (code >> 1) and 2p+1 (code & 1) of the vector<bool> image; A0 C1 T2 G3."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[:8].tobytes() == b"BLFLAT01"
    u32 = raw[8:40].view(np.uint32)
    u64 = raw[40:40 + 88].view(np.uint64)
    h = dict(k=int(u32[0]), m=int(u32[1]), n_log2=int(u32[2]), s_log2=int(u32[3]), b=int(u32[4]), n_buckets=int(u64[0]),
             n_mphf=int(u64[1]), number_kmer=int(u64[2]), total_nuc=int(u64[4]), seq_words=int(u64[6]))
    off = 128
    nb = h["n_buckets"]
    bucket_start = raw[off:off + 8 * nb].view(np.uint64).copy()
    off += 8 * nb
    bucket_nuc = raw[off:off + 4 * nb].view(np.uint32).copy()
    off += (4 * nb + 7) // 8 * 8
    off += h["n_mphf"] * (10 * 8 + 16 * 8)
    seq = raw[off:off + 8 * h["seq_words"]]
    bits = np.unpackbits(seq, bitorder="little")[: 2 * h["total_nuc"]]
    code = bits[0::2] * 2 + bits[1::2]
    return np.frombuffer(b"ACTG", dtype=np.uint8)[code], bucket_start, bucket_nuc, h


def kmers_of(text: np.ndarray, k: int):
    """(canonical k-mer of every window of an ASCII text, as uint64) — numpy, first base in the high bits."""
    code = (text >> 1) & 3
    n = len(text) - k + 1
    f = np.zeros(n, dtype=np.uint64)
    r = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        c = code[j:j + n].astype(np.uint64)
        f = (f << np.uint64(2)) | c
        r |= (c ^ np.uint64(2)) << np.uint64(2 * j)
    return np.minimum(f, r)


# Shapes whose buckets are tiny, so that keys answered only through a window PAST the end of the bucket they are routed to
# (blight.cpp:729-740 never re-checks the bucket length) exist: (genome length, genome seed, unitig mean, unitig seed, k, m, n, b)
BUCKET_END_SHAPES = [(40000, 4, 60, 5, 31, 7, 5, 8), (30000, 7, 60, 8, 31, 5, 3, 8), (30000, 4, 60, 5, 15, 3, 0, 8), (30000, 2, 40, 3, 21, 5, 2, 7)]
# the first shape's keys and the identifiers the live reference returns for them (re-derived by
# tests/test_oracle.py::test_bucket_end_keys_against_live_reference)
KNOWN_BUCKET_END_IDS = {"0x25df24897e8df79": 18341, "0x9225fa37de45b07": 18312, "0x977c9225fa37de4": 18507,
                        "0x17a4fb1d7605f22c": 18614, "0x293ec75d817c8b25": 18408}


def kmers_to_ascii(kmers: np.ndarray, k: int) -> np.ndarray:
    """2-bit k-mers (first base in the high bits, A0 C1 T2 G3) -> ASCII, back to back."""
    kmers = np.asarray(kmers, dtype=np.uint64)
    sh = (np.uint64(2) * np.arange(k - 1, -1, -1, dtype=np.uint64))[None, :]
    code = ((kmers[:, None] >> sh) & np.uint64(3)).astype(np.int64)
    return np.frombuffer(b"ACTG", dtype=np.uint8)[code].reshape(-1)


def bucket_end_case(shape, workdir: str):
    """(flat index, index text as ASCII, C port, [(key, id)] of the keys the reference finds ONLY through a window that no
    own-bucket lookup answers — the keys a filter built from own-bucket answers would lose)."""
    G, gs, um, us, k, m, n, b = shape
    g = synth.random_genome(G, seed=gs)
    st, ln = synth.cut_unitigs(g, k, um, seed=us)
    ub, uo = synth.concat_sequences(g, st, ln)
    flat = api.FlatIndex.build_seqs(ub, uo, k=k, m=m, n=n, s=0, b=b, threads=0)
    blob = os.path.join(workdir, "be_%s.blflat" % "_".join(map(str, shape)))
    flat.save(blob)
    port = oracle.CPort(blob)
    text, bstart, bnuc, h = read_blob_text(blob)
    x = kmers_of(text, k)
    ids = port.query_kmers(x)
    bucket_of = np.searchsorted(bstart, np.arange(len(x)), side="right") - 1
    own_found = np.zeros(len(x), dtype=bool)
    for i in np.nonzero(ids >= 0)[0]:
        own_found[i] = port.query_get_hash(int(x[i]), int(bucket_of[i])) >= 0
    lost = sorted(set(x[ids >= 0].tolist()) - set(x[own_found].tolist()))
    first = {int(v): int(ids[np.nonzero(x == np.uint64(v))[0][0]]) for v in lost}
    return flat, text, port, [(v, first[v]) for v in lost]


def read_blob_fallback(path: str):
    """(fallback keys, fallback ranks, per-group (fb_off, fb_count, id_offset)) of a BLFLAT01 blob: the keys no BBHash level
    accommodated (bbhash.h:567-575)."""
    raw = np.fromfile(path, dtype=np.uint8)
    u64 = raw[40:40 + 88].view(np.uint64)
    nb, nm = int(u64[0]), int(u64[1])
    seq_words, pos_words, bits_words, ranks_total, fb_total = (int(u64[i]) for i in (6, 7, 8, 9, 10))
    off = 128 + 8 * nb + (4 * nb + 7) // 8 * 8
    recs = raw[off:off + nm * 208].view(np.uint64).reshape(nm, 26)
    off += nm * 208 + 8 * (seq_words + pos_words + bits_words + ranks_total)
    keys = raw[off:off + 8 * fb_total].view(np.uint64).copy()
    vals = raw[off + 8 * fb_total:off + 16 * fb_total].view(np.uint64).copy()
    return keys, vals, [(int(r[7]), int(r[8]), int(r[0])) for r in recs]
