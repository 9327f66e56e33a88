"""Generates the golden fixtures from the reference itself (oracle/_ref = /root/reference + fixes P1/P2).
Run once in the build container:  python -m tests.golden.make_golden
Writes lambda_unitigs.npz, golden_answers.json, small_reads.npz next to this file."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from blight_b200 import synth  # noqa: E402
from tests.golden import fixtures  # noqa: E402

LAMBDA = "/root/reference/lambda_virus.unitigs.fa"
SHAPES = [(7, 5, 3, 6), (9, 5, 3, 6), (11, 5, 3, 8), (7, 13, 3, 6), (7, 0, 0, 6), (7, 5, 3, 0)]


def pack(bases: np.ndarray) -> np.ndarray:
    codes = ((bases >> 1) & 3).astype(np.uint8)
    pad = (-len(codes)) % 4
    codes = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)])
    return (codes[0::4] | (codes[1::4] << 2) | (codes[2::4] << 4) | (codes[3::4] << 6)).astype(np.uint8)


def main():
    oracle.build_reference()
    seqs = [l.strip().encode() for l in open(LAMBDA) if not l.startswith(">") and l.strip()]
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "lambda_unitigs.npz"), lengths=lens, packed=pack(bases))
    b2, offs = fixtures.lambda_unitigs()
    assert np.array_equal(b2, bases)

    ans = {"lambda": {}, "absent": {}, "small": {}}
    absent = synth.random_canonical_kmers(100000, 31, seed=12345)
    for (m, n, s, b) in SHAPES:
        ref = oracle.Reference(31, m, n, s, 1, b)
        ref.construct_index(LAMBDA)
        ids = np.concatenate([ref.query_sequence(q) for q in seqs])
        key = f"m{m}_n{n}_s{s}_b{b}"
        ans["lambda"][key] = {
            "number_kmer": ref.number_kmer, "number_super_kmer": ref.number_super_kmer,
            "first8": [int(x) for x in ids[:8]], "sha256": fixtures.digest(ids),
            "bijection": bool(np.array_equal(np.sort(ids), np.arange(len(ids)))),
        }
        with tempfile.TemporaryDirectory() as td:
            ref.export(os.path.join(td, "x.blflat"))
            ans["lambda"][key]["blob_sha256"] = hashlib.sha256(open(os.path.join(td, "x.blflat"), "rb").read()).hexdigest()
        a = ref.query_kmers(absent)
        ans["absent"][key] = {"n": len(absent), "n_found": int((a >= 0).sum()), "sha256": fixtures.digest(a)}
        print(key, ans["lambda"][key]["number_super_kmer"], ans["lambda"][key]["first8"], ans["absent"][key]["n_found"])

    # small synthetic case with error-bearing reads (the workload that can expose the b>0 phantom windows)
    g = synth.random_genome(200_000, seed=7)
    st, ln = synth.cut_unitigs(g, 31, 500, seed=8)
    ub, uo = synth.concat_sequences(g, st, ln)
    rb, ro = synth.simulate_reads(g, 3000, 150, 0.02, 0.5, seed=9)
    # a few ragged reads: shorter than k, exactly k, k+1
    extra = [g[100:120], g[5000:5031], g[7000:7032], g[9000:9030], g[11000:11300]]
    rb = np.concatenate([rb] + extra)
    ro = np.concatenate([ro, ro[-1] + np.cumsum([len(e) for e in extra]).astype(np.uint64)])
    out = {"unitig_bases_packed": pack(ub), "unitig_offsets": uo, "read_bases_packed": pack(rb), "read_offsets": ro}
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "u.fa")
        open(fa, "wb").write(synth.fasta_bytes(ub, uo))
        for (m, n, b) in [(7, 5, 0), (7, 5, 3), (7, 5, 6), (7, 5, 8), (9, 8, 6), (11, 4, 6), (5, 9, 4)]:
            ref = oracle.Reference(31, m, n, min(n, 3), 1, b)
            ref.construct_index(fa)
            ids, f, nf, _ = ref.query_reads(rb, ro, threads=1)
            key = f"m{m}_n{n}_b{b}"
            ref.export(os.path.join(td, "x.blflat"))
            blob_sha = hashlib.sha256(open(os.path.join(td, "x.blflat"), "rb").read()).hexdigest()
            out["ids_" + key] = ids.astype(np.int32)
            ans["small"][key] = {"found": f, "not_found": nf, "number_kmer": ref.number_kmer,
                                 "number_super_kmer": ref.number_super_kmer, "sha256": fixtures.digest(ids),
                                 "blob_sha256": blob_sha}
            print(key, f, nf)
    np.savez_compressed(os.path.join(HERE, "small_reads.npz"), **out)
    with open(os.path.join(HERE, "golden_answers.json"), "w") as f:
        json.dump(ans, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
