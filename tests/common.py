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
