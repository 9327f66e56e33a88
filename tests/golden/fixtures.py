"""Loaders of the committed golden fixtures (made by tests/golden/make_golden.py from the reference itself)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LUT = np.frombuffer(b"ACTG", dtype=np.uint8)  # code (c>>1)&3 -> letter


def _unpack(packed: np.ndarray, n: int) -> np.ndarray:
    codes = np.empty(len(packed) * 4, dtype=np.uint8)
    for j in range(4):
        codes[j::4] = (packed >> (2 * j)) & 3
    return _LUT[codes[:n]]


def lambda_unitigs():
    """The reference's lambda_virus.unitigs.fa sample (4 unitigs, 48,462 31-mers) as (bases uint8, offsets uint64)."""
    z = np.load(os.path.join(HERE, "lambda_unitigs.npz"))
    lens = z["lengths"].astype(np.int64)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    return _unpack(z["packed"], int(offs[-1])), offs


def lambda_fasta() -> bytes:
    bases, offs = lambda_unitigs()
    b = bases.tobytes()
    out = bytearray()
    for i in range(len(offs) - 1):
        out += b">%d\n" % i + b[int(offs[i]):int(offs[i + 1])] + b"\n"
    return bytes(out)


def answers() -> dict:
    with open(os.path.join(HERE, "golden_answers.json")) as f:
        return json.load(f)


def small_reads():
    """Synthetic 200 kbp genome case: dict with unitig bases/offsets, read bases/offsets and, per (m,n,b) shape,
    the reference's ids for every read k-mer."""
    z = np.load(os.path.join(HERE, "small_reads.npz"))
    return {k: z[k] for k in z.files}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).astype("<i8").tobytes()).hexdigest()
