"""Record cut of the streaming file_query (stream_query.cu / flat_index.cpp: fasta_chunk_lines + fasta_chunk_records) on
the CPU: for any text and any chunk size it must produce the records the reference's getline loop produces
(blight.cpp:760-772: header line skipped whatever it holds, an empty header swallows the next line, an empty sequence
line drops the record, the last line needs no newline)."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from blight_b200 import api


def reference_records(text: bytes):
    """The reference's loop, restated line by line."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # a trailing newline does not open another line
    pos, starts = 0, []
    for ln in lines:
        starts.append(pos)
        pos += len(ln) + 1
    out, i = [], 0
    while i < len(lines):
        header = lines[i]
        if i + 1 >= len(lines):
            break  # getline at EOF returns an empty sequence
        if len(header) == 0:
            i += 2  # the empty header swallows the next line
            continue
        seq = lines[i + 1]
        if len(seq):
            out.append((starts[i + 1], starts[i + 1] + len(seq)))
        i += 2
    return out


def cut(text: bytes, chunk: int, slices: int = 0):
    """slices > 0: the newline search as the streaming reader does it (every chunk's new bytes scanned in that many slices,
    lists merged behind the carried tail: flat_index.cpp fasta_chunk_lines_merge)."""
    L = api.lib()
    buf = np.frombuffer(text, dtype=np.uint8) if text else np.zeros(0, dtype=np.uint8)
    cap = len(text) // 2 + 2
    beg = np.zeros(cap, dtype=np.uint64)
    end = np.zeros(cap, dtype=np.uint64)
    n = C.c_uint64()
    rc = L.blight_fasta_cut_stream_parts(buf.ctypes.data if len(buf) else None, len(buf), chunk, slices, beg.ctypes.data, end.ctypes.data, cap, C.byref(n))
    assert rc == 0, L.blight_last_error()
    return [(int(beg[i]), int(end[i])) for i in range(n.value)]


def test_known_cases():
    cases = [b"", b"\n", b">a\nACGT\n", b">a\nACGT", b"A\nB\nC", b"A\n\nB\nC\n", b"\nA\nB\nC\n", b">x\n\n>y\nAC\n",
             b"hdr-without-gt\nACGT\n\nSWALLOWED\n>x\n\n>y\nGG", b"\n\n\n\n", b">only-header", b">h\n" + b"A" * 1000 + b"\n>h2\n" + b"C" * 7]
    for t in cases:
        want = reference_records(t)
        for chunk in (1, 2, 3, 5, 7, 64, 1 << 20):
            assert cut(t, chunk) == want, (t[:40], chunk)
            for slices in (1, 3, 12):
                assert cut(t, chunk, slices) == want, (t[:40], chunk, slices)


@settings(max_examples=300, deadline=None)
@given(st.lists(st.sampled_from([b"", b">r", b"ACGT", b"A", b"GGGTTTAAACCC", b">", b"x"]), max_size=40), st.booleans(), st.integers(1, 50))
def test_any_text_any_chunk(lines, trailing_newline, chunk):
    text = b"\n".join(lines) + (b"\n" if trailing_newline and lines else b"")
    assert cut(text, chunk) == reference_records(text)


@settings(max_examples=300, deadline=None)
@given(st.lists(st.sampled_from([b"", b">r", b"ACGT", b"A", b"GGGTTTAAACCC" * 9, b">", b"x"]), max_size=40), st.booleans(), st.integers(1, 400),
       st.integers(1, 16))
def test_any_text_any_chunk_reader_slices(lines, trailing_newline, chunk, slices):
    text = b"\n".join(lines) + (b"\n" if trailing_newline and lines else b"")
    assert cut(text, chunk, slices) == reference_records(text)


def test_large_text_parallel_cut():
    """Sizes at which the cut runs its multi-threaded paths (4096 pairs and up), with and without reader slices."""
    rng = np.random.default_rng(5)
    recs = []
    for i in range(20000):
        recs.append(b">r%d" % i if rng.random() > 0.01 else b"")
        recs.append(b"ACGT" * int(rng.integers(0, 40)))
    text = b"\n".join(recs) + b"\n"
    want = reference_records(text)
    for chunk, slices in ((1 << 30, 0), (1 << 30, 12), (200_000, 0), (200_000, 5), (65_537, 16)):
        assert cut(text, chunk, slices) == want, (chunk, slices)


def test_cut_plus_oracle_equals_the_reference_file_query(tmp_path):
    """The reference's own file_query (oracle/_ref, one thread) on FASTA files with pairing quirks: its Good / Erroneous
    recap must equal the counts obtained by cutting the same bytes with the product's record cut and querying every
    record with the C port of the lookup."""
    import os
    import subprocess
    import sys
    import pytest
    import oracle
    from tests import common
    from tests.golden import fixtures
    if not oracle.reference_available():
        pytest.skip("oracle/_ref was not built here")
    flat = common.build_lambda(7, 5, 3, 6)
    blob = str(tmp_path / "lambda.blflat")
    flat.save(blob)
    port = oracle.CPort(blob)
    bases, offs = fixtures.lambda_unitigs()
    s1, s2, s3 = bases[2000:2300].tobytes(), bases[30000:30100].tobytes(), bases[5:45].tobytes()
    mutated = bytearray(bases[7000:7200].tobytes())
    mutated[100] = ord("A") if mutated[100] != ord("A") else ord("C")
    texts = {
        "plain": b">a\n" + s1 + b"\n>b\n" + s2 + b"\n>c\n" + bytes(mutated) + b"\n",
        "quirks": b"hdr-without-gt\n" + s1 + b"\n\n" + b"SWALLOWED\n" + b">x\n\n" + b">y\n" + s2 + b"\n>short\nACGT\n>z\n" + s3,  # no trailing newline
        "empty_lines": b"\n" + s1 + b"\n>q\n" + s2 + b"\n\n\n>r\n" + bytes(mutated) + b"\n",
    }
    code = r"""
import sys, oracle
ref = oracle.Reference.from_blob(sys.argv[1], 31, 7, cores=1)
ref.file_query(sys.argv[2], quiet=False)
"""
    for name, text in texts.items():
        path = tmp_path / (name + ".fa")
        path.write_bytes(text)
        r = subprocess.run([sys.executable, "-c", code, blob, str(path)], capture_output=True, text=True,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stderr[-2000:]
        good = bad = None
        for line in r.stdout.splitlines():
            if line.startswith("Good kmer:"):
                good = int(line.split(":")[1].replace(",", "").strip().split()[0])
            if line.startswith("Erroneous kmers:"):
                bad = int(line.split(":")[1].replace(",", "").strip().split()[0])
        assert good is not None and bad is not None, r.stdout
        f = nf = 0
        for chunk in (7, 1 << 20):
            f = nf = 0
            for b, e in cut(text, chunk):
                ids = port.query_sequence(np.frombuffer(text[b:e], dtype=np.uint8))
                f += int((ids >= 0).sum())
                nf += int((ids < 0).sum())
            assert (f, nf) == (good, bad), (name, chunk, f, nf, good, bad)
