"""Parity at the sizes BASELINE.json quotes, id for id (not only through properties):

  configs[1]  5 Mbp genome, k31 m7 n5 s3 b6, 2 M simulated 150 bp reads with errors (240 M k-mers) against the multi-threaded
              C restatement — and, where oracle/_ref travelled with the repo, a 200 k-read sample against the reference itself
  configs[2]  100 M-k-mer index, b in {0, 6, 8} and m in {7, 11}: a 100 k-read sample (12 M k-mers), plus the index's own
              BBHash fallback-map keys (bbhash.h:567-575) queried directly and as reads
  the 64-bit bit arithmetic that only MPHF groups of 2^32 bits or more need, forced on (BLIGHT_FORCE_WIDE)
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from blight_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THREADS = os.cpu_count() or 1


def _index(genome_len, m, n, s, b, tmp_path, name):
    g = synth.random_genome(genome_len, seed=42)
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    flat = api.FlatIndex.build_spans(g, st, ln, 31, m, n, s, b, threads=THREADS)
    blob = os.path.join(str(tmp_path), name)
    flat.save(blob)
    return g, flat, blob


def test_config1_5mbp_2m_reads_id_for_id(tmp_path):
    g, flat, blob = _index(5_000_000, 7, 5, 3, 6, tmp_path, "c1.blflat")
    port = oracle.CPort(blob)
    idx = flat.upload(0)
    rb, ro = synth.simulate_reads(g, 2_000_000, 150, 0.01, 0.5, seed=44)
    want, wctr = port.query_reads(rb, ro, threads=THREADS)
    ids, ctr = idx.query_reads_host(rb, ro)
    assert np.array_equal(ids, want)
    assert [int(c) for c in ctr[:3]] == [int(c) for c in wctr[:3]]
    _, ctr2 = idx.query_reads_host(rb, ro, want_ids=False)
    assert np.array_equal(ctr, ctr2)
    assert 0.70 < int(ctr[0]) / int(ctr[2]) < 0.85  # ~79 % of the k-mers of 1 %-error reads are in the graph
    if oracle.reference_available():
        ref = oracle.Reference.from_blob(blob, 31, 7)
        n_s = 200_000
        rids, f, nf, _ = ref.query_reads(rb[:n_s * 150], ro[:n_s + 1], threads=THREADS)
        assert np.array_equal(ids[:n_s * 120], rids)
    # self-query: a bijection on [0, N)
    N = flat.info()["number_kmer"]
    st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
    ub, uo = synth.concat_sequences(g, st, ln)
    sids, sctr = idx.query_reads_host(ub, uo)
    assert int(sctr[0]) == N and np.array_equal(np.sort(sids), np.arange(N))


@pytest.mark.parametrize("shape", [(7, 5, 0), (7, 5, 6), (7, 5, 8), (11, 5, 6)])
def test_config2_100m_index_sample_and_fallback_keys(shape, tmp_path):
    m, n, b = shape
    g, flat, blob = _index(100_000_000, m, n, 3, b, tmp_path, "c2.blflat")
    info = flat.info()
    assert info["number_kmer"] == 100_000_000 - 30
    port = oracle.CPort(blob)
    idx = flat.upload(0)
    rb, ro = synth.simulate_reads(g, 100_000, 150, 0.01, 0.5, seed=44)
    want, wctr = port.query_reads(rb, ro, threads=THREADS)
    ids, ctr = idx.query_reads_host(rb, ro)
    assert np.array_equal(ids, want)
    assert [int(c) for c in ctr[:3]] == [int(c) for c in wctr[:3]]
    _, ctr2 = idx.query_reads_host(rb, ro, want_ids=False)
    assert np.array_equal(ctr, ctr2)
    # the keys no BBHash level accommodated: answered through the sorted fallback arrays (lookup.cuh: fallback_rank)
    keys, vals, groups = common.read_blob_fallback(blob)
    assert len(keys) == info["fallback_keys"]
    if len(keys):
        expect = np.empty(len(keys), dtype=np.int64)
        for fb_off, fb_count, id_offset in groups:
            expect[fb_off:fb_off + fb_count] = vals[fb_off:fb_off + fb_count].astype(np.int64) + id_offset
        assert np.array_equal(port.query_kmers(keys), expect)  # a key of the index: found, id = rank + group offset
        assert np.array_equal(idx.query_kmers_host(keys), expect)
        kb = common.kmers_to_ascii(keys, 31)
        ko = np.arange(len(keys) + 1, dtype=np.uint64) * np.uint64(31)
        ids2, ctr3 = idx.query_reads_host(kb, ko)
        assert np.array_equal(ids2, expect) and int(ctr3[0]) == len(keys)
    absent = synth.random_canonical_kmers(1_000_000, 31, seed=99)
    assert np.array_equal(idx.query_kmers_host(absent), port.query_kmers(absent))


def test_config2_has_fallback_keys(tmp_path):
    """At 100 M k-mers the BBHash fallback map is not empty (r01: 6 entries at m7 n5), so the test above does exercise it;
    a small index with a tiny gamma-2 level budget would not. Checked on the flat image alone."""
    g, flat, blob = _index(100_000_000, 7, 5, 3, 6, tmp_path, "c2f.blflat")
    assert flat.info()["fallback_keys"] > 0


def test_wide_bit_arithmetic_forced(tmp_path):
    """SMALL=false instantiations of every kernel (64-bit level-bit arithmetic, only needed by MPHF groups of 2^32 bits or
    more) on the 5 Mbp configuration, in a fresh process (the knob is read at upload)."""
    code = r"""
import sys, os, json, numpy as np
sys.path.insert(0, os.getcwd())
import torch, oracle, tempfile
from blight_b200 import api, synth
from blight_b200 import dist as bdist
g = synth.random_genome(5_000_000, seed=42)
st, ln = synth.cut_unitigs(g, 31, 2000, seed=43)
flat = api.FlatIndex.build_spans(g, st, ln, 31, 9, 6, 3, 6, threads=os.cpu_count())
td = tempfile.mkdtemp(); blob = os.path.join(td, "w.blflat"); flat.save(blob)
port = oracle.CPort(blob)
rb, ro = synth.simulate_reads(g, 100_000, 150, 0.02, 0.5, seed=7)
want, wctr = port.query_reads(rb, ro, threads=os.cpu_count())
idx = flat.upload(0)
res = {}
ids, ctr = idx.query_reads_host(rb, ro)
res["reads"] = bool(np.array_equal(ids, want))
_, c2 = idx.query_reads_host(rb, ro, want_ids=False)
res["count"] = [int(c) for c in c2[:3]] == [int(c) for c in wctr[:3]]
d_b = torch.from_numpy(rb).cuda(); d_o = torch.from_numpy(ro.astype(np.int64)).cuda()
koff = synth.kmer_offsets(ro, 31); d_k = torch.from_numpy(koff.astype(np.int64)).cuda()
canon, mini, _ = api.reads_to_kmers(31, 9, d_b, d_o, d_k, int(koff[-1]))
res["kmers"] = bool(np.array_equal(idx.query_kmers(canon.contiguous()).cpu().numpy(), want))
plan = bdist.PartitionPlan([0, flat.info()["n_mphf"]], 2 * 9 - 1 - 6)
ps = bdist.PartitionedSet(plan, flat, 0, 31, 9)
ps.enable_fused(sub_positions=1 << 20)
ids3, _ = ps.query_reads_fused(d_b, d_o, d_k, int(koff[-1]))
torch.cuda.synchronize()
res["partition"] = bool(np.array_equal(ids3.cpu().numpy(), want))
print(json.dumps(res))
"""
    for kern in ({"BLIGHT_FORCE_WIDE": "1"}, {"BLIGHT_FORCE_WIDE": "1", "BLIGHT_READS_KERNEL": "plain"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **kern), cwd=ROOT)
        assert r.returncode == 0, r.stderr[-3000:]
        res = json.loads(r.stdout.strip().splitlines()[-1])
        assert all(res.values()), (kern, res)
