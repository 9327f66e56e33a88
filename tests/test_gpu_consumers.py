"""The id consumers fused behind the lookup (what Abundance_De_Bruijn_graph_snippet.cpp:118-151 and
Colored_De_Bruijn_graph_snippet.cpp:117-151 do with the ids of query_sequence_hash) against the same operations
applied on the host to the oracle's ids."""
import numpy as np
import pytest

from blight_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu


def test_abundance_colors_and_gather(tmp_path):
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    g, ub, uo, rb, ro = common.synthetic(500_000, 6000, seed=71, sub_rate=0.02)
    flat = api.FlatIndex.build_seqs(ub, uo, k=31, m=9, n=6, s=0, b=6, threads=0)
    N = flat.info()["number_kmer"]
    port = common.cport_of(flat, tmp_path)
    want, wctr = port.query_reads(rb, ro)
    idx = flat.upload(0)
    koff = synth.kmer_offsets(ro, 31)
    d_b = torch.from_numpy(rb).cuda()
    d_o = torch.from_numpy(ro.astype(np.int64)).cuda()
    d_k = torch.from_numpy(koff.astype(np.int64)).cuda()
    total = int(koff[-1])

    # abundance[id]++ over two passes of the reads
    table = torch.zeros(N, dtype=torch.int32, device="cuda")
    c1 = idx.count_reads(d_b, d_o, table)
    idx.count_reads(d_b, d_o, table)
    torch.cuda.synchronize()
    expect = np.bincount(want[want >= 0], minlength=N)
    assert np.array_equal(table.cpu().numpy().astype(np.int64), 2 * expect)
    assert (int(c1[0]), int(c1[1]), int(c1[2])) == (int(wctr[0]), int(wctr[1]), int(wctr[2]))

    # color[id * n_colors + c] = true, the reads cut into three "files"
    n_colors = 3
    bits = torch.zeros((N * n_colors + 31) // 32, dtype=torch.int32, device="cuda")
    n_reads = len(ro) - 1
    cuts = [0, n_reads // 3, 2 * n_reads // 3, n_reads]
    exp_bits = np.zeros(N * n_colors, dtype=bool)
    for c in range(n_colors):
        lo, hi = cuts[c], cuts[c + 1]
        pb = torch.from_numpy(rb[int(ro[lo]):int(ro[hi])]).cuda()
        po = torch.from_numpy((ro[lo:hi + 1] - ro[lo]).astype(np.int64)).cuda()
        idx.color_reads(pb, po, bits, n_colors, c)
        ids_c = want[int(koff[lo]):int(koff[hi])]
        ids_c = ids_c[ids_c >= 0]
        exp_bits[ids_c * n_colors + c] = True
    torch.cuda.synchronize()
    got_bits = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[: N * n_colors].astype(bool)
    assert np.array_equal(got_bits, exp_bits)

    # the read side: abundance of every k-mer of every read, -1 for k-mers the index does not hold
    out, _ = idx.gather_reads(d_b, d_o, d_k, total, table)
    torch.cuda.synchronize()
    exp_out = np.where(want >= 0, (2 * expect)[np.maximum(want, 0)], -1)
    assert np.array_equal(out.cpu().numpy().astype(np.int64), exp_out)
