"""Seeded synthetic inputs for the Blight query path: random genome, unitigs cut from it,
simulated reads.  No network, so BCALM2 is unavailable; the unitigs are consecutive genome slices that overlap
by k-1, which is what a compacted de Bruijn graph of a repeat-free random genome looks like.

numpy versions serve the CPU tests and small cases; the torch versions build the large benchmark inputs
directly on the GPU (10 M reads = 1.5 GB of bases) and are only data preparation, never timed.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
# complement table over ASCII (upper case)
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def random_genome(n: int, seed: int = 42) -> np.ndarray:
    """i.i.d. uniform ACGT, as uint8 ASCII."""
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def cut_unitigs(genome: np.ndarray, k: int = 31, mean_len: int = 2000, seed: int = 43):
    """Consecutive slices overlapping by k-1 with lengths k + Geometric(mean mean_len).
    Returns (starts, lengths) int64 arrays; every unitig has length >= k and together they cover every k-mer of
    the genome exactly once."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    est = int(G / mean_len * 1.3) + 16
    starts, lens = [], []
    pos = 0
    while pos + k <= G:
        ext = rng.geometric(1.0 / mean_len, size=est)
        for e in ext:
            L = min(k + int(e), G - pos)
            if L < k:
                break
            starts.append(pos)
            lens.append(L)
            pos += L - (k - 1)
            if pos + k > G:
                break
    return np.asarray(starts, dtype=np.int64), np.asarray(lens, dtype=np.int64)


def concat_sequences(genome: np.ndarray, starts: np.ndarray, lens: np.ndarray):
    """Returns (bases, offsets): the sequences back to back without separators, offsets has n+1 entries."""
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    idx = np.arange(total, dtype=np.int64)
    seq_id = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    src = starts[seq_id] + (idx - offsets[:-1].astype(np.int64)[seq_id])
    return genome[src], offsets


def fasta_bytes(bases: np.ndarray, offsets: np.ndarray) -> bytes:
    """2-line FASTA records ('>i' header, sequence) like a BCALM unitig file."""
    out = bytearray()
    b = bases.tobytes()
    for i in range(len(offsets) - 1):
        out += b">%d\n" % i
        out += b[int(offsets[i]):int(offsets[i + 1])]
        out += b"\n"
    return bytes(out)


def simulate_reads(genome: np.ndarray, n_reads: int, read_len: int = 150, sub_rate: float = 0.01,
                   rc_fraction: float = 0.5, seed: int = 44):
    """Uniform starts, fixed length, each base replaced with probability sub_rate by a uniform ACGT draw,
    rc_fraction of the reads reverse-complemented.  Returns (bases uint8 [n_reads*read_len], offsets uint64)."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    st = rng.integers(0, G - read_len + 1, size=n_reads, dtype=np.int64)
    idx = st[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
    reads = genome[idx]
    sub = rng.random(reads.shape) < sub_rate
    repl = ACGT[rng.integers(0, 4, size=reads.shape, dtype=np.uint8)]
    reads = np.where(sub, repl, reads)
    rc = rng.random(n_reads) < rc_fraction
    reads[rc] = _COMP[reads[rc][:, ::-1]]
    offsets = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return np.ascontiguousarray(reads.reshape(-1)), offsets


def kmer_offsets(read_offsets: np.ndarray, k: int) -> np.ndarray:
    """Exclusive prefix of max(0, len-k+1): where each read's ids land in the output."""
    lens = np.diff(read_offsets.astype(np.int64))
    nk = np.maximum(lens - (k - 1), 0)
    out = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(nk, out=out[1:])
    return out


def random_canonical_kmers(n: int, k: int = 31, seed: int = 12345) -> np.ndarray:
    """Uniform random k-mers, canonicalised (min of forward and reverse complement)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << (2 * k), size=n, dtype=np.uint64)
    return np.minimum(x, revcomp_u64(x, k))


def revcomp_u64(x: np.ndarray, k: int) -> np.ndarray:
    """Reverse complement of 2-bit packed k-mers (A0 C1 T2 G3, first base in the high field)."""
    x = x.astype(np.uint64) ^ np.uint64(0xAAAAAAAAAAAAAAAA)
    x = x.byteswap()
    x = ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4)) | ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F))
    x = ((x & np.uint64(0x3333333333333333)) << np.uint64(2)) | ((x >> np.uint64(2)) & np.uint64(0x3333333333333333))
    return x >> np.uint64(64 - 2 * k)


def encode_kmers(bases: np.ndarray, k: int) -> np.ndarray:
    """All forward k-mers of one ASCII sequence as uint64 (for tests)."""
    codes = ((bases >> 1) & 3).astype(np.uint64)
    n = len(codes) - k + 1
    out = np.zeros(max(n, 0), dtype=np.uint64)
    for j in range(k):
        out = (out << np.uint64(2)) | codes[j:j + n]
    return out


# ---- torch (GPU) generators for the large benchmark inputs --------------------------------------------------

def torch_random_genome(n: int, seed: int, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = lut[torch.randint(0, 4, (e - s,), generator=g, device=device)]
    return out


def torch_simulate_reads(genome, n_reads: int, read_len: int, sub_rate: float, rc_fraction: float, seed: int,
                         chunk: int = 1 << 20):
    """Same model as simulate_reads, on the genome's device, in chunks. Returns bases uint8 [n_reads*read_len]."""
    import torch
    dev = genome.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.arange(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    G = genome.numel()
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    for s in range(0, n_reads, chunk):
        n = min(chunk, n_reads - s)
        st = torch.randint(0, G - read_len + 1, (n,), generator=g, device=dev)
        reads = genome[st[:, None] + ar[None, :]]
        sub = torch.rand((n, read_len), generator=g, device=dev) < sub_rate
        repl = lut[torch.randint(0, 4, (n, read_len), generator=g, device=dev)]
        reads = torch.where(sub, repl, reads)
        rc = torch.rand((n,), generator=g, device=dev) < rc_fraction
        rcr = comp[reads.flip(1).long()]
        reads = torch.where(rc[:, None], rcr, reads)
        out[s * read_len:(s + n) * read_len] = reads.reshape(-1)
    return out
