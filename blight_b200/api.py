"""ctypes binding of the C ABI (include/blight_b200.h) and a Python mirror of the reference's
`kmer_Set_Light` surface (blight.h:15-136) on top of it.

PyTorch is only plumbing here: device buffers, streams, torch.distributed.  Every query goes through the
in-tree CUDA library; if it is missing or no GPU is present the calls fail loudly (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLIGHT_B200_LIB") or os.path.join(_HERE, "lib", "libblight_b200.so")  # the override is for kernel experiments

OK = 0
ERR_INVALID_ARG, ERR_IO, ERR_INVALID_BASE, ERR_CUDA, ERR_NO_DEVICE, ERR_FORMAT, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7
CTR_FOUND, CTR_NOT_FOUND, CTR_QUERIES, CTR_INVALID, N_CTR = 0, 1, 2, 3, 4

# every symbol include/blight_b200.h declares
SYMBOLS = [
    "blight_version", "blight_last_error", "blight_check_params", "blight_flat_build_file", "blight_flat_build_seqs", "blight_flat_build_spans", "blight_flat_build_gpu", "blight_flat_build_file_gpu",
    "blight_flat_save", "blight_flat_load", "blight_flat_free", "blight_flat_info", "blight_flat_compare",
    "blight_flat_slice", "blight_flat_group_sizes", "blight_index_upload", "blight_index_upload_opts", "blight_index_free", "blight_index_info",
    "blight_query_kmers", "blight_query_kmers_mini", "blight_reads_to_kmers", "blight_query_reads", "blight_query_reads_packed",
    "blight_query_sequence_bool_host", "blight_transfer_bytes", "blight_host_pack_stats",
    "blight_query_fasta_host", "blight_query_file_host", "blight_query_sequence_host", "blight_query_reads_host",
    "blight_query_kmers_host", "blight_owner_count", "blight_owner_scatter", "blight_scatter_ids", "blight_launch_count",
    "blight_consume_reads", "blight_gather_reads", "blight_fasta_cut_stream", "blight_fasta_cut_stream_parts", "blight_part_dispatch", "blight_part_lookup", "blight_part_lookup_direct", "blight_part_scatter",
    "blight_part_session_create", "blight_part_session_free", "blight_part_session_handles", "blight_part_session_connect_ipc",
    "blight_part_session_connect_local", "blight_part_session_ids", "blight_part_session_sub_batches", "blight_part_session_query", "blight_part_session_status",
    "blight_comm_init", "blight_comm_free", "blight_comm_describe", "blight_comm_query_reads_host", "blight_comm_query_fasta_host",
    "blight_comm_query_file_host", "blight_comm_query_sequence_host", "blight_peer_alloc", "blight_peer_open", "blight_peer_close", "blight_peer_free",
]
MAX_RANKS = 16
RUN_RECORD_BYTES = 32


class PartRoute(C.Structure):
    """blight_part_route (include/blight_b200.h)"""
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("lb", C.c_uint32), ("reserved", C.c_uint32),
                ("cuts", C.c_uint32 * (MAX_RANKS + 1)), ("inbox", C.c_void_p * MAX_RANKS), ("cap", C.c_uint64),
                ("kcap", C.c_uint64), ("side", C.c_void_p)]


class PartConfig(C.Structure):
    """blight_part_config (include/blight_b200.h)"""
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("lb", C.c_uint32), ("order", C.c_uint32),
                ("cuts", C.c_uint32 * (MAX_RANKS + 1)), ("sub_positions", C.c_uint64), ("cap", C.c_uint64),
                ("ids_capacity", C.c_uint64), ("return_path", C.c_uint32), ("reserved", C.c_uint32), ("ret_kmers", C.c_uint64)]


PART_OVERFLOW, PART_TIMEOUT = 1, 2
PART_RETURNS = {None: 0, "": 0, "default": 0, "session": 0, "stream": 1, "direct": 2, "pull": 3}
PART_ORDERS = {None: 0, "": 0, "default": 0, "serial": 1, "ahead": 2, "overlap": 3}


class BlightError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"blight_b200 error {code}: {msg}")
        self.code = code


class InvalidBase(BlightError, ValueError):
    """std::domain_error("Invalid char in DNA") of the reference (kmer.h:68)."""


class Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("k", "m", "n_log2", "s_log2", "b", "layout")] + [
        (n, C.c_uint64) for n in ("n_buckets", "n_mphf", "number_kmer", "number_super_kmer", "total_nuc", "positions_bits",
                                  "mphf_bits", "fallback_keys", "largest_mphf", "largest_bucket", "device_bytes", "id_base")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


LAYOUT_POS_ID, LAYOUT_FILTER, LAYOUT_EXACT_POS = 1, 2, 4
COMM_REPLICA, COMM_PARTITION = 0, 1


class CommInfo(C.Structure):
    """blight_comm_info (include/blight_b200.h)"""
    _fields_ = [("n_gpus", C.c_uint32), ("mode", C.c_uint32), ("devices", C.c_int32 * MAX_RANKS), ("device_bytes", C.c_uint64 * MAX_RANKS),
                ("kmers", C.c_uint64 * MAX_RANKS), ("cuts", C.c_uint32 * (MAX_RANKS + 1)), ("whole", Info)]


class UploadOptions(C.Structure):
    """blight_upload_options (include/blight_b200.h): 1 = on, 0 = off, -1 = library default."""
    _fields_ = [("struct_size", C.c_uint32), ("pos_id", C.c_int32), ("filter_bits", C.c_int32), ("exact_pos", C.c_int32),
                ("filter_anchors", C.c_int32)]

    def __init__(self, pos_id=-1, filter_bits=-1, exact_pos=-1, filter_anchors=-1):
        super().__init__(C.sizeof(UploadOptions), pos_id, filter_bits, exact_pos, filter_anchors)

    @classmethod
    def compact(cls) -> "UploadOptions":
        """Only the reference's own arrays re-laid out (+ 1 bit per base): ~30 bits per k-mer instead of ~160."""
        return cls(0, 0, 0, 0)


_lib = None


def lib() -> C.CDLL:
    """Loads the in-tree shared library (built by __graft_entry__.build() / make -C blight_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(blight_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, cp = C.c_void_p, C.c_uint32, C.c_uint64, C.c_char_p
    L.blight_version.restype = cp
    L.blight_last_error.restype = cp
    L.blight_check_params.argtypes = [u32] * 5
    L.blight_flat_build_file.argtypes = [cp] + [u32] * 6 + [C.POINTER(vp)]
    L.blight_flat_build_seqs.argtypes = [vp, vp, u64] + [u32] * 6 + [C.POINTER(vp)]
    L.blight_flat_build_spans.argtypes = [vp, vp, vp, u64] + [u32] * 6 + [C.POINTER(vp)]
    L.blight_flat_build_gpu.argtypes = [vp, vp, vp, u64] + [u32] * 5 + [C.c_int, C.POINTER(vp), C.POINTER(C.c_double)]
    L.blight_flat_build_file_gpu.argtypes = [cp] + [u32] * 5 + [C.c_int, C.POINTER(vp)]
    L.blight_flat_save.argtypes = [vp, cp]
    L.blight_flat_load.argtypes = [cp, C.POINTER(vp)]
    L.blight_flat_free.argtypes = [vp]
    L.blight_flat_free.restype = None
    L.blight_flat_info.argtypes = [vp, C.POINTER(Info)]
    L.blight_flat_compare.argtypes = [vp, vp]
    L.blight_flat_slice.argtypes = [vp, u64, u64, C.POINTER(vp)]
    L.blight_flat_group_sizes.argtypes = [vp, vp]
    L.blight_index_upload.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.blight_index_upload_opts.argtypes = [vp, C.c_int, vp, C.POINTER(vp)]
    L.blight_index_free.argtypes = [vp]
    L.blight_index_free.restype = None
    L.blight_index_info.argtypes = [vp, C.POINTER(Info)]
    L.blight_query_kmers.argtypes = [vp, vp, u64, vp, vp]
    L.blight_query_kmers_mini.argtypes = [vp, vp, vp, u64, vp, vp]
    L.blight_reads_to_kmers.argtypes = [u32, u32, vp, vp, vp, u64, u64, vp, vp, vp, vp]
    L.blight_query_reads.argtypes = [vp, vp, vp, vp, u64, u64, u64, vp, vp, vp]
    L.blight_query_reads_packed.argtypes = [vp, vp, vp, vp, u64, u64, vp, vp, vp]
    L.blight_query_sequence_bool_host.argtypes = [vp, vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.blight_transfer_bytes.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    L.blight_transfer_bytes.restype = None
    L.blight_host_pack_stats.argtypes = [C.POINTER(u64), C.POINTER(u64), C.POINTER(u32)]
    L.blight_host_pack_stats.restype = None
    L.blight_query_fasta_host.argtypes = [vp, vp, u64, vp]
    L.blight_query_file_host.argtypes = [vp, cp, vp]
    L.blight_query_sequence_host.argtypes = [vp, vp, u64, vp, C.POINTER(u64)]
    L.blight_query_reads_host.argtypes = [vp, vp, vp, u64, vp, vp]
    L.blight_query_kmers_host.argtypes = [vp, vp, u64, vp]
    L.blight_owner_count.argtypes = [vp, u64, vp, u32, u32, vp, vp]
    L.blight_owner_scatter.argtypes = [vp, vp, u64, vp, u32, u32, vp, vp, vp, vp, vp]
    L.blight_scatter_ids.argtypes = [vp, vp, u64, vp, vp]
    L.blight_launch_count.restype = u64
    L.blight_fasta_cut_stream.argtypes = [vp, u64, u64, vp, vp, u64, C.POINTER(u64)]
    L.blight_fasta_cut_stream_parts.argtypes = [vp, u64, u64, u32, vp, vp, u64, C.POINTER(u64)]
    L.blight_consume_reads.argtypes = [vp, vp, vp, u64, u64, C.c_int, vp, u32, u32, vp, vp]
    L.blight_gather_reads.argtypes = [vp, vp, vp, vp, u64, u64, vp, vp, vp, vp]
    L.blight_part_dispatch.argtypes = [u32, u32, vp, vp, vp, u64, u64, u64, u64, C.POINTER(PartRoute), vp, vp, vp, vp]
    L.blight_part_lookup.argtypes = [vp, u32, C.POINTER(vp), vp, C.POINTER(vp), u64, u64, vp, vp]
    L.blight_part_scatter.argtypes = [vp, u64, vp, vp, u64, u32, u64, vp, vp, vp]
    L.blight_part_lookup_direct.argtypes = [vp, u32, C.POINTER(vp), vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64), u64, u64, vp, vp]
    L.blight_part_session_create.argtypes = [vp, C.POINTER(PartConfig), C.POINTER(vp)]
    L.blight_part_session_free.argtypes = [vp]
    L.blight_part_session_free.restype = None
    L.blight_part_session_handles.argtypes = [vp, C.c_char_p]
    L.blight_part_session_connect_ipc.argtypes = [vp, u32, C.c_char_p, u64, u64]
    L.blight_part_session_connect_local.argtypes = [vp, u32, vp]
    L.blight_part_session_ids.argtypes = [vp]
    L.blight_part_session_ids.restype = vp
    L.blight_part_session_query.argtypes = [vp, vp, vp, vp, u64, u64, u64, vp, vp]
    L.blight_part_session_sub_batches.argtypes = [vp, u64, C.c_int]
    L.blight_part_session_sub_batches.restype = u64
    L.blight_part_session_status.argtypes = [vp, C.POINTER(u32), C.c_int, vp]
    L.blight_comm_init.argtypes = [vp, C.POINTER(C.c_int), u32, C.c_int, vp, C.POINTER(vp)]
    L.blight_comm_free.argtypes = [vp]
    L.blight_comm_free.restype = None
    L.blight_comm_describe.argtypes = [vp, C.POINTER(CommInfo)]
    L.blight_comm_query_reads_host.argtypes = [vp, vp, vp, u64, vp, vp]
    L.blight_comm_query_fasta_host.argtypes = [vp, vp, u64, vp]
    L.blight_comm_query_file_host.argtypes = [vp, cp, vp]
    L.blight_comm_query_sequence_host.argtypes = [vp, vp, u64, vp, C.POINTER(u64)]
    L.blight_peer_alloc.argtypes = [u64, C.POINTER(vp), C.c_char_p]
    L.blight_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.blight_peer_close.argtypes = [vp]
    L.blight_peer_free.argtypes = [vp]
    _lib = L
    return L


def _check(rc: int):
    if rc == OK:
        return
    msg = lib().blight_last_error().decode(errors="replace")
    if rc == ERR_INVALID_BASE:
        raise InvalidBase(rc, msg)
    raise BlightError(rc, msg)


def launch_count() -> int:
    return int(lib().blight_launch_count())


def transfer_bytes():
    """(host->device, device->host) bytes copied by the host-buffer entry points since load."""
    a, b = C.c_uint64(), C.c_uint64()
    lib().blight_transfer_bytes(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def host_pack_stats():
    """(bases packed to 2 bits by the host packer, seconds it spent packing, threads it uses) since load."""
    a, b, t = C.c_uint64(), C.c_uint64(), C.c_uint32()
    lib().blight_host_pack_stats(C.byref(a), C.byref(b), C.byref(t))
    return int(a.value), b.value * 1e-9, int(t.value)


class FlatIndex:
    """Host-side flat index image (BLFLAT01)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def build_file(cls, path: str, k=31, m=9, n=17, s=6, b=6, threads=0) -> "FlatIndex":
        h = C.c_void_p()
        _check(lib().blight_flat_build_file(os.fsencode(path), k, m, n, s, b, threads, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_seqs(cls, bases: np.ndarray, offsets: np.ndarray, k=31, m=9, n=17, s=6, b=6, threads=0) -> "FlatIndex":
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().blight_flat_build_seqs(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k, m, n, s, b,
                                            threads, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_spans(cls, bases: np.ndarray, starts: np.ndarray, lengths: np.ndarray, k=31, m=9, n=17, s=6, b=6, threads=0) -> "FlatIndex":
        """Sequences are (possibly overlapping) spans of one base buffer: bases[starts[i] : starts[i]+lengths[i]]."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().blight_flat_build_spans(bases.ctypes.data, starts.ctypes.data, lengths.ctypes.data, len(starts),
                                             k, m, n, s, b, threads, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_gpu(cls, bases: np.ndarray, starts: np.ndarray, lengths: np.ndarray, k=31, m=9, n=17, s=6, b=6, device=0) -> "FlatIndex":
        """construct_index on the GPU (csrc/gpu_builder.cu): same flat image as the host builders, word for word.
        The time between the first H2D copy and the last kernel is left in .gpu_build_seconds."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        h = C.c_void_p()
        sec = C.c_double()
        _check(lib().blight_flat_build_gpu(bases.ctypes.data, starts.ctypes.data, lengths.ctypes.data, len(starts), k, m, n, s, b, device,
                                           C.byref(h), C.byref(sec)))
        out = cls(h.value)
        out.gpu_build_seconds = float(sec.value)
        return out

    @classmethod
    def build_file_gpu(cls, path: str, k=31, m=9, n=17, s=6, b=6, device=0) -> "FlatIndex":
        h = C.c_void_p()
        _check(lib().blight_flat_build_file_gpu(os.fsencode(path), k, m, n, s, b, device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def load(cls, path: str) -> "FlatIndex":
        h = C.c_void_p()
        _check(lib().blight_flat_load(os.fsencode(path), C.byref(h)))
        return cls(h.value)

    def save(self, path: str):
        _check(lib().blight_flat_save(self._h, os.fsencode(path)))

    def info(self) -> dict:
        i = Info()
        _check(lib().blight_flat_info(self._h, C.byref(i)))
        return i.as_dict()

    def equals(self, other: "FlatIndex") -> bool:
        rc = lib().blight_flat_compare(self._h, other._h)
        if rc < 0:
            _check(rc)
        return rc == 0

    def difference(self, other: "FlatIndex") -> str:
        rc = lib().blight_flat_compare(self._h, other._h)
        return "" if rc == 0 else lib().blight_last_error().decode()

    def group_sizes(self) -> np.ndarray:
        out = np.zeros(self.info()["n_mphf"], dtype=np.uint64)
        _check(lib().blight_flat_group_sizes(self._h, out.ctypes.data))
        return out

    def slice(self, g_begin: int, g_end: int) -> "FlatIndex":
        h = C.c_void_p()
        _check(lib().blight_flat_slice(self._h, g_begin, g_end, C.byref(h)))
        return FlatIndex(h.value)

    def upload(self, device: int = 0, options: Optional["UploadOptions"] = None) -> "DeviceIndex":
        h = C.c_void_p()
        if options is None:
            _check(lib().blight_index_upload(self._h, device, C.byref(h)))
        else:
            _check(lib().blight_index_upload_opts(self._h, device, C.addressof(options), C.byref(h)))
        return DeviceIndex(h.value, device)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.blight_flat_free(self._h)
            self._h = C.c_void_p()


def _ptr(t) -> int:
    """data pointer of a torch tensor / numpy array / None"""
    if t is None:
        return 0
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def _stream_handle(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


class DeviceIndex:
    """Index resident in HBM. Device-buffer methods take torch CUDA tensors and are asynchronous on the given
    (default: torch's current) stream; *_host methods take numpy / bytes and are synchronous."""

    def __init__(self, handle: int, device: int):
        self._h = C.c_void_p(handle)
        self.device = device
        i = Info()
        _check(lib().blight_index_info(self._h, C.byref(i)))
        self.info = i.as_dict()
        self.k = self.info["k"]
        self.m = self.info["m"]

    # -- device buffers ----------------------------------------------------------------------------------
    def query_kmers(self, canon, out=None, mini=None, stream=None):
        import torch
        assert canon.is_cuda and canon.dtype in (torch.int64, torch.uint64) and canon.is_contiguous()
        n = canon.numel()
        if out is None:
            out = torch.empty(n, dtype=torch.int64, device=canon.device)
        if mini is None:
            _check(lib().blight_query_kmers(self._h, _ptr(canon), n, _ptr(out), _stream_handle(stream)))
        else:
            assert mini.is_cuda and mini.numel() == n and mini.element_size() == 4 and mini.is_contiguous()
            _check(lib().blight_query_kmers_mini(self._h, _ptr(canon), _ptr(mini), n, _ptr(out), _stream_handle(stream)))
        return out

    def query_reads(self, bases, read_off, kmer_off=None, total_kmers=0, ids=None, ctr=None, want_ids=True, stream=None):
        """bases: uint8 CUDA tensor; read_off: int64/uint64 CUDA tensor (n+1); returns (ids or None, ctr[4])."""
        import torch
        n_reads = read_off.numel() - 1
        if ctr is None:
            ctr = torch.zeros(N_CTR, dtype=torch.int64, device=bases.device)
        if want_ids and ids is None:
            ids = torch.empty(max(int(total_kmers), 1), dtype=torch.int64, device=bases.device)
        _check(lib().blight_query_reads(self._h, _ptr(bases), _ptr(read_off), _ptr(kmer_off), n_reads, bases.numel(),
                                        int(total_kmers), _ptr(ids) if want_ids else 0, _ptr(ctr), _stream_handle(stream)))
        return (ids if want_ids else None), ctr

    # -- id consumers fused behind the lookup (the reference's snippet applications) -------------------------
    def count_reads(self, bases, read_off, table, ctr=None, stream=None):
        """table[id] += 1 (uint32/int32 CUDA tensor of number_kmer entries) for every k-mer of every read: abundance[id]++."""
        import torch
        if ctr is None:
            ctr = torch.zeros(N_CTR, dtype=torch.int64, device=bases.device)
        _check(lib().blight_consume_reads(self._h, _ptr(bases), _ptr(read_off), read_off.numel() - 1, bases.numel(), 0, _ptr(table), 0, 0,
                                          _ptr(ctr), _stream_handle(stream)))
        return ctr

    def color_reads(self, bases, read_off, bits, n_colors: int, color: int, ctr=None, stream=None):
        """Sets bit id * n_colors + color of `bits` (int32 CUDA tensor of ceil(number_kmer * n_colors / 32) words)."""
        import torch
        if ctr is None:
            ctr = torch.zeros(N_CTR, dtype=torch.int64, device=bases.device)
        _check(lib().blight_consume_reads(self._h, _ptr(bases), _ptr(read_off), read_off.numel() - 1, bases.numel(), 1, _ptr(bits), n_colors,
                                          color, _ptr(ctr), _stream_handle(stream)))
        return ctr

    def gather_reads(self, bases, read_off, kmer_off, total_kmers, table, out=None, ctr=None, stream=None):
        """out[kmer_off[r] + pos] = table[id] (0xFFFFFFFF = -1 as int32 for absent k-mers)."""
        import torch
        if ctr is None:
            ctr = torch.zeros(N_CTR, dtype=torch.int64, device=bases.device)
        if out is None:
            out = torch.empty(max(int(total_kmers), 1), dtype=torch.int32, device=bases.device)
        _check(lib().blight_gather_reads(self._h, _ptr(bases), _ptr(read_off), _ptr(kmer_off), read_off.numel() - 1, bases.numel(),
                                         _ptr(table), _ptr(out), _ptr(ctr), _stream_handle(stream)))
        return out[:total_kmers], ctr

    # -- host buffers ------------------------------------------------------------------------------------
    def query_fasta_host(self, text) -> np.ndarray:
        """file_query on a text buffer -> [found, not_found, queries, invalid]"""
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        if isinstance(text, (bytes, bytearray)):
            buf = np.frombuffer(text, dtype=np.uint8)
        else:
            buf = text
        _check(lib().blight_query_fasta_host(self._h, _ptr(buf), len(buf) if isinstance(buf, np.ndarray) else buf.numel(), ctr.ctypes.data))
        return ctr

    def query_file_host(self, path: str) -> np.ndarray:
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        _check(lib().blight_query_file_host(self._h, os.fsencode(path), ctr.ctypes.data))
        return ctr

    def query_sequence_host(self, seq) -> np.ndarray:
        if isinstance(seq, str):
            seq = seq.encode()
        buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.empty(max(len(buf) - self.k + 1, 0), dtype=np.int64)
        n = C.c_uint64()
        _check(lib().blight_query_sequence_host(self._h, buf.ctypes.data, len(buf), out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def query_sequence_bool_host(self, seq):
        """query_sequence_bool: (found, not_found) of one sequence, counted on the device."""
        if isinstance(seq, str):
            seq = seq.encode()
        buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        f, nf = C.c_uint64(), C.c_uint64()
        _check(lib().blight_query_sequence_bool_host(self._h, buf.ctypes.data, len(buf), C.byref(f), C.byref(nf)))
        return int(f.value), int(nf.value)

    def query_reads_packed(self, packed, read_off, total_bases: int, kmer_off=None, total_kmers=0, ids=None, ctr=None, want_ids=True, stream=None):
        """Reads held as 2-bit codes on the device (int32 CUDA tensor, 16 bases per word, first base in the high bits)."""
        import torch
        if ctr is None:
            ctr = torch.zeros(N_CTR, dtype=torch.int64, device=packed.device)
        if want_ids and ids is None:
            ids = torch.empty(max(int(total_kmers), 1), dtype=torch.int64, device=packed.device)
        _check(lib().blight_query_reads_packed(self._h, _ptr(packed), _ptr(read_off), _ptr(kmer_off), read_off.numel() - 1, int(total_bases),
                                               _ptr(ids) if want_ids else 0, _ptr(ctr), _stream_handle(stream)))
        return (ids if want_ids else None), ctr

    def query_reads_host(self, bases: np.ndarray, read_off: np.ndarray, want_ids=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        n = len(read_off) - 1
        ids = None
        if want_ids:
            lens = np.diff(read_off.astype(np.int64))
            ids = np.empty(int(np.maximum(lens - (self.k - 1), 0).sum()), dtype=np.int64)
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        _check(lib().blight_query_reads_host(self._h, bases.ctypes.data, read_off.ctypes.data, n, _ptr(ids), ctr.ctypes.data))
        return ids, ctr

    def query_kmers_host(self, canon: np.ndarray) -> np.ndarray:
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        out = np.empty(len(canon), dtype=np.int64)
        _check(lib().blight_query_kmers_host(self._h, canon.ctypes.data, len(canon), out.ctypes.data))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.blight_index_free(self._h)
            self._h = C.c_void_p()


class Comm:
    """Several GPUs of the box driven from this process (blight_comm, csrc/comm.cu): replica or bucket-partitioned; the
    host-buffer queries of a DeviceIndex with the same results."""

    def __init__(self, flat: FlatIndex, devices: Sequence[int], mode: int = COMM_REPLICA, options: Optional[UploadOptions] = None):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().blight_comm_init(flat._h, devs, len(devices), mode, C.addressof(options) if options is not None else None, C.byref(h)))
        self._h = h
        ci = CommInfo()
        _check(lib().blight_comm_describe(self._h, C.byref(ci)))
        self.k = int(ci.whole.k)
        self.n_gpus, self.mode = int(ci.n_gpus), int(ci.mode)
        self.cuts = [int(ci.cuts[i]) for i in range(self.n_gpus + 1)]
        self.device_bytes = [int(ci.device_bytes[i]) for i in range(self.n_gpus)]
        self.kmers = [int(ci.kmers[i]) for i in range(self.n_gpus)]

    def query_reads_host(self, bases: np.ndarray, read_off: np.ndarray, want_ids=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        ids = None
        if want_ids:
            lens = np.diff(read_off.astype(np.int64))
            ids = np.empty(int(np.maximum(lens - (self.k - 1), 0).sum()), dtype=np.int64)
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        _check(lib().blight_comm_query_reads_host(self._h, bases.ctypes.data, read_off.ctypes.data, len(read_off) - 1, _ptr(ids), ctr.ctypes.data))
        return ids, ctr

    def query_fasta_host(self, text) -> np.ndarray:
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        _check(lib().blight_comm_query_fasta_host(self._h, _ptr(buf), len(buf), ctr.ctypes.data))
        return ctr

    def query_file_host(self, path: str) -> np.ndarray:
        ctr = np.zeros(N_CTR, dtype=np.uint64)
        _check(lib().blight_comm_query_file_host(self._h, os.fsencode(path), ctr.ctypes.data))
        return ctr

    def query_sequence_host(self, seq) -> np.ndarray:
        if isinstance(seq, str):
            seq = seq.encode()
        buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        out = np.empty(max(len(buf) - self.k + 1, 0), dtype=np.int64)
        n = C.c_uint64()
        _check(lib().blight_comm_query_sequence_host(self._h, buf.ctypes.data, len(buf), out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.blight_comm_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerBuffer:
    """Device buffer that the other processes of the box can map (CUDA IPC). `handle` (64 bytes) goes to the peers,
    which call PeerBuffer.open(handle). Exposes __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int, handle: bytes, owner: bool):
        self.ptr, self.nbytes, self.handle, self._owner = ptr, nbytes, handle, owner

    @classmethod
    def alloc(cls, nbytes: int) -> "PeerBuffer":
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(lib().blight_peer_alloc(nbytes, C.byref(p), h))
        return cls(p.value, nbytes, h.raw, True)

    @classmethod
    def open(cls, handle: bytes, nbytes: int) -> "PeerBuffer":
        p = C.c_void_p()
        _check(lib().blight_peer_open(handle, C.byref(p)))
        return cls(p.value, nbytes, handle, False)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self, dtype, device):
        """torch view of the whole buffer (no copy); the PeerBuffer must outlive it."""
        import torch
        return torch.as_tensor(self, device=device).view(dtype)

    def close(self):
        if self.ptr and _lib is not None:
            (_lib.blight_peer_free if self._owner else _lib.blight_peer_close)(C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DeviceView:
    """A device pointer torch can view without a copy (the owner of the memory must outlive the tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PartSession:
    """One rank of the bucket-partitioned path (blight_part_session, csrc/part_session.cu): peer-visible inbox / mailbox /
    id array, the peers' buffers, and the per-batch pipeline ordered by device-side flags."""

    def __init__(self, index: "DeviceIndex", world: int, rank: int, lb: int, cuts: Sequence[int], sub_positions: int, cap: int,
                 ids_capacity: int = 0, order: Optional[str] = None, return_path: Optional[str] = None, ret_kmers: int = 0):
        cfg = PartConfig()
        cfg.return_path = PART_RETURNS[return_path]
        cfg.ret_kmers = ret_kmers
        cfg.world, cfg.rank, cfg.lb, cfg.sub_positions, cfg.cap, cfg.ids_capacity = world, rank, lb, sub_positions, cap, ids_capacity
        cfg.order = PART_ORDERS[order]
        for i, c in enumerate(cuts):
            cfg.cuts[i] = c
        h = C.c_void_p()
        _check(lib().blight_part_session_create(index._h, C.byref(cfg), C.byref(h)))
        self._h, self.index, self.world, self.rank = h, index, world, rank
        self.sub_positions, self.cap, self.ids_capacity = sub_positions, cap, ids_capacity

    def handles(self) -> bytes:
        buf = C.create_string_buffer(256)
        _check(lib().blight_part_session_handles(self._h, buf))
        return buf.raw

    def connect_ipc(self, peer: int, handles: bytes, peer_ids_capacity: int, peer_id_base: int):
        _check(lib().blight_part_session_connect_ipc(self._h, peer, handles, peer_ids_capacity, peer_id_base))

    def connect_local(self, peer: int, other: "PartSession"):
        _check(lib().blight_part_session_connect_local(self._h, peer, other._h))

    def ids_tensor(self, device):
        """int64 torch view of this rank's id array (no copy)."""
        import torch
        p = lib().blight_part_session_ids(self._h)
        if not p:
            return None
        return torch.as_tensor(_DeviceView(p, self.ids_capacity * 8), device=device).view(torch.int64)

    def sub_batches(self, total_bases: int, want_ids: bool) -> int:
        """Sub-batches this rank cuts a batch of total_bases positions into (the n_sub of query() is the maximum over the ranks)."""
        return int(lib().blight_part_session_sub_batches(self._h, int(total_bases), int(bool(want_ids))))

    def query(self, bases, read_off, kmer_off, n_sub: int, ctr, stream=None):
        _check(lib().blight_part_session_query(self._h, _ptr(bases), _ptr(read_off), _ptr(kmer_off), read_off.numel() - 1, bases.numel(),
                                               n_sub, _ptr(ctr), _stream_handle(stream)))

    def status(self, reset: bool = True, stream=None) -> int:
        f = C.c_uint32()
        _check(lib().blight_part_session_status(self._h, C.byref(f), int(reset), _stream_handle(stream)))
        return int(f.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.blight_part_session_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def part_dispatch(k, m, bases, read_off, kmer_off, route: PartRoute, counts, ctr, err, pos_begin=0, pos_end=None, stream=None):
    """Source side of the fused partitioned path: super-k-mer records of the k-mers starting in [pos_begin, pos_end)
    stored into the owners' inboxes. kmer_off None = counting mode."""
    total = bases.numel()
    _check(lib().blight_part_dispatch(k, m, _ptr(bases), _ptr(read_off), _ptr(kmer_off), read_off.numel() - 1, total, pos_begin,
                                      total if pos_end is None else pos_end, C.byref(route), _ptr(counts), _ptr(ctr), _ptr(err),
                                      _stream_handle(stream)))


def part_lookup(index: "DeviceIndex", regions: Sequence[int], counts, ret_ptrs: Optional[Sequence[int]], cap: int, kcap: int, ctr, stream=None):
    """Owner side: looks the received runs up and stores 32-bit ids into its return region at every source
    (ret_ptrs None = counting mode)."""
    world = len(regions)
    reg = (C.c_void_p * world)(*regions)
    retp = (C.c_void_p * world)(*ret_ptrs) if ret_ptrs is not None else None
    _check(lib().blight_part_lookup(index._h, world, reg, _ptr(counts), retp, cap, kcap, _ptr(ctr), _stream_handle(stream)))


def part_scatter(side_ptr: int, cap: int, counts, ret_ptr: int, kcap: int, world: int, max_records: int, ids, id_bases=None, stream=None):
    """Source side, after the owners answered: return regions -> int64 ids in read order. id_bases: first identifier of
    every owner's slice (numpy uint64, world entries); owners return slice-local 32-bit ids."""
    _check(lib().blight_part_scatter(side_ptr, cap, _ptr(counts), ret_ptr, kcap, world, max_records, _ptr(id_bases), _ptr(ids), _stream_handle(stream)))


def reads_to_kmers(k: int, m: int, bases, read_off, kmer_off, total_kmers: int, stream=None):
    """Front end only: (canon uint64-as-int64, minimizer int32, ctr) for every k-mer of every read."""
    import torch
    canon = torch.empty(max(total_kmers, 1), dtype=torch.int64, device=bases.device)
    mini = torch.empty(max(total_kmers, 1), dtype=torch.int32, device=bases.device)
    ctr = torch.zeros(N_CTR, dtype=torch.int64, device=bases.device)
    _check(lib().blight_reads_to_kmers(k, m, _ptr(bases), _ptr(read_off), _ptr(kmer_off), read_off.numel() - 1,
                                       bases.numel(), _ptr(canon), _ptr(mini), _ptr(ctr), _stream_handle(stream)))
    return canon[:total_kmers], mini[:total_kmers], ctr


class KmerSetLight:
    """Python mirror of `kmer_Set_Light` (blight.h:15-136): same constructor arguments, `construct_index`,
    `file_query`, `query_sequence_bool`, `query_sequence_hash`, `query_kmer_bool`, `query_kmer_hash`, and the public
    counters `number_kmer`, `number_super_kmer`, `number_query`."""

    def __init__(self, k: int, m: int, log2_mphfs: int, log2_superbuckets: int, cores: int, bits_to_save: int, device: int = 0):
        rc = lib().blight_check_params(k, m, log2_mphfs, log2_superbuckets, bits_to_save)
        if rc != OK:
            raise ValueError(lib().blight_last_error().decode())  # std::invalid_argument (blight.h:75-92)
        self.k, self.m, self.n, self.s, self.cores, self.b, self.device = k, m, log2_mphfs, log2_superbuckets, cores, bits_to_save, device
        self.flat: Optional[FlatIndex] = None
        self.index: Optional[DeviceIndex] = None
        self.number_kmer = 0
        self.number_super_kmer = 0
        self.number_query = 0

    def construct_index(self, input_file: str):
        self.flat = FlatIndex.build_file(input_file, self.k, self.m, self.n, self.s, self.b, self.cores)
        self._after_build()

    def import_flat(self, flat: FlatIndex):
        """Adopts an index built elsewhere (e.g. exported from the reference's own construction)."""
        self.flat = flat
        self._after_build()

    def _after_build(self):
        i = self.flat.info()
        self.number_kmer, self.number_super_kmer = i["number_kmer"], i["number_super_kmer"]
        self.index = self.flat.upload(self.device)

    def file_query(self, query_file: str):
        ctr = self.index.query_file_host(query_file)
        self.number_query += int(ctr[CTR_QUERIES])
        return int(ctr[CTR_FOUND]), int(ctr[CTR_NOT_FOUND])

    def query_sequence_hash(self, query: str) -> np.ndarray:
        ids = self.index.query_sequence_host(query)
        self.number_query += len(ids)
        return ids

    def query_sequence_bool(self, query: str):
        f, nf = self.index.query_sequence_bool_host(query)
        self.number_query += f + nf
        return f, nf

    def query_kmer_hash(self, canon: int) -> int:
        self.number_query += 1
        return int(self.index.query_kmers_host(np.array([canon], dtype=np.uint64))[0])

    def query_kmer_bool(self, canon: int) -> bool:
        return self.query_kmer_hash(canon) >= 0
