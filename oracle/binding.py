"""ctypes bindings of the two CPU oracles (test infrastructure, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CPORT_SO = os.path.join(_HERE, "libblight_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libblight_ref.so")


def build_cport(force: bool = False) -> str:
    src = os.path.join(_HERE, "blight_oracle.c")
    if force or not os.path.exists(CPORT_SO) or os.path.getmtime(CPORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-march=x86-64-v3", "-o", CPORT_SO, src])
    return CPORT_SO


def build_reference() -> bool:
    """Builds oracle/_ref/libblight_ref.so when /root/reference is present; keeps a prebuilt one otherwise."""
    subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return os.path.exists(REF_SO)


def reference_available() -> bool:
    return os.path.exists(REF_SO)


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        return np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


class CPort:
    """The plain-C restatement, on a BLFLAT01 blob."""

    def __init__(self, blob_path: str):
        L = C.CDLL(build_cport())
        vp, u64 = C.c_void_p, C.c_uint64
        L.blo_load.restype = vp
        L.blo_load.argtypes = [C.c_char_p]
        L.blo_free.argtypes = [vp]
        L.blo_free.restype = None
        L.blo_k.argtypes = [vp]; L.blo_k.restype = C.c_uint32
        L.blo_m.argtypes = [vp]; L.blo_m.restype = C.c_uint32
        L.blo_number_kmer.argtypes = [vp]; L.blo_number_kmer.restype = u64
        L.blo_minimizer.argtypes = [u64, C.c_uint, C.c_uint]; L.blo_minimizer.restype = C.c_uint32
        L.blo_query_kmers_hash.argtypes = [vp, vp, u64, vp]; L.blo_query_kmers_hash.restype = None
        L.blo_query_get_hash.argtypes = [vp, u64, C.c_uint32]; L.blo_query_get_hash.restype = C.c_int64
        L.blo_query_sequence_hash.argtypes = [vp, vp, u64, vp, vp, vp]; L.blo_query_sequence_hash.restype = C.c_int64
        L.blo_query_reads.argtypes = [vp, vp, vp, u64, vp, vp, vp]; L.blo_query_reads.restype = C.c_int
        L.blo_query_reads_mt.argtypes = [vp, vp, vp, u64, vp, vp, vp, C.c_int]; L.blo_query_reads_mt.restype = C.c_int
        self.L = L
        self.h = L.blo_load(os.fsencode(blob_path))
        if not self.h:
            raise IOError(f"cannot load {blob_path}")
        self.k = int(L.blo_k(self.h))
        self.m = int(L.blo_m(self.h))
        self.number_kmer = int(L.blo_number_kmer(self.h))

    def minimizer(self, canon: int) -> int:
        return int(self.L.blo_minimizer(int(canon), self.k, self.m))

    def query_kmers(self, canon: np.ndarray) -> np.ndarray:
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        out = np.empty(len(canon), dtype=np.int64)
        self.L.blo_query_kmers_hash(self.h, canon.ctypes.data, len(canon), out.ctypes.data)
        return out

    def query_get_hash(self, canon: int, minimizer: int) -> int:
        return int(self.L.blo_query_get_hash(self.h, int(canon), int(minimizer)))

    def query_sequence(self, seq, with_kmers: bool = False):
        buf = _u8(seq)
        n = max(len(buf) - self.k + 1, 0)
        ids = np.empty(n, dtype=np.int64)
        canon = np.empty(n, dtype=np.uint64) if with_kmers else None
        mini = np.empty(n, dtype=np.uint32) if with_kmers else None
        got = self.L.blo_query_sequence_hash(self.h, buf.ctypes.data, len(buf), ids.ctypes.data,
                                             canon.ctypes.data if with_kmers else None, mini.ctypes.data if with_kmers else None)
        if got < 0:
            raise ValueError("Invalid char in DNA")
        if with_kmers:
            return ids[:got], canon[:got], mini[:got]
        return ids[:got]

    def query_reads(self, bases, read_off, want_ids: bool = True, threads: int = 1):
        bases = _u8(bases)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        lens = np.diff(read_off.astype(np.int64))
        nk = np.maximum(lens - (self.k - 1), 0)
        koff = np.zeros(len(lens) + 1, dtype=np.uint64)
        np.cumsum(nk, out=koff[1:])
        ids = np.empty(int(koff[-1]), dtype=np.int64) if want_ids else None
        ctr = np.zeros(3, dtype=np.uint64)
        rc = self.L.blo_query_reads_mt(self.h, bases.ctypes.data, read_off.ctypes.data, len(lens),
                                       ids.ctypes.data if want_ids else None, koff.ctypes.data, ctr.ctypes.data, int(threads))
        if rc != 0:
            raise ValueError("Invalid char in DNA")
        return ids, ctr

    def __del__(self):
        if getattr(self, "h", None):
            self.L.blo_free(self.h)
            self.h = None


class Reference:
    """The reference's own kmer_Set_Light (+P1+P2) behind oracle/ref_harness.cpp."""

    _L = None

    @classmethod
    def lib(cls):
        if cls._L is None:
            L = C.CDLL(REF_SO)
            vp, u64, ui = C.c_void_p, C.c_uint64, C.c_uint
            L.blref_create.restype = vp; L.blref_create.argtypes = [ui] * 6
            L.blref_destroy.argtypes = [vp]; L.blref_destroy.restype = None
            L.blref_construct.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
            for f in ("number_kmer", "number_super_kmer", "number_query", "largest_mphf", "largest_bucket"):
                getattr(L, "blref_" + f).argtypes = [vp]
                getattr(L, "blref_" + f).restype = u64
            L.blref_minimizer.argtypes = [u64, ui, ui]; L.blref_minimizer.restype = C.c_uint32
            L.blref_query_sequence_hash.argtypes = [vp, vp, u64, vp, u64]; L.blref_query_sequence_hash.restype = C.c_int64
            L.blref_query_sequence_bool.argtypes = [vp, vp, u64, vp, vp]
            L.blref_query_kmers_hash.argtypes = [vp, vp, u64, vp, C.c_int]; L.blref_query_kmers_hash.restype = None
            L.blref_query_reads.argtypes = [vp, vp, vp, u64, C.c_int, vp, vp, vp, vp]; L.blref_query_reads.restype = C.c_double
            L.blref_file_query.argtypes = [vp, C.c_char_p, C.c_int]
            L.blref_export.argtypes = [vp, C.c_char_p]
            L.blref_import.argtypes = [C.c_char_p, ui]; L.blref_import.restype = vp
            L.blref_max_threads.restype = C.c_int
            cls._L = L
        return cls._L

    def __init__(self, k=31, m=9, n=17, s=6, cores=1, b=6, _handle=None):
        self.L = self.lib()
        self.k, self.m = k, m
        self.h = _handle if _handle is not None else self.L.blref_create(k, m, n, s, cores, b)
        if not self.h:
            raise ValueError("std::invalid_argument from kmer_Set_Light constructor")

    @classmethod
    def from_blob(cls, path: str, k: int, m: int, cores: int = 1) -> "Reference":
        L = cls.lib()
        h = L.blref_import(os.fsencode(path), cores)
        if not h:
            raise IOError(f"cannot import {path}")
        return cls(k=k, m=m, _handle=h)

    def construct_index(self, unitig_path: str, quiet: bool = True):
        with tempfile.TemporaryDirectory() as wd:  # the reference drops _out<i> temp files in the CWD (blight.cpp:132)
            rc = self.L.blref_construct(self.h, os.fsencode(os.path.abspath(unitig_path)), os.fsencode(wd), int(quiet))
        if rc != 0:
            raise RuntimeError(f"reference construct_index failed ({rc})")

    @property
    def number_kmer(self):
        return int(self.L.blref_number_kmer(self.h))

    @property
    def number_super_kmer(self):
        return int(self.L.blref_number_super_kmer(self.h))

    def minimizer(self, canon: int) -> int:
        return int(self.L.blref_minimizer(int(canon), self.k, self.m))

    def export(self, path: str):
        if self.L.blref_export(self.h, os.fsencode(path)) != 0:
            raise IOError(f"export to {path} failed")

    def query_sequence(self, seq) -> np.ndarray:
        buf = _u8(seq)
        out = np.empty(max(len(buf) - self.k + 1, 0) + 1, dtype=np.int64)
        got = self.L.blref_query_sequence_hash(self.h, buf.ctypes.data, len(buf), out.ctypes.data, len(out))
        if got == -1:
            raise ValueError("Invalid char in DNA")
        return out[:got].copy()

    def query_kmers(self, canon: np.ndarray, threads: int = 1) -> np.ndarray:
        canon = np.ascontiguousarray(canon, dtype=np.uint64)
        out = np.empty(len(canon), dtype=np.int64)
        self.L.blref_query_kmers_hash(self.h, canon.ctypes.data, len(canon), out.ctypes.data, threads)
        return out

    def query_reads(self, bases, read_off, threads: int = 1, want_ids: bool = True):
        """Returns (ids or None, found, not_found, seconds)."""
        bases = _u8(bases)
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        lens = np.diff(read_off.astype(np.int64))
        nk = np.maximum(lens - (self.k - 1), 0)
        koff = np.zeros(len(lens) + 1, dtype=np.uint64)
        np.cumsum(nk, out=koff[1:])
        ids = np.empty(int(koff[-1]), dtype=np.int64) if want_ids else None
        f, nf = C.c_uint64(), C.c_uint64()
        sec = self.L.blref_query_reads(self.h, bases.ctypes.data, read_off.ctypes.data, len(lens), threads,
                                       ids.ctypes.data if want_ids else None, koff.ctypes.data, C.byref(f), C.byref(nf))
        return ids, int(f.value), int(nf.value), float(sec)

    def file_query(self, path: str, quiet: bool = True):
        return self.L.blref_file_query(self.h, os.fsencode(path), int(quiet))

    def max_threads(self) -> int:
        return int(self.L.blref_max_threads())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.blref_destroy(self.h)
            self.h = None
