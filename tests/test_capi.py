"""The C ABI library: loads, exports every declared symbol, fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from blight_b200 import api
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "blight_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|uint64_t|const char\*)\s+(blight_[a-z_0-9]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations parsed"
    L = api.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/blight_b200.h but not exported"
    assert sorted(api.SYMBOLS) == declared


def test_version_and_error_string():
    L = api.lib()
    assert b"blight_b200" in L.blight_version()
    assert L.blight_check_params(31, 8, 5, 3, 6) == api.ERR_INVALID_ARG
    assert b"odd" in L.blight_last_error()


def test_null_arguments_are_rejected():
    L = api.lib()
    assert L.blight_flat_load(None, None) == api.ERR_INVALID_ARG
    assert L.blight_index_upload(None, 0, None) == api.ERR_INVALID_ARG
    assert L.blight_query_kmers(None, None, 1, None, None) == api.ERR_INVALID_ARG


def test_no_cpu_fallback():
    """Without a CUDA device the query path must refuse to run (never silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    flat = common.build_lambda(7, 5, 3, 6)
    with pytest.raises(api.BlightError) as e:
        flat.upload(0)
    assert e.value.code == api.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under blight_b200/ may reference it."""
    pkg = os.path.join(ROOT, "blight_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
