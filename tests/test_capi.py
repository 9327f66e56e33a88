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
    declared = sorted(set(re.findall(r"^(?:int|void\*?|uint64_t|const char\*)\s+(blight_[a-z_0-9]+)\s*\(", hdr, re.M)))
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


CLI = os.path.join(ROOT, "blight_b200", "lib", "bench_blight_b200")


def test_cli_usage_and_no_device(tmp_path):
    """The C++ drop-in (kmer_set_light.hpp behind the reference's bench_blight.cpp option set): usage text without
    arguments; without a GPU the run must fail loudly, never fall back to the CPU."""
    import subprocess
    import torch
    from tests.golden import fixtures
    assert os.path.exists(CLI), "build the cli target (python -c 'import __graft_entry__ as g; g.build()')"
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and "Mandatory arguments" in r.stdout and "-b bit saved" in r.stdout
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = tmp_path / "lambda.fa"
    fa.write_bytes(fixtures.lambda_fasta())
    r = subprocess.run([CLI, "-g", str(fa), "-k", "31", "-m", "7", "-n", "5", "-s", "3", "-b", "6"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "Good kmer" not in r.stdout


@pytest.mark.gpu
def test_cli_lambda_self_query(tmp_path):
    """bench_blight -g lambda_virus.unitigs.fa -k 31 -m 7 -n 5 -s 3 -b 6 (BASELINE configs[0]) through the C++ class:
    Kmer in graph 48,462 / Good kmer 48,462 / Erroneous 0 / Query performed 48,462."""
    import subprocess
    from tests.golden import fixtures
    fa = tmp_path / "lambda.fa"
    fa.write_bytes(fixtures.lambda_fasta())
    r = subprocess.run([CLI, "-g", str(fa), "-k", "31", "-m", "7", "-n", "5", "-s", "3", "-b", "6", "-t", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for line in ("Kmer in graph: 48462", "Super Kmer in graph: 3708", "Good kmer: 48462", "Erroneous kmers: 0", "Query performed: 48462"):
        assert line in r.stdout, (line, r.stdout)
