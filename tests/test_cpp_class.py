"""The C++ drop-in class (kmer_set_light.hpp) used like the reference's kmer_Set_Light: a small C++ program is compiled
against the in-tree library and run; its facts are checked against the reference's behaviour (exceptions, blight.h:75-92,
kmer.h:68, blight.cpp:188-189) and, on the GPU, against the oracle's ids for the same read."""
import os
import subprocess

import numpy as np
import pytest

from tests import common
from tests.golden import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "class_check")
    lib = os.path.join(ROOT, "blight_b200", "lib")
    subprocess.check_call(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "class_check.cpp"), "-o", exe,
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "blight_b200", "csrc"),
                           "-L" + lib, "-lblight_b200", "-Wl,-rpath," + lib])
    return exe


def _facts(out: str):
    d = {}
    for line in out.strip().splitlines():
        k, _, v = line.partition(" ")
        d[k] = v
    return d


def test_class_exceptions_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    fa = tmp_path / "lambda.fa"
    fa.write_bytes(fixtures.lambda_fasta())
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe, "cpu", str(fa)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    f = _facts(r.stdout)
    for key in ("invalid_even_m", "invalid_big_n", "valid_params", "query_before_index", "missing_file", "no_device_is_an_error"):
        assert f[key] == "1", (key, r.stdout)


def _devices(n_sim=3):
    """Every GPU of the box, or GPU 0 listed n_sim times on a one-GPU box (the ranks are then separate indices / sessions
    on one device: same code, same flags, no NVLink)."""
    import torch
    n = torch.cuda.device_count()
    return [str(i) for i in range(n)] if n > 1 else ["0"] * n_sim


@pytest.mark.gpu
@pytest.mark.parametrize("comm", [None, "replica", "partition"])
def test_class_on_lambda(tmp_path, comm):
    """comm None: one device. replica / partition: the same object spread over several devices by ONE process
    (kmer_Set_Light::use_devices -> blight_comm, csrc/comm.cu): every fact must be unchanged."""
    exe = _build(tmp_path)
    fa = tmp_path / "lambda.fa"
    fa.write_bytes(fixtures.lambda_fasta())
    r = subprocess.run([exe, "gpu", str(fa)] + ([comm] + _devices() if comm else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    f = _facts(r.stdout)
    for key in ("invalid_even_m", "invalid_big_n", "valid_params", "query_before_index", "missing_file", "short_read_empty",
                "domain_error_on_N", "kmer_hash_equals_first_id", "kmer_bool"):
        assert f[key] == "1", (key, r.stdout)
    assert f["number_kmer"] == "48462" and f["number_super_kmer"] == "3708"
    # the same read through the oracle
    flat = common.build_lambda(7, 5, 3, 6)
    port = common.cport_of(flat, tmp_path)
    bases, offs = fixtures.lambda_unitigs()
    u = int(f["unitig"])
    assert int(offs[u + 1] - offs[u]) >= 250 and all(int(offs[i + 1] - offs[i]) < 250 for i in range(u))
    want = port.query_sequence(bases[int(offs[u]) + 100:int(offs[u]) + 250])
    assert f["ids_n"] == "120" and int(f["ids_first"]) == int(want[0]) and int(f["ids_last"]) == int(want[-1])
    assert f["bool"] == "120 0" and f["absent_kmer"] == "-1"
    assert f["file_query"] == "48462 0"
    # number_query: 120 (hash) + 120 (bool) + 0 (short read) + 3 k-mer queries + 48462 (file_query); the failed N query adds none
    assert int(f["number_query"]) == 120 + 120 + 3 + 48462
