"""CPU-side checks of the product boundary: the C-ABI library loads and exports every symbol
include/hjb200.h declares, struct layouts match, the relation file format round-trips, the
numpy generator mirror has the reference generator's properties, and the product fails loudly
without a GPU.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import hash_join_codes_knl_b200 as hj
from hash_join_codes_knl_b200 import _lib, api, build, datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hjb200.h")


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(hjb_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hjb_version() == 100


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "hjb200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   "sizeof(hjb_rel),sizeof(hjb_opts),sizeof(hjb_result),sizeof(hjb_split),sizeof(hjb_gen));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_lib.Rel), C.sizeof(_lib.Opts), C.sizeof(_lib.Result), C.sizeof(_lib.Split),
                     C.sizeof(_lib.Gen)]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = C.c_void_p()
    assert lib.hjb_create(0, C.byref(ctx)) == -5          # HJB_E_NODEVICE
    assert b"no CPU fallback" in lib.hjb_last_error(None)
    with pytest.raises(hj.HjbError):
        hj.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hash_join_codes_knl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hj_oracle" not in text and "liboracle" not in text and "_oracle" not in text, f


def test_relation_files_roundtrip_and_match_reference_format(lib, tmp_path):
    from _oracle import lib as olib, _p
    k = np.arange(1, 1001, dtype=np.uint32) * np.uint32(2654435761)
    v = k * np.uint32(datagen.INNER_FACTOR)
    api.relation_write(tmp_path, False, k, v)
    api.relation_write(tmp_path, True, v, k)
    # names and bytes of write.cpp:1824-1865
    assert sorted(os.listdir(tmp_path)) == ["ik_1000.txt", "iv_1000.txt", "ok_1000.txt", "ov_1000.txt"]
    assert (np.fromfile(tmp_path / "ik_1000.txt", dtype="<u4") == k).all()
    k2, v2 = api.relation_read(tmp_path, False, 1000)
    assert (k2 == k).all() and (v2 == v).all()
    # the oracle's reader (restating npj.cpp:1013-1039) reads what the product wrote
    ok, ov = np.empty(1000, np.uint32), np.empty(1000, np.uint32)
    assert olib().hjo_relation_read(str(tmp_path).encode(), b"o", 1000, _p(ok), _p(ov)) == 0
    assert (ok == v).all() and (ov == k).all()
    with pytest.raises(hj.HjbError):
        api.relation_read(tmp_path, False, 999)          # no such file
    os.rename(tmp_path / "ik_1000.txt", tmp_path / "ik_999.txt")
    os.rename(tmp_path / "iv_1000.txt", tmp_path / "iv_999.txt")
    with pytest.raises(hj.HjbError):
        api.relation_read(tmp_path, False, 999)          # size != 4 * tuples


@pytest.mark.parametrize("n", [1, 7, 1000, 1 << 16, 100003])
def test_generator_unique_keys(n):
    rk, rv = datagen.generate(0, n, n, 42, 1, datagen.INNER_FACTOR)
    sk, sv = datagen.generate(0, n, n, 42, 2, datagen.OUTER_FACTOR)
    assert np.unique(rk).size == n and (rk != 0).all()
    assert (np.sort(rk) == np.sort(sk)).all()            # same key set, different order
    assert n < 64 or (rk != sk).any()
    assert (rv == rk * np.uint32(datagen.INNER_FACTOR)).all()
    # sharded generation = slices of the whole
    a, _ = datagen.generate(0, n - n // 3, n, 42, 1, datagen.INNER_FACTOR, first=n // 3, total=n)
    assert (a == rk[n // 3:]).all()


def test_generator_foreign_keys():
    nr, ns = 1000, 50000
    rk, _ = datagen.generate(0, nr, nr, 7, 1, datagen.INNER_FACTOR)
    sk, _ = datagen.generate(1, ns, nr, 7, 2, datagen.OUTER_FACTOR)
    assert np.isin(sk, rk).all()
    assert np.unique(sk).size == nr                       # every build key at least once (cpra2.cpp:1639-1646)
    counts = np.unique(sk, return_counts=True)[1]
    assert counts.max() < 5 * ns / nr                     # the rest uniform


def test_workloads_named_in_baseline():
    assert datagen.workload("phj_cfg2") == (1 << 27, 1 << 27, 0)
    assert datagen.workload("npj_cfg1") == (1 << 24, 1 << 28, 1)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: the reference's own CPU join (oracle/_ref when it was
    compiled here, else the oracle port) on the host cores, one JSON line with the contract's keys"""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "tuples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("PHJ 2^27 x 2^27")
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
