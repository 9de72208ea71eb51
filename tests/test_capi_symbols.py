"""The C-ABI library builds, loads, and exports every symbol include/*.h declares; struct
layouts seen from Python match the headers.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("ivs.h", "ivs_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(ivs_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    declared = declared_symbols()
    assert len(declared) >= 50
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding covers the same set
    assert declared == set(capi.SIGNATURES)


def test_struct_layouts_match_headers(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ivs.h"\n#include "ivs_host.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ivs_ring_params), sizeof(ivs_derived),'
                   'sizeof(ivs_sim_desc), offsetof(ivs_ring_params, padding), offsetof(ivs_derived, dE),'
                   'offsetof(ivs_sim_desc, rf_offset));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    from inovesa_b200.params import Derived, RingParams, SimDesc
    want = [C.sizeof(RingParams), C.sizeof(Derived), C.sizeof(SimDesc), RingParams.padding.offset,
            Derived.dE.offset, SimDesc.rf_offset.offset]
    assert got == want


def test_no_device_is_a_loud_error(capi):
    """without a GPU the context cannot be created -- and nothing falls back to a CPU path"""
    lib = capi.load()
    if lib.ivs_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.IvsError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """nothing under inovesa_b200/ may import, link or call oracle/"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "inovesa_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(pyoracle|vfps_oracle|liboracle|libinovesa_ref|from oracle|import oracle|orc_[a-z]+\()", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
