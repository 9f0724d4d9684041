"""CPU: the C-ABI library loads, exports every symbol include/vdjgraph.h declares, agrees with the
ctypes structs, and fails loudly (no CPU fallback) when there is no GPU.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from vdjer_b200 import graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vdjgraph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vdjgraph_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    lib = graph.load_library()
    names = _declared()
    assert set(names) == set(graph.EXPORTS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vdjgraph.h but not exported"
    assert lib.vdjgraph_version() == 5


def test_struct_layouts_match_header(built):
    # compile a tiny C program against the header and compare sizeof/offsetof with ctypes
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "vdjgraph.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(vdjgraph_params), sizeof(vdjgraph_result), sizeof(vdjgraph_pre_table),
   offsetof(vdjgraph_params, table_capacity), offsetof(vdjgraph_result, n_records), offsetof(vdjgraph_result, ms_stage),
   offsetof(vdjgraph_result, kernel_launches));
 return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(graph._Params), C.sizeof(graph._Result), C.sizeof(graph._PreTable),
            graph._Params.table_capacity.offset, graph._Result.n_records.offset, graph._Result.ms_stage.offset,
            graph._Result.kernel_launches.offset]
    assert got == want


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", graph.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_param_validation(built):
    lib = graph.load_library()
    ctx = C.c_void_p()
    for L, k in [(50, 51), (50, 0), (300, 35), (30, 35), (0, 1)]:
        p = graph._Params(L, k, 3, 90, -1, 0, 0, 0, 0)
        assert lib.vdjgraph_create(C.byref(p), C.byref(ctx)) == -1
        assert lib.vdjgraph_last_error()
    assert lib.vdjgraph_create(None, C.byref(ctx)) == -1
    for rounds in (3, 512):     # not a power of two / more than the 256 hash buckets
        p = graph._Params(50, 35, 3, 90, -1, 0, 0, 0, 0, rounds, 0)
        assert lib.vdjgraph_create(C.byref(p), C.byref(ctx)) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu(built):
    with pytest.raises(graph.VdjGraphError) as e:
        graph.GraphBuilder(50, 35, 3, 90)
    assert e.value.code == -5  # VDJGRAPH_ERR_CUDA


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_host_alloc_fails_loudly_without_gpu(built):
    lib = graph.load_library()
    p = C.c_void_p()
    assert lib.vdjgraph_host_alloc(1 << 20, C.byref(p)) < 0 and not p.value
    assert lib.vdjgraph_last_error()
    assert lib.vdjgraph_host_free(None) == 0


def test_product_never_touches_oracle():
    """Nothing under vdjer_b200/ or include/ may import, link or mention oracle/."""
    for base in ("vdjer_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")) or f == "Makefile":
                    text = open(os.path.join(dp, f)).read()
                    assert "oracle" not in text.lower(), os.path.join(dp, f)
