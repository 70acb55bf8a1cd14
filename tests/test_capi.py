"""The C-ABI library loads and exports every symbol include/direct_ddp.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol():
    from direct_b200 import capi
    lib = capi.load_library()
    hdr = open(os.path.join(ROOT, "include", "direct_ddp.h")).read()
    declared = sorted(set(re.findall(r"\b(direct_ddp_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in direct_ddp.h but not exported"
    assert set(declared) == set(capi.EXPORTS)
    lib.direct_ddp_version.restype = C.c_int
    assert lib.direct_ddp_version() == 200


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (checked against a tiny C program compiled on the fly)."""
    import subprocess, tempfile
    from direct_b200 import capi
    src = r'''
#include <stdio.h>
#include "direct_ddp.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(direct_ddp_opts), sizeof(direct_ddp_batch), sizeof(direct_ddp_result),
         sizeof(direct_ddp_two_stage), sizeof(direct_ddp_stats), sizeof(direct_ddp_trace_row));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "sz.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [C.sizeof(t) for t in (capi.Opts, capi.Batch, capi.ResultC, capi.TwoStage, capi.Stats, capi.TraceRow)]
    assert sizes == mine


@pytest.mark.skipif(has_gpu(), reason="checks the loud failure on a GPU-less host")
def test_no_cpu_fallback_without_gpu():
    from direct_b200 import capi
    with pytest.raises(capi.DirectDdpError, match="no usable CUDA device"):
        capi.Solver(0, "fp64")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under direct_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "direct_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("oracle_py", "ipddp_oracle", "libemu", "gddp_py", "libgddp_oracle", "voxel_py", "libvoxel_oracle", "libvoxel_ref",
                               "from oracle", "import oracle"):
                    assert needle not in txt, (f, needle)
                for ln in txt.splitlines():   # no source under direct_b200/ includes anything from oracle/ (comments may cite it)
                    assert not (ln.lstrip().startswith("#include") and "oracle" in ln), (f, ln)


@pytest.mark.skipif(has_gpu(), reason="checks the loud failure on a GPU-less host")
def test_dropin_translation_unit_fails_loudly_without_gpu(oracle):
    """The replacement for the reference's ddp_optimizer.cpp compiles against the reference's unchanged header
    (where /root/reference is mounted) and, with no B200 visible, returns -100 instead of computing on the CPU."""
    oracle.build(ref=True)
    if not oracle.dropin_available():
        pytest.skip("the reference's headers are not mounted here")
    from direct_b200.problems import make_batch
    r = oracle.solve_batch(make_batch(2, 4, "box"), use_dropin=True, infeas=1, zero_init=1)
    assert (r.rtn == -100).all() and not r.poly_coeff.any()
