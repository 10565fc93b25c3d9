"""CPU-side checks of the drop-in boundary: libwbem.so loads, exports every symbol that
include/wbem.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "wbem.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wbem_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(wb):
    lib = wb.lib()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wbem.h but not exported"
    assert sorted(wb.ABI_SYMBOLS) == declared


def test_struct_layouts_match_header(wb):
    # sizes the C compiler gives the two ABI structs
    src = '#include "wbem.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n",sizeof(wbem_params),sizeof(wbem_timings));return 0;}\n'
    exe = "/tmp/wbem_sizeof"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src.encode(), check=True)
    a, b = map(int, subprocess.check_output([exe]).split())
    assert a == C.sizeof(wb.Params) and b == C.sizeof(wb.Timings)


def test_default_params_are_the_prm_defaults(wb):
    p = wb.default_params()
    assert (p.quad_order, p.sing_order) == (4, 5)            # prm-files/default.prm:218-220
    assert p.gmres_tol == 1e-16 and p.gmres_max_steps == 200  # prm-files/default.prm:228-229
    assert p.gmres_n_tmp_vectors == 100 and p.preconditioner_band == 100  # bem_problem.cc:68, 826
    assert p.world_size == 1 and p.rank == 0


def test_no_cpu_fallback(wb):
    """Without a CUDA device wbem_create must fail loudly; nothing can be computed."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wb.WbemError, match="no CUDA device|CPU fallback"):
        wb.Context()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wavebem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|oracle\.|liboracle|orc_[a-z]", text), f
