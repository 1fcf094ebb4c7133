"""C-ABI checks that need no GPU: the product library loads, exports every symbol include/hamers_b200.h
declares, answers the size / count queries like the reference's containers would, and FAILS LOUDLY (no CPU
fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import pytest

from hamers_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hamers_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _declared_symbols() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol(product_lib):
    missing = [s for s in _declared_symbols() if not hasattr(product_lib, s)]
    assert not missing, missing


def _desc(dim, n, model=abi.SINGLE_SPECIES, gam=(1.4,)):
    d = abi.PatchDescC()
    d.dim = dim
    for a in range(3):
        d.n[a] = n[a] if a < dim else 1
        d.dx[a] = 0.1
    d.flow_model = model
    d.num_species = len(gam)
    for i, g in enumerate(gam):
        d.species_gamma[i] = g
    d.weno_p, d.math, d.device = 2, abi.MATH_EXACT, -1
    return d


def test_size_queries_follow_samrai_layouts(product_lib):
    """CellData ghost 4 / SideData ghost 0 extents and FlowModel::getNumberOfEquations."""
    d = _desc(3, (16, 12, 10))
    v = C.c_int32()
    assert product_lib.hb2_num_eqn(C.byref(d), C.byref(v)) == 0 and v.value == 5
    assert product_lib.hb2_num_comp(C.byref(d), C.byref(v)) == 0 and v.value == 5
    assert product_lib.hb2_cell_ghost_size(C.byref(d)) == 24 * 20 * 18
    assert product_lib.hb2_cell_size(C.byref(d)) == 16 * 12 * 10
    assert [product_lib.hb2_side_size(C.byref(d), a) for a in range(3)] == [17 * 12 * 10, 16 * 13 * 10, 16 * 12 * 11]
    g = (C.c_int32 * 3)()
    assert product_lib.hb2_num_ghosts(C.byref(d), g) == 0 and list(g) == [4, 4, 4]
    d2 = _desc(2, (8, 8), abi.FIVE_EQN_ALLAIRE, (1.6, 1.4))
    assert product_lib.hb2_num_eqn(C.byref(d2), C.byref(v)) == 0 and v.value == 6      # d + 2*ns
    assert product_lib.hb2_num_comp(C.byref(d2), C.byref(v)) == 0 and v.value == 7     # + stored Z_last
    assert product_lib.hb2_num_ghosts(C.byref(d2), g) == 0 and list(g) == [4, 4, 0]


def test_descriptor_validation_messages(product_lib):
    h = C.c_void_p()
    bad = _desc(1, (8, 1, 1))
    assert product_lib.hb2_plan_create(C.byref(bad), C.byref(h)) != 0
    assert b"dim" in product_lib.hb2_last_error()
    bad = _desc(3, (8, 8, 8), abi.FIVE_EQN_ALLAIRE, (1.6, 1.4, 1.3))
    assert product_lib.hb2_plan_create(C.byref(bad), C.byref(h)) != 0
    bad = _desc(3, (8, 8, 8), gam=(0.9,))
    assert product_lib.hb2_plan_create(C.byref(bad), C.byref(h)) != 0
    assert b"gamma" in product_lib.hb2_last_error()


def test_no_cpu_fallback_without_a_device(product_lib):
    """Without a CUDA device plan creation must fail with a message -- never compute on the CPU."""
    n = C.c_int32()
    product_lib.hb2_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(abi.HamersB200Error) as ei:
        abi.Plan(3, (8, 8, 8))
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(abi.HamersB200Error) as ei:
        abi.DiffusivePlan(3, (8, 8, 8), (0.1, 0.1, 0.1), 1.4, 2.5, 0.05, 0.0, 3.5, 0.72)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "hamers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "hamers_oracle" not in text, f


def test_descriptor_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of hb2_patch_desc / hb2_diffusive_desc have the C structs' size and the offsets of their last
    members (a C probe compiled against include/hamers_b200.h)."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hamers_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu\\n", sizeof(hb2_patch_desc), offsetof(hb2_patch_desc, num_ghosts),'
                   ' sizeof(hb2_diffusive_desc), offsetof(hb2_diffusive_desc, device)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(abi.PatchDescC), abi.PatchDescC.num_ghosts.offset, C.sizeof(abi.DiffusiveDescC),
                   abi.DiffusiveDescC.device.offset]
