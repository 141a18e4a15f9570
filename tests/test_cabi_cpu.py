"""CPU tests of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/petsyn.h declares; the ctypes table covers them all; descriptors are validated before any device work."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "petsyn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(petsyn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(petsyn):
    from petsyn_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(_cabi.lib, n), f"libpetsyn.so does not export {n}"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_cabi.SIGNATURES) == names
    assert _cabi.lib.petsyn_version() == 100
    assert _cabi.lib.petsyn_launch_count() == 0


def test_struct_layout_matches_header(petsyn):
    import ctypes
    from petsyn_b200 import _cabi
    src = open(os.path.join(ROOT, "include", "petsyn.h")).read()
    body = src[src.index("typedef struct petsyn_conv_desc {"):src.index("} petsyn_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for line in body.splitlines()[1:]:
        m = re.match(r"\s*(int32_t|float)\s+([^;]+);", line)
        if m:
            fields += [f.strip() for f in m.group(2).split(",")]
    assert fields == [f[0] for f in _cabi.ConvDesc._fields_]
    assert ctypes.sizeof(_cabi.ConvDesc) == 4 * len(fields)


def _parse_struct(name):
    """[(field name, ctypes type)] of a struct in include/petsyn.h, in declaration order."""
    import ctypes as C
    src = open(os.path.join(ROOT, "include", "petsyn.h")).read()
    body = src[src.index("typedef struct %s {" % name):src.index("} %s;" % name)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"(const )?(void|float|double|int32_t|int64_t)\s*(\*)?\s*(.+)", stmt)
        assert m, stmt
        ptr, base = m.group(3), m.group(2)
        ctype = C.c_void_p if ptr else {"float": C.c_float, "int32_t": C.c_int32, "int64_t": C.c_int64}[base]
        out += [(f.strip(), ctype) for f in m.group(4).split(",")]
    return out


@pytest.mark.parametrize("cname,attr", [("petsyn_normact_desc", "NormActDesc"), ("petsyn_volume_src", "VolumeSrc")])
def test_descriptor_structs_match_header(petsyn, cname, attr):
    """Field order and types of the ctypes mirrors follow the header exactly (a mismatch would shift every later field)."""
    from petsyn_b200 import _cabi
    want = _parse_struct(cname)
    got = list(getattr(_cabi, attr)._fields_)
    assert [n for n, _ in got] == [n for n, _ in want]
    for (n, a), (_, b) in zip(got, want):
        assert a is b, (n, a, b)


def test_bad_descriptor_is_rejected_without_a_gpu(petsyn):
    """Validation happens before any CUDA call, so the ValueError contract is testable on a CPU box."""
    import ctypes as C
    from petsyn_b200 import _cabi
    d = _cabi.ConvDesc(_cabi.OP_CONV, 1, 8, 8, 8, 60, 64, 3, 1, 1, 60, 0, 64, 0, 64, 0, 60, 0, 0, 0.2, 0)
    h = C.c_void_p()
    rc = _cabi.lib.petsyn_conv_plan_create(C.byref(d), C.byref(h))
    assert rc == _cabi.E_INVAL and "multiples of 8" in _cabi.last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, "conv_plan_create")


def test_no_cpu_fallback(petsyn):
    import torch
    m = petsyn.UnetGenerator3d(1, 1, num_downs=4, ngf=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 16, 16, 16))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            petsyn.ops.ConvPlan(petsyn.ops.OP_CONV, 1, 8, 8, 8, 64, 64, 3, 1, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "causality-informed-pet-synthesis-from-multi-modal-data_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"


def test_tools_and_bench_gpu_arms_never_import_the_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may use oracle/: tools/ and petsyn.py must not mention it, and in
    bench.py every ``oracle`` import sits inside a ``cpu_*`` function (the CPU arms) or ``reference_*`` (the loader of the
    unmodified reference class from baseline/_ref, which needs the MONAI stub; used by the CPU arm and the same-GPU
    PyTorch/cuDNN incumbent leg -- baselines measured BESIDE the product, never on its path)."""
    import ast
    for rel in [os.path.join("tools", f) for f in os.listdir(os.path.join(ROOT, "tools"))] + ["petsyn.py"]:
        if rel.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, rel)).read().lower(), f"{rel} mentions the oracle"
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").startswith("oracle"):
                assert fn.name.startswith(("cpu_", "reference_")), f"bench.py:{fn.name} imports the oracle"
    for node in tree.body:
        assert not (isinstance(node, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(node)), "module-level oracle import"
