"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/sparenet_b200.h declares, and the ctypes table matches it.  No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sparenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from sparenet_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_table_matches_header():
    from sparenet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_version_and_strerror_without_gpu():
    from sparenet_b200 import _lib
    lib = _lib.load()
    assert lib.snb_version() >= 100
    assert lib.snb_strerror(0) == b"ok"
    assert b"invalid" in lib.snb_strerror(-1)
    tree = 2 * (8192 + 2 * 8192 // 32 + 2 * 8192 // 512) * 16     # both clouds: sorted points + leaf boxes + super-boxes (float4)
    assert lib.snb_emd_workspace_bytes(32, 8192) == 32 * (tree + 8192 * 32 + 8192 // 32 * 4 + 64)
    assert lib.snb_emd_workspace_bytes(2, 32768) == 2 * (32768 * 40 + 64)            # above 16384 points: exhaustive Bid only
    assert lib.snb_knn_workspace_bytes(2, 128) == 2 * 128 * 128 * 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from sparenet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    from sparenet_b200 import functional as F_
    with pytest.raises(ValueError):
        F_.chamfer_forward(torch.rand(1, 8, 3), torch.rand(1, 8, 3))
    with pytest.raises(ValueError):
        F_.knn_indices(torch.rand(1, 3, 8), 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sparenet_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(dp, f)


def test_integration_snippet_signature_matches_header():
    """INTEGRATION.md section B / docs/integration_snippet_chamfer.py: the ctypes argument list a maintainer would copy must be the
    one include/sparenet_b200.h declares (round 1 shipped a 10-argument example for the 12-argument function)."""
    import importlib.util
    from sparenet_b200 import _lib
    spec = importlib.util.spec_from_file_location("integration_snippet_chamfer", os.path.join(ROOT, "docs", "integration_snippet_chamfer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # loads the library and sets argtypes; no compute
    res, args = _lib.SIGNATURES["snb_chamfer_fwd"]
    assert list(mod._lib.snb_chamfer_fwd.argtypes) == list(args) and mod._lib.snb_chamfer_fwd.restype is res
    assert len(args) == 12
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = open(os.path.join(ROOT, "docs", "integration_snippet_chamfer.py")).read()
    assert body[body.index("import ctypes"):] in doc, "INTEGRATION.md must quote docs/integration_snippet_chamfer.py verbatim"


def test_gemm_desc_struct_matches_header():
    """The ctypes mirror of snb_gemm_desc must list the header's fields in order."""
    from sparenet_b200 import _lib
    src = open(os.path.join(ROOT, "include", "sparenet_b200.h")).read()
    body = src[src.index("typedef struct snb_gemm_desc {"):src.index("} snb_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        head, *rest = decl.split(",")
        names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", head.strip())[0])
        names += [r.strip().lstrip("*") for r in rest]
    assert names == [f[0] for f in _lib.GemmDesc._fields_]
