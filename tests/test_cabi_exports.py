"""The C-ABI library loads on a CPU-only box and exports every symbol include/altformer_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import altformer_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ab._lib.lib()
    names = ab._lib.exported_names()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.afb_version() == 100


def test_ctypes_signatures_cover_the_header():
    lib = ab._lib.lib()
    declared = set(ab._lib.exported_names()) - {"afb_last_error"}
    assert declared == set(lib._afb_signatures), declared ^ set(lib._afb_signatures)


def test_header_has_no_torch_types_and_cites_reference():
    txt = open(os.path.join(ROOT, "include", "altformer_b200.h")).read()
    assert "at::" not in txt and "torch::" not in txt and "#include <torch" not in txt
    assert 'extern "C"' in txt
    for cite in ("model/unit_agcn.py:73-93", "model/net.py", "model_ST.py", "train_sttran.py"):
        assert cite in txt, cite


def test_struct_layouts_match_header_field_order():
    txt = open(os.path.join(ROOT, "include", "altformer_b200.h")).read()

    def fields(struct_name):
        body = re.search(r"typedef struct \{([^}]*)\} " + struct_name + ";", txt).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                name = re.sub(r"\[.*?\]", "", part.strip().split()[-1].lstrip("*"))
                out.append(name)
        return out

    for cname, cls in (("afb_gemm_tn_t", ab._lib.GemmTn), ("afb_gemm_dw_t", ab._lib.GemmDw), ("afb_gemm_simt_t", ab._lib.GemmSimt),
                       ("afb_gcn0_fwd_t", ab._lib.Gcn0Fwd), ("afb_gcn0_bwd_t", ab._lib.Gcn0Bwd)):
        assert fields(cname) == [f[0] for f in cls._fields_], cname
    assert ctypes.sizeof(ab._lib.GemmTn) % 8 == 0


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    monkeypatch.setattr(ab._lib, "_lib", None)
    monkeypatch.setattr(ab._lib, "LIB_PATH", "/nonexistent/libaltformer_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        ab._lib.lib()
