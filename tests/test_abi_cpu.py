"""CPU: the C-ABI shared library loads and exports every symbol include/sequoia_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sequoia_b200.h")).read()
    return sorted(set(re.findall(r"SQ_API\s+[^;(]*?\b(sq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from sequoia_pub_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().sq_version() >= 100


def test_struct_mirrors_have_the_header_field_counts():
    from sequoia_pub_b200 import _lib
    text = open(os.path.join(ROOT, "include", "sequoia_b200.h")).read()
    body = re.search(r"typedef struct sq_gemm_desc \{(.*?)\} sq_gemm_desc;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f for stmt in body.split(";") for f in stmt.split(",") if f.strip()]
    assert len(fields) == len(_lib.GemmDesc._fields_)
    body = re.search(r"typedef struct sq_vis_config \{(.*?)\} sq_vis_config;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    assert len([f for f in body.split(";") if f.strip()]) == len(_lib.VisConfig._fields_)


def test_vis_layout_is_contiguous_per_stage_and_aligned():
    import ctypes as C
    from sequoia_pub_b200 import _lib
    cfg = _lib.VisConfig(2048, 6, 16, 100, 20530)
    L = _lib.lib()
    n = L.sq_vis_param_table_len(C.byref(cfg))
    assert n == 1 + 18 * 6 + 4
    table = (C.c_longlong * n)()
    total = C.c_longlong()
    assert L.sq_vis_param_layout(C.byref(cfg), table, n, C.byref(total)) == 0
    offs = list(table)
    assert offs == sorted(offs) and offs[0] == 0 and all(o % 64 == 0 for o in offs)
    assert 131_246_130 <= total.value < 131_246_130 + 64 * n          # SURVEY §8a A1 parameter count + alignment padding
    assert L.sq_vis_act_bytes(C.byref(cfg), 32) > 0 and L.sq_vis_bwd_bytes(C.byref(cfg), 32) > 0
    bad = _lib.VisConfig(2000, 6, 16, 100, 10)
    assert L.sq_vis_param_table_len(C.byref(bad)) < 0 and b"multiple of 64" in L.sq_last_error()
