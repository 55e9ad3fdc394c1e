"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/vrod_knn.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vrod_knn.h")).read()
    return sorted(set(re.findall(r"VROD_API\s+[^;(]*?\b(vrod_\w+)\s*\(", src)))


def test_header_symbol_list_matches_binding():
    from vrod_b200 import ffi
    assert declared_symbols() == sorted(ffi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from vrod_b200 import ffi
    assert os.path.exists(ffi.LIB_PATH), "libvrod_knn.so is not built: run `make` (or __graft_entry__.build())"
    L = ctypes.CDLL(ffi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} is declared in include/vrod_knn.h but not exported"
    assert b"sm_100a" in ctypes.cast(L.vrod_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_no_cpu_fallback():
    """Without a CUDA device the library refuses to create a context instead of computing on the CPU."""
    import torch
    from vrod_b200 import ffi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ffi.VrodError) as e:
        ffi.Context(0)
    assert e.value.status == ffi.ENOGPU


def test_product_does_not_touch_the_oracle():
    """Nothing under vrod_b200/ or include/ may import, link, load or include anything of oracle/
    (comments may cite it); the oracle is the checker only."""
    banned = [r"^\s*(import|from)\s+oracle", r"libvrod_oracle", r"#include\s+[\"<][^\">]*oracle", r"oracle\.py",
              r"-loracle", r"vrod_oracle_\w+\s*\("]
    for base in ("vrod_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    for pat in banned:
                        assert not re.search(pat, text, flags=re.M), f"{dirpath}/{f} matches {pat}"
