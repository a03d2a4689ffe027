"""The C-ABI shared library loads and exports every entry point include/pmoe_b200.h declares, and the ctypes
signature table covers them (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pmoe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmoe_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_families():
    syms = declared_symbols()
    for must in ("pmoe_conv_tc", "pmoe_conv_wgrad_tc", "pmoe_bn_finalize", "pmoe_gate_mixture_fwd", "pmoe_gate_mixture_bwd",
                 "pmoe_moe_loss", "pmoe_segloss_fwd", "pmoe_segloss_bwd", "pmoe_mt_adam", "pmoe_last_error", "pmoe_version"):
        assert must in syms
    assert len(syms) >= 35


def test_library_exports_every_declared_symbol():
    from pmoe_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.pmoe_version.restype = ctypes.c_int
    assert lib.pmoe_version() >= 100
    lib.pmoe_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.pmoe_last_error(), bytes)


def test_ctypes_table_matches_header():
    from pmoe_b200 import _sigs
    explicit = {"pmoe_conv_tc", "pmoe_dbg_umma_view", "pmoe_segloss_workspace_floats", "pmoe_version", "pmoe_last_error",
                "pmoe_device_check"}
    syms = set(declared_symbols())
    assert set(_sigs.SIGS) <= syms, sorted(set(_sigs.SIGS) - syms)
    assert syms <= set(_sigs.SIGS) | explicit, sorted(syms - set(_sigs.SIGS) - explicit)


def test_struct_layouts_match_the_header():
    from pmoe_b200 import _lib, optim
    assert ctypes.sizeof(_lib.View4) == 48
    assert ctypes.sizeof(_lib.Seg) == 8
    assert optim._DT.itemsize == 48  # PmoeMtChunk
    # PmoeConvTc: field offsets follow natural C alignment; its size is what the library was compiled with
    assert ctypes.sizeof(_lib.ConvTc) % 8 == 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pmoe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_ctypes_argument_counts_match_the_header():
    """Every entry point's ctypes argtypes list has as many entries as the C declaration has parameters (a drifted signature would
    otherwise only show up as a crash on the GPU box)."""
    from pmoe_b200 import _sigs
    text = open(os.path.join(ROOT, "include", "pmoe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(pmoe_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        decl[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    bad = {n: (len(a), decl[n]) for n, a in _sigs.SIGS.items() if n in decl and len(a) != decl[n]}
    assert not bad, bad
    assert len([n for n in _sigs.SIGS if n in decl]) >= 50
