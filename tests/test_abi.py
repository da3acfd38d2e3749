"""The C-ABI library loads without a GPU and exports every symbol include/ktn.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ktn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ktn_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("ktn_create", "ktn_load_begin", "ktn_add_rows", "ktn_load_end", "ktn_separate", "ktn_fetch_cuts", "ktn_gencut_rows",
              "ktn_get_g", "ktn_comm_init", "ktn_allgather_cuts_async", "ktn_synth_rows"):
        assert s in syms


def test_cuda_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "katana.jl_b200", "libktn.so")
    assert os.path.exists(path), "libktn.so missing: run __graft_entry__.build()"
    dll = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, missing
    dll.ktn_backend.restype = ctypes.c_char_p
    assert dll.ktn_backend() == b"cuda"


def test_oracle_exports_the_same_abi(oracle_lib):
    synth = {"ktn_synth_rows", "ktn_synth_point"}
    missing = [s for s in declared_symbols() if s not in synth and not hasattr(oracle_lib.dll, s)]
    assert not missing, missing
    assert oracle_lib.backend == "oracle"


def test_binding_matches_header():
    from katana_jl_b200.binding import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == declared_symbols()


def test_product_has_no_cpu_fallback():
    """The product package never references the oracle or the emulator."""
    pkg = os.path.join(ROOT, "katana.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libktn_oracle" not in src and "libktn_emu" not in src and "oracle/" not in src.replace("oracle/ktn_oracle.c", ""), f


def test_julia_shim_binds_exported_symbols_and_matches_the_guide():
    """julia/gpu_separator.jl (the reference-side ccall binding; it cannot run here, there is no julia) names only symbols the
    library exports, passes as many arguments as the header declares, and is the code INTEGRATION.md prints."""
    shim = open(os.path.join(ROOT, "julia", "gpu_separator.jl")).read()
    names = sorted(set(re.findall(r"ccall\(\(:(\w+)", shim)))
    assert len(names) >= 8
    dll = ctypes.CDLL(os.path.join(ROOT, "katana.jl_b200", "libktn.so"))
    assert all(n in declared_symbols() and hasattr(dll, n) for n in names), names
    # argument counts: the ccall's type tuple against the C prototype
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ktn.h")).read(), flags=re.S)
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libktn\),\s*\w+,\s*\(([^)]*)\)", shim):
        name, types = m.group(1), [t for t in m.group(2).split(",") if t.strip()]
        proto = re.search(r"\b%s\s*\(([^)]*)\)" % name, hdr)
        assert proto, name
        nargs = 0 if proto.group(1).strip() in ("", "void") else len(proto.group(1).split(","))
        assert len(types) == nargs, (name, types, proto.group(1))
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```julia\n(.*?)```", guide, flags=re.S)
    for b in blocks:
        if b.count("\n") > 5:
            assert b.strip() in shim
