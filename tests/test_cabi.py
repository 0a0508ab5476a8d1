"""CPU: the C-ABI library loads, exports every symbol include/graphbix_cuda.h declares, and fails loudly (no CPU
fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "graphbix_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gbx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from graphbix_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/graphbix_cuda.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"
    assert lib.gbx_version() >= 100


def test_struct_layouts_match_header():
    from graphbix_b200 import _lib
    opt = _lib.SbmOptions()
    _lib.load().gbx_sbm_default_options(ctypes.byref(opt))
    assert (opt.degreecorrection, opt.priorlearning, opt.itrmax, opt.check_every) == (1, 1, 256, 1)
    assert opt.learningrate == pytest.approx(0.3) and opt.bp_conv_threshold == pytest.approx(1e-6)
    assert opt.damping == 1.0 and opt.maxCrs == 0.0
    lopt = _lib.LsbmOptions()
    _lib.load().gbx_lsbm_default_options(ctypes.byref(lopt))
    assert (lopt.dc, lopt.learning, lopt.priorlearning, lopt.itrmax) == (1, 1, 1, 512)
    assert lopt.learningrate == pytest.approx(0.3) and lopt.bp_conv_threshold == pytest.approx(1e-6)
    assert lopt.REDfrac == 0.0 and lopt.damping == 1.0


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must raise; nothing may quietly run on the host."""
    from graphbix_b200 import _lib
    lib = _lib.load()
    if lib.gbx_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from graphbix_b200.engine import EM, SbmEngine
    links = np.array([[1, 2], [2, 1]], dtype=np.int64)
    with pytest.raises(_lib.GbxError):
        SbmEngine(2, links, 2)
    with pytest.raises(_lib.GbxError):
        EM(2, 2, links, 2, "uniformAssortative", True, 4, True, 0.3)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under graphbix_b200/ may import or execute it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "graphbix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "oracle/" not in src and "sbm_oracle" not in src, os.path.join(dirpath, f)


def _header_structs():
    """{struct name: [(ctype, field), ...]} of the typedef'd structs of include/graphbix_cuda.h, in declaration order."""
    text = open(os.path.join(ROOT, "include", "graphbix_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s+\w+\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(unsigned long long|uint64_t|int64_t|double|int)\s+(.*)$", decl, flags=re.S)
            assert m, f"unparsed declaration in {name}: {decl!r}"
            fields += [(m.group(1), f.strip()) for f in m.group(2).split(",")]
        out[name] = fields
    return out


def _julia_structs():
    """{struct name: [(type, field), ...]} of integration/graphbix_cuda.jl."""
    text = open(os.path.join(ROOT, "integration", "graphbix_cuda.jl")).read()
    out = {}
    for name, body in re.findall(r"^(?:mutable\s+)?struct\s+(\w+)[^\n]*\n(.*?)\nend", text, flags=re.S | re.M):
        body = re.sub(r"#[^\n]*", "", body)
        out[name] = [(t, f) for f, t in re.findall(r"(\w+)::(\w+)", body)]
    return out


def test_julia_binding_matches_header():
    """integration/graphbix_cuda.jl cannot be run here (no Julia): at least its struct layouts and the symbols it
    `ccall`s must be the header's."""
    hs, js = _header_structs(), _julia_structs()
    ctype = {"Cint": "int", "Cdouble": "double", "Int64": "int64_t", "UInt64": "uint64_t"}
    pairs = {"GbxSbmOptions": "gbx_sbm_options", "GbxSbmResult": "gbx_sbm_result", "GbxModOptions": "gbx_mod_options",
             "GbxModResult": "gbx_mod_result", "GbxLsbmOptions": "gbx_lsbm_options", "GbxLsbmResult": "gbx_lsbm_result"}
    for jl, c in pairs.items():
        assert jl in js and c in hs, (jl, c)
        assert [(ctype[t], f) for t, f in js[jl]] == hs[c], f"{jl} and struct {c} differ"
    text = open(os.path.join(ROOT, "integration", "graphbix_cuda.jl")).read()
    called = set(re.findall(r"ccall\(\(:(\w+),\s*libgbx\)", text))
    assert {"gbx_sbm_create", "gbx_sbm_em", "gbx_mod_em", "gbx_lsbm_create_batch", "gbx_lsbm_em"} <= called
    assert called <= set(_header_symbols()), called - set(_header_symbols())


def test_julia_ccall_argument_counts_match_header():
    """Every `ccall` of the Julia binding passes as many arguments as the header's prototype takes, and pointer
    arguments where the prototype has pointers."""
    htext = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "graphbix_cuda.h")).read(), flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(gbx_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", htext):
        args = " ".join(args.split())
        protos[name] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    jtext = open(os.path.join(ROOT, "integration", "graphbix_cuda.jl")).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libgbx\),\s*\w+,\s*\(", jtext):
        depth, i = 1, m.end()
        while depth:                      # the balanced end of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(jtext[i], 0)
            i += 1
        calls.append((m.group(1), jtext[m.end():i - 1]))
    assert len(calls) >= 20
    for name, tup in calls:
        jargs = [a.strip() for a in " ".join(tup.split()).split(",") if a.strip()]
        assert name in protos, name
        assert len(jargs) == len(protos[name]), f"{name}: ccall passes {len(jargs)} arguments, the header takes {len(protos[name])}"
        for ja, ha in zip(jargs, protos[name]):
            is_ptr = ja.startswith(("Ptr{", "Ref{")) or ja == "Cstring"
            assert is_ptr == ("*" in ha), f"{name}: {ja} against `{ha}`"
