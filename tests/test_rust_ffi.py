"""bindings/rust/cuda_ffi.rs is generated from include/egobox_gpu.h (tools/gen_rust_ffi.py).  No Rust toolchain exists in
the build image, so these tests hold the file to the header instead: the committed file equals a fresh generation, every
exported egx_* symbol is declared exactly once with the arity the ctypes binding uses for the same symbol, the #[repr(C)]
structs carry the fields of the C structs in order, and the hand-written wrapper only calls declared functions."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_rust_ffi  # noqa: E402

from egobox_b200 import _lib  # noqa: E402

RS = os.path.join(ROOT, "bindings", "rust", "cuda_ffi.rs")
WRAPPER = os.path.join(ROOT, "bindings", "rust", "cuda_backend.rs")


def _rust_functions():
    src = open(RS).read()
    block = src[src.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (egx_[a-z0-9_]+)\(([^)]*)\)( -> [^;]+)?;", block):
        args = [a for a in m.group(2).split(",") if a.strip()]
        assert m.group(1) not in out, "declared twice: " + m.group(1)
        out[m.group(1)] = (len(args), m.group(3))
    return out


def test_committed_file_is_a_fresh_generation():
    consts, structs, funcs = gen_rust_ffi.parse(open(gen_rust_ffi.HEADER).read())
    assert gen_rust_ffi.emit(consts, structs, funcs) == open(RS).read(), "run python tools/gen_rust_ffi.py"


def test_every_header_symbol_is_declared_with_the_ctypes_arity():
    rust = _rust_functions()
    assert set(rust) == set(_lib.SIGNATURES), (set(rust) ^ set(_lib.SIGNATURES))
    for name, (restype, argtypes) in _lib.SIGNATURES.items():
        nargs, ret = rust[name]
        assert nargs == len(argtypes), name
        assert (ret is None) == (restype is None), name


def test_constants_and_structs_follow_the_header():
    header = gen_rust_ffi.strip_comments(open(gen_rust_ffi.HEADER).read())
    src = open(RS).read()
    for name, value in re.findall(r"^#define (EGX_[A-Z0-9_]+)\s+(-?\d+)\s*$", header, flags=re.M):
        assert "pub const %s: c_int = %s;" % (name, value) in src
    assert "pub const EGX_NUM_STAGES: c_int = %d;" % _lib.NUM_STAGES in src
    _, structs, _ = gen_rust_ffi.parse(open(gen_rust_ffi.HEADER).read())
    assert {s for s, _ in structs} == {"egx_gp_params", "egx_sgp_params"}
    for cname, fields in structs:
        body = re.search(r"pub struct %s \{(.*?)\}" % gen_rust_ffi.rust_name(cname), src, flags=re.S).group(1)
        assert re.findall(r"pub ([a-z_0-9]+):", body) == [n for _, n in fields]
    # the ctypes mirror of the same structs has the same field order
    for cname, pyname in (("egx_gp_params", "GpParams"), ("egx_sgp_params", "SgpParams")):
        py = getattr(_lib, pyname, None)
        if py is not None:
            assert [f[0] for f in py._fields_] == [n for _, n in dict(structs)[cname]]


def test_wrapper_calls_only_declared_functions():
    rust = _rust_functions()
    used = set(re.findall(r"\b(egx_[a-z0-9_]+)\(", open(WRAPPER).read()))
    assert used and used <= set(rust), used - set(rust)
    consts = set(re.findall(r"\b(EGX_[A-Z0-9_]*[A-Z0-9])\b(?!_)", open(WRAPPER).read()))
    declared = set(re.findall(r"pub const (EGX_[A-Z0-9_]+)", open(RS).read()))
    assert consts <= declared, consts - declared
