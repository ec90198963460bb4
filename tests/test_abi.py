"""The drop-in boundary (SURVEY section 8 row b) without a GPU: ``include/vlpet.h`` is plain C, ``libvlpet.so`` exports exactly
the entry points it declares, and the ctypes mirror in ``vlpet_b200/_lib.py`` (what INTEGRATION.md's reference-side stub
copies) agrees with the header on every argument list and on every structure's size and member offsets -- the latter
measured by compiling a probe against the header with gcc.  No compute entry point is called here."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "vlpet.h")
STRUCTS = {"VlpetK1Desc": "K1Desc", "VlpetK1Params": "K1Params", "VlpetK1Grads": "K1Grads", "VlpetK2Desc": "K2Desc",
           "VlpetK2Params": "K2Params", "VlpetK2Grads": "K2Grads", "VlpetK3Desc": "K3Desc", "VlpetK3Params": "K3Params",
           "VlpetK3Grads": "K3Grads", "VlpetK3LRDesc": "K3LRDesc", "VlpetK3LRParams": "K3LRParams",
           "VlpetK3LRGrads": "K3LRGrads", "VlpetWgradPair": "WgradPair"}


@pytest.fixture(scope="module")
def L():
    try:
        import vlpet_b200._lib as L_
    except Exception as e:                       # the package needs libvlpet.so (built in-tree by build())
        pytest.skip(f"vlpet_b200 not importable: {e}")
    return L_


def _header_text():
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _declarations():
    """name -> parameter strings of every ``VLPET_API`` function the header declares."""
    out = {}
    for m in re.finditer(r"VLPET_API\s+[\w\s\*]+?\b(vlpet_\w+)\s*\(([^;]*?)\)\s*;", _header_text(), flags=re.S):
        params = [p.strip() for p in m.group(2).split(",")]
        out[m.group(1)] = [] if params in ([""], ["void"]) else params
    return out


def test_header_is_plain_c_and_cpp():
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        r = subprocess.run(["gcc", "-x", lang, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", HEADER],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = _header_text()
    assert 'extern "C"' in text and "torch" not in text.lower() and "#include <cuda" not in text


def test_library_exports_exactly_the_declared_symbols(L):
    declared = set(_declarations())
    assert len(declared) >= 40
    assert declared == set(L.SYMBOLS), sorted(declared ^ set(L.SYMBOLS))
    for name in declared:
        assert getattr(L.lib, name) is not None
    nm = shutil.which("nm")
    if nm is None:
        pytest.skip("binutils nm not available")
    sym = subprocess.run([nm, "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in sym.splitlines() if " T " in ln}
    # vlpet_debug_*: developer hooks of tools/ (phase traces, stress self-test); not part of the boundary, not in the header
    assert {s for s in exported if s.startswith("vlpet_") and not s.startswith("vlpet_debug_")} == declared
    # nothing but the C ABI leaks out of the library (kernels, helpers and C++ symbols stay hidden)
    leaked = {s for s in exported if not s.startswith("vlpet_") and s not in ("_init", "_fini")}
    assert not leaked, sorted(leaked)[:10]


def _kind(param: str) -> str:
    if "*" in param:
        return "ptr"
    t = param.split()
    for name, kind in (("int64_t", "i64"), ("uint64_t", "u64"), ("int32_t", "i32"), ("size_t", "size"), ("float", "f32"), ("int", "int")):
        if name in t:
            return kind
    raise AssertionError(f"unparsed parameter {param!r}")


def _ckind(t) -> str:
    if t is C.c_void_p or t is C.c_char_p or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
        return "ptr"
    return {C.c_int64: "i64", C.c_uint64: "u64", C.c_int32: "i32", C.c_size_t: "size", C.c_float: "f32", C.c_int: "int"}[t]


def test_ctypes_signatures_match_the_header(L):
    same = {"int": "i32", "size": "u64"}                         # c_int is c_int32, c_size_t is c_uint64 on this ABI
    for name, params in _declarations().items():
        res, args = L.SYMBOLS[name]
        want = [_kind(p) for p in params]
        got = [_ckind(a) for a in args]
        assert [same.get(k, k) for k in got] == [same.get(k, k) for k in want], (name, got, want)
    assert L.SYMBOLS["vlpet_last_error"][0] is C.c_char_p and L.SYMBOLS["vlpet_launch_count"][0] is C.c_uint64


def test_struct_layouts_match_the_header(L, tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, pyname in STRUCTS.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in getattr(L, pyname)._fields_:
            lines.append(f'  printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "probe.c", tmp_path / "probe"
    src.write_text("\n".join(lines))
    r = subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr                           # a member the mirror names but the header lacks fails here
    probe = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, pyname in STRUCTS.items():
        st = getattr(L, pyname)
        assert C.sizeof(st) == int(probe[cname]), (cname, C.sizeof(st), probe[cname])
        for field, _ in st._fields_:
            assert getattr(st, field).offset == int(probe[f"{cname}.{field}"]), (cname, field)
    # and the header has no structure the mirror does not know
    assert set(re.findall(r"typedef struct (\w+)", _header_text())) == set(STRUCTS)


def test_host_only_entry_points(L):
    m = re.search(r"#define\s+VLPET_VERSION\s+(\d+)", _header_text())
    assert L.lib.vlpet_version() == int(m.group(1))
    assert isinstance(L.lib.vlpet_last_error(), bytes)
    assert L.launch_count() >= 0


def test_integration_stub_is_current(L):
    """The hand-written reference-side binding INTEGRATION.md shows is executed against the built library (definitions only,
    no call) and compared with the complete mirror."""
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        md = f.read()
    code = re.search(r"```python\n(# src/vlpet_ffi\.py.*?)```", md, flags=re.S).group(1)
    assert '"/path/to/libvlpet.so"' in code
    ns = {}
    exec(compile(code.replace('"/path/to/libvlpet.so"', repr(L.LIB_PATH)), "INTEGRATION.md", "exec"), ns)
    for name in ("K1Desc", "K1Params"):
        stub, full = ns[name], getattr(L, name)
        assert [n for n, _ in stub._fields_] == [n for n, _ in full._fields_]
        assert C.sizeof(stub) == C.sizeof(full)
        assert all(getattr(stub, n).offset == getattr(full, n).offset for n, _ in full._fields_)
    assert len(ns["lib"].vlpet_k1_fwd.argtypes) == len(L.SYMBOLS["vlpet_k1_fwd"][1])
    assert callable(ns["k1_forward"])
