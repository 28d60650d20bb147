"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, refuses to run without a B200 (no CPU fallback), and the host-side mirror of the
reference's storages resolves to the same Between(affected, affecting) as the oracle's statement of
storage.rs:207-241.  No compute calls happen here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT, uniform_cloud

HEADER = os.path.join(ROOT, "include", "particular_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcuda_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from particular_b200 import _ffi
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_ffi.lib, s), f"libparticular_cuda.so does not export {s}"
        assert s in _ffi.SIGNATURES, f"{s} is declared in the header but not bound in _ffi.py"
    assert _ffi.lib.pcuda_abi_version() == 1


def test_rust_ffi_declares_every_symbol_with_matching_arity():
    """rust/particular-cuda/src/ffi.rs cannot be compiled here (no Rust toolchain), so at least keep
    it in lock-step with the header: same symbol set, same number of arguments, and build.rs lists
    the translation units build.py compiles."""
    def arities(text, pat):
        out = {}
        for name, args in re.findall(pat, text, flags=re.S):
            args = args.strip()
            out[name] = 0 if args in ("", "void") else args.count(",") + 1
        return out
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    c_ar = arities(hdr, r"\b(pcuda_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;")
    rs = open(os.path.join(ROOT, "rust", "particular-cuda", "src", "ffi.rs")).read()
    rs = re.sub(r"//.*", "", rs)
    rs_ar = arities(rs, r"pub fn (pcuda_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]*)?;")
    assert set(c_ar) == set(declared_symbols())
    missing = set(c_ar) - set(rs_ar)
    assert not missing, f"ffi.rs lacks {sorted(missing)}"
    assert not set(rs_ar) - set(c_ar), f"ffi.rs declares unknown {sorted(set(rs_ar) - set(c_ar))}"
    for name, n in c_ar.items():
        assert rs_ar[name] == n, f"{name}: header has {n} arguments, ffi.rs {rs_ar[name]}"
    from particular_b200.build import SOURCES
    build_rs = open(os.path.join(ROOT, "rust", "particular-cuda", "build.rs")).read()
    for unit in SOURCES:
        assert f'"{unit}"' in build_rs, f"build.rs does not compile {unit}"


def _c_structs(hdr):
    """{name: [(field, type, array_len)]} of the plain structs in the header."""
    out = {}
    for body, name in re.findall(r"typedef struct \w+ \{(.*?)\}\s*(\w+);", hdr, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(.*?)\s*(\*?)\s*(\w+)\s*(?:\[(\d+)\])?$", decl, flags=re.S)
            ctype = (m.group(1).strip() + m.group(2)).replace(" *", "*")
            fields.append((m.group(3), ctype, int(m.group(4)) if m.group(4) else 0))
        out[name] = fields
    return out


def _rs_structs(rs):
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\](?:\s*#\[derive\([^)]*\)\])?\s*pub struct (\w+)\s*\{(.*?)\}", rs, flags=re.S):
        fields = []
        for decl in body.split(","):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(?:pub\s+)?(\w+)\s*:\s*(.+)$", decl, flags=re.S)
            fields.append((m.group(1), m.group(2).strip()))
        out[name] = fields
    return out


def test_rust_repr_c_structs_match_the_header_layouts():
    """Every `#[repr(C)]` struct with fields in ffi.rs has the same field names, in the same order, with
    the Rust type of the same size and signedness as the C field (`[T; N]` for arrays, `*mut c_void`
    for `void *`) — the lock-step check that symbol names and argument counts alone do not give."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    rs = re.sub(r"//.*", "", open(os.path.join(ROOT, "rust", "particular-cuda", "src", "ffi.rs")).read())
    cmap = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "float": "f32", "double": "f64",
            "void*": "*mut c_void", "int": "c_int", "size_t": "usize"}
    c_structs, rs_structs = _c_structs(hdr), _rs_structs(rs)
    checked = 0
    for name, fields in rs_structs.items():
        if [f for f in fields if f[0] == "_private"]:
            continue  # opaque handles
        assert name in c_structs, f"ffi.rs declares {name}, the header does not"
        want = [(f, f"[{cmap[t]}; {n}]" if n else cmap[t]) for f, t, n in c_structs[name]]
        assert fields == want, f"{name}: ffi.rs {fields} != header {want}"
        checked += 1
    for name in ("pcuda_config", "pcuda_timings", "pcuda_tree_info", "pcuda_sim_config", "pcuda_sim_info_t"):
        assert name in rs_structs, f"ffi.rs lacks {name}"
    assert checked >= 5
    # the flag / enum constants the Rust side mirrors
    for cname, value in re.findall(r"#define (PCUDA_FLAG_\w+) (\d+)u", hdr):
        if cname == "PCUDA_FLAG_NONE":
            continue
        assert re.search(rf"pub const {cname}: u32 = {value};", rs), f"ffi.rs lacks {cname} = {value}"


def test_ctypes_structs_match_the_header_layouts():
    """The Python binding's ctypes structures: header field names in header order, and the sizes that
    tests/cpp/test_host_api.cpp pins with static_assert."""
    from particular_b200 import _ffi
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    c_structs = _c_structs(hdr)
    for py, cname, size in ((_ffi.Config, "pcuda_config", 16), (_ffi.Timings, "pcuda_timings", 28),
                            (_ffi.TreeInfo, "pcuda_tree_info", 56), (_ffi.SimConfig, "pcuda_sim_config", 48),
                            (_ffi.SimInfo, "pcuda_sim_info_t", 56)):
        assert [f for f, _ in py._fields_] == [f for f, _, _ in c_structs[cname]], cname
        assert C.sizeof(py) == size, (cname, C.sizeof(py))


def test_integration_md_lists_the_real_files():
    """INTEGRATION.md sections 1 and 2 are rust/particular-cuda/build.rs and src/ffi.rs verbatim
    (scripts/gen_integration.py regenerates them)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_integration.py"), "--check"])
    assert r.returncode == 0, "INTEGRATION.md is out of date: run python scripts/gen_integration.py"


def test_header_compiles_as_c():
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import particular_b200 as pb
    from particular_b200 import _ffi
    with pytest.raises(pb.CudaError) as e:
        pb.CudaContext(0)
    assert e.value.status == _ffi.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)
    n = C.c_int(-1)
    assert _ffi.lib.pcuda_device_count(C.byref(n)) != 0 and n.value == 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "particular_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "oracle/" not in txt, f


def test_status_strings():
    from particular_b200 import _ffi
    assert _ffi.lib.pcuda_status_string(0) == b"ok"
    assert b"sm_100" in _ffi.lib.pcuda_status_string(_ffi.ERR_NO_DEVICE)


def test_storage_resolution_matches_reference_semantics():
    import oracle
    import particular_b200.interface as pi
    p = uniform_cloud(40, massive_ratio=0.3)
    p = p[np.random.default_rng(1).permutation(40)]
    # &[P]  => Between(slice, slice); affected None means "alias of affecting"
    aff, src = pi._resolve(p)
    assert aff is None and np.array_equal(src, p)
    # &Reordered => Between(unordered, affecting copy)
    ro = pi.Reordered.new(p)
    aff, src = pi._resolve(ro)
    a2, s2 = oracle.between_of_reordered(p)
    assert np.array_equal(aff, a2) and np.array_equal(src, s2)
    assert ro.affecting_len() == 12 and len(ro.non_affecting()) == 28
    assert np.array_equal(ro.reordered()[:12], s2)
    # &Ordered => Between(ordered_all, affecting prefix)
    od = pi.Ordered.new(p)
    aff, src = pi._resolve(od)
    a3, s3 = oracle.between_of_ordered(p)
    assert np.array_equal(aff, a3) and np.array_equal(src, s3)
    # Ordered::with: affecting_len = first index failing the predicate (storage.rs:71-74)
    od2 = pi.Ordered.with_(p[:5], p[5:])
    first_massless = int(np.flatnonzero(p[:, 3] == 0)[0])
    assert od2.affecting_len() == first_massless
    # Between(&P1, &[P2]) and particles given as affected
    aff, src = pi._resolve(pi.Between(p[3], p))
    assert aff.shape == (1, 3) and np.array_equal(aff[0], p[3, :3])
    aff, src = pi._resolve(pi.Between(p[:7], p[7:]))
    assert aff.shape == (7, 3)


def test_interactions_mirror_reference_constructors():
    import particular_b200 as pb
    assert pb.Acceleration.checked().is_checked and pb.Acceleration.checked().softening == 0.0
    assert not pb.Acceleration.unchecked().is_checked
    s = pb.AccelerationSoftened.checked(100.0)
    assert s.softening == 100.0 and s.is_checked
    assert not pb.AccelerationSoftened.unchecked(1.0).is_checked


def test_unsupported_shapes_raise():
    import particular_b200.interface as pi
    with pytest.raises(TypeError):
        pi._as_particles(np.zeros((4, 6), np.float32))
    assert pi._suffix(np.zeros((4, 3), np.float64)) == "f64x2"  # DVec2
    with pytest.raises(NotImplementedError):
        pi._suffix(np.zeros((4, 3), np.float16))  # no kernel, like unimplemented!()


def test_cpp_host_api_builds_and_refuses_without_gpu():
    """include/particular_cuda.hpp (the C++ mirror of the reference's operator interface) compiles
    against the C ABI; without a GPU its test program stops at context creation (exit 77)."""
    import torch
    import __graft_entry__ as ge
    exe = ge.build_cpp_host_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
        assert "no CPU fallback" in r.stdout
