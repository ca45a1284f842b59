"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/otmb.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side logic of the shim (cleaning,
vertex permutation, topology detection, synthetic generator invariants) behaves like the reference."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import otmb_b200
import otmb_b200.api as A
from otmb_b200 import _lib, synthetic

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "otmb.h").read_text()
    declared = set(re.findall(r"\b(otmb_[a-z0-9_]+)\s*\(", header))
    declared -= {"otmb_ctx", "otmb_tm_params"}
    assert len(declared) >= 30
    lib = C.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in otmb.h but not exported by libotmb.so"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert all(re.search(rf"\bT {n}\b", out) for n in declared)


def test_library_is_sm100a_only_and_has_no_cpu_fallback():
    lib = _lib.load()
    assert lib.otmb_version() >= 100
    assert b"no CPU fallback" in lib.otmb_status_string(_lib.ERR_NO_GPU)
    n = C.c_int(-1)
    assert lib.otmb_device_count(C.byref(n)) == 0
    if n.value == 0:                       # this container: creating a context must fail loudly
        h = C.c_void_p()
        assert lib.otmb_create(C.byref(h), 0) == _lib.ERR_NO_GPU and not h.value
        with pytest.raises(A.OTMBError) as e:
            A.Context(0)
        assert e.value.code == _lib.ERR_NO_GPU
    sass = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and not re.search(r"sm_(5|6|7|8|9)\d", sass)


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "oceantransportmatrixbuilder.jl_b200"
    for p in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        txt = p.read_text(errors="ignore")
        assert not re.search(r"^\s*(from|import)\s+oracle\b|oracle/|otmb_oracle|pyoracle", txt, re.M), \
            f"{p} reaches into oracle/"


def test_error_messages_are_the_references():
    lib = _lib.load()
    want = {1: "Tadv contains NaNs.", 2: "TκH contains NaNs.", 3: "TκVML contains NaNs.", 4: "TκVdeep contains NaNs.",
            5: "ρ contains NaNs", 6: "Unknown grid type"}
    for code, msg in want.items():
        assert lib.otmb_status_string(code).decode() == msg


def test_clean_missing_and_fillvalue():
    f = A.Field(np.array([[0.0, -0.0, 5.0], [1e20, np.nan, 2.0]]), {"_FillValue": 1e20})
    a = A._clean_missing(f, [1e20])
    assert np.isnan(a[0, 0]) and a[0, 1] == 0.0 and np.signbit(a[0, 1]) and np.isnan(a[1, 0]) and np.isnan(a[1, 1])
    assert a.flags.f_contiguous and a[0, 2] == 5.0
    m = np.ma.masked_array([[1.0, 2.0]], mask=[[False, True]])
    assert np.isnan(A._clean_missing(m, [])[0, 1])


@pytest.mark.parametrize("topo", ["bipolar", "tripolar"])
def test_topology_and_vertex_permutation(topo):
    oc = synthetic.make_ocean(16, 10, 4, topo, seed=1)
    t = A.getgridtopology(oc.lon_vertices, oc.lat_vertices, oc.lev)
    assert (t.kind, t.nx, t.ny, t.nz) == (topo, 16, 10, 4)
    assert A.vertexpermutation(oc.lon_vertices, oc.lat_vertices) == [0, 1, 2, 3]
    perm = [3, 2, 0, 1]
    p = A.vertexpermutation(oc.lon_vertices[perm], oc.lat_vertices[perm])
    assert np.array_equal(oc.lon_vertices[perm][p], oc.lon_vertices)
    lonv = oc.lon_vertices.copy()
    if topo == "tripolar":
        lonv[2, 2, -1] += 11.0
        with pytest.warns(UserWarning):
            assert A.getgridtopology(lonv, oc.lat_vertices, oc.lev).kind == "unknown"
        # periodic longitudes still count as equal (isapprox_lon, src/gridtopology.jl:23-26)
        lonv = oc.lon_vertices.copy()
        lonv[2, 2, -1] += 360.0
        assert A.getgridtopology(lonv, oc.lat_vertices, oc.lev).kind == "tripolar"


@pytest.mark.parametrize("cfg", ["C1", "C1t"])
def test_synthetic_generator_invariants(cfg):
    oc = synthetic.make_config(cfg, seed=0)
    oc2 = synthetic.make_config(cfg, seed=0)
    assert np.array_equal(oc.umo, oc2.umo) and np.array_equal(oc.volcello, oc2.volcello)       # seeded
    nx, ny, nz = oc.nx, oc.ny, oc.nz
    assert oc.volcello.shape == (nx, ny, nz) and oc.volcello.flags.f_contiguous
    assert oc.lon_vertices.shape == (4, nx, ny)
    wet = oc.volcello > 0
    assert (wet[:, 0, :] == False).all()                                                      # southern row is land
    assert (np.diff(wet.astype(int), axis=2) <= 0).all()                                       # wet from the surface down
    assert ((oc.umo == oc.fill) == ~wet).all()
    # shared corners are bitwise identical between neighbouring cells
    assert np.array_equal(oc.lon_vertices[1, :-1, :], oc.lon_vertices[0, 1:, :])
    assert np.array_equal(oc.lat_vertices[2, :, :-1], oc.lat_vertices[1, :, 1:])
    if oc.topology == "tripolar":
        i = np.arange(nx)
        assert np.array_equal(oc.lat_vertices[2, i, -1], oc.lat_vertices[3, nx - 1 - i, -1])  # fold
        assert wet[0, -1, 0] and wet[nx - 1, -1, 0] and wet[nx // 2 - 1, -1, 0] and wet[nx // 2, -1, 0]
    else:
        assert (oc.lat_vertices[2:4, :, -1] == 90).all()
    assert (oc.areacello[wet[:, :, 0]] > 0).all()


def test_dump_writes_raw_float64(tmp_path):
    oc = synthetic.make_ocean(6, 5, 3, "bipolar", seed=0)
    oc.dump(tmp_path)
    a = np.fromfile(tmp_path / "umo.f64", dtype="<f8").reshape((6, 5, 3), order="F")
    assert np.array_equal(a, oc.umo)


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "C1t"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "nnz(T)/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("threads", [0, 1, 3, 8])
def test_host_widening_of_the_fetch_pipeline(threads):
    """The host half of otmb_transportmatrix_fetch_all (csrc/fetch.cu): Int32 indices that crossed the link are
    sign-extended into the caller's Int64 arrays by a pool of threads (AVX2 + scalar head / tail).  No GPU involved."""
    lib = _lib.load()
    rng = np.random.default_rng(threads)
    for n in (0, 1, 7, 65535, 65536, 65537, (1 << 22) + 13, (7 << 20) + 5):
        src = rng.integers(-2**31, 2**31 - 1, size=n, dtype=np.int64).astype(np.int32)
        if n > 2:
            src[:3] = (np.iinfo(np.int32).max, np.iinfo(np.int32).min, -1)
        for shift in (0, 1):                        # destination 32-byte aligned or not
            buf = np.full(n + 4, -99, np.int64)
            dst = buf[shift:shift + n]
            assert lib.otmb_host_widen(src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n, threads) == 0
            assert np.array_equal(dst, src.astype(np.int64))
            assert (buf[:shift] == -99).all() and (buf[shift + n:] == -99).all()      # nothing written outside
    assert lib.otmb_host_widen(None, None, -1, 0) != 0


def _abi_driver(tmp_path):
    exe = tmp_path / "abi_driver"
    r = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-o", str(exe), str(ROOT / "tests" / "abi_driver.c"), "-ldl", "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return subprocess.run([str(exe), str(_lib.LIB_PATH)], capture_output=True, text=True, timeout=120)


def test_c_driver_resolves_the_abi_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/abi_driver.c: the ABI from plain C (dlopen).  Here (no GPU) it must resolve every symbol it needs and
    report the missing GPU as OTMB_ERR_NO_GPU; on the GPU box the same program checks a hand-derived matrix."""
    n = C.c_int(-1)
    _lib.load().otmb_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present: covered by test_c_driver_on_gpu")
    r = _abi_driver(tmp_path)
    assert r.returncode == 0 and "ABI-DRIVER: no GPU" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_driver_on_gpu(tmp_path):
    r = _abi_driver(tmp_path)
    assert r.returncode == 0 and "ABI-DRIVER: OK" in r.stdout, r.stdout + r.stderr


def _bench(*args, timeout=600):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "bench.py must print exactly ONE JSON line on stdout"
    import json
    return json.loads(lines[0])


def test_bench_reference_arm_line():
    """`bench.py --impl reference`: the oracle port timed on host cores (no Julia in the image), one JSON line with the
    contract's keys; run here on the test-suite-sized grid."""
    d = _bench("--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["unit"] == "nnz(T)/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C1 90x45x20") and d["config"]["N_wet"] > 0 and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_bench_cuda_arm_line_and_same_config_as_reference_arm():
    d = _bench("--workload", "C1", "--steps", "3", "--warmup", "3", "--mode", "batch")
    r = _bench("--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0")
    assert d["config"] == r["config"], "both arms must describe the same configuration"
    assert d["metric"] == r["metric"] and d["unit"] == r["unit"] and d["dtype"] == "f64" and d["vs_baseline"] is None
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert rf["frac_step"] <= rf["frac"] * 1.05 and rf["traffic"] is None
    assert d["gpu_launches"] >= 3 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_julia_shim_binds_every_export_with_matching_types():
    """Julia is not in the image, so the shim is never executed: every entry point of include/otmb.h must be bound
    there, and every `ccall` must list, argument by argument, a Julia type that matches the C parameter."""
    import re
    root = Path(__file__).resolve().parent.parent
    header = re.sub(r"/\*.*?\*/", " ", (root / "include" / "otmb.h").read_text(), flags=re.S)

    def c_kind(param):                                  # (base type, levels of indirection) of one C parameter
        param = re.sub(r"\s+", " ", param.strip())
        levels = 1 if "[" in param else 0
        param = re.sub(r"\[[^\]]*\]", "", param)
        toks = param.split(" ")
        typ = " ".join(toks[:-1]) if len(toks) > 1 and not toks[-1].endswith("*") else param
        typ = typ.replace("const ", "").replace(" const", "").strip()
        return typ.replace("*", "").strip(), levels + typ.count("*")

    decl = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(otmb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        params = m.group(2).strip()
        decl[m.group(1)] = [] if params in ("", "void") else [c_kind(x) for x in params.split(",")]
    assert len(decl) > 60
    julia_for = {
        ("otmb_ctx", 1): {"PV"}, ("otmb_ctx", 2): {"Ref{PV}"}, ("void", 2): {"Ref{PV}"}, ("void", 1): {"PV", "Ptr{UInt8}"},
        ("double", 0): {"Float64"}, ("double", 1): {"PF", "Ptr{Float64}"}, ("double", 2): {"Ptr{PF}", "PV"},
        ("int64_t", 0): {"Int64"}, ("int64_t", 1): {"PI", "Ref{Int64}", "Ptr{Int64}"}, ("int64_t", 2): {"Ptr{PI}", "PV"},
        ("int32_t", 0): {"Int32", "Cint"}, ("int32_t", 1): {"Ptr{Int32}", "Ref{Int32}", "Ref{Cint}", "Ptr{Cint}"},
        ("int", 0): {"Cint", "Int32"}, ("int", 1): {"Ref{Cint}", "Ptr{Cint}"},
        ("uint64_t", 0): {"UInt64"}, ("uint64_t", 1): {"Ptr{UInt64}"}, ("float", 1): {"Ref{Cfloat}"},
        ("char", 1): {"Cstring"}, ("unsigned char", 1): {"Ptr{UInt8}"}, ("uint8_t", 1): {"Ptr{UInt8}"},
        ("otmb_tm_params", 1): {"Ref{TMParams}"},
    }
    shim = (root / "oceantransportmatrixbuilder.jl_b200" / "julia" / "OceanTransportMatrixBuilderB200.jl").read_text()
    assert not re.search(r"ccall\(\(\s*[a-z_]\w*\s*,", shim), "a ccall names its function through a variable: Julia needs a constant"
    seen = set()
    for m in re.finditer(r"ccall\(\(:(otmb_[a-z0-9_]+), LIBOTMB\),\s*\w+,\s*\(", shim):
        name, i, depth = m.group(1), m.end(), 1
        start = i
        while depth:                                   # the type tuple, up to its closing parenthesis
            depth += {"(": 1, ")": -1}.get(shim[i], 0)
            i += 1
        types, cur, braces = [], "", 0
        for ch in shim[start:i - 1]:                   # split on commas outside {...}
            braces += {"{": 1, "}": -1}.get(ch, 0)
            if ch == "," and braces == 0:
                types.append(cur.strip())
                cur = ""
            else:
                cur += ch
        types = [t for t in types + [cur.strip()] if t]
        assert name in decl, f"{name} is not declared in include/otmb.h"
        assert len(types) == len(decl[name]), f"ccall of {name}: {len(types)} argument types, the header declares {len(decl[name])}"
        for k, (jt, ck) in enumerate(zip(types, decl[name])):
            assert ck in julia_for, f"{name} parameter {k}: no rule for C type {ck}"
            assert jt in julia_for[ck], f"ccall of {name}, argument {k}: Julia type {jt} for C parameter {ck}"
        seen.add(name)
    assert set(decl) - seen == set(), f"not bound in the Julia shim: {sorted(set(decl) - seen)}"
