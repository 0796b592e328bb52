"""CPU tests of the host side: block decomposition (vs the oracle's restatement of
parapp/parallelini), the C-ABI library (loads, exports every declared symbol), and the
algebra of the partitioned Thomas solve the CUDA kernel implements."""
import ctypes
import os
import re

import numpy as np
import pytest

import astr_b200
from astr_b200 import decompose, mpisizedis, refcal
from astr_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dims,size,homo", [
    ((32, 32, 32), (1, 1, 1), (True, True, True)),
    ((32, 32, 32), (2, 2, 2), (True, True, True)),
    ((33, 31, 30), (2, 3, 1), (True, False, True)),
    ((64, 48, 32), (4, 1, 2), (False, False, False)),
    ((32, 32, 32), (1, 1, 2), (True, True, True)),
])
def test_decompose_matches_oracle(oracle, dims, size, homo):
    blocks = decompose(dims, size, homo)
    c = oracle.Case(*dims, blocks=size, homo=homo)
    assert c.nblocks == len(blocks)
    for ib, b in enumerate(blocks):
        info = c.block_info(ib)
        assert b.rank == ib
        assert b.dims == (info["im"], info["jm"], info["km"])
        assert list(b.npdc) == info["npdc"]
        assert [b.s[0], b.e[0], b.s[1], b.e[1], b.s[2], b.e[2]] == info["is_ie"]
        assert list(b.g0) == info["g0"]
        assert b.nbr == info["nb"]
    c.close()


def test_mpisizedis():
    assert mpisizedis(1, (512, 512, 512)) == (1, 1, 1)
    assert mpisizedis(8, (512, 512, 512)) == (2, 2, 2)
    assert mpisizedis(8, (512, 512, 512), strict3d=True) == (2, 2, 2)   # what the reference picks
    assert mpisizedis(2, (512, 512, 512)) == (1, 1, 2)                  # k slabs (extension)
    assert mpisizedis(4, (512, 512, 512)) == (1, 2, 2)
    with pytest.raises(ValueError):
        mpisizedis(2, (512, 512, 512), strict3d=True)                   # src/parallel.F90:289-297


def test_refcal_matches_oracle_constants():
    th = refcal(1600.0, 0.1)
    assert th["const2"] == 1.4 * (0.1 * 0.1)
    assert th["const6"] == 1.0 / (1.4 - 1.0)
    assert th["tempconst"] == 110.3 / 273.15


def test_library_builds_and_exports_every_declared_symbol():
    so = astr_b200.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    header = open(os.path.join(ROOT, "include", "astr_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(astr_gpu_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/astr_gpu.h but not exported"
    assert sorted(L.SYMBOLS) == declared
    # the ctypes mirror of struct astr_cfg has the C struct's size (58 ints + 29 doubles)
    assert lib.astr_gpu_sizeof_cfg() == ctypes.sizeof(L.AstrCfg) == 4 * 58 + 8 * 29


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    b = decompose((16, 16, 16), (1, 1, 1), (True, True, True))[0]
    with pytest.raises(astr_b200.AstrGpuError):
        astr_b200.RhsEngine(b, (16, 16, 16), (True, True, True), refcal(1600.0, 0.1))


# ---- the algebra of the partitioned solve (astr_b200/csrc/sweep.cu) ------------------------------
def _partitioned_thomas(ac1, ac2, ac3, d, C):
    """Chunked forward/backward with zero carries + exact carry recovery, as the kernel does."""
    N = d.size
    def chunk_start(c):   # astr_b200/csrc/common.cuh
        if c <= 0:
            return 0
        if c >= C:
            return N
        s = (c * N) // C
        return (s | 1) if (c & 1) else (s & ~1)
    cs = [chunk_start(c) for c in range(C + 1)]
    ac2 = ac2.copy(); ac3 = ac3.copy(); ac2[0] = 1.0; ac3[0] = 0.0
    pf = np.zeros(N); qb = np.zeros(N)
    for c in range(C):
        p = 1.0
        for r in range(cs[c], cs[c + 1]):
            p *= -ac3[r]; pf[r] = p
        q = 1.0
        for r in range(cs[c + 1] - 1, cs[c] - 1, -1):
            q *= -ac1[r]; qb[r] = q
    e = np.zeros(N); ee = np.zeros(C)
    for c in range(C):
        prev = 0.0
        for r in range(cs[c], cs[c + 1]):
            prev = d[r] * ac2[r] - prev * ac3[r]; e[r] = prev
        ee[c] = prev
    cin = np.zeros(C)
    for c in range(1, C):
        cin[c] = ee[c - 1] + pf[cs[c] - 1] * cin[c - 1]
    g = np.zeros(N); gs = np.zeros(C)
    for c in range(C):
        nxt = 0.0
        for r in range(cs[c + 1] - 1, cs[c] - 1, -1):
            nxt = (e[r] + pf[r] * cin[c]) - ac1[r] * nxt; g[r] = nxt
        gs[c] = nxt
    xin = np.zeros(C)
    for c in range(C - 2, -1, -1):
        xin[c] = gs[c + 1] + qb[cs[c + 1]] * xin[c + 1]
    x = np.array([g[r] + qb[r] * xin[np.searchsorted(cs, r, side="right") - 1] for r in range(N)])
    return x


@pytest.mark.parametrize("is_filter", [False, True])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
@pytest.mark.parametrize("C", [1, 2, 4, 8])
def test_partitioned_thomas_is_exact(oracle, is_filter, ntype, C):
    n = 128
    first, a, c, ac1, ac2, ac3 = oracle.scheme_tables(is_filter, ntype, n)
    N = a.size
    rng = np.random.default_rng(C + 10 * ntype)
    d = rng.standard_normal(N)
    # sequential Thomas (src/commfunc.F90:790-813)
    dd = d.copy()
    for i in range(1, N):
        dd[i] = dd[i] * ac2[i] - dd[i - 1] * ac3[i]
    x = np.zeros(N); x[-1] = dd[-1]
    for i in range(N - 2, -1, -1):
        x[i] = dd[i] - ac1[i] * x[i + 1]
    xp = _partitioned_thomas(ac1, ac2, ac3, d, C)
    scale = np.abs(x).max()
    assert np.abs(xp - x).max() <= (2e-13 if is_filter else 2e-15) * scale
    # residual of the original system
    res = x.copy()
    res[1:] += a[1:] * x[:-1]
    res[:-1] += c[:-1] * x[1:]
    assert np.abs(res - d).max() < 1e-12 * max(1.0, scale)


def test_refcal_dimensional_matches_the_oracle(oracle):
    # nondimen=f branch of refcal (src/solver.F90:124-148) with the HBL reference state (input.M3)
    from astr_b200 import refcal_dimensional
    ref = (226.65, 900.0, 1.0, 0.0180119)
    th = refcal_dimensional(*ref)
    c = oracle.Case(16, 16, 16)
    c.set_dimensional(*ref)
    want = c.thermo()
    c.close()
    assert want["nondimen"] == 0.0 and th["nondimen"] == 0
    for k in ("reynolds", "mach", "const1", "const2", "const5", "const6", "rgas", "cp", "cv", "pinf"):
        assert abs(th[k] - want[k]) <= 1e-14 * abs(want[k]), k
    assert 2.9 < th["mach"] < 3.1      # "M3"


# ---- the algebra of the register engine (astr_b200/csrc/linecore.h, driven by tests/emul_sweep2.cpp) -------
_EMUL = None


def _emul_lib(tmp_path_factory):
    """g++ build of tests/emul_sweep2.cpp: the device functions of linecore.h, run on the CPU element by element
    in the kernels' phase order (passes 1-2 and publication of S / S', barrier, truncated reduced scan, pass 3)."""
    global _EMUL
    if _EMUL is None:
        import subprocess
        out = tmp_path_factory.mktemp("emul") / "libemul_sweep2.so"
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(out),
                               os.path.join(ROOT, "tests", "emul_sweep2.cpp")])
        _EMUL = ctypes.CDLL(str(out))
    return _EMUL


@pytest.mark.parametrize("op", [0, 1, 2, 3], ids=["deriv", "filter", "flux+", "flux-"])
@pytest.mark.parametrize("ntype", [1, 2, 3, 4])
def test_register_engine_algebra_matches_the_oracle(oracle, tmp_path_factory, op, ntype):
    # fresh-start chunk factorisation + head / tail blocks with neutral slots + reduced scan truncated where the
    # dropped coupling is below 1e-20, against the reference's single Thomas sweep (df_compact, compact_filter,
    # flux_compact restated in oracle/lineops.cpp) -- every line length class: no / small / large head remainder,
    # 1 to 15 regular chunks, with and without the 16-byte alignment rule of the i direction
    lib = _emul_lib(tmp_path_factory)
    rng = np.random.default_rng(10 * op + ntype)
    covered = 0
    shorts = 0
    for n in (44, 48, 64, 99, 100, 128, 255, 256, 300, 500, 512, 513, 527):
        for align in (0, 1):
            f = rng.standard_normal(n + 11)
            alfa = 0.49 if op == 1 else (0.3 if op >= 2 else 0.0)
            out = np.zeros(n + 20)
            info = (ctypes.c_int * 7)()
            rc = lib.emul_line(op, ntype, n, ctypes.c_double(alfa), align, f.ctypes.data_as(ctypes.c_void_p),
                               out.ctypes.data_as(ctypes.c_void_p), info)
            if rc:
                assert n < 48, f"no plan for n={n}"      # short lines stay on the shared-memory engine
                continue
            first, nrows, nw, sh, st, w, ls0 = list(info)
            assert 1 <= nw <= 15 and sh <= 40 and st <= 6 and 1 <= w <= 8 and 2 <= ls0 <= 34
            # the head block stays on the 8-slot path whenever a SHORT first chunk can take the left-over rows
            assert sh <= 8 or nw == 15 or ls0 == 34, list(info)
            assert sh + ls0 + (nw - 1) * 34 + st == nrows
            shorts += ls0 < 34
            if op == 0:
                ref, lo = oracle.df_compact(f, ntype), 0
            elif op == 1:
                ref, lo = oracle.compact_filter(f, ntype, 0.49), 0
            else:
                ref, lo = oracle.flux_compact(f, ntype, op == 2, 0.3), -1
            got = np.array([out[node - first] for node in range(lo, n + 1)])
            if op == 1:      # physical-boundary nodes are left unfiltered by the caller (src/filter.F90:141-142)
                if ntype in (1, 4):
                    got[0] = ref[0]
                if ntype in (2, 4):
                    got[-1] = ref[-1]
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err < 2e-14, (op, ntype, n, align, list(info), err)
            covered += 1
    assert covered >= 18 and shorts >= 4


def test_register_engine_falls_back_when_the_coupling_decays_slowly(tmp_path_factory):
    # alfa -> 1/2: the elimination factors of a whole chunk no longer contract below 1e-20 within 8 elements;
    # the plan refuses (the shared-memory engine, which does no truncation, takes the line)
    lib = _emul_lib(tmp_path_factory)
    n = 512
    f = np.zeros(n + 11); out = np.zeros(n + 20); info = (ctypes.c_int * 7)()
    rc = lib.emul_line(1, 3, n, ctypes.c_double(0.4999), 0, f.ctypes.data_as(ctypes.c_void_p),
                       out.ctypes.data_as(ctypes.c_void_p), info)
    assert rc == 1
    rc = lib.emul_line(1, 3, n, ctypes.c_double(0.49), 0, f.ctypes.data_as(ctypes.c_void_p),
                       out.ctypes.data_as(ctypes.c_void_p), info)
    assert rc == 0 and info[5] <= 8


def test_fortran_module_binds_every_stage_operator():
    # fortran/astr_gpu_mod.F90 is the bind(C) layer a maintainer adds: every entry point of the header except the
    # introspection / profiling helpers (bench and tests only) must have an interface there, and the derived type
    # must list the members of struct astr_cfg in the same order
    hdr = open(os.path.join(ROOT, "include", "astr_gpu.h")).read()
    f90 = open(os.path.join(ROOT, "fortran", "astr_gpu_mod.F90")).read()
    declared = set(re.findall(r"\b(astr_gpu_[a-z_0-9]+)\s*\(", hdr))
    bound = set(re.findall(r"name='(astr_gpu_[a-z_0-9]+)'", f90))
    helpers = {"astr_gpu_bench_sweep", "astr_gpu_device_ptr", "astr_gpu_get_profile", "astr_gpu_kernel_launches",
               "astr_gpu_rk_steps_timed", "astr_gpu_set_profile"}
    assert declared - bound <= helpers, sorted(declared - bound - helpers)
    struct = hdr[hdr.index("typedef struct astr_cfg {"):hdr.index("} astr_cfg;")]
    c_members = []
    for line in struct.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        m = re.match(r"(int|double)\s+(.*);", line)
        if m:
            c_members += [re.sub(r"\[\d+\]", "", v).strip() for v in m.group(2).split(",")]
    ftype = f90[f90.index("type, bind(c) :: astr_cfg"):f90.index("end type astr_cfg")]
    f_members = []
    for line in ftype.splitlines()[1:]:
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\))\s*::\s*(.*)", line)
        if m:
            f_members += [re.sub(r"\(\d+\)", "", v).strip() for v in m.group(2).split(",")]
    assert [m.lower() for m in c_members] == [m.lower() for m in f_members]
    assert [f[0] for f in L.AstrCfg._fields_] == [("is_" if m == "is" else m) for m in c_members]
