"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol include/b200newton.h declares,
refuses to run without a GPU (no fallback), and the host-side set-up pieces (colouring, sharding) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200newton.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(nls):
    L = nls.abi.lib()
    syms = _header_symbols()
    assert len(syms) >= 80
    for s in syms:
        assert hasattr(L, s), "missing export: %s" % s
    # and the ctypes table covers exactly the header
    assert sorted(nls.abi.SIGNATURES) == syms


def test_struct_layouts_match_header(nls, po):
    # the oracle includes the same header; its ctypes mirrors must agree with the product's
    for name in ("GmresOpts", "GmresStats", "NewtonOpts", "NewtonResult", "TraceRec", "EnsResult"):
        a, b = getattr(nls.abi, name), getattr(po, name)
        assert C.sizeof(a) == C.sizeof(b)
        assert [f[0] for f in a._fields_] == [f[0] for f in b._fields_]
    o = nls.abi.NewtonOpts()
    nls.abi.lib().b200_newton_opts_default(C.byref(o))
    assert o.maxiters == 1000 and o.max_shrink_times == 32 and o.gmres.memory == 20 and o.gmres.restart == 0
    assert o.ew_eta0 == 0.5 and o.ew_eta_max == 0.9 and o.ew_gamma == 0.9 and o.ew_alpha == 2.0


def test_no_cpu_fallback(nls):
    cnt = C.c_int32(-1)
    nls.abi.lib().b200_device_count(C.byref(cnt))
    if cnt.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nls.abi.B200Error) as e:
        nls.Context(0)
    assert e.value.code == nls.abi.ERR_NO_DEVICE
    with pytest.raises(nls.abi.B200Error):
        nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(10), np.ones(10), 2.0), nls.NewtonRaphson())


def test_product_does_not_touch_oracle():
    # the product path must never import / link / call the oracle
    pkg = os.path.join(ROOT, "nonlinearsolve.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle.h" not in txt, f


def test_coloring_bit_exact_vs_oracle(nls, po, golden):
    # host-side set-up: GreedyColoringAlgorithm(LargestFirst()) column colouring, bit-exact with the oracle and the fixture
    for N in (8, 32):
        colptr, rowval = golden["colptr_%d" % N], golden["rowval_%d" % N]
        n = len(colptr) - 1
        c_lib, nc_lib = nls.coloring_column(n, colptr, rowval, 1)
        c_orc, nc_orc = po.coloring_column(n, colptr, rowval, 1)
        assert nc_lib == nc_orc and np.array_equal(c_lib, c_orc) and np.array_equal(c_lib, golden["colors_%d" % N])
    P3 = po.OracleProblem.bruss3d(12)
    cp, rv = P3.pattern(0)
    a, na = nls.coloring_column(P3.n, cp, rv, 0)
    b, nb = po.coloring_column(P3.n, cp, rv, 0)
    assert na == nb and np.array_equal(a, b)
    T = po.OracleProblem.tridiag_quad(np.ones(50))
    cp, rv = T.pattern(1)
    for order in (0, 1):
        a, na = nls.coloring_column(T.n, cp, rv, 1, order)
        b, nb = po.coloring_column(T.n, cp, rv, 1, order)
        assert na == nb == 3 and np.array_equal(a, b)


def test_shard_range(nls):
    for K, W in ((8192, 8), (10, 3), (5, 8), (1, 1)):
        spans = [nls.shard_range(K, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == K
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
