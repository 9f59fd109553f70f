"""Host-side logic of the reference-interface mirror (api.py) that needs no GPU: option mapping onto the C structs
(through the library's own `b200_newton_opts_default`, a pure host entry point), return-code helpers, trajectory sharding,
and the loud failure of the product path without a device."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def nls():
    import nonlinearsolve_jl_b200 as m
    return m


def _opts(nls, alg, f=None, **kw):
    from nonlinearsolve_jl_b200 import api
    prob = nls.NonlinearProblem(f or nls.Brusselator2D(8), np.zeros(128), (3.4, 1.0, 10.0))
    args = dict(abstol=1e-8, reltol=None, maxiters=50, termination_condition=None, store_trace=True)
    args.update(kw)
    return api._build_opts(prob, alg, **args)


def test_defaults_match_the_reference(nls):
    abi = nls.abi
    o = _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()))
    assert (o.linsolve, o.globalization, o.descent, o.forcing, o.precond) == (abi.LINSOLVE_GMRES, abi.GLOB_NONE, abi.DESCENT_NEWTON, 0, abi.PRECOND_NONE)
    assert o.termination == 0 and o.abstol == 1e-8 and o.reltol == 0.0 and o.maxiters == 50 and o.store_trace == 1
    assert o.gmres.memory == 20 and o.gmres.restart == 0 and o.gmres.itmax == 0 and o.gmres.warm_start == 0      # KrylovJL_GMRES(): memory 20, no restart
    assert o.gmres.atol == 0.0 and o.gmres.rtol == 0.0                                                          # inherit abstol / reltol (solve.jl:203)
    assert o.max_shrink_times == 32 and (o.ew_eta0, o.ew_eta_max, o.ew_gamma, o.ew_alpha) == (0.5, 0.9, 0.9, 2.0)  # eisenstat_walker.jl:18-30
    assert o.jvp_mode == abi.JVP_EXACT
    # linsolve = nothing: dense LU for a dense J, the sparse direct factorisation for a sparse one (LinearSolve's defaults)
    assert _opts(nls, nls.NewtonRaphson()).linsolve == abi.LINSOLVE_DENSE_LU
    fs = nls.NonlinearFunction(nls.Brusselator2D(8), sparsity=nls.TracerSparsityDetector())
    assert _opts(nls, nls.NewtonRaphson(), f=fs).linsolve == abi.LINSOLVE_SPARSE_LU
    assert _opts(nls, nls.NewtonRaphson(linsolve=nls.KLUFactorization()), f=fs).linsolve == abi.LINSOLVE_SPARSE_LU
    assert _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), f=fs).linsolve == abi.LINSOLVE_SPARSE_GMRES
    assert _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(), concrete_jac=True)).linsolve == abi.LINSOLVE_SPARSE_GMRES


def test_algorithm_options_reach_the_struct(nls):
    abi = nls.abi
    o = _opts(nls, nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(), radius_update_scheme=nls.RadiusUpdateSchemes.Fan, max_trust_radius=7.0, shrink_factor=0.3))
    assert (o.globalization, o.tr_scheme, o.tr_max_trust_radius, o.tr_shrink_factor) == (abi.GLOB_TRUST_REGION, abi.TR_FAN, 7.0, 0.3)
    assert o.tr_step_threshold == 0.0 and o.tr_expand_factor == 0.0          # 0 -> the scheme's own default (trust_region.jl:320-328)
    o = _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(), linesearch=nls.BackTracking(c_1=1e-3, maxiters=7)))
    assert (o.globalization, o.ls_c1, o.ls_maxiters) == (abi.GLOB_LINESEARCH, 1e-3, 7)
    o = _opts(nls, nls.PseudoTransient(alpha_initial=0.25, linsolve=nls.KrylovJL_GMRES(precs=nls.BlockJacobi("right"), orth="mgs", gmres_restart=30)))
    assert (o.descent, o.pt_alpha_initial, o.precond) == (abi.DESCENT_PSEUDO_TRANSIENT, 0.25, abi.PRECOND_BLOCK_JACOBI_RIGHT)
    assert (o.gmres.orth, o.gmres.restart) == (abi.ORTH_MGS, 30)
    o = _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.Multigrid("right"))), maxtime=2.5,
              termination_condition=nls.RelNormSafeBestTerminationMode(norm="l2", max_stalled_steps=None))
    assert (o.precond, o.maxtime, o.termination, o.term_norm, o.term_max_stalled_steps) == (abi.PRECOND_MULTIGRID_RIGHT, 2.5, abi.TERM_REL_NORM_SAFE_BEST, abi.NORM_L2, -1)
    o = _opts(nls, nls.TrustRegion(radius_update_scheme=nls.RadiusUpdateSchemes.Bastin))
    assert o.tr_scheme == abi.TR_BASTIN == 6
    o = _opts(nls, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(), forcing=nls.EisenstatWalkerForcing2(), jvp_autodiff=nls.AutoFiniteDiff()),
              termination_condition=nls.AbsNormTerminationMode())
    assert o.forcing == abi.FORCING_EW2 and o.jvp_mode == abi.JVP_FINITE_DIFF and o.termination == nls.AbsNormTerminationMode().code
    o = _opts(nls, nls.Broyden(init_jacobian="true_jacobian", update_rule="bad_broyden", max_resets=7, reset_tolerance=1e-9, alpha=0.5))
    assert (o.descent, o.qn_init_jacobian, o.qn_update_rule, o.qn_max_resets, o.qn_reset_tolerance, o.qn_alpha, o.linsolve) == (
        abi.DESCENT_BROYDEN, abi.QN_INIT_TRUE_JACOBIAN, abi.QN_UPDATE_BAD_BROYDEN, 7, 1e-9, 0.5, abi.LINSOLVE_DENSE_LU)
    o = _opts(nls, nls.Broyden())
    assert (o.descent, o.qn_init_jacobian, o.qn_update_rule, o.qn_max_resets, o.qn_reset_tolerance, o.qn_alpha) == (abi.DESCENT_BROYDEN, 0, 0, 100, 0.0, 0.0)
    assert nls.ReturnCode.name(nls.ReturnCode.ConvergenceFailure) == "ConvergenceFailure"
    with pytest.raises(ValueError):
        nls.Broyden(update_rule="diagonal")
    with pytest.raises(TypeError):
        _opts(nls, nls.NewtonRaphson(linsolve=object()))


def test_retcodes_and_sharding(nls):
    assert nls.successful_retcode(nls.ReturnCode.Success) and nls.successful_retcode(nls.ReturnCode.StalledSuccess)
    assert not nls.successful_retcode(nls.ReturnCode.MaxIters) and nls.ReturnCode.name(nls.ReturnCode.Stalled) == "Stalled"
    from nonlinearsolve_jl_b200.api import shard_range
    for K in (1, 7, 8192, 8193):
        for R in (1, 2, 3, 8):
            blocks = [shard_range(K, r, R) for r in range(R)]
            assert blocks[0][0] == 0 and blocks[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))                 # contiguous, no gap, no overlap
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_product_fails_loudly_without_a_gpu(nls):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nls.abi.B200Error) as e:
        nls.Context(0)
    assert e.value.code == nls.abi.ERR_NO_DEVICE
    with pytest.raises((nls.abi.B200Error, RuntimeError)):
        nls.solve(nls.NonlinearProblem(nls.Brusselator2D(8), np.zeros(128), (3.4, 1.0, 10.0)), nls.NewtonRaphson())
