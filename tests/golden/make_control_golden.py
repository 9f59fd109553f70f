"""Generates tests/golden/control_sequences.json from oracle/newton_numpy.py (the independent NumPy restatement of the
reference's scalar control logic).  Deterministic; tests/test_control_sequences.py regenerates it and compares with the
committed file, then holds BOTH oracle/oracle.c and the CUDA driver to these sequences.

    python tests/golden/make_control_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import newton_numpy as nn  # noqa: E402

SCHEMES = ["Simple", "NLsolve", "NocedalWright", "Hei", "Yuan", "Fan", "Bastin"]
MODES = ["AbsNormSafeBest", "AbsNorm", "AbsNormSafe", "Norm", "Rel", "RelNorm", "Abs", "RelNormSafe", "RelNormSafeBest"]


def cases():
    out = []
    # the reference start, and starts far enough out (10x, 30x) that steps get rejected and the line search backtracks
    for N, scale in ((6, 1.0), (6, 10.0), (6, 30.0)):
        base = dict(problem="bruss2d", N=N, u0_scale=scale, abstol=1e-8, reltol=1e-8)
        out.append(dict(base, name="newton_N%d_x%g" % (N, scale)))
        for s in SCHEMES:
            if (scale <= 10.0 or s == "Simple") and not (s == "Bastin" and scale > 1.0):  # Bastin from 10x takes 269 steps
                out.append(dict(base, name="tr_%s_N%d_x%g" % (s, N, scale), globalization="trust_region", tr_scheme=s))
        out.append(dict(base, name="linesearch_N%d_x%g" % (N, scale), globalization="linesearch"))
        if scale <= 10.0:
            out.append(dict(base, name="pseudotransient_N%d_x%g" % (N, scale), descent="pseudo_transient", alpha_initial=1.0, maxiters=200))
    for scale in (1.0, 5.0, 30.0):   # 5x: two geodesic rejections; 30x: a long accept / reject pattern with the damping doubling and tripling
        out.append(dict(problem="bruss2d", N=6, u0_scale=scale, name="levenberg_marquardt_N6_x%g" % scale, descent="levenberg_marquardt", abstol=1e-8, reltol=1e-8, maxiters=300))
    out.append(dict(problem="bruss2d", N=6, u0_scale=10.0, name="levenberg_marquardt_nogeodesic_N6_x10", descent="levenberg_marquardt", disable_geodesic=True, abstol=1e-8,
                    reltol=1e-8, maxiters=300))
    out.append(dict(problem="quadratic", n=10, name="levenberg_marquardt_quadratic", descent="levenberg_marquardt", abstol=1e-9, reltol=1e-9))
    out.append(dict(problem="quadratic", n=10, name="pseudotransient_quadratic_alpha10", descent="pseudo_transient", alpha_initial=10.0, abstol=1e-9, reltol=1e-9))
    for m in MODES:
        for norm in ("inf", "l2"):
            out.append(dict(problem="bruss2d", N=6, u0_scale=1.0, name="term_%s_%s" % (m, norm), termination=m, term_norm=norm, abstol=1e-7, reltol=1e-9))
    out.append(dict(problem="bruss2d", N=6, u0_scale=1.0, name="maxiters_2", maxiters=2, abstol=1e-12, reltol=1e-12))
    # Broyden (quasi-Newton): both initialisations, both update rules, a reset, the max_resets exit
    q = dict(problem="quadratic", n=10, u0_linspace=[0.6, 3.0], descent="broyden", abstol=1e-11, reltol=1e-11, maxiters=200)
    out.append(dict(q, name="broyden_quadratic_good"))
    out.append(dict(q, name="broyden_quadratic_bad", update_rule="bad_broyden"))
    out.append(dict(q, name="broyden_quadratic_true_jacobian_with_reset", init_jacobian="true_jacobian"))
    out.append(dict(q, name="broyden_quadratic_max_resets", reset_tolerance=0.05, max_resets=1))
    b = dict(problem="bruss2d", N=6, u0_scale=1.0, descent="broyden", init_jacobian="true_jacobian", abstol=1e-8, reltol=1e-8, maxiters=200)
    out.append(dict(b, name="broyden_bruss2d_true_jacobian_good"))
    out.append(dict(b, name="broyden_bruss2d_true_jacobian_bad", update_rule="bad_broyden"))
    # LimitedMemoryBroyden: threshold 10 (the circular buffer wraps), and threshold 3 (every update after the third evicts one)
    out.append(dict(q, name="lbroyden_quadratic", init_jacobian="low_rank", max_resets=3))
    out.append(dict(q, name="lbroyden_quadratic_threshold3", init_jacobian="low_rank", threshold=3, max_resets=3, maxiters=60))
    # Klement's diagonal rule divides a numerator that cancels to rounding level by J^2 du^2 ~ 1e-15 once |f| ~ 1e-7: past that point
    # its iterates depend on the last bit of the residual (a fused multiply-add in f is enough).  The sequence is pinned up to 1e-6.
    out.append(dict(q, name="klement_quadratic", descent="klement", abstol=1e-6, reltol=1e-6))
    return out


def run(case):
    if case["problem"] == "bruss2d":
        prob = nn.Brusselator2D(case["N"])
        u0 = prob.u0() * case["u0_scale"]
    else:
        prob = nn.Quadratic(case["n"])
        u0 = np.linspace(*case["u0_linspace"], case["n"]) if "u0_linspace" in case else np.ones(case["n"])
    term = nn.Termination(mode=case.get("termination", "AbsNormSafeBest"), norm=case.get("term_norm", "inf"), abstol=case["abstol"], reltol=case["reltol"])
    if case.get("descent") == "levenberg_marquardt":
        r = nn.solve_lm(prob, u0, disable_geodesic=case.get("disable_geodesic", False), termination=term, maxiters=case.get("maxiters", 1000))
        u = r.pop("u")
        r["u_norm2"] = float(np.linalg.norm(u))
        r["u_first"] = [float(x) for x in u[:4]]
        return r
    if case.get("descent") == "klement":
        r = nn.solve_klement(prob, u0, termination=term, maxiters=case.get("maxiters", 1000))
        u = r.pop("u")
        r["u_norm2"] = float(np.linalg.norm(u))
        r["u_first"] = [float(x) for x in u[:4]]
        return r
    if case.get("descent") == "broyden":
        r = nn.solve_broyden(prob, u0, init_jacobian=case.get("init_jacobian", "identity"), update_rule=case.get("update_rule", "good_broyden"),
                             max_resets=case.get("max_resets", 100), reset_tolerance=case.get("reset_tolerance"), termination=term, maxiters=case.get("maxiters", 1000),
                             threshold=case.get("threshold", 10))
        u = r.pop("u")
        r["u_norm2"] = float(np.linalg.norm(u))
        r["u_first"] = [float(x) for x in u[:4]]
        r["reset"] = [int(x) for x in r["reset"]]
        return r
    r = nn.solve(prob, u0, globalization=case.get("globalization", "none"), tr_scheme=case.get("tr_scheme", "Simple"), descent=case.get("descent", "newton"),
                 alpha_initial=case.get("alpha_initial", 1e-3), termination=term, maxiters=case.get("maxiters", 1000))
    u = r.pop("u")
    r["u_norm2"] = float(np.linalg.norm(u))
    r["u_first"] = [float(x) for x in u[:4]]
    return r


def generate():
    return [dict(case=c, expect=run(c)) for c in cases()]


if __name__ == "__main__":
    data = generate()
    path = os.path.join(ROOT, "tests", "golden", "control_sequences.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    for d in data:
        e = d["expect"]
        print("%-40s rc=%d nsteps=%d nf=%d njacs=%d acc=%s" % (d["case"]["name"], e["retcode"], e["nsteps"], e["nf"], e["njacs"], "".join(map(str, e["accepted"]))[:40]))
