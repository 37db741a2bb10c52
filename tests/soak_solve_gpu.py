"""Randomised GPU-vs-oracle soak of the refractive / in-air marker-pose solves (not collected by pytest; run on a B200).
Noise levels from 0 to 3e-2 normalised units (far beyond the 2e-4 of the logs: the corner scatter matrix then loses the clear
separation of its smallest eigenvalue, the hard case for the closed-form plane normal), a fraction of out-of-range markers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import orc  # noqa: E402
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402

cfg = capi.config_default()
f = BatchFilter(cfg, batch=1)
rng = np.random.default_rng(99)
bad = 0
for it, noise in enumerate([0.0, 2e-4, 2e-3, 1e-2, 3e-2] * 2):
    n = int(rng.integers(1000, 20000))
    Rm, p = synth.random_marker_poses(n, rng, far_fraction=0.05)
    under = it < 5
    corners = synth.marker_corners_from_pose(cfg, Rm, p, noise=noise, rng=rng) if under else synth.marker_corners_inair(cfg, Rm, p, noise=noise, rng=rng)
    if under:
        pose, c3, valid = f.RefractSolve(corners)
        po, co, vo = orc.refract_solve(cfg, corners)
    else:
        pose, c3, valid = f.InAirSolve(corners)
        po, co, vo = orc.inair_solve(cfg, corners)
    same_valid = np.array_equal(valid, vo)
    good = (vo == 1) & np.isfinite(po).all(axis=0)
    # quaternion sign is fixed by the reference's rule; compare directly
    e_p = float(np.abs(pose[:3, good] - po[:3, good]).max()) if good.any() else 0.0
    e_q = float(np.abs(pose[3:, good] - po[3:, good]).max()) if good.any() else 0.0
    e_c = float(np.abs(c3[:, good] - co[:, good]).max()) if good.any() else 0.0
    ok = same_valid and e_p <= 1e-8 and e_q <= 1e-8 and e_c <= 1e-8
    bad += 0 if ok else 1
    print(f"{it:2d} {'refract' if under else 'in-air '} n={n:5d} noise={noise:.0e} valid={'ok' if same_valid else 'DIFF'} ({int(good.sum())} good) "
          f"pos={e_p:.1e} quat={e_q:.1e} corners={e_c:.1e} {'' if ok else '  <-- FAIL'}", flush=True)
print(f"solve soak: {10 - bad}/10 agree with the oracle")

# ---- Gauss-Newton refinement against the NumPy restatement (bracketing root finder, LAPACK), a few dozen markers (slow oracle) ----
import fbus_oracle_np as onp  # noqa: E402
k = onp.Consts(onp.Config(tsc_left=np.array(cfg.tsc_left).reshape(4, 4), tsc_right=np.array(cfg.tsc_right).reshape(4, 4)))
bad_gn = 0
for noise in (0.0, 2e-4, 2e-3):
    Rm, p = synth.random_marker_poses(24, rng)
    corners = synth.marker_corners_from_pose(cfg, Rm, p, noise=noise, rng=rng)
    pose, cost, valid = f.RefractSolveGN(corners, iters=5)
    worst = 0.0
    for i in range(corners.shape[1]):
        if not valid[i]:
            continue
        pg, qg, _ = onp.refract_solve_gn(k, corners[:, i].astype(np.float64), 5)
        worst = max(worst, float(np.abs(pg - pose[:3, i]).max()), float(np.abs(qg - pose[3:, i]).max()))
    ok = worst <= 1e-8
    bad_gn += 0 if ok else 1
    print(f"GN 5 iterations, noise={noise:.0e}: max |GPU - NumPy oracle| = {worst:.1e} {'' if ok else '  <-- FAIL'}", flush=True)
sys.exit(1 if (bad or bad_gn) else 0)
