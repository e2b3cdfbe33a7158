#!/usr/bin/env python3
"""Golden vectors for the interpolating pose front-end (the TSDF_Python variant).

Run in the development container, where /root/reference exists: the reference's own `slerp`
(src/TSDF_Python/tsdf_utils.py:80-100) is extracted from its source file with `ast` (the module itself cannot be
imported: it pulls in SDL2 / PyOpenGL at the top) and evaluated together with the translation lerp of
src/TSDF_Python/main.py:127-138 on seeded random pose pairs.  Writes tests/golden/pose_interp.npz
(inputs + the reference's outputs); nothing of the reference's text is stored."""
import ast
import math
import os

import numpy as np

REF = "/root/reference/src/TSDF_Python/tsdf_utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_interp.npz")


def reference_slerp():
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "slerp")
    ns = {"np": np, "math": math}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["slerp"]


def main():
    slerp = reference_slerp()
    rng = np.random.default_rng(20261018)
    a, b, ts, out = [], [], [], []
    for case in range(200):
        t0 = rng.uniform(0, 1e5)
        t1 = t0 + rng.uniform(1e-3, 0.2)
        qa = rng.standard_normal(4)
        if case % 4 == 0:      # nearly identical rotations: the linear branch (dot > 0.9995)
            qb = qa + 1e-3 * rng.standard_normal(4)
        elif case % 4 == 1:    # opposite hemisphere: the sign flip
            qb = -qa + 0.3 * rng.standard_normal(4)
        else:
            qb = rng.standard_normal(4)
        if case % 7 == 0:      # un-normalised input quaternions (the function normalises)
            qa, qb = qa * 3.0, qb * 0.5
        else:
            qa, qb = qa / np.linalg.norm(qa), qb / np.linalg.norm(qb)
        pa = np.concatenate([[t0], rng.uniform(-2, 2, 3), qa])
        pb = np.concatenate([[t1], rng.uniform(-2, 2, 3), qb])
        stamp = t0 + (t1 - t0) * (rng.uniform(0, 1) if case % 10 else (0.0 if case % 20 else 1.0))
        t = (stamp - pa[0]) / (pb[0] - pa[0])                      # main.py:133
        pos = (pb[1:4] - pa[1:4]) * t + pa[1:4]                    # main.py:135
        q = slerp(pa[-4:].copy(), pb[-4:].copy(), t)               # main.py:136
        a.append(pa); b.append(pb); ts.append(stamp); out.append(np.concatenate([pos, q]))
    np.savez(OUT, pose_a=np.array(a), pose_b=np.array(b), stamp=np.array(ts), pose_out=np.array(out))
    print("wrote", OUT, len(a), "cases")


if __name__ == "__main__":
    main()
