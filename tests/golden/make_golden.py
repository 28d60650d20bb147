"""Generates tests/golden/*.json.

The reference (Rust) cannot be executed in this image (no cargo/rustc), so the golden vectors are
the known-answer data its OWN tests define, evaluated in float64 exactly as those tests do:

  fixture_{2d,3d}   the six-particle `acceleration_error!` fixture and its closed form
                    (reference particular/src/gravity/newtonian/mod.rs:228-277):
                    acc_i = sum_j dir * mu_j * sqrt(1/|dir|^2) / |dir|^2 over the three massive ones.
  doctest_forces    the three-body force identities of the crate-level doctest (lib.rs:247-261).

plus REGRESSION vectors of the CPU oracle on a small seeded cloud (`oracle_cloud_*`), which pin the
oracle build against accidental change; they are outputs of oracle/, not of the reference.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def fixture(dim):
    splat = lambda v: [float(v)] * dim  # noqa: E731
    massive = [(splat(0.0), 20.0), (splat(1.0), 30.0), (splat(-3.0), 40.0)]
    particles = [(splat(10.0), 0.0), massive[0], massive[1], massive[2], (splat(30.0), 0.0),
                 (splat(-45.0), 0.0)]
    expected = []
    for pos1, _ in particles:
        acc = np.zeros(dim)
        for pos2, mu2 in massive:
            d = np.array(pos2) - np.array(pos1)
            mag2 = float(d @ d)
            if mag2 != 0.0:
                acc += d * mu2 * np.sqrt(1.0 / mag2) / mag2
        expected.append(acc.tolist())
    return {"particles": [list(p) + [m] for p, m in particles], "expected": expected,
            "tolerance": {"brute_force": 1e-2, "barnes_hut_theta0": 1e-2, "barnes_hut_theta05": 5e-1},
            "source": "gravity/newtonian/mod.rs:228-277, 385-418"}


def doctest_forces():
    G = 4.0 * np.pi * np.pi
    bodies = [([0.0, 0.0, 0.0], 1.0), ([1.0, 0.0, 0.0], 3.0027e-6), ([5.2, 0.0, 0.0], 0.000954588)]

    def force(a, b):
        pa, ma = np.array(a[0]), a[1]
        pb_, mb = np.array(b[0]), b[1]
        if (pa == pb_).all():
            return np.zeros(3)
        r = pb_ - pa
        l2 = float(r @ r)
        return r * (G * ma * mb / (l2 * np.sqrt(l2)))

    se, sj, ej = force(bodies[0], bodies[1]), force(bodies[0], bodies[2]), force(bodies[1], bodies[2])
    return {"G": G, "bodies": [list(p) + [m] for p, m in bodies],
            "forces": [(se + sj).tolist(), (-se + ej).tolist(), (-sj - ej).tolist()],
            "source": "lib.rs:247-261"}


def oracle_cloud():
    import oracle
    rng = np.random.default_rng(1808)
    out = {}
    for name, dt, d in (("f32x3", np.float32, 3), ("f32x2", np.float32, 2), ("f64x3", np.float64, 3)):
        n = 48
        p = np.concatenate([rng.uniform(-5e3, 5e3, (n, d)), rng.uniform(1e3, 1e9, (n, 1))], axis=1)
        p[5, -1] = 0.0
        p[17, :d] = p[3, :d]  # a coincident pair
        p = p.astype(dt)
        out[name] = {
            "particles": p.astype(np.float64).tolist(),
            "brute_force": oracle.brute_force(p[:, :d], p).astype(np.float64).tolist(),
            "brute_force_softened_1.5": oracle.brute_force(p[:, :d], p, 1.5).astype(np.float64).tolist(),
            "barnes_hut_0.5": oracle.barnes_hut(p[:, :d], p, 0.5).astype(np.float64).tolist(),
        }
    return out


if __name__ == "__main__":
    gold = {"fixture_3d": fixture(3), "fixture_2d": fixture(2), "doctest_forces": doctest_forces()}
    json.dump(gold, open(os.path.join(HERE, "reference_kat.json"), "w"), indent=1)
    json.dump(oracle_cloud(), open(os.path.join(HERE, "oracle_regression.json"), "w"))
    print("wrote", os.listdir(HERE))
