"""Pins the CPU oracle against every known answer the reference's own tests hold for this path
 — the reference itself (Rust) cannot be executed in this image.

  acceleration_error!  gravity/newtonian/mod.rs:228-277   (six-particle fixture + closed form)
  circular_orbit!      gravity/newtonian/mod.rs:281-347   (drift over many orbits)
  doctest              lib.rs:247-261                     (fold identities)
  barnes_hut theta=0   gravity/newtonian/mod.rs:409-413   (== brute force tolerance)
"""
import json
import os

import numpy as np
import pytest

import oracle
from tests.conftest import rel_err, uniform_cloud

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))
REGR = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_regression.json")))


def fixture_error(computed, expected):
    """The reference's assertion: || 1 - computed / expected || (component-wise division)."""
    return np.linalg.norm(1.0 - np.asarray(computed, np.float64) / np.asarray(expected), axis=1)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fixture_brute_force(dim, dtype):
    fx = GOLD[f"fixture_{dim}d"]
    p = np.array(fx["particles"], dtype=dtype)
    aff, src = oracle.between_of_reordered(p)  # the test wraps the particles in Reordered
    assert len(src) == 3
    got = oracle.brute_force(aff, src)
    err = fixture_error(got, fx["expected"])
    assert err.max() <= fx["tolerance"]["brute_force"]
    # the restatement is far tighter than the reference's own 1e-2 bound
    assert err.max() <= (1e-6 if dtype == np.float32 else 1e-14)
    got_par = oracle.brute_force_parallel(aff, src)
    assert np.array_equal(got, got_par)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("theta,key", [(0.0, "barnes_hut_theta0"), (0.5, "barnes_hut_theta05")])
def test_fixture_barnes_hut(dim, dtype, theta, key):
    fx = GOLD[f"fixture_{dim}d"]
    p = np.array(fx["particles"], dtype=dtype)
    aff, src = oracle.between_of_reordered(p)
    got = oracle.barnes_hut(aff, src, theta)
    assert fixture_error(got, fx["expected"]).max() <= fx["tolerance"][key]
    assert np.array_equal(got, oracle.barnes_hut(aff, src, theta, parallel=True))


def test_fixture_simd_baseline():
    """brute_force_simd_8 (gravity/newtonian/mod.rs:389-399): same fixture, 1e-2."""
    fx = GOLD["fixture_3d"]
    p = np.array(fx["particles"], dtype=np.float32)
    aff, src = oracle.between_of_reordered(p)
    got = oracle.brute_force_simd8_parallel(aff, src)
    assert fixture_error(got, fx["expected"]).max() <= 1e-2


def semi_implicit_orbit(compute, dtype, orbits):
    """circular_orbit! with `compute(particles) -> accelerations`."""
    dt = dtype(1.0 / 60.0)
    particles = np.array([[0, 0, 0, 1e6], [100, 0, 0, 0]], dtype=dtype)
    vel = np.array([[0, 0, 0], [0, 100, 0]], dtype=dtype)
    dist0 = np.linalg.norm(particles[0, :3] - particles[1, :3])
    energy = lambda r: -(1e6 + 0.0) / (r + r)  # noqa: E731
    period = 2 * np.pi * np.sqrt(dist0 ** 3 / 1e6)
    steps = int(round(period / float(dt)))
    assert steps == 377
    for _ in range(steps * orbits):
        acc = compute(particles)
        vel = (vel + acc * dt).astype(dtype)
        particles[:, :3] = particles[:, :3] + vel * dt
    dist1 = np.linalg.norm(particles[0, :3] - particles[1, :3])
    return abs(1.0 - dist0 / dist1), abs(1.0 - energy(dist0) / energy(dist1))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_circular_orbit_brute_force(dtype):
    e_d, e_e = semi_implicit_orbit(lambda p: oracle.brute_force(p[:, :3], p), dtype, 60)
    assert e_d < 1e-2 and e_e < 1e-2


def test_circular_orbit_barnes_hut():
    e_d, e_e = semi_implicit_orbit(lambda p: oracle.barnes_hut(p[:, :3], p, 0.5), np.float32, 30)
    assert e_d < 1e-1 and e_e < 1e-1


def test_semi_implicit_euler_helper_is_the_reference_loop():
    """oracle.semi_implicit_euler (the checker of the device-resident stepping tests) reproduces
    the circular_orbit! loop above step for step, bit for bit."""
    for dtype in (np.float32, np.float64):
        dt = dtype(1.0 / 60.0)
        particles = np.array([[0, 0, 0, 1e6], [100, 0, 0, 0]], dtype=dtype)
        vel = np.array([[0, 0, 0], [0, 100, 0]], dtype=dtype)
        p_ref, v_ref = particles.copy(), vel.copy()
        for _ in range(50):
            acc = oracle.brute_force(p_ref[:, :3], p_ref)
            v_ref = (v_ref + acc * dt).astype(dtype)
            p_ref[:, :3] = p_ref[:, :3] + v_ref * dt
        p, v, a = oracle.semi_implicit_euler(lambda aff, src: oracle.brute_force(aff, src),
                                             particles, vel, float(dt), 50)
        assert p.dtype == dtype and np.array_equal(p, p_ref) and np.array_equal(v, v_ref)
    # massive_only: massless particles are affected but do not affect (Reordered storage)
    p0 = np.array([[0, 0, 0, 1.0], [1, 0, 0, 0.0], [0, 2, 0, 0.0]], dtype=np.float64)
    _, _, a = oracle.semi_implicit_euler(lambda aff, src: oracle.brute_force(aff, src), p0, None,
                                         0.1, 1, massive_only=True)
    assert np.array_equal(a[0], [0, 0, 0]) and np.allclose(a[1], [-1, 0, 0]) and np.allclose(a[2], [0, -0.25, 0])


def test_doctest_fold_identities():
    """lib.rs:247-261: forces[i] is the left fold 0 + f(i,0) + f(i,1) + f(i,2), bit-exactly, in f64.
    Stated for accelerations: out[i] == ((0 + t(i,0)) + t(i,1)) + t(i,2) with t the pair term."""
    b = np.array(GOLD["doctest_forces"]["bodies"], dtype=np.float64)
    out = oracle.brute_force(b[:, :3], b)
    for i in range(3):
        acc = np.zeros(3)
        for j in range(3):
            acc = acc + oracle.brute_force(b[i:i + 1, :3], b[j:j + 1])[0]
        assert np.array_equal(out[i], acc)
    # the doctest's own numbers: force = G * m_i * acceleration_i / G-free mu; check to 1e-15
    G = GOLD["doctest_forces"]["G"]
    mu = b.copy()
    mu[:, 3] *= G
    forces = oracle.brute_force(mu[:, :3], mu) * b[:, 3:4]
    assert np.allclose(forces, GOLD["doctest_forces"]["forces"], rtol=1e-14, atol=0)


def test_barnes_hut_theta0_equals_brute_force():
    p = uniform_cloud(600)
    bf = oracle.brute_force(p[:, :3], p)
    bh = oracle.barnes_hut(p[:, :3], p, 0.0)
    assert rel_err(bh, bf).max() < 2e-5  # same terms, different summation order


def test_checked_and_unchecked_coincident():
    p = np.array([[1, 2, 3, 5.0], [1, 2, 3, 7.0], [4, 4, 4, 1.0]], dtype=np.float32)
    assert np.isfinite(oracle.brute_force(p[:, :3], p, 0.0, True)).all()
    assert np.isnan(oracle.brute_force(p[:, :3], p, 0.0, False)).any()  # 0 * inf, as the reference
    assert np.isfinite(oracle.brute_force(p[:, :3], p, 0.5, False)).all()


def test_empty_inputs():
    p = uniform_cloud(5)
    assert oracle.brute_force(np.zeros((0, 3), np.float32), p).shape == (0, 3)
    z = oracle.brute_force(p[:, :3], np.zeros((0, 4), np.float32))
    assert z.shape == (5, 3) and not z.any()


def test_storage_semantics():
    """storage.rs:61-95, 153-163, 207-241."""
    p = uniform_cloud(20, massive_ratio=0.5)
    rng = np.random.default_rng(0)
    p = p[rng.permutation(20)]
    aff, src = oracle.between_of_reordered(p)
    assert np.array_equal(aff, p[:, :3]) and (src[:, 3] != 0).all() and len(src) == 10
    assert np.array_equal(src, p[p[:, 3] != 0])  # stable
    aff_o, src_o = oracle.between_of_ordered(p)
    assert np.array_equal(aff_o[:10], src[:, :3]) and np.array_equal(src_o, src)
    aff_s, src_s = oracle.between_of_slice(p)
    assert np.array_equal(aff_s, p[:, :3]) and src_s is not None and len(src_s) == 20


@pytest.mark.parametrize("name", ["f32x3", "f32x2", "f64x3"])
def test_oracle_regression(name):
    r = REGR[name]
    dt = np.float64 if name.startswith("f64") else np.float32
    p = np.array(r["particles"], dtype=dt)
    d = p.shape[1] - 1
    assert np.array_equal(oracle.brute_force(p[:, :d], p).astype(np.float64), np.array(r["brute_force"]))
    assert np.array_equal(oracle.brute_force(p[:, :d], p, 1.5).astype(np.float64),
                          np.array(r["brute_force_softened_1.5"]))
    assert np.array_equal(oracle.barnes_hut(p[:, :d], p, 0.5).astype(np.float64),
                          np.array(r["barnes_hut_0.5"]))


def test_exact_mode_agrees():
    p = uniform_cloud(300)
    assert rel_err(oracle.brute_force(p[:, :3], p), oracle.brute_force_exact(p[:, :3], p)).max() < 1e-5


def test_octree_spec_invariants():
    """Our own tree specification (parity unpinned by the reference): keys sorted, permutation
    stable, children partition their parent, root moments = total mass / centre of mass."""
    p = uniform_cloud(5000)
    p[100:110, :3] = p[100, :3]  # duplicates share a key
    t = oracle.Octree(p, nleaf=8)
    assert (np.diff(t.keys.astype(np.uint64)) >= 0).all()
    assert sorted(t.perm.tolist()) == list(range(5000))
    same = t.keys[1:] == t.keys[:-1]
    assert (t.perm[1:][same] > t.perm[:-1][same]).all()
    keys = oracle.morton_keys(p[:, :3], t.origin, t.inv)
    assert np.array_equal(keys[t.perm], t.keys)
    for j in range(t.n_nodes):
        if t.n_child[j]:
            c0, nc = t.first_child[j], t.n_child[j]
            assert t.begin[c0] == t.begin[j]
            assert t.count[c0:c0 + nc].sum() == t.count[j]
            assert (t.level[c0:c0 + nc] == t.level[j] + 1).all()
        else:
            assert t.count[j] <= 8 or t.level[j] == 21
    m = p[:, 3].astype(np.float64)
    assert np.isclose(t.commass[0, 3], m.sum(), rtol=1e-6)
    com = (p[:, :3].astype(np.float64) * m[:, None]).sum(0) / m.sum()
    assert np.allclose(t.commass[0, :3], com, rtol=1e-5, atol=1e-2)
