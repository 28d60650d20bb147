"""Device-resident stepping (pcuda_sim_*) against the CPU oracle's statement
of the caller loop: accelerations -> velocity += a*dt -> position += velocity*dt
(examples/simple/src/main.rs:45-59; circular_orbit!, gravity/newtonian/mod.rs:281-347)."""
import numpy as np
import pytest

import oracle
from tests.conftest import plummer_cloud, rel_err, uniform_cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import particular_b200 as pb
    return pb


def _velocities(n, d, dtype, seed=3, scale=10.0):
    return (np.random.default_rng(seed).normal(size=(n, d)) * scale).astype(dtype)


@pytest.mark.parametrize("dim,dtype", [(3, np.float32), (2, np.float32), (3, np.float64), (2, np.float64)])
def test_one_step_integrator_is_bit_exact(pb, ctx, dim, dtype):
    """Given the accelerations the device itself produced, velocities and positions are the
    reference's unfused `v += a*dt; p += v*dt`, bit for bit."""
    n = 1537
    p0 = uniform_cloud(n, d=dim, dtype=dtype, seed=11)
    v0 = _velocities(n, dim, dtype)
    dt = 1.0 / 60.0
    with pb.Simulation(pb.BruteForce(ctx, pb.Acceleration.checked()), p0, v0, dt=dt) as sim:
        sim.step(1)
        p1, v1, a1 = sim.read(True, True, True)
    dts = dtype(dt)
    v_ref = v0 + a1 * dts
    p_ref = p0[:, :dim] + v_ref * dts
    assert np.array_equal(v1, v_ref)
    assert np.array_equal(p1[:, :dim], p_ref)
    assert np.array_equal(p1[:, dim], p0[:, dim])  # masses untouched
    # and the accelerations are the one-shot operator's
    a_op = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p0)
    assert np.array_equal(a1, a_op)


@pytest.mark.parametrize("dim,dtype,tol", [(3, np.float32, 2e-4), (2, np.float32, 2e-4),
                                           (3, np.float64, 1e-10), (2, np.float64, 1e-10)])
@pytest.mark.parametrize("graph", [True, False])
def test_brute_force_trajectory_matches_oracle(pb, ctx, dim, dtype, tol, graph):
    """40 steps (eager step + CUDA-graph blocks + eager tail) against the oracle loop.  Softened
    so that no close encounter amplifies the per-step rounding differences."""
    n, steps, dt = 700, 43, 1e-3
    p0 = uniform_cloud(n, d=dim, dtype=dtype, seed=5)
    p0[:, :dim] *= 1e-3            # box of 10 units
    p0[:, dim] *= 1e-6             # mu up to 1e3
    v0 = _velocities(n, dim, dtype, scale=1.0)
    eps = 0.5
    it = pb.AccelerationSoftened.checked(eps)
    with pb.Simulation(pb.BruteForce(ctx, it), p0, v0, dt=dt, graph=graph) as sim:
        sim.step(steps)
        p1, v1, a1 = sim.read(True, True, True)
        info = sim.info()
    assert info["steps_done"] == steps
    assert bool(info["graph_active"]) == graph
    pr, vr, ar = oracle.semi_implicit_euler(
        lambda aff, src: oracle.brute_force(aff, src, eps, True), p0, v0, dt, steps)
    span = np.abs(pr[:, :dim] - p0[:, :dim]).max()
    assert np.abs(p1[:, :dim] - pr[:, :dim]).max() <= tol * max(span, 1.0)
    assert rel_err(v1, vr).max() <= tol * 10
    assert np.percentile(rel_err(a1, ar), 99) <= tol * 10


def test_graph_and_eager_agree_bitwise(pb, ctx):
    n = 2048
    p0 = uniform_cloud(n, seed=9)
    v0 = _velocities(n, 3, np.float32)
    outs = []
    for graph in (True, False):
        with pb.Simulation(pb.BruteForce(ctx, pb.Acceleration.checked()), p0, v0, dt=1e-3,
                           graph=graph) as sim:
            sim.step(20)
            outs.append(sim.read(True, True, True))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_massive_only_is_the_reordered_storage(pb, ctx):
    """affecting='massive' == Reordered: everyone is affected, only mu != 0 affects
    (storage.rs:153-163, 219-229); massless particles are test particles."""
    n = 3000
    p0 = uniform_cloud(n, seed=21, massive_ratio=0.1)
    rng = np.random.default_rng(0)
    p0 = p0[rng.permutation(n)]  # interleave massive and massless
    v0 = _velocities(n, 3, np.float32)
    dt, steps = 1e-3, 11
    bf = pb.BruteForce(ctx, pb.Acceleration.checked())
    with pb.Simulation(bf, p0, v0, dt=dt, affecting="massive") as sim:
        assert sim.info()["n_affecting"] == int((p0[:, 3] != 0).sum())
        sim.step(1)
        p1, v1, a1 = sim.read(True, True, True)
        a_op = bf.compute(pb.Reordered(p0))
        assert np.array_equal(a1, a_op)
        sim.step(steps - 1)
        p2 = sim.particles()
    pr, _, _ = oracle.semi_implicit_euler(lambda aff, src: oracle.brute_force(aff, src), p0, v0, dt,
                                          steps, massive_only=True)
    assert np.abs(p2[:, :3] - pr[:, :3]).max() <= 1e-3 * np.abs(pr[:, :3] - p0[:, :3]).max() + 1e-2


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim", [3, 2])
def test_barnes_hut_stepping(pb, ctx, dim, dtype):
    """Barnes-Hut steps: first-step accelerations are the one-shot operator's; the trajectory
    stays within the theta-approximation of the exact (brute-force) trajectory."""
    n, steps, dt, theta = 4000, 5, 1e-4, 0.5
    p0 = plummer_cloud(n, d=dim, seed=4, dtype=dtype)
    v0 = np.zeros((n, dim), dtype)
    bh = pb.BarnesHut(ctx, theta, pb.AccelerationSoftened.checked(0.01))
    with pb.Simulation(bh, p0, v0, dt=dt) as sim:
        sim.step(1)
        a1 = sim.accelerations()
        assert np.array_equal(a1, bh.compute(p0))
        sim.step(steps - 1)
        p1 = sim.particles()
        assert not sim.info()["graph_active"]
    pr, _, _ = oracle.semi_implicit_euler(
        lambda aff, src: oracle.brute_force(aff, src, 0.01, True), p0, v0, dt, steps)
    move = np.linalg.norm(pr[:, :dim] - p0[:, :dim], axis=1)
    err = np.linalg.norm(p1[:, :dim] - pr[:, :dim], axis=1)
    assert np.median(err / np.maximum(move, 1e-30)) <= 2e-2


def test_circular_orbit_on_device(pb, ctx):
    """circular_orbit! (gravity/newtonian/mod.rs:281-347) run entirely on the device: 60 orbits of
    377 steps, distance and energy drift below the reference's 1e-2."""
    dt = np.float32(1.0 / 60.0)
    p0 = np.array([[0, 0, 0, 1e6], [100, 0, 0, 0]], dtype=np.float32)
    v0 = np.array([[0, 0, 0], [0, 100, 0]], dtype=np.float32)
    steps = 377 * 60
    with pb.Simulation(pb.BruteForce(ctx, pb.Acceleration.checked()), p0, v0, dt=float(dt)) as sim:
        sim.step(steps)
        p1 = sim.particles()
    d0 = 100.0
    d1 = np.linalg.norm(p1[0, :3].astype(np.float64) - p1[1, :3].astype(np.float64))
    energy = lambda r: -1e6 / (r + r)  # noqa: E731
    assert abs(1.0 - d0 / d1) < 1e-2
    assert abs(1.0 - energy(d0) / energy(d1)) < 1e-2
    # the oracle loop, same dt, lands in the same place to f32 accumulation noise
    pr, _, _ = oracle.semi_implicit_euler(lambda aff, src: oracle.brute_force(aff, src), p0, v0,
                                          float(dt), 377 * 3)
    with pb.Simulation(pb.BruteForce(ctx, pb.Acceleration.checked()), p0, v0, dt=float(dt)) as sim:
        sim.step(377 * 3)
        p3 = sim.particles()
    assert np.abs(p3[:, :3] - pr[:, :3]).max() < 0.05  # of a 100-unit orbit


def test_edge_cases_and_errors(pb, ctx):
    bf = pb.BruteForce(ctx, pb.Acceleration.checked())
    # empty system: steps are no-ops (CPU-path semantics: empty in, empty out)
    with pb.Simulation(bf, np.zeros((0, 4), np.float32), dt=0.1) as sim:
        sim.step(3)
        p, v, a = sim.read(True, True, True)
        assert p.shape == (0, 4) and v.shape == (0, 3) and a.shape == (0, 3)
        assert sim.info()["steps_done"] == 3
    # a single particle feels nothing and moves uniformly
    with pb.Simulation(bf, np.array([[1, 2, 3, 5]], np.float32), np.array([[1, 0, 0]], np.float32),
                       dt=0.5) as sim:
        sim.step(4)
        p, v, _ = sim.read()
        assert np.array_equal(p, np.array([[3, 2, 3, 5]], np.float32))
    # all massless + massive_only: no sources, zero accelerations
    with pb.Simulation(bf, np.array([[0, 0, 0, 0], [1, 0, 0, 0]], np.float32), dt=0.5,
                       affecting="massive") as sim:
        sim.step(9)
        assert np.array_equal(sim.accelerations(), np.zeros((2, 3), np.float32))
    # f64 Barnes-Hut with the massive-only storage: the tree holds the massive subset
    q = plummer_cloud(3000, seed=8, dtype=np.float64)
    q[::3, 3] = 0.0
    bh64 = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(0.01))
    with pb.Simulation(bh64, q, dt=1e-4, affecting="massive") as sim:
        sim.step(1)
        a = sim.accelerations()
    assert a.dtype == np.float64
    assert np.array_equal(a, bh64.compute(pb.Reordered.new(q)))
    # dt can be changed on a live simulation
    p0 = uniform_cloud(64, seed=1)
    with pb.Simulation(bf, p0, dt=1e-3) as sim:
        sim.step(10)
        sim.configure(dt=2e-3)
        sim.step(10)
        assert sim.info()["steps_done"] == 20
        assert np.isfinite(sim.particles()).all()
