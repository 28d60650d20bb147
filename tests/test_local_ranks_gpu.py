"""Every multi-GPU path on ONE GPU: `world` ranks as threads over the library's in-process communicator
(tests/local_ranks.py), results against the CPU oracle — the restated sequential::BruteForce for the
sharded brute force and the split, the restated sequential::BarnesHut at equal theta (error statistics
against the extended-precision sum) for every Barnes-Hut build: locally essential trees
(the default), the partitioned build, the replicated build, with both result routings."""
import numpy as np
import pytest

import oracle
from tests.conftest import assert_bruteforce_parity, plummer_cloud, rel_err, uniform_cloud
from tests.local_ranks import LocalWorld

pytestmark = pytest.mark.gpu


def stats(e):
    return np.array([np.median(e), np.percentile(e, 99), e.max()])


def set_debug(**kv):
    from particular_b200._ffi import lib
    for k, v in kv.items():
        assert lib.pcuda_debug_set(k.encode(), v) == 0, (k, v)


# "let": the own tree is walked while the others' trees travel (the default from 8 ranks on); "let_one_walk": one
# walk at the end (the default below 8 ranks); "let_stop": the first phase stops when the others' trees are here
BUILDS = {"let": dict(bh_forest=3, bh_let_overlap=1), "let_one_walk": dict(bh_forest=3, bh_let_overlap=0),
          "let_stop": dict(bh_forest=3, bh_let_overlap=1, bh_let_stop=1),
          "partitioned": dict(bh_forest=1), "replicated": dict(bh_forest=2)}
RESET = dict(bh_forest=0, bh_route=0, bh_let_overlap=-1, bh_let_stop=0)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_brute_force_and_split(world):
    p = uniform_cloud(20011, seed=3)
    r = uniform_cloud(16001, seed=6, massive_ratio=0.01)
    with LocalWorld(world) as w:
        pb = w.pb
        got = w.brute_force(p, pb.AccelerationSoftened.checked(2.0))
        assert_bruteforce_parity(got, oracle.brute_force_parallel(p[:, :3], p, 2.0), p[:, :3], p, 2.0)
        gots = w.between(pb.Reordered.new(r), pb.AccelerationSoftened.checked(1.0))
        aff, src = oracle.between_of_reordered(r)
        assert_bruteforce_parity(gots, oracle.brute_force_parallel(aff, src, 1.0), aff, src, 1.0)


@pytest.mark.parametrize("build", list(BUILDS))
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("cloud", ["plummer", "uniform"])
def test_barnes_hut_same_theta_error_as_reference(world, build, cloud):
    n = 60000 if cloud == "plummer" else 50000
    p = plummer_cloud(n, seed=5) if cloud == "plummer" else uniform_cloud(n, seed=5)
    exact = oracle.brute_force_exact(p[:, :3], p)
    for theta in (0.5, 1.0):
        s_ref = stats(rel_err(oracle.barnes_hut(p[:, :3], p, theta, parallel=True), exact))
        try:
            set_debug(**BUILDS[build])
            with LocalWorld(world) as w:
                outs = []
                for route in (1, 2):
                    set_debug(bh_route=route)
                    got = w.barnes_hut(p, theta)
                    assert got.shape == (n, 3) and np.isfinite(got).all()
                    outs.append(got)
                    s = stats(rel_err(got, exact))
                    assert (s <= 1.1 * s_ref + 2e-6).all(), (build, world, theta, route, s, s_ref)
                if not build.startswith("let"):  # (the locally essential trees have one routing)
                    assert np.array_equal(outs[0], outs[1])
        finally:
            set_debug(**RESET)


@pytest.mark.parametrize("build", list(BUILDS))
@pytest.mark.parametrize("world", [2, 5, 8])
def test_barnes_hut_theta0_is_brute_force(world, build):
    p = uniform_cloud(9000, seed=8)
    ref = oracle.brute_force_parallel(p[:, :3], p)
    try:
        set_debug(**BUILDS[build])
        with LocalWorld(world) as w:
            got = w.barnes_hut(p, 0.0)
        assert_bruteforce_parity(got, ref, p[:, :3], p, aggregate=False, plain=False)
    finally:
        set_debug(**RESET)


@pytest.mark.parametrize("overlap", [1, 0])
@pytest.mark.parametrize("world", [2, 8])
def test_let_degenerate_inputs(world, overlap):
    """Locally essential trees with inputs that leave ranks empty or put everything into one cell: all
    particles at one point, two tight clumps far apart, fewer distinct keys than ranks, and a block
    order that is already sorted in space (every rank's block goes to one destination)."""
    rng = np.random.default_rng(11)
    cases = {}
    one = uniform_cloud(4000, seed=2)
    one[:, :3] = one[0, :3]
    cases["one point"] = one
    two = uniform_cloud(6000, seed=3)
    two[:3000, :3] = two[0, :3] + rng.normal(scale=1e-3, size=(3000, 3)).astype(np.float32)
    two[3000:, :3] = -two[0, :3] + rng.normal(scale=1e-3, size=(3000, 3)).astype(np.float32)
    cases["two clumps"] = two
    srt = plummer_cloud(30000, seed=4)
    srt = srt[np.lexsort((srt[:, 2], srt[:, 1], srt[:, 0]))]
    cases["sorted blocks"] = srt
    try:
        set_debug(bh_forest=3, bh_let_overlap=overlap)
        with LocalWorld(world) as w:
            for name, p in cases.items():
                got = w.barnes_hut(p, 0.5)
                assert np.isfinite(got).all(), name
                exact = oracle.brute_force_exact(p[:, :3], p)
                s_ref = stats(rel_err(oracle.barnes_hut(p[:, :3], p, 0.5, parallel=True), exact))
                s = stats(rel_err(got, exact))
                assert (s <= 1.1 * s_ref + 2e-6).all(), (name, s, s_ref)
    finally:
        set_debug(**RESET)


@pytest.mark.parametrize("overlap", [-1, 1])
def test_let_is_the_default_and_matches_one_gpu_at_size(overlap):
    """N = 2M Plummer on 4 ranks, the automatic build (locally essential trees from 65536 particles per
    rank on), with the automatic choice of the walk (one walk below 8 ranks) and with the two-phase walk:
    same error statistics as the single tree on one GPU, sampled against the exact sum."""
    import particular_b200 as pb
    n = 2_000_000
    p = plummer_cloud(n, seed=1808)
    idx = np.sort(np.random.default_rng(1).choice(n, 1024, replace=False))
    exact = oracle.brute_force_exact(p[idx, :3], p)
    with pb.CudaContext(0) as c1:
        single = pb.BarnesHut(c1, 0.5, pb.Acceleration.checked()).compute(p)
    try:
        set_debug(bh_let_overlap=overlap)
        with LocalWorld(4) as w:
            got = w.barnes_hut(p, 0.5)
            comm = [c.timings() for c in w.ctxs]
    finally:
        set_debug(**RESET)
    s1, s4 = stats(rel_err(single[idx], exact)), stats(rel_err(got[idx], exact))
    print("one GPU", s1, "4 ranks, locally essential trees", s4, "timings rank 0", comm[0])
    assert (s4 <= 1.1 * s1 + 2e-6).all(), (s4, s1)
