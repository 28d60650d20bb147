"""GPU parity of the Barnes-Hut path (K2-K5) through the C ABI.

  * tree build: Morton keys, sort permutation and every node array are BIT-EXACT against the CPU
    statement of our tree specification (oracle.Octree; the reference has no Morton code, so this
    part is pinned CPU-vs-GPU only — "parity unpinned" by the reference);
  * traversal: same theta-approximation error as the reference algorithm at equal theta
    (SURVEY.md 8c): with a_exact the extended-precision brute-force sum, the median / p99 / max of
    ||a - a_exact|| / ||a_exact|| of the GPU must be <= 1.1 x those of the oracle restatement of
    sequential::BarnesHut (sequential.rs:466-505); theta = 0 must meet the brute-force tolerance.
"""
import json
import os

import numpy as np
import pytest

import oracle
from tests.conftest import assert_bruteforce_parity, plummer_cloud, rel_err, uniform_cloud

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))


@pytest.fixture(scope="module")
def pb():
    import particular_b200 as pb
    return pb


def stats(e):
    return np.array([np.median(e), np.percentile(e, 99), e.max()])


def assert_same_theta_error(a_gpu, a_ref, a_exact, slack=1.1, floor=2e-6):
    e_gpu, e_ref = rel_err(a_gpu, a_exact), rel_err(a_ref, a_exact)
    s_gpu, s_ref = stats(e_gpu), stats(e_ref)
    assert (s_gpu <= slack * s_ref + floor).all(), (s_gpu, s_ref)
    return s_gpu, s_ref


def check_tree(pb, ctx, p, nleaf):
    from particular_b200 import _ffi
    import particular_b200.interface as pi
    c2 = pi.CudaContext(0, leaf_size=nleaf)
    try:
        t = pi.RootedOrthtree(c2, p)
        o = oracle.Octree(p, nleaf=nleaf)
        assert t.n_nodes == o.n_nodes and t.n_levels == o.n_levels
        d = p.shape[1] - 1
        assert np.array_equal(np.array(t.info.origin[:d], np.float32), o.origin)
        assert np.float32(t.info.extent) == np.float32(o.ext)
        assert np.float32(t.info.inv) == np.float32(o.inv)
        assert np.array_equal(t.read(_ffi.TREE_KEYS), o.keys)
        assert np.array_equal(t.read(_ffi.TREE_PERM), o.perm)
        assert np.array_equal(t.read(_ffi.TREE_NODE_BEGIN), o.begin)
        assert np.array_equal(t.read(_ffi.TREE_NODE_COUNT), o.count)
        assert np.array_equal(t.read(_ffi.TREE_NODE_LEVEL), o.level)
        assert np.array_equal(t.read(_ffi.TREE_NODE_FIRST_CHILD), o.first_child)
        assert np.array_equal(t.read(_ffi.TREE_NODE_NUM_CHILDREN), o.n_child)
        cm = t.read(_ffi.TREE_NODE_COM_MASS)
        assert np.array_equal(cm.view(np.uint32), o.commass.view(np.uint32))  # bit-exact
        t.close()
    finally:
        c2.close()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n,nleaf", [(1, 16), (2, 16), (17, 16), (1000, 1), (5000, 8), (50000, 16),
                                     (200000, 32)])
def test_tree_bit_exact_uniform(pb, ctx, dim, n, nleaf):
    check_tree(pb, ctx, uniform_cloud(n, d=dim, seed=n), nleaf)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nleaf", [1, 16, 32])
def test_tree_bit_exact_one_pass_build(pb, ctx, dim, nleaf):
    """The one-pass (radix) build of bh_radix_build.cu against the CPU statement of the tree
    specification: clustered cloud above the single-block size, long runs of coincident particles
    (leaves of unbounded size on the last level), massless particles, keys that differ in the last
    digit only; then the same build forced at sizes where build_small normally runs (tile and
    window edges: n around the leaf size, around one warp, around one 2048-particle tile), and the
    level-wise build it replaces as a cross-check."""
    from particular_b200._ffi import lib
    p = plummer_cloud(120_000, d=dim, seed=31 + nleaf)
    p[500:3700, :dim] = p[500, :dim]              # 3200 coincident particles
    p[9000:9033, :dim] = p[9000, :dim]            # 33: one more than the widest leaf
    p[20000:20100, dim] = 0.0                     # massless
    p[30000:30040, :dim] = p[30000, :dim] * (1 + np.arange(40)[:, None] * 2e-7)  # last-digit neighbours
    check_tree(pb, ctx, p, nleaf)
    try:
        assert lib.pcuda_debug_set(b"bh_level_build", 2) == 0
        for n in (1, 2, nleaf, nleaf + 1, 31, 32, 33, 2047, 2048, 2049, 4100, 30000):
            check_tree(pb, ctx, uniform_cloud(n, d=dim, seed=n + 7), nleaf)
        q = uniform_cloud(5000, d=dim, seed=3)
        q[:, :dim] = q[0, :dim]                   # every particle at one point: a single deep chain
        check_tree(pb, ctx, q, nleaf)
        assert lib.pcuda_debug_set(b"bh_level_build", 1) == 0
        check_tree(pb, ctx, p, nleaf)
    finally:
        assert lib.pcuda_debug_set(b"bh_level_build", 0) == 0


@pytest.mark.parametrize("dim", [2, 3])
def test_tree_bit_exact_clustered_and_degenerate(pb, ctx, dim):
    p = plummer_cloud(30000, d=dim, seed=3)
    p[100:140, :dim] = p[100, :dim]           # 40 coincident particles: a leaf at the last level
    p[200:210, dim] = 0.0                     # massless particles
    check_tree(pb, ctx, p, 16)
    q = uniform_cloud(50, d=dim, seed=1)
    q[:, :dim] = q[0, :dim]                   # every particle at one point: extent 0
    check_tree(pb, ctx, q, 16)
    z = uniform_cloud(300, d=dim, seed=2)
    z[:, dim] = 0.0                           # a tree of massless particles
    check_tree(pb, ctx, z, 4)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("theta,key", [(0.0, "barnes_hut_theta0"), (0.5, "barnes_hut_theta05")])
def test_reference_fixture(pb, ctx, dim, theta, key):
    """acceleration_error! for barnes_hut / barnes_hut_05 (gravity/newtonian/mod.rs:409-420)."""
    fx = GOLD[f"fixture_{dim}d"]
    p = np.array(fx["particles"], dtype=np.float32)
    got = pb.BarnesHut(ctx, theta, pb.Acceleration.checked()).compute(pb.Reordered.new(p))
    err = np.linalg.norm(1.0 - got.astype(np.float64) / np.array(fx["expected"]), axis=1)
    assert err.max() <= fx["tolerance"][key]
    got2 = pb.cuda_barnes_hut(pb.Reordered.new(p), ctx, theta, pb.Acceleration.checked())
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("cloud", ["uniform", "plummer"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("theta", [0.3, 0.5, 0.7, 1.0])
def test_same_theta_error_as_reference(pb, ctx, cloud, dim, theta):
    n = 20000
    p = uniform_cloud(n, d=dim, seed=5) if cloud == "uniform" else plummer_cloud(n, d=dim, seed=5)
    exact = oracle.brute_force_exact(p[:, :dim], p)
    ref = oracle.barnes_hut(p[:, :dim], p, theta, parallel=True)
    got = pb.BarnesHut(ctx, theta, pb.Acceleration.checked()).compute(p)
    assert got.shape == (n, dim) and np.isfinite(got).all()
    assert_same_theta_error(got, ref, exact)


@pytest.mark.parametrize("dim", [2, 3])
def test_theta0_is_brute_force(pb, ctx, dim):
    p = uniform_cloud(6000, d=dim, seed=8)
    exact = oracle.brute_force_exact(p[:, :dim], p)
    got = pb.BarnesHut(ctx, 0.0, pb.Acceleration.checked()).compute(p)
    ref32 = oracle.brute_force_parallel(p[:, :dim], p)
    assert_bruteforce_parity(got, ref32, p[:, :dim], p, aggregate=False, plain=False)
    c = pb.BarnesHut(ctx, 0.0, pb.Acceleration.checked()).last_counters()
    assert c["particle_interactions"] == 6000 * 6000 and c["node_interactions"] == 0


@pytest.mark.parametrize("cloud", ["uniform", "plummer"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("theta", [0.3, 0.7])
def test_f64_same_theta_error_as_reference(pb, ctx, cloud, dim, theta):
    """DVec2 / DVec3 (pcuda_barneshut_f64x*): same error-distribution criterion, against the f64
    restatement of sequential::BarnesHut."""
    n = 12000
    p = (uniform_cloud(n, d=dim, seed=6, dtype=np.float64) if cloud == "uniform"
         else plummer_cloud(n, d=dim, seed=6, dtype=np.float64))
    exact = oracle.brute_force_exact(p[:, :dim], p)
    ref = oracle.barnes_hut(p[:, :dim], p, theta, parallel=True)
    got = pb.BarnesHut(ctx, theta, pb.Acceleration.checked()).compute(p)
    assert got.dtype == np.float64 and got.shape == (n, dim) and np.isfinite(got).all()
    assert_same_theta_error(got, ref, exact)


@pytest.mark.parametrize("dim", [2, 3])
def test_f64_theta0_is_f64_brute_force(pb, ctx, dim):
    """theta = 0 opens every cell: the result must meet the f64 brute-force bound (1e-12), i.e.
    nothing single precision enters an acceleration; separate targets and softening included."""
    p = uniform_cloud(5000, d=dim, seed=9, dtype=np.float64)
    p[100:110, :dim] = p[100, :dim]                      # coincident particles
    p[200, :dim] = p[201, :dim] * (1.0 + 1e-12)          # distinct in f64, equal in f32
    got = pb.BarnesHut(ctx, 0.0, pb.Acceleration.checked()).compute(p)
    ref = oracle.brute_force_parallel(p[:, :dim], p)
    assert_bruteforce_parity(got, ref, p[:, :dim], p, aggregate=False, plain=False)
    aff = uniform_cloud(777, d=dim, seed=10, dtype=np.float64)[:, :dim]
    got = pb.BarnesHut(ctx, 0.0, pb.AccelerationSoftened.checked(2.0)).compute(pb.Between(aff, p))
    ref = oracle.brute_force_parallel(aff, p, 2.0)
    assert_bruteforce_parity(got, ref, aff, p, 2.0, aggregate=False, plain=False)


def test_f64_device_api_reference_fixture_and_empty(pb, ctx):
    import torch
    p = plummer_cloud(9000, seed=12, dtype=np.float64)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    host = bh.compute(p)
    d_p = torch.from_numpy(p).cuda()
    d_o = torch.zeros((len(p), 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    bh.compute_device(None, len(p), d_p.data_ptr(), len(p), d_o.data_ptr(), "f64x3")
    ctx.sync()
    assert np.array_equal(d_o.cpu().numpy(), host)
    assert np.array_equal(bh.compute(p), host)            # deterministic
    # the reference's own fixture (gravity/newtonian/mod.rs:228-277, 407-418) in f64
    for dim, theta, key in ((3, 0.0, "barnes_hut_theta0"), (3, 0.5, "barnes_hut_theta05"), (2, 0.5, "barnes_hut_theta05")):
        fx = GOLD[f"fixture_{dim}d"]
        q = np.array(fx["particles"], dtype=np.float64)
        got = pb.BarnesHut(ctx, theta, pb.Acceleration.checked()).compute(pb.Reordered.new(q))
        err = np.linalg.norm(1.0 - got / np.array(fx["expected"]), axis=1)
        assert err.max() <= fx["tolerance"][key]
    # empty inputs
    assert bh.compute(np.zeros((0, 4), np.float64)).shape == (0, 3)
    z = bh.compute(pb.Between(p[:9, :3], np.zeros((0, 4), np.float64)))
    assert z.shape == (9, 3) and not z.any()
    one = bh.compute(p[:1])
    assert one.shape == (1, 3) and not one.any()


@pytest.mark.parametrize("cloud,dim", [("plummer", 3), ("uniform", 3), ("uniform", 2)])
def test_quadrupole_nodes(pb, ctx, cloud, dim):
    """pcuda_config.expansion_order = 2 (an accuracy / speed knob beyond the
    reference's centre-of-mass nodes): at equal theta the error against the exact sum drops (the
    leading error term goes from the quadrupole, O(theta^2), to the octupole, O(theta^3)), so it
    is a fortiori inside the reference's error distribution; theta = 0 is still brute force."""
    import particular_b200.interface as pi
    n = 20000
    p = plummer_cloud(n, d=dim, seed=21) if cloud == "plummer" else uniform_cloud(n, d=dim, seed=21)
    exact = oracle.brute_force_exact(p[:, :dim], p)
    cq = pi.CudaContext(0, expansion_order=2)
    try:
        for theta in (0.5, 0.8):
            mono = pi.BarnesHut(ctx, theta, pi.Acceleration.checked()).compute(p)
            quad = pi.BarnesHut(cq, theta, pi.Acceleration.checked()).compute(p)
            assert np.isfinite(quad).all()
            s_m, s_q = stats(rel_err(mono, exact)), stats(rel_err(quad, exact))
            assert s_q[0] <= 0.8 * s_m[0] and s_q[1] <= 0.7 * s_m[1], (theta, s_m, s_q)
            ref = oracle.barnes_hut(p[:, :dim], p, theta, parallel=True)
            assert_same_theta_error(quad, ref, exact)
        # theta = 0: no node is ever accepted
        small = p[:3000]
        got = pi.BarnesHut(cq, 0.0, pi.Acceleration.checked()).compute(small)
        assert_bruteforce_parity(got, oracle.brute_force_parallel(small[:, :dim], small), small[:, :dim], small,
                                 aggregate=False, plain=False)
        # separate targets, softening, tiny inputs, a prebuilt tree
        aff = uniform_cloud(501, d=dim, seed=3)[:, :dim] * 1e-3
        a_q = pi.BarnesHut(cq, 0.5, pi.AccelerationSoftened.checked(0.01)).compute(pi.Between(aff, p))
        a_x = oracle.brute_force_exact(aff, p, 0.01)
        a_m = pi.BarnesHut(ctx, 0.5, pi.AccelerationSoftened.checked(0.01)).compute(pi.Between(aff, p))
        assert np.median(rel_err(a_q, a_x)) <= np.median(rel_err(a_m, a_x))
        for k in (1, 2, 17):
            assert np.isfinite(pi.BarnesHut(cq, 0.5, pi.Acceleration.checked()).compute(p[:k])).all()
        tree = pi.RootedOrthtree(cq, p)
        a_t = pi.BarnesHut(cq, 0.5, pi.Acceleration.checked()).compute(pi.Between(p[:, :dim], tree))
        assert np.array_equal(a_t, pi.BarnesHut(cq, 0.5, pi.Acceleration.checked()).compute(p))
        tree.close()
    finally:
        cq.close()
    with pytest.raises(pi.CudaError):
        pi.CudaContext(0, expansion_order=3)


def test_two_tight_clusters_beyond_the_key_resolution(pb, ctx):
    """Depth is limited by the key resolution (extent / 2^21 per axis; the reference subdivides until
    positions differ, tree/mod.rs:112-134).  Two clumps of 2000 particles, 2e-5 wide, 100 apart: the
    deepest cells (4.8e-5 wide) hold hundreds of particles each — leaves far beyond leaf_size that are
    summed directly whenever they are opened.  The result must stay as accurate as the reference's
    (the cost is the documented O(k^2) per clump, include/particular_cuda.h)."""
    from particular_b200 import _ffi
    rng = np.random.default_rng(17)
    p = uniform_cloud(4000, seed=17)
    p[:2000, :3] = np.float32(50.0) + rng.normal(scale=2e-5, size=(2000, 3)).astype(np.float32)
    p[2000:, :3] = np.float32(-50.0) + rng.normal(scale=2e-5, size=(2000, 3)).astype(np.float32)
    t = pb.RootedOrthtree(ctx, p)
    leaves = t.read(_ffi.TREE_NODE_NUM_CHILDREN) == 0
    assert t.n_levels == 22 and t.read(_ffi.TREE_NODE_COUNT)[leaves].max() > 100
    t.close()
    exact = oracle.brute_force_exact(p[:, :3], p)
    got = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(p)
    ref = oracle.barnes_hut(p[:, :3], p, 0.5, parallel=True)
    assert np.isfinite(got).all()
    assert_same_theta_error(got, ref, exact)


def test_softened_and_unchecked(pb, ctx):
    p = plummer_cloud(15000, seed=6)
    eps = 0.01
    exact = oracle.brute_force_exact(p[:, :3], p, eps)
    ref = oracle.barnes_hut(p[:, :3], p, 0.5, eps, parallel=True)
    got = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(eps)).compute(p)
    assert_same_theta_error(got, ref, exact)
    got_u = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.unchecked(eps)).compute(p)
    assert np.array_equal(got, got_u)
    # unchecked without softening: the traversal never evaluates a zero-distance pair
    # (sequential.rs:485-487), so the result stays finite, as in the reference
    got_u0 = pb.BarnesHut(ctx, 0.5, pb.Acceleration.unchecked()).compute(p)
    ref_u0 = oracle.barnes_hut(p[:, :3], p, 0.5, 0.0, False, parallel=True)
    assert np.isfinite(got_u0).all() and np.isfinite(ref_u0).all()


def test_separate_targets_and_reordered(pb, ctx):
    src = plummer_cloud(20000, seed=9)
    rng = np.random.default_rng(4)
    aff = rng.normal(size=(7777, 3)).astype(np.float32) * 3.0   # some far outside the root cube
    exact = oracle.brute_force_exact(aff, src)
    ref = oracle.barnes_hut(aff, src, 0.5, parallel=True)
    got = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(pb.Between(aff, src))
    assert_same_theta_error(got, ref, exact)
    # Reordered: tree over the massive particles only (storage.rs:219-229), all particles affected
    p = uniform_cloud(30000, seed=10, massive_ratio=0.1)
    p = p[rng.permutation(len(p))]
    a2, s2 = oracle.between_of_reordered(p)
    exact = oracle.brute_force_exact(a2, s2)
    ref = oracle.barnes_hut(a2, s2, 0.5, parallel=True)
    got = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(pb.Reordered.new(p))
    assert_same_theta_error(got, ref, exact)


def test_split_phase_matches_one_shot(pb, ctx):
    p = plummer_cloud(12000, seed=12)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    one = bh.compute(pb.Between(p[:, :3], p))
    tree = pb.RootedOrthtree(ctx, p)
    two = bh.compute(pb.Between(p[:, :3], tree))
    assert np.array_equal(one, two)
    c = bh.last_counters()
    assert c["node_tests"] > 0 and c["node_interactions"] > 0 and c["particle_interactions"] > 0
    alias = bh.compute(p)   # affected == affecting: traversal in the tree's own order
    assert np.array_equal(alias, one)
    tree.close()


def test_device_api_and_empty(pb, ctx):
    import torch
    p = plummer_cloud(9000, seed=13)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    host = bh.compute(p)
    d_src = torch.from_numpy(p).cuda()
    d_out = torch.zeros((len(p), 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    bh.compute_device(None, len(p), d_src.data_ptr(), len(p), d_out.data_ptr())
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy(), host)
    t = ctx.timings()
    assert t["build_ms"] > 0 and t["compute_ms"] > 0 and t["kernel_launches"] > 10
    assert bh.compute(np.zeros((0, 4), np.float32)).shape == (0, 3)
    z = bh.compute(pb.Between(p[:5, :3], np.zeros((0, 4), np.float32)))
    assert z.shape == (5, 3) and not z.any()
    one = bh.compute(p[:1])
    assert one.shape == (1, 3) and not one.any()


def test_circular_orbit(pb, ctx):
    from tests.test_oracle_golden import semi_implicit_orbit
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    e_d, e_e = semi_implicit_orbit(lambda p: bh.compute(p), np.float32, 5)
    assert e_d < 1e-1 and e_e < 1e-1


def test_full_size_plummer(pb, ctx):
    """BASELINE configs[3] size: N = 10M Plummer sphere, theta = 0.5.  The reference algorithm is
    too slow to run on every target here, so both are judged on a fixed sample of targets against
    the extended-precision sum over all 10M sources; plus structural invariants of the tree."""
    from particular_b200 import _ffi
    n = 10_000_000
    p = plummer_cloud(n, seed=1808)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    got = bh.compute(p)
    assert np.isfinite(got).all()
    c = bh.last_counters()
    per_target = (c["node_interactions"] + c["particle_interactions"]) / n
    assert 500 < per_target < 50000, per_target
    idx = np.sort(np.random.default_rng(1).choice(n, 1024, replace=False))  # >= 1024 at N = 10M
    exact = oracle.brute_force_exact(p[idx, :3], p)
    e_gpu = rel_err(got[idx], exact)
    tree = oracle.Tree(p)                       # the reference's recursive build, single thread
    ref = tree.traverse(p[idx, :3], 0.5, parallel=True)
    e_ref = rel_err(ref, exact)
    print(f"N=10M theta=0.5, 1024 targets: median / p99 / max  gpu {stats(e_gpu)}  reference {stats(e_ref)}")
    assert (stats(e_gpu) <= 1.1 * stats(e_ref) + 2e-6).all(), (stats(e_gpu), stats(e_ref))
    # sortedness + permutation (checksum of the index set) at full size
    t = pb.RootedOrthtree(ctx, p)
    keys = t.read(_ffi.TREE_KEYS)
    assert (keys[1:] >= keys[:-1]).all()
    perm = t.read(_ffi.TREE_PERM)
    assert int(perm.astype(np.uint64).sum()) == n * (n - 1) // 2
    cm = t.read(_ffi.TREE_NODE_COM_MASS)
    assert np.isclose(cm[0, 3], p[:, 3].astype(np.float64).sum(), rtol=1e-6)
    t.close()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n", [1, 2, 1000, 100003])
def test_morton_entry_bit_exact(pb, ctx, dim, n):
    """pcuda_morton_f32x3 / _f32x2: keys, stable permutation and root cube equal the CPU
    specification (oracle/oracle_octree.inc) bit for bit, and the keys the tree build uses."""
    p = plummer_cloud(n, d=dim, seed=n + dim) if n > 2 else uniform_cloud(n, d=dim, seed=n)
    if n > 100:
        p[10:30, :dim] = p[10, :dim]  # equal keys: ties must keep input order
    keys, perm, info = pb.morton_keys(ctx, p)
    o = oracle.Octree(p, nleaf=16)
    assert np.array_equal(keys, o.keys) and np.array_equal(perm, o.perm)
    assert np.array_equal(np.array(info.origin[:dim], np.float32), o.origin)
    assert np.float32(info.extent) == np.float32(o.ext) and np.float32(info.inv) == np.float32(o.inv)
    assert np.array_equal(oracle.morton_keys(p[:, :dim], o.origin, o.inv)[perm], keys)
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    ties = np.flatnonzero(np.diff(keys.astype(np.int64)) == 0)
    assert (perm[ties] < perm[ties + 1]).all()
    k0, p0, _ = pb.morton_keys(ctx, np.zeros((0, dim + 1), np.float32))
    assert k0.shape == (0,) and p0.shape == (0,)


def test_sharded_entry_single_rank(pb, ctx):
    """The multi-GPU step degenerates to the plain evaluation without a communicator."""
    p = plummer_cloud(10000, seed=14)
    got = pb.ShardedBarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(p)
    ref = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(pb.Between(p[:, :3], p))
    assert np.array_equal(got, ref)
    from particular_b200 import _ffi
    import ctypes as C
    out = np.zeros((5, 3), np.float32)
    st = _ffi.lib.pcuda_barneshut_f32x3_sharded(ctx.handle, p.ctypes.data_as(C.c_void_p), 5, 10000, 0.5,
                                                0.0, 1, out.ctypes.data_as(C.c_void_p))
    assert st == _ffi.ERR_INVALID_ARGUMENT  # a rank must own its whole block


def test_full_size_2d_quadtree(pb, ctx):
    """BASELINE configs[4b]: 2-D f32 Barnes-Hut quadtree, N = 4,194,304, theta = 0.5, softening 100
    (the particle-toy shape, examples/particle-toy/src/nbody.rs:30).  Sampled error statistics
    against the extended-precision sum and the reference algorithm."""
    n = 4_194_304
    p = uniform_cloud(n, d=2, seed=1808)
    it = pb.AccelerationSoftened.checked(100.0)
    bh = pb.BarnesHut(ctx, 0.5, it)
    got = bh.compute(p)
    assert got.shape == (n, 2) and np.isfinite(got).all()
    idx = np.sort(np.random.default_rng(3).choice(n, 256, replace=False))
    exact = oracle.brute_force_exact(p[idx, :2], p, 100.0)
    tree = oracle.Tree(p)
    ref = tree.traverse(p[idx, :2], 0.5, 100.0, parallel=True)
    assert_same_theta_error(got[idx], ref, exact, slack=1.25)


# ---- key-range-partitioned build (the multi-GPU "forest" path run as virtual ranks on one GPU) ----
def test_partitioned_one_part_is_the_ordinary_tree(pb, ctx):
    p = plummer_cloud(30000, seed=21)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    assert np.array_equal(bh.compute_partitioned(p, 1), bh.compute(p))


@pytest.mark.parametrize("cloud", ["uniform", "plummer"])
@pytest.mark.parametrize("parts", [2, 3, 8, 16])
def test_partitioned_same_theta_error_as_reference(pb, ctx, cloud, parts):
    n = 20000
    p = uniform_cloud(n, seed=5) if cloud == "uniform" else plummer_cloud(n, seed=5)
    exact = oracle.brute_force_exact(p[:, :3], p)
    ref = oracle.barnes_hut(p[:, :3], p, 0.5, parallel=True)
    got = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute_partitioned(p, parts)
    assert got.shape == (n, 3) and np.isfinite(got).all()
    assert_same_theta_error(got, ref, exact)


@pytest.mark.parametrize("parts", [2, 5])
def test_partitioned_theta0_is_brute_force(pb, ctx, parts):
    p = uniform_cloud(6000, seed=8)
    got = pb.BarnesHut(ctx, 0.0, pb.Acceleration.checked()).compute_partitioned(p, parts)
    ref32 = oracle.brute_force_parallel(p[:, :3], p)
    assert_bruteforce_parity(got, ref32, p[:, :3], p, aggregate=False, plain=False)


def test_partitioned_degenerate_inputs(pb, ctx):
    """Fewer particles than parts, coincident particles (every key equal: one part gets them all),
    softened interaction, and the argument checks."""
    bh = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(0.5))
    tiny = uniform_cloud(3, seed=2)
    assert np.allclose(bh.compute_partitioned(tiny, 8), oracle.brute_force(tiny[:, :3], tiny, 0.5),
                       rtol=1e-5, atol=0)
    same = np.tile(np.array([[1.0, 2.0, 3.0, 5.0]], np.float32), (500, 1))
    assert np.array_equal(bh.compute_partitioned(same, 4), np.zeros((500, 3), np.float32))
    two = np.concatenate([same, same + np.array([[10.0, 0, 0, 0]], np.float32)])
    got = bh.compute_partitioned(two, 4)
    ref = oracle.brute_force(two[:, :3], two, 0.5)
    assert np.allclose(got, ref, rtol=1e-4, atol=0)
    assert bh.compute_partitioned(np.zeros((0, 4), np.float32), 2).shape == (0, 3)
    with pytest.raises(pb.CudaError):
        bh.compute_partitioned(tiny, 0)
    with pytest.raises(pb.CudaError):
        bh.compute_partitioned(tiny, 17)


def test_partitioned_full_size_counts(pb, ctx):
    """N = 2M Plummer, 8 parts: the forest walk costs about the same work as the single tree
    (partial top cells add a few node interactions) and gives the same error distribution."""
    n = 2_000_000
    p = plummer_cloud(n, seed=31)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    one = bh.compute(p)
    forest = bh.compute_partitioned(p, 8)
    idx = np.random.default_rng(3).choice(n, 1024, replace=False)
    exact = oracle.brute_force_exact(p[idx, :3], p)
    e1, e8 = rel_err(one[idx], exact), rel_err(forest[idx], exact)
    assert np.median(e8) <= 1.1 * np.median(e1) + 2e-6, (np.median(e8), np.median(e1))
    assert np.percentile(e8, 99) <= 1.25 * np.percentile(e1, 99) + 2e-6
    assert np.isfinite(forest).all()
