"""Property-based parity (hypothesis): random small problems — ragged sizes, 2-D / 3-D, f32 / f64,
softened or not, checked or not, coincident and massless particles, extreme coordinate and mass
scales — through the C ABI against the CPU oracle, plus metamorphic properties that need no oracle
(exact scaling laws, order of the outputs, determinism).  Complements the fixed cases of
test_bruteforce_gpu.py / test_barneshut_gpu.py."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from tests.conftest import assert_bruteforce_parity, rel_err, uniform_cloud

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def pb():
    import particular_b200 as pb
    return pb


SETTINGS = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@st.composite
def clouds(draw, max_n=300, dims=(2, 3), dtypes=(np.float32, np.float64)):
    dim = draw(st.sampled_from(dims))
    dtype = draw(st.sampled_from(dtypes))
    n_src = draw(st.integers(0, max_n))
    n_aff = draw(st.integers(0, max_n))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    pos_scale = draw(st.sampled_from([1e-3, 1.0, 5e3, 1e6]))
    mu_scale = draw(st.sampled_from([1e-6, 1.0, 1e9]))
    rng = np.random.default_rng(seed)
    src = np.concatenate([rng.uniform(-pos_scale, pos_scale, (n_src, dim)),
                          rng.uniform(0.1 * mu_scale, mu_scale, (n_src, 1))], axis=1).astype(dtype)
    aff = rng.uniform(-pos_scale, pos_scale, (n_aff, dim)).astype(dtype)
    if n_src >= 4 and draw(st.booleans()):  # coincident sources and massless sources
        src[1, :dim] = src[0, :dim]
        src[2, dim] = 0.0
    if n_src and n_aff and draw(st.booleans()):  # affected particles sitting on sources
        k = min(n_src, n_aff, 5)
        aff[:k] = src[:k, :dim]
    return dim, dtype, np.ascontiguousarray(src), np.ascontiguousarray(aff)


@settings(max_examples=80, **SETTINGS)
@given(c=clouds(), soft=st.sampled_from([0.0, 0.0, 1e-3, 2.5]), alias=st.booleans())
def test_bruteforce_matches_oracle(pb, ctx, c, soft, alias):
    dim, dtype, src, aff = c
    pos_scale = max(float(np.abs(src[:, :dim]).max()) if len(src) else 1.0, 1e-30)
    soft = dtype(soft * pos_scale / 5e3)  # softening in units of the cloud
    inter = pb.AccelerationSoftened.checked(float(soft)) if soft else pb.Acceleration.checked()
    if alias:  # &[P] => Between(slice, slice)
        got = pb.BruteForce(ctx, inter).compute(src)
        a = np.ascontiguousarray(src[:, :dim])
    else:
        got = pb.BruteForce(ctx, inter).compute(pb.Between(aff, src))
        a = aff
    assert got.shape == (len(a), dim) and got.dtype == dtype
    if len(a) == 0:
        return
    if len(src) == 0:
        assert not got.any()
        return
    ref = oracle.brute_force(a, src, float(soft))
    assert np.isfinite(got).all()
    assert_bruteforce_parity(got, ref, a, src, float(soft), aggregate=False)


@settings(max_examples=25, **SETTINGS)
@given(c=clouds(max_n=200, dtypes=(np.float32,)))
def test_unchecked_nan_pattern(pb, ctx, c):
    """Acceleration::unchecked(): a zero-distance pair gives 0 * inf = NaN in the reference
    (gravity/impls/mod.rs:160-165) — the GPU must poison exactly the same outputs."""
    dim, dtype, src, aff = c
    if len(src) == 0 or len(aff) == 0:
        return
    got = pb.BruteForce(ctx, pb.Acceleration.unchecked()).compute(pb.Between(aff, src))
    ref = oracle.brute_force(aff, src, 0.0, False)
    assert np.array_equal(np.isnan(got).any(axis=1), np.isnan(ref).any(axis=1))


@settings(max_examples=30, **SETTINGS)
@given(c=clouds(max_n=400, dtypes=(np.float32,)), k=st.sampled_from([2.0, 0.25, 1024.0]))
def test_exact_scaling_laws(pb, ctx, c, k):
    """Multiplying every mu by a power of two multiplies every acceleration by it, bit for bit
    (all products stay exact), for brute force and Barnes-Hut alike; scaling positions by k and
    masses by k^3 scales accelerations by k (same tree: the keys are scale invariant)."""
    dim, dtype, src, _ = c
    if len(src) < 2 or np.abs(src[:, :dim]).max() < 0.5:
        return  # (tiny clouds with huge masses approach the documented floor of `checked`)
    src = src.copy()
    src[:, dim] = np.abs(src[:, dim]) + dtype(1e-3)
    bf = pb.BruteForce(ctx, pb.Acceleration.checked())
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    heavier = src.copy()
    heavier[:, dim] *= dtype(k)
    for algo in (bf, bh):
        a0, a1 = algo.compute(src), algo.compute(heavier)
        assert np.array_equal(a1, a0 * dtype(k))
    bigger = src.copy()
    bigger[:, :dim] *= dtype(k)
    bigger[:, dim] *= dtype(k) ** 3
    for algo in (bf, bh):
        a0, a1 = algo.compute(src), algo.compute(bigger)
        assert rel_err(a1, a0 * dtype(k)).max() <= 2e-6


@settings(max_examples=40, **SETTINGS)
@given(c=clouds(max_n=500), theta=st.sampled_from([0.0, 0.3, 0.5, 1.0]),
       leaf=st.sampled_from([1, 4, 16]))
def test_barneshut_small_random(pb, ctx, c, theta, leaf):
    """theta = 0 is brute force; theta > 0 stays inside the reference algorithm's own error at the
    same theta (restated sequential::BarnesHut), on ragged / degenerate small inputs, for several
    leaf sizes; separate affected sets included."""
    import particular_b200.interface as pi
    dim, dtype, src, aff = c
    if len(src) == 0:
        return
    src = src.copy()
    src[:, dim] = np.abs(src[:, dim])
    c2 = pi.CudaContext(0, leaf_size=leaf)
    try:
        for a in ([None] if len(aff) == 0 else [None, aff]):
            tgt = np.ascontiguousarray(src[:, :dim]) if a is None else a
            storage = src if a is None else pi.Between(a, src)
            got = pi.BarnesHut(c2, theta, pi.Acceleration.checked()).compute(storage)
            assert got.shape == (len(tgt), dim) and np.isfinite(got).all()
            exact = oracle.brute_force_exact(tgt, src)
            if theta == 0.0:
                ref = oracle.brute_force(tgt, src)
                assert_bruteforce_parity(got, ref, tgt, src, aggregate=False)
            else:
                ref = oracle.barnes_hut(tgt, src, theta)
                e_gpu, e_ref = rel_err(got, exact), rel_err(ref, exact)
                # per-problem maxima are noisy at these sizes: bound by the reference's own
                # worst error with slack, and by its fixture bound for theta <= 0.5
                assert e_gpu.max() <= max(2.0 * e_ref.max(), 3e-2 * theta * theta) + 1e-5
    finally:
        c2.close()


def test_outputs_follow_affected_order(pb, ctx):
    """Permuting the affected particles permutes the outputs (sequential.rs:101-106), bit for bit
    for brute force (each target's fold over the sources is unchanged)."""
    rng = np.random.default_rng(11)
    src = np.concatenate([rng.uniform(-1, 1, (700, 3)), rng.uniform(1, 2, (700, 1))], axis=1).astype(np.float32)
    aff = rng.uniform(-1, 1, (900, 3)).astype(np.float32)
    perm = rng.permutation(len(aff))
    bf = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(0.01))
    a0 = bf.compute(pb.Between(aff, src))
    a1 = bf.compute(pb.Between(np.ascontiguousarray(aff[perm]), src))
    assert np.array_equal(a1, a0[perm])
    bh = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(0.01))
    b0 = bh.compute(pb.Between(aff, src))
    b1 = bh.compute(pb.Between(np.ascontiguousarray(aff[perm]), src))
    assert np.array_equal(b1, b0[perm])  # the target groups come from the sorted keys: same groups
    assert np.array_equal(bh.compute(pb.Between(aff, src)), b0)  # deterministic


def test_sharded_device_step_validates_and_orders_streams(ctx):
    """step_device refuses tensors the C ABI would misread (dtype, shape, contiguity,
    device, alignment of the records) and orders the library's stream against torch's current
    stream in both directions: a producer and a consumer on torch's stream need no host sync."""
    import torch
    import particular_b200 as pb
    from particular_b200.sharded import _check_tensor
    good = torch.from_numpy(uniform_cloud(4096, seed=2)).cuda()
    for bad, exc in ((good.double(), TypeError), (good[:, :3], ValueError), (good.t().contiguous().t(), ValueError),
                     (good.reshape(-1)[1:-3].reshape(-1, 4), ValueError)):
        with pytest.raises(exc):
            _check_tensor(bad, 4, "local", ctx.device)
    sh = pb.ShardedBruteForce(ctx, pb.AccelerationSoftened.checked(1.0), init_comm=False)
    with pytest.raises(TypeError):
        sh.step_device(good.double(), 4096)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        src = torch.zeros_like(good)
        for _ in range(50):                      # a long producer chain on torch's stream
            src = src * 0.5 + good * 0.5
        src = src * 0 + good
        out = sh.step_device(src, 4096)
        total = out.abs().sum()                  # consumer on torch's stream, no host synchronisation
    side.synchronize()
    ref = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(1.0)).compute(good.cpu().numpy())
    assert np.isclose(float(total), np.abs(ref).sum(), rtol=1e-4)
