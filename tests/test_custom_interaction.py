"""User-defined interactions compiled at run time (pcuda_interaction_*): the
CUDA counterpart of implementing InteractionShader for the reference's wgpu operator
(gpu/mod.rs:40-82).  CPU part: the sources compile with NVRTC (no device needed).  GPU part: the
reference's own examples of a non-acceleration interaction evaluated on the device against numpy /
the oracle."""
import numpy as np
import pytest

import oracle
from tests.conftest import rel_err, uniform_cloud

# The reference's Acceleration<CHECKED> / AccelerationSoftened pair term, as a user would write it
# (gravity/impls/mod.rs:151-166): d = p2 - p1; n = |d|^2; n == 0 -> nothing; else
# ns = n + eps^2; out += d * (mu / (ns * sqrt(ns))).
ACCELERATION_SRC = """
struct Affected { float x, y, z; };
struct Affecting { float x, y, z, mu; };
struct Interaction { float ax, ay, az; };
struct Push { float softening; };
__device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out) {
    const float dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
    const float n = dx * dx + dy * dy + dz * dz;
    if (n != 0.f) {
        const float ns = n + push.softening * push.softening;
        const float s = p2.mu / (ns * sqrtf(ns));
        out.ax += dx * s; out.ay += dy * s; out.az += dz * s;
    }
}
"""

# The doc example of the crate (lib.rs:205-261): a user-defined GravitationalForce interaction
# between bodies {position, mass}: F = G m1 m2 d / |d|^3, zero for the body itself.
FORCE_SRC = """
struct Body { double x, y, z, mass; };
typedef Body Affected;
typedef Body Affecting;
struct Interaction { double fx, fy, fz; };
struct Push { double G; };
__device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out) {
    const double dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
    const double n = dx * dx + dy * dy + dz * dz;
    if (n == 0.0) return;
    const double s = push.G * p1.mass * p2.mass / (n * sqrt(n));
    out.fx += dx * s; out.fy += dy * s; out.fz += dz * s;
}
"""

# Not gravity at all: neighbours within a cut-off radius and a Lennard-Jones-like energy.
NEIGHBOUR_SRC = """
struct Affected { float x, y; };
struct Affecting { float x, y; };
struct Interaction { uint32_t count; float energy; };
struct Push { float cutoff2; float sigma2; };
__device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out) {
    const float dx = p2.x - p1.x, dy = p2.y - p1.y;
    const float r2 = dx * dx + dy * dy;
    if (r2 == 0.f || r2 > push.cutoff2) return;
    const float q = push.sigma2 / r2, q3 = q * q * q;
    out.count += 1u;
    out.energy += 4.f * (q3 * q3 - q3);
}
"""


@pytest.fixture(scope="module")
def pb():
    import particular_b200 as pb
    return pb


@pytest.mark.parametrize("src", [ACCELERATION_SRC, FORCE_SRC, NEIGHBOUR_SRC])
def test_sources_compile_without_a_device(pb, src):
    assert pb.check_interaction_source(src) == ""


def test_compile_error_carries_the_compiler_log(pb):
    with pytest.raises(pb.CudaError) as e:
        pb.check_interaction_source(ACCELERATION_SRC.replace("sqrtf(ns)", "sqrtf(undefined_name)"))
    assert "undefined_name" in str(e.value) and "interaction.cu(11)" in str(e.value)
    with pytest.raises(pb.CudaError) as e:  # a struct the template needs is missing
        pb.check_interaction_source(ACCELERATION_SRC.replace("struct Push { float softening; };", ""))
    assert "Push" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("softening", [0.0, 25.0])
def test_acceleration_written_as_a_custom_interaction(pb, ctx, softening):
    """Same arithmetic as the reference's scalar pair kernel (IEEE sqrt and division, no
    contraction), same fold order => bit-identical to the oracle's sequential::BruteForce."""
    n = 3001
    p = uniform_cloud(n, seed=17)
    aff_t = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4")])
    src_t = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("mu", "f4")])
    out_t = np.dtype([("ax", "f4"), ("ay", "f4"), ("az", "f4")])
    it = pb.CustomInteraction(ctx, ACCELERATION_SRC, aff_t, src_t, out_t, np.dtype([("softening", "f4")]),
                              push=(softening,))
    assert it.sizes == (12, 16, 12, 4)
    got = pb.BruteForce(ctx, it).compute(pb.Between(p[:, :3].copy().view(aff_t).reshape(-1),
                                                    p.view(src_t).reshape(-1)))
    got = got.view(np.float32).reshape(n, 3)
    ref = oracle.brute_force(p[:, :3], p, softening, True)
    assert np.array_equal(got, ref)
    # and it agrees with the tuned kernel to the stated tolerance
    inter = pb.AccelerationSoftened.checked(softening) if softening else pb.Acceleration.checked()
    tuned = pb.BruteForce(ctx, inter).compute(p)
    assert np.percentile(rel_err(tuned, got), 99) <= 1e-5
    it.close()


@pytest.mark.gpu
def test_doc_example_gravitational_force_f64(pb, ctx):
    """lib.rs:247-261: forces == [se + sj, -se + ej, -sj - ej] for sun / earth / jupiter."""
    G = 6.67430e-11
    bodies = np.array([[0.0, 0.0, 0.0, 1.989e30], [1.496e11, 0.0, 0.0, 5.972e24],
                       [0.0, 7.785e11, 0.0, 1.898e27]])
    body_t = np.dtype([("x", "f8"), ("y", "f8"), ("z", "f8"), ("mass", "f8")])
    out_t = np.dtype([("fx", "f8"), ("fy", "f8"), ("fz", "f8")])
    it = pb.CustomInteraction(ctx, FORCE_SRC, body_t, body_t, out_t, np.dtype([("G", "f8")]), push=(G,))
    forces = pb.BruteForce(ctx, it).compute(bodies.view(body_t).reshape(-1)).view(np.float64).reshape(3, 3)

    def pair(i, j):
        d = bodies[j, :3] - bodies[i, :3]
        n = d @ d
        return d * (G * bodies[i, 3] * bodies[j, 3] / (n * np.sqrt(n)))
    se, sj, ej = pair(0, 1), pair(0, 2), pair(1, 2)
    expect = np.array([se + sj, -se + ej, -sj - ej])
    assert np.allclose(forces, expect, rtol=1e-14, atol=0)
    # Newton's third law: the forces sum to zero to rounding
    assert np.abs(forces.sum(axis=0)).max() <= 1e-10 * np.abs(forces).max()
    it.close()


@pytest.mark.gpu
def test_non_gravity_interaction_with_integer_output(pb, ctx):
    rng = np.random.default_rng(2)
    n_a, n_b = 1500, 5000
    a = rng.uniform(0, 100, (n_a, 2)).astype(np.float32)
    b = rng.uniform(0, 100, (n_b, 2)).astype(np.float32)
    pos_t = np.dtype([("x", "f4"), ("y", "f4")])
    out_t = np.dtype([("count", "u4"), ("energy", "f4")])
    push_t = np.dtype([("cutoff2", "f4"), ("sigma2", "f4")])
    it = pb.CustomInteraction(ctx, NEIGHBOUR_SRC, pos_t, pos_t, out_t, push_t)
    got = it.brute_force(a.view(pos_t).reshape(-1), b.view(pos_t).reshape(-1), push=(9.0, 1.0))
    d = b[None, :, :] - a[:, None, :]
    r2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
    mask = (r2 != 0) & (r2 <= np.float32(9.0))
    assert np.array_equal(got["count"], mask.sum(axis=1).astype(np.uint32))
    q3 = np.where(mask, (1.0 / np.where(mask, r2, 1.0)) ** 3, 0.0)
    energy = (4.0 * (q3 * q3 - q3)).sum(axis=1)
    assert np.allclose(got["energy"], energy, rtol=2e-4, atol=1e-3)
    # empty affecting: Interaction() for everyone; empty affected: nothing
    zero = it.brute_force(a.view(pos_t).reshape(-1), np.zeros(0, pos_t), push=(9.0, 1.0))
    assert not zero["count"].any() and not zero["energy"].any()
    assert len(it.brute_force(np.zeros(0, pos_t), b.view(pos_t).reshape(-1), push=(9.0, 1.0))) == 0
    it.close()


@pytest.mark.gpu
def test_dtype_size_mismatch_is_rejected(pb, ctx):
    bad = np.dtype([("x", "f4"), ("y", "f4")])
    src_t = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("mu", "f4")])
    out_t = np.dtype([("ax", "f4"), ("ay", "f4"), ("az", "f4")])
    with pytest.raises(TypeError):
        pb.CustomInteraction(ctx, ACCELERATION_SRC, bad, src_t, out_t)
    with pytest.raises(pb.CudaError):
        pb.CustomInteraction(ctx, "this is not CUDA", bad, src_t, out_t)
