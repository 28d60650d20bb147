// C++ host-API test: the reference's own unit tests for this path, written against
// particular_cuda.hpp with a user-defined particle type.
//   acceleration_error!  particular/src/gravity/newtonian/mod.rs:228-277 (tolerances :385-418)
//   circular_orbit!      particular/src/gravity/newtonian/mod.rs:281-347
// Exit code 0 = all passed, 77 = no usable GPU (pcuda_create failed), 1 = a check failed.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <vector>

#include "particular_cuda.hpp"

using namespace particular;

// Layout of every struct that crosses the C ABI by value or by pointer: the numbers a binding in
// another language (rust/particular-cuda/src/ffi.rs, particular_b200/_ffi.py) relies on.
static_assert(sizeof(pcuda_config) == 16 && offsetof(pcuda_config, device) == 0 &&
                  offsetof(pcuda_config, flags) == 4 && offsetof(pcuda_config, leaf_size) == 8 &&
                  offsetof(pcuda_config, expansion_order) == 12,
              "pcuda_config layout");
static_assert(sizeof(pcuda_timings) == 28 && offsetof(pcuda_timings, upload_ms) == 0 &&
                  offsetof(pcuda_timings, comm_ms) == 4 && offsetof(pcuda_timings, build_ms) == 8 &&
                  offsetof(pcuda_timings, compute_ms) == 12 && offsetof(pcuda_timings, download_ms) == 16 &&
                  offsetof(pcuda_timings, kernel_launches) == 20 && offsetof(pcuda_timings, reserved) == 24,
              "pcuda_timings layout");
static_assert(sizeof(pcuda_tree_info) == 56 && offsetof(pcuda_tree_info, n_particles) == 0 &&
                  offsetof(pcuda_tree_info, n_nodes) == 8 && offsetof(pcuda_tree_info, n_levels) == 16 &&
                  offsetof(pcuda_tree_info, leaf_size) == 20 && offsetof(pcuda_tree_info, dim) == 24 &&
                  offsetof(pcuda_tree_info, bits) == 28 && offsetof(pcuda_tree_info, origin) == 32 &&
                  offsetof(pcuda_tree_info, extent) == 44 && offsetof(pcuda_tree_info, inv) == 48 &&
                  offsetof(pcuda_tree_info, reserved) == 52,
              "pcuda_tree_info layout");
static_assert(sizeof(pcuda_sim_config) == 48 && offsetof(pcuda_sim_config, dim) == 0 &&
                  offsetof(pcuda_sim_config, scalar) == 4 && offsetof(pcuda_sim_config, algorithm) == 8 &&
                  offsetof(pcuda_sim_config, flags) == 12 && offsetof(pcuda_sim_config, theta) == 16 &&
                  offsetof(pcuda_sim_config, softening) == 24 && offsetof(pcuda_sim_config, dt) == 32 &&
                  offsetof(pcuda_sim_config, checked) == 40 && offsetof(pcuda_sim_config, reserved) == 44,
              "pcuda_sim_config layout");
static_assert(sizeof(pcuda_sim_info_t) == 56 && offsetof(pcuda_sim_info_t, n_particles) == 0 &&
                  offsetof(pcuda_sim_info_t, n_affecting) == 8 && offsetof(pcuda_sim_info_t, steps_done) == 16 &&
                  offsetof(pcuda_sim_info_t, d_particles) == 24 && offsetof(pcuda_sim_info_t, d_velocities) == 32 &&
                  offsetof(pcuda_sim_info_t, d_accelerations) == 40 && offsetof(pcuda_sim_info_t, graph_active) == 48 &&
                  offsetof(pcuda_sim_info_t, launches_per_step) == 52,
              "pcuda_sim_info_t layout");
static_assert(PCUDA_UNIQUE_ID_BYTES == 128, "ncclUniqueId size");

// What `#[derive(Position, Mass)] struct Body { position: Vec3, mu: f32 }` gives a Rust user.
template <class S, std::size_t D>
struct Body {
    std::array<S, D> pos;
    S mu_;
    const std::array<S, D> &position() const { return pos; }
    S mu() const { return mu_; }
};

static int failures = 0;
#define CHECK(cond, ...)                                      \
    do {                                                      \
        if (!(cond)) {                                        \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__);  \
            std::printf(__VA_ARGS__);                         \
            std::printf("\n");                                \
            ++failures;                                       \
        }                                                     \
    } while (0)

template <class S, std::size_t D>
static std::array<S, D> splat(S v) {
    std::array<S, D> a;
    a.fill(v);
    return a;
}

template <class S, std::size_t D, class Algo>
static void acceleration_error(Algo &&algo, double epsilon, const char *name) {
    using B = Body<S, D>;
    std::vector<B> massive = {{splat<S, D>(0), 20}, {splat<S, D>(1), 30}, {splat<S, D>(-3), 40}};
    std::vector<B> particles = {{splat<S, D>(10), 0}, massive[0], massive[1], massive[2],
                                {splat<S, D>(30), 0}, {splat<S, D>(-45), 0}};
    Reordered<B> reordered(particles, [](const B &b) { return b.mu() != S(0); });
    auto computed = algo.compute(reordered);
    CHECK(computed.size() == particles.size(), "%s: %zu outputs", name, computed.size());
    for (std::size_t i = 0; i < particles.size(); ++i) {
        double acc[D] = {};
        for (const B &m : massive) {
            double dir[D], mag2 = 0;
            for (std::size_t k = 0; k < D; ++k) {
                dir[k] = (double)m.pos[k] - (double)particles[i].pos[k];
                mag2 += dir[k] * dir[k];
            }
            if (mag2 != 0)
                for (std::size_t k = 0; k < D; ++k) acc[k] += dir[k] * m.mu() * std::sqrt(1.0 / mag2) / mag2;
        }
        double err = 0;
        for (std::size_t k = 0; k < D; ++k) err += std::pow(1.0 - computed[i][k] / acc[k], 2);
        err = std::sqrt(err);
        CHECK(err <= epsilon, "%s: particle %zu error %.3e > %.1e", name, i, err, epsilon);
    }
}

template <class Algo>
static void circular_orbit(Algo &&algo, int orbits, double epsilon, const char *name) {
    using B = Body<float, 3>;
    const float DT = 1.0f / 60.0f;
    std::vector<B> particles = {{{0, 0, 0}, 1e6f}, {{100, 0, 0}, 0.f}};
    std::array<float, 3> vel[2] = {{0, 0, 0}, {0, 100, 0}};
    auto dist = [&] {
        double d2 = 0;
        for (int k = 0; k < 3; ++k) d2 += std::pow((double)particles[0].pos[k] - particles[1].pos[k], 2);
        return std::sqrt(d2);
    };
    const double before = dist();
    const int steps = (int)std::lround(2 * M_PI * std::sqrt(before * before * before / 1e6) / DT);
    for (int s = 0; s < steps * orbits; ++s) {
        auto acc = algo.compute(particles);
        for (int i = 0; i < 2; ++i)
            for (int k = 0; k < 3; ++k) {
                vel[i][k] += acc[i][k] * DT;
                particles[i].pos[k] += vel[i][k] * DT;
            }
    }
    const double after = dist();
    CHECK(std::fabs(1.0 - before / after) < epsilon, "%s: distance drift %.3e", name, std::fabs(1.0 - before / after));
    const double e0 = -1e6 / (2 * before), e1 = -1e6 / (2 * after);
    CHECK(std::fabs(1.0 - e0 / e1) < epsilon, "%s: energy drift %.3e", name, std::fabs(1.0 - e0 / e1));
}

int main() {
    try {
        cuda::CudaContext ctx(0);
        // tests_algorithms! (gravity/newtonian/mod.rs:351-423), CUDA flavour
        acceleration_error<float, 3>(cuda::BruteForce(ctx, Acceleration<true>{}), 1e-2, "brute_force f32x3");
        acceleration_error<float, 2>(cuda::BruteForce(ctx, Acceleration<true>{}), 1e-2, "brute_force f32x2");
        acceleration_error<double, 3>(cuda::BruteForce(ctx, Acceleration<true>{}), 1e-2, "brute_force f64x3");
        acceleration_error<float, 3>(cuda::BruteForce(ctx, AccelerationSoftened<true>(0.0)), 1e-2, "softened(0)");
        acceleration_error<float, 3>(cuda::BarnesHut(ctx, 0.0, Acceleration<true>{}), 1e-2, "barnes_hut f32x3");
        acceleration_error<float, 3>(cuda::BarnesHut(ctx, 0.5, Acceleration<true>{}), 5e-1, "barnes_hut_05 f32x3");
        acceleration_error<float, 2>(cuda::BarnesHut(ctx, 0.0, Acceleration<true>{}), 1e-2, "barnes_hut f32x2");
        acceleration_error<float, 2>(cuda::BarnesHut(ctx, 0.5, Acceleration<true>{}), 5e-1, "barnes_hut_05 f32x2");
        circular_orbit(cuda::BruteForce(ctx, Acceleration<true>{}), 5, 1e-2, "orbit brute_force");
        circular_orbit(cuda::BarnesHut(ctx, 0.5, Acceleration<true>{}), 2, 1e-1, "orbit barnes_hut_05");

        // storages: &[P], Between, Ordered give consistent answers (storage.rs:207-241)
        using B = Body<float, 3>;
        std::vector<B> ps;
        unsigned state = 1808;
        auto rnd = [&] { state = state * 1664525u + 1013904223u; return (state >> 8) / 16777216.0f; };
        for (int i = 0; i < 1500; ++i) ps.push_back({{rnd() * 100, rnd() * 100, rnd() * 100}, i % 3 ? rnd() * 1e6f : 0.f});
        cuda::BruteForce bf(ctx, Acceleration<true>{});
        auto a_slice = bf.compute(ps);
        auto a_between = bf.compute(Between<const std::vector<B> &, const std::vector<B> &>{ps, ps});
        auto a_reordered = bf.compute(Reordered<B>(ps, [](const B &b) { return b.mu() != 0.f; }));
        CHECK(a_slice.size() == 1500 && a_between.size() == 1500 && a_reordered.size() == 1500, "sizes");
        double worst = 0;
        for (int i = 0; i < 1500; ++i)
            for (int k = 0; k < 3; ++k) {
                worst = std::max(worst, (double)std::fabs(a_slice[i][k] - a_between[i][k]));
                const double den = std::fabs(a_slice[i][k]) + 1e-3;
                worst = std::max(worst, std::fabs(a_slice[i][k] - a_reordered[i][k]) / den * 1e-3);
            }
        CHECK(worst < 1e-3, "storages disagree: %.3e", worst);
        auto ordered = Ordered<B>::create(ps, [](const B &b) { return b.mu() != 0.f; });
        CHECK(ordered.affecting_len() == 1000, "affecting_len %zu", ordered.affecting_len());
        CHECK(bf.compute(ordered).size() == 1500, "ordered size");
        cuda::RootedOrthtree tree(ctx, ps);
        auto a_tree = cuda::BarnesHut(ctx, 0.5, Acceleration<true>{})
                          .compute(Between<const std::vector<B> &, const cuda::RootedOrthtree &>{ps, tree});
        CHECK(a_tree.size() == 1500 && tree.info().n_particles == 1500, "tree traversal");
        // circular_orbit! run on the device (cuda::Simulation): 20 orbits of 377 steps, and a
        // Reordered-style simulation in which the massless particles do not affect anything
        {
            std::vector<B> two = {{{0, 0, 0}, 1e6f}, {{100, 0, 0}, 0.f}};
            std::vector<std::array<float, 3>> vel = {{0, 0, 0}, {0, 100, 0}};
            cuda::Simulation<float, 3> sim(ctx, bf, Acceleration<true>{}, two, vel, 1.0 / 60.0,
                                           cuda::Affecting::Massive);
            sim.step(377 * 20);
            auto after = sim.particles();
            double d2 = 0;
            for (int k = 0; k < 3; ++k) d2 += std::pow((double)after[0].position_[k] - after[1].position_[k], 2);
            CHECK(std::fabs(1.0 - 100.0 / std::sqrt(d2)) < 1e-2, "device orbit drift %.3e",
                  std::fabs(1.0 - 100.0 / std::sqrt(d2)));
            CHECK(after[0].position_[0] == 0.f && after[0].position_[1] == 0.f,
                  "the massless satellite must not move the primary");
            CHECK(sim.info().steps_done == 377 * 20 && sim.info().n_affecting == 1, "sim info");
            cuda::BarnesHut bh(ctx, 0.5, AccelerationSoftened<true>(1.0));
            cuda::Simulation<float, 3> sim_bh(ctx, bh, 0.5, AccelerationSoftened<true>(1.0), ps, {}, 1e-3);
            sim_bh.step(2);
            CHECK(sim_bh.accelerations().size() == 1500 && sim_bh.velocities().size() == 1500, "bh sim sizes");
        }
        // a user-defined interaction (InteractionShader counterpart): neighbour count within a radius
        {
            struct Pos { float x, y, z; };
            struct Count { uint32_t n; };
            struct Cut { float r2; };
            cuda::CustomInteraction<Pos, Pos, Count, Cut> neighbours(ctx, R"(
                struct Affected { float x, y, z; };
                typedef Affected Affecting;
                struct Interaction { uint32_t n; };
                struct Push { float r2; };
                __device__ void compute(const Affected &a, const Affecting &b, Interaction &out) {
                    const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z;
                    const float d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 != 0.f && d2 <= push.r2) out.n += 1u;
                })");
            std::vector<Pos> pts;
            for (const B &b : ps) pts.push_back({b.pos[0], b.pos[1], b.pos[2]});
            auto counts = neighbours.brute_force(pts, pts, Cut{100.f});
            uint32_t expect0 = 0;
            for (const Pos &q : pts) {
                const float dx = q.x - pts[0].x, dy = q.y - pts[0].y, dz = q.z - pts[0].z;
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 != 0.f && d2 <= 100.f) ++expect0;
            }
            CHECK(counts.size() == pts.size() && counts[0].n == expect0, "custom interaction: %u vs %u",
                  counts.empty() ? 0u : counts[0].n, expect0);
        }
        // empty input: CPU-path semantics (the wgpu path panics, gpu/resources.rs:24)
        std::vector<B> none;
        CHECK(bf.compute(none).empty(), "empty slice");
        CHECK(ctx.timings().kernel_launches == 0, "no launch for empty input");
    } catch (const cuda::Error &e) {
        std::printf("cuda::Error: %s\n", e.what());
        return e.status == PCUDA_ERR_NO_DEVICE ? 77 : 1;
    }
    std::printf(failures ? "%d check(s) failed\n" : "all host-API checks passed\n", failures);
    return failures ? 1 : 0;
}
