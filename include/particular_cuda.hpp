// particular_cuda.hpp — header-only C++17 host API above the C ABI (particular_cuda.h).
//
// The reference is a Rust crate and this build environment has no Rust toolchain, so the host side
// of the drop-in is written in C++ and mirrors the reference's operator interface for the hot path
// name by name (paths relative to /root/reference/particular/src):
//
//   Between<S1, S2>                          lib.rs:299-300
//   Interaction::compute(storage)            lib.rs:364-370      -> Algorithm::compute(storage)
//   Position / Mass accessors                gravity/mod.rs:40-58, 133-148  -> p.position(), p.mu()
//   GravitationalField<V, S>                 gravity/mod.rs:12-35
//   Acceleration<CHECKED>                    gravity/newtonian/acceleration.rs:17-58
//   AccelerationSoftened<S, CHECKED>         gravity/newtonian/acceleration_softened.rs:17-63
//   Ordered<P>, Reordered<P, F>, &[P]        storage.rs:48-241 (same Between mapping)
//   gpu::BruteForce<'a, T>                   gpu/mod.rs:149-208  -> cuda::BruteForce<T>
//   sequential::BarnesHut<S, T>              sequential.rs:439-543 -> cuda::BarnesHut<T>
//   RootedOrthtree                           storage.rs:11-46    -> cuda::RootedOrthtree
//   GpuResources + wgpu::Device + Queue      gpu/mod.rs:85-159   -> cuda::CudaContext
//   the caller's integration loop            examples/simple/src/main.rs:45-59 -> cuda::Simulation
//
// A user particle type needs `position()` returning something indexable with operator[] for D
// components and `mu()` returning the gravitational parameter — what `#[derive(Position, Mass)]`
// generates in the reference (particular_derive/src/gravity.rs:1-108).  Results come back as a
// std::vector of std::array<S, D> in affected order (the reference returns a Vec's IntoIter,
// gpu/mod.rs:184).  Errors throw cuda::Error (the reference panics: gpu/resources.rs:24, 38-39).
#ifndef PARTICULAR_CUDA_HPP
#define PARTICULAR_CUDA_HPP

#include <algorithm>
#include <array>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <cstdint>
#include <type_traits>
#include <utility>
#include <vector>

#include "particular_cuda.h"

namespace particular {

// ---- core vocabulary -------------------------------------------------------------------------------
template <class S1, class S2>
struct Between {  // lib.rs:299-300: the first is acted upon, the second acts
    S1 affected;
    S2 affecting;
};
template <class S1, class S2>
Between(S1, S2) -> Between<S1, S2>;

template <class S, std::size_t D>
struct GravitationalField {  // gravity/mod.rs:12-18, #[repr(C)] {position, m}
    std::array<S, D> position_;
    S m;
    const std::array<S, D> &position() const { return position_; }
    S mu() const { return m; }
    bool is_affecting() const { return m != S(0); }  // gravity/mod.rs:29-34
};

template <bool CHECKED = true>
struct Acceleration {  // acceleration.rs:17-58; softening = Default (0)
    static constexpr bool checked = CHECKED;
    double softening() const { return 0.0; }
};
template <bool CHECKED = true>
struct AccelerationSoftened {  // acceleration_softened.rs:17-63
    static constexpr bool checked = CHECKED;
    double eps = 0.0;
    explicit AccelerationSoftened(double softening) : eps(softening) {}
    double softening() const { return eps; }
};

// ---- storages (storage.rs) ---------------------------------------------------------------------------
template <class P>
class Ordered {  // storage.rs:48-138: affecting particles first
public:
    template <class F>
    static Ordered create(const std::vector<P> &unordered, F is_affecting) {  // Ordered::new :85-95
        Ordered o;
        for (const P &p : unordered)
            if (is_affecting(p)) o.particles_.push_back(p);
        for (const P &p : unordered)
            if (!is_affecting(p)) o.particles_.push_back(p);
        // Ordered::with :71-74: affecting_len = first index failing the predicate
        std::size_t k = 0;
        while (k < o.particles_.size() && is_affecting(o.particles_[k])) ++k;
        o.affecting_len_ = k;
        return o;
    }
    std::size_t affecting_len() const { return affecting_len_; }
    const std::vector<P> &particles() const { return particles_; }
    const P *affecting() const { return particles_.data(); }
    const P *non_affecting() const { return particles_.data() + affecting_len_; }

private:
    std::vector<P> particles_;
    std::size_t affecting_len_ = 0;
};

template <class P>
class Reordered {  // storage.rs:141-205: borrow of the original slice + an Ordered copy
public:
    template <class F>
    Reordered(const std::vector<P> &unordered, F is_affecting)
        : unordered_(&unordered), ordered_(Ordered<P>::create(unordered, is_affecting)) {}
    const std::vector<P> &unordered() const { return *unordered_; }
    const Ordered<P> &ordered() const { return ordered_; }
    std::size_t affecting_len() const { return ordered_.affecting_len(); }
    const P *affecting() const { return ordered_.affecting(); }

private:
    const std::vector<P> *unordered_;
    Ordered<P> ordered_;
};

namespace cuda {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &msg)
        : std::runtime_error(std::string(pcuda_status_string(s)) + ": " + msg), status(s) {}
};

// Owns one device, one stream, grow-only buffers ("should not be recreated for every iteration",
// gpu/mod.rs:150-151).
class CudaContext {
public:
    // extra_flags: PCUDA_FLAG_BH_PARTITIONED_BUILD / PCUDA_FLAG_BH_REPLICATED_BUILD force the
    // multi-GPU Barnes-Hut tree build (default: partitioned from 4 GPUs on); PCUDA_FLAG_EXACT_CHECKED
    // makes the f32 brute-force kernels test r^2 == 0 exactly at every problem size.
    explicit CudaContext(int device = 0, unsigned leaf_size = 0, bool phase_timings = true,
                         unsigned expansion_order = 1, unsigned extra_flags = 0) {
        pcuda_config cfg{device, (phase_timings ? PCUDA_FLAG_NONE : PCUDA_FLAG_NO_PHASE_TIMINGS) | extra_flags,
                         leaf_size, expansion_order};
        int s = pcuda_create(&cfg, &ctx_);
        if (s != PCUDA_OK) throw Error(s, pcuda_last_error(nullptr));
    }
    ~CudaContext() { pcuda_destroy(ctx_); }
    CudaContext(const CudaContext &) = delete;
    CudaContext &operator=(const CudaContext &) = delete;
    pcuda_ctx *handle() const { return ctx_; }
    void check(int s) const {
        if (s != PCUDA_OK) throw Error(s, pcuda_last_error(ctx_));
    }
    pcuda_timings timings() const {
        pcuda_timings t{};
        pcuda_get_timings(ctx_, &t);
        return t;
    }

private:
    pcuda_ctx *ctx_ = nullptr;
};

namespace detail {

template <class P>
using position_t = std::decay_t<decltype(std::declval<const P &>().position())>;
template <class P>
using scalar_t = std::decay_t<decltype(std::declval<const P &>().mu())>;

template <class V>
constexpr std::size_t dim_of() {
    if constexpr (std::is_array_v<V>) return std::extent_v<V>;
    else return std::tuple_size<V>::value;
}

// pack_affecting / pack_affected play the role of InteractionShader::write_affecting /
// write_affected (gpu/mod.rs:60-64; acceleration.rs:166-183): GravitationalField::from(&p)
// (gravity/mod.rs:150-161: m = p.mu()) into the wire layout of particular_cuda.h.
template <class S, std::size_t D, class P>
std::vector<S> pack_affecting(const P *p, std::size_t n) {
    std::vector<S> out(n * (D + 1));
    for (std::size_t i = 0; i < n; ++i) {
        const auto &pos = p[i].position();
        for (std::size_t k = 0; k < D; ++k) out[i * (D + 1) + k] = static_cast<S>(pos[k]);
        out[i * (D + 1) + D] = static_cast<S>(p[i].mu());
    }
    return out;
}
template <class S, std::size_t D, class P>
std::vector<S> pack_affected(const P *p, std::size_t n) {
    std::vector<S> out(n * D);
    for (std::size_t i = 0; i < n; ++i) {
        const auto &pos = p[i].position();
        for (std::size_t k = 0; k < D; ++k) out[i * D + k] = static_cast<S>(pos[k]);
    }
    return out;
}
template <class S, std::size_t D>
std::vector<std::array<S, D>> unpack(const std::vector<S> &flat, std::size_t n) {
    std::vector<std::array<S, D>> out(n);
    for (std::size_t i = 0; i < n; ++i)
        for (std::size_t k = 0; k < D; ++k) out[i][k] = flat[i * D + k];
    return out;
}

template <class S, std::size_t D>
struct Kernels;  // which C entry points serve (scalar, dimension); others: unimplemented, as the
                 // reference's shaders are for D outside {2, 3} (gravity/impls/mod.rs:362, 374)
template <>
struct Kernels<float, 3> {
    static constexpr auto brute = pcuda_bruteforce_f32x3;
    static constexpr auto barnes = pcuda_barneshut_f32x3;
};
template <>
struct Kernels<float, 2> {
    static constexpr auto brute = pcuda_bruteforce_f32x2;
    static constexpr auto barnes = pcuda_barneshut_f32x2;
};
template <>
struct Kernels<double, 3> {
    static constexpr auto brute = pcuda_bruteforce_f64x3;
    static constexpr auto barnes = pcuda_barneshut_f64x3;
};
template <>
struct Kernels<double, 2> {
    static constexpr auto brute = pcuda_bruteforce_f64x2;
    static constexpr auto barnes = pcuda_barneshut_f64x2;
};

}  // namespace detail

// A tree built on the device over the affecting particles (storage.rs:11-46).  f32.
class RootedOrthtree {
public:
    template <class P>
    RootedOrthtree(CudaContext &ctx, const std::vector<P> &affecting) : ctx_(&ctx) {
        using V = detail::position_t<P>;
        constexpr std::size_t D = detail::dim_of<V>();
        static_assert(std::is_same_v<detail::scalar_t<P>, float>, "device trees are f32");
        auto src = detail::pack_affecting<float, D>(affecting.data(), affecting.size());
        ctx.check(pcuda_tree_build_f32(ctx.handle(), (uint32_t)D, src.data(), affecting.size(), &tree_));
        dim_ = D;
    }
    ~RootedOrthtree() { pcuda_tree_destroy(ctx_->handle(), tree_); }
    RootedOrthtree(const RootedOrthtree &) = delete;
    RootedOrthtree &operator=(const RootedOrthtree &) = delete;
    pcuda_tree *handle() const { return tree_; }
    std::size_t dim() const { return dim_; }
    pcuda_tree_info info() const {
        pcuda_tree_info i{};
        pcuda_tree_info_get(tree_, &i);
        return i;
    }

private:
    CudaContext *ctx_;
    pcuda_tree *tree_ = nullptr;
    std::size_t dim_ = 3;
};

// Brute force on the GPU: the CUDA counterpart of gpu::BruteForce (gpu/mod.rs:149-208).
template <class T>
class BruteForce {
public:
    BruteForce(CudaContext &ctx, T interaction) : ctx_(&ctx), interaction_(std::move(interaction)) {}

    // Interaction<Between<&[P1], &[P2]>> (gpu/mod.rs:179-208)
    template <class P1, class P2>
    auto compute(const Between<const std::vector<P1> &, const std::vector<P2> &> &b) {
        return run(b.affected.data(), b.affected.size(), b.affecting.data(), b.affecting.size(), false);
    }
    // &[P] => Between(slice, slice) (storage.rs:231-241); the targets alias the sources on the device
    template <class P>
    auto compute(const std::vector<P> &slice) {
        return run(slice.data(), slice.size(), slice.data(), slice.size(), true);
    }
    // &Ordered => Between(all ordered, affecting prefix) (storage.rs:207-217)
    template <class P>
    auto compute(const Ordered<P> &o) {
        return run(o.particles().data(), o.particles().size(), o.affecting(), o.affecting_len(), false);
    }
    // &Reordered => Between(unordered original, affecting copy) (storage.rs:219-229)
    template <class P>
    auto compute(const Reordered<P> &r) {
        return run(r.unordered().data(), r.unordered().size(), r.affecting(), r.affecting_len(), false);
    }

private:
    template <class P1, class P2>
    auto run(const P1 *aff, std::size_t na, const P2 *src, std::size_t nb, bool alias) {
        using S = detail::scalar_t<P2>;
        constexpr std::size_t D = detail::dim_of<detail::position_t<P2>>();
        auto s = detail::pack_affecting<S, D>(src, nb);
        std::vector<S> a;
        if (!alias) a = detail::pack_affected<S, D>(aff, na);
        std::vector<S> out(na * D);
        ctx_->check(detail::Kernels<S, D>::brute(ctx_->handle(), alias ? nullptr : a.data(), na, s.data(),
                                                 nb, static_cast<S>(interaction_.softening()),
                                                 T::checked ? 1 : 0, out.data()));
        return detail::unpack<S, D>(out, na);
    }
    CudaContext *ctx_;
    T interaction_;
};

// Barnes-Hut on the GPU (sequential.rs:439-543 semantics: the tree is rebuilt on every call unless
// the storage is Between(affected, RootedOrthtree), :508-524).
template <class T>
class BarnesHut {
public:
    BarnesHut(CudaContext &ctx, double theta, T interaction)
        : ctx_(&ctx), theta_(theta), interaction_(std::move(interaction)) {}

    template <class P1, class P2>
    auto compute(const Between<const std::vector<P1> &, const std::vector<P2> &> &b) {
        return run(b.affected.data(), b.affected.size(), b.affecting.data(), b.affecting.size(), false);
    }
    template <class P>
    auto compute(const std::vector<P> &slice) {
        return run(slice.data(), slice.size(), slice.data(), slice.size(), true);
    }
    template <class P>
    auto compute(const Ordered<P> &o) {
        return run(o.particles().data(), o.particles().size(), o.affecting(), o.affecting_len(), false);
    }
    template <class P>
    auto compute(const Reordered<P> &r) {
        return run(r.unordered().data(), r.unordered().size(), r.affecting(), r.affecting_len(), false);
    }
    // Between<&[P1], &RootedOrthtree> (sequential.rs:508-524)
    template <class P1>
    auto compute(const Between<const std::vector<P1> &, const RootedOrthtree &> &b) {
        constexpr std::size_t D = detail::dim_of<detail::position_t<P1>>();
        auto a = detail::pack_affected<float, D>(b.affected.data(), b.affected.size());
        std::vector<float> out(b.affected.size() * D);
        ctx_->check(pcuda_tree_traverse_f32(ctx_->handle(), b.affecting.handle(), a.data(),
                                            b.affected.size(), (float)theta_,
                                            (float)interaction_.softening(), T::checked ? 1 : 0,
                                            out.data()));
        return detail::unpack<float, D>(out, b.affected.size());
    }

private:
    template <class P1, class P2>
    auto run(const P1 *aff, std::size_t na, const P2 *src, std::size_t nb, bool alias) {
        using S = detail::scalar_t<P2>;
        constexpr std::size_t D = detail::dim_of<detail::position_t<P2>>();
        auto s = detail::pack_affecting<S, D>(src, nb);
        std::vector<S> a;
        if (!alias) a = detail::pack_affected<S, D>(aff, na);
        std::vector<S> out(na * D);
        ctx_->check(detail::Kernels<S, D>::barnes(ctx_->handle(), alias ? nullptr : a.data(), na, s.data(),
                                                  nb, (S)theta_, (S)interaction_.softening(),
                                                  T::checked ? 1 : 0, out.data()));
        return detail::unpack<S, D>(out, na);
    }
    CudaContext *ctx_;
    double theta_;
    T interaction_;
};

// Device-resident stepping (particular_cuda.h "device-resident stepping"): the loop every caller of
// the reference writes around compute() — accelerations, then velocity += a * dt and
// position += velocity * dt (examples/simple/src/main.rs:45-59) — with the particles kept on the
// device between steps.  `Algorithm` is cuda::BruteForce<T> or cuda::BarnesHut<T>.
enum class Affecting { All, Massive };  // Massive == the Reordered storage (storage.rs:153-163)

template <class S, std::size_t D>
class Simulation {
public:
    using Vector = std::array<S, D>;

    template <class T, class P>
    Simulation(CudaContext &ctx, const BruteForce<T> &, T interaction, const std::vector<P> &particles,
               const std::vector<Vector> &velocities, double dt, Affecting affecting = Affecting::All)
        : ctx_(&ctx) {
        create(PCUDA_BRUTE_FORCE, 0.0, interaction.softening(), T::checked, particles, velocities, dt,
               affecting);
    }
    template <class T, class P>
    Simulation(CudaContext &ctx, const BarnesHut<T> &, double theta, T interaction,
               const std::vector<P> &particles, const std::vector<Vector> &velocities, double dt,
               Affecting affecting = Affecting::All)
        : ctx_(&ctx) {
        create(PCUDA_BARNES_HUT, theta, interaction.softening(), T::checked, particles, velocities, dt,
               affecting);
    }
    ~Simulation() { pcuda_sim_destroy(ctx_->handle(), sim_); }
    Simulation(const Simulation &) = delete;
    Simulation &operator=(const Simulation &) = delete;

    void step(unsigned n_steps = 1) { ctx_->check(pcuda_sim_step(ctx_->handle(), sim_, n_steps)); }
    std::vector<GravitationalField<S, D>> particles() {
        std::vector<S> flat(n_ * (D + 1));
        ctx_->check(pcuda_sim_read(ctx_->handle(), sim_, flat.data(), nullptr, nullptr));
        std::vector<GravitationalField<S, D>> out(n_);
        for (std::size_t i = 0; i < n_; ++i) {
            for (std::size_t k = 0; k < D; ++k) out[i].position_[k] = flat[i * (D + 1) + k];
            out[i].m = flat[i * (D + 1) + D];
        }
        return out;
    }
    std::vector<Vector> velocities() {
        std::vector<S> flat(n_ * D);
        ctx_->check(pcuda_sim_read(ctx_->handle(), sim_, nullptr, flat.data(), nullptr));
        return detail::unpack<S, D>(flat, n_);
    }
    std::vector<Vector> accelerations() {
        std::vector<S> flat(n_ * D);
        ctx_->check(pcuda_sim_read(ctx_->handle(), sim_, nullptr, nullptr, flat.data()));
        return detail::unpack<S, D>(flat, n_);
    }
    pcuda_sim_info_t info() const {
        pcuda_sim_info_t i{};
        pcuda_sim_info(sim_, &i);
        return i;
    }

private:
    template <class P>
    void create(uint32_t algorithm, double theta, double softening, bool checked,
                const std::vector<P> &particles, const std::vector<Vector> &velocities, double dt,
                Affecting affecting) {
        static_assert(std::is_same_v<S, float> || std::is_same_v<S, double>, "f32 or f64");
        n_ = particles.size();
        auto flat = detail::pack_affecting<S, D>(particles.data(), n_);
        std::vector<S> vel;
        if (!velocities.empty()) {
            if (velocities.size() != n_) throw Error(PCUDA_ERR_INVALID_ARGUMENT, "one velocity per particle");
            vel.resize(n_ * D);
            for (std::size_t i = 0; i < n_; ++i)
                for (std::size_t k = 0; k < D; ++k) vel[i * D + k] = velocities[i][k];
        }
        pcuda_sim_config cfg{};
        cfg.dim = (uint32_t)D;
        cfg.scalar = std::is_same_v<S, double> ? PCUDA_F64 : PCUDA_F32;
        cfg.algorithm = algorithm;
        cfg.flags = affecting == Affecting::Massive ? PCUDA_SIM_AFFECTING_MASSIVE_ONLY : 0u;
        cfg.theta = theta;
        cfg.softening = softening;
        cfg.dt = dt;
        cfg.checked = checked ? 1 : 0;
        ctx_->check(pcuda_sim_create(ctx_->handle(), &cfg, flat.data(), vel.empty() ? nullptr : vel.data(),
                                     n_, &sim_));
    }
    CudaContext *ctx_;
    pcuda_sim *sim_ = nullptr;
    std::size_t n_ = 0;
};

// A user-defined pair interaction compiled for the device at run time: the counterpart of
// implementing InteractionShader<P1, P2> for the wgpu operator (gpu/mod.rs:40-82).  `source` is CUDA
// C++ defining struct Affected / Affecting / Interaction / Push and
// `__device__ void compute(const Affected &, const Affecting &, Interaction &)`; the host-side
// template arguments must be trivially copyable types with the same layouts (sizes are checked
// against the device compiler's sizeof, the reference's AFFECTED_SIZE / AFFECTING_SIZE /
// INTERACTION_SIZE).
struct NoPush {};
template <class Affected, class Affecting, class Interaction, class Push = NoPush>
class CustomInteraction {
public:
    CustomInteraction(CudaContext &ctx, const std::string &source) : ctx_(&ctx) {
        static_assert(std::is_trivially_copyable_v<Affected> && std::is_trivially_copyable_v<Affecting> &&
                      std::is_trivially_copyable_v<Interaction> && std::is_trivially_copyable_v<Push>);
        ctx.check(pcuda_interaction_create(ctx.handle(), source.c_str(), &it_));
        uint32_t sz[4];
        pcuda_interaction_sizes(it_, sz);
        if (sz[0] != sizeof(Affected) || sz[1] != sizeof(Affecting) || sz[2] != sizeof(Interaction) ||
            (!std::is_same_v<Push, NoPush> && sz[3] < sizeof(Push))) {
            pcuda_interaction_destroy(ctx.handle(), it_);
            throw Error(PCUDA_ERR_INVALID_ARGUMENT, "host types do not match the device structs");
        }
    }
    ~CustomInteraction() { pcuda_interaction_destroy(ctx_->handle(), it_); }
    CustomInteraction(const CustomInteraction &) = delete;
    CustomInteraction &operator=(const CustomInteraction &) = delete;

    // Interaction<Between<&[P1], &[P2]>> with the brute-force algorithm
    std::vector<Interaction> brute_force(const std::vector<Affected> &affected,
                                         const std::vector<Affecting> &affecting, const Push &push = Push{}) {
        std::vector<Interaction> out(affected.size());
        const bool has_push = !std::is_same_v<Push, NoPush>;
        ctx_->check(pcuda_interaction_brute_force(ctx_->handle(), it_, affected.data(), affected.size(),
                                                  affecting.data(), affecting.size(), has_push ? &push : nullptr,
                                                  has_push ? sizeof(Push) : 0, out.data()));
        return out;
    }

private:
    CudaContext *ctx_;
    pcuda_interaction *it_ = nullptr;
};

// Extension-trait sugar (GpuCompute, gpu/mod.rs:13-37).
template <class Storage, class T>
auto cuda_brute_force(const Storage &storage, CudaContext &ctx, T interaction) {
    return BruteForce<T>(ctx, std::move(interaction)).compute(storage);
}
template <class Storage, class T>
auto cuda_barnes_hut(const Storage &storage, CudaContext &ctx, double theta, T interaction) {
    return BarnesHut<T>(ctx, theta, std::move(interaction)).compute(storage);
}

}  // namespace cuda
}  // namespace particular

#endif  // PARTICULAR_CUDA_HPP
