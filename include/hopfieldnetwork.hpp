// hopfieldnetwork.hpp — header-only C++17 host-side mirror of the reference's Go package
// `hopfieldnetwork` (HopfieldNetworkBuilder.go, HopfieldNetwork.go, LearningRule.go, LearningMethod.go)
// over the C ABI of libhopfield_b200.so.  Same names, argument meaning and error behaviour: the four
// Build() panics become std::invalid_argument with the reference's text; every other failure throws
// std::runtime_error carrying hn_last_error().  No numerics live here.
//
// The reference is compiled Go whose toolchain is absent from the build image; this mirror is what a C++
// host links, and the cgo package (hopfield_network_go_b200/go/hopfieldnetwork) has the same shape.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hopfield_b200.h"

namespace hopfieldnetwork {

using State = std::vector<double>;  // one mat.VecDense

enum DomainEnum { BipolarDomain = 0, BinaryDomain = 1 };                       // domain/DomainEnum.go:6-9
enum LearningRuleEnum {                                                        // LearningRule.go:32-39
    HebbianLearningRule = 0, BipolarMappedHebbianLearningRule, DeltaLearningRule, BipolarMappedDeltaLearningRule,
    ThermalDeltaLearningRule, BipolarMappedThermalDeltaLearningRule };
enum LearningMethodEnum { FullSetMethod = 0, IterativeBatchMethod = 1 };       // LearningMethod.go:22-25
enum NoiseApplicationEnum { None = 0, MaximalInversion, RandomSubMaximalInversion, GaussianApplication };

inline void check(int rc) {
    if (rc != HN_OK) throw std::runtime_error("libhopfield_b200: [" + std::to_string(rc) + "] " + hn_last_error());
}

struct LearnStateData {  // datacollector/LearnStateData.go:14-19
    int Epoch; int TargetStateIndex; std::vector<double> EnergyProfile; bool Stable;
};
struct HopfieldNetworkSummary {  // HopfieldNetwork.go:124-134 (consumed by main.go:262-279)
    std::vector<double> Matrix; int Dimension; bool ForceSymmetric, ForceZeroDiagonal; int Epochs;
    int MaximumRelaxationUnstableUnits, MaximumRelaxationIterations; double LearningNoiseScale; int UnitsUpdatedPerStep;
};
struct RelaxationResult {  // HopfieldNetwork.go:268-273 (last history entry only, plus NumSteps)
    bool Stable; int NumSteps; std::vector<double> DistancesToTargets; State FinalState; std::vector<double> EnergyProfile;
};

class HopfieldNetwork {
public:
    ~HopfieldNetwork() { if (handle_) hn_destroy(handle_); if (rng_) hn_rng_destroy(rng_); }
    HopfieldNetwork(const HopfieldNetwork&) = delete;
    HopfieldNetwork& operator=(const HopfieldNetwork&) = delete;

    int GetDimension() const { return cfg_.dimension; }
    std::vector<double> GetMatrix() const {  // row-major copy (HopfieldNetwork.go:93-95 returns the live matrix)
        std::vector<double> w((size_t)cfg_.dimension * cfg_.dimension);
        check(hn_get_weights(handle_, w.data()));
        return w;
    }
    void SetMatrix(const std::vector<double>& w) { check(hn_set_weights(handle_, w.data())); }
    HopfieldNetworkSummary GetNetworkSummary() const {  // HopfieldNetwork.go:137-149
        return HopfieldNetworkSummary{GetMatrix(), cfg_.dimension, cfg_.force_symmetric != 0, cfg_.force_zero_diagonal != 0, epochs_,
                                      cfg_.max_relax_unstable_units, cfg_.max_relax_iterations, noise_scale_, units_updated_};
    }

    std::vector<double> AllUnitEnergies(const State& s) { return profiles({s}, true).energies[0]; }   // :198-200
    // :185-187 — the reference's scalar loop (Binary: -1 per term, BinaryDomainManager.go:43-52), not AllUnitEnergies(s)[i]
    double UnitEnergy(const State& s, int i) {
        std::vector<uint64_t> packed((size_t)hn_words(cfg_.dimension));
        check(hn_pack_states_f64(cfg_.dimension, 1, s.data(), packed.data()));
        double e = 0.0;
        check(hn_unit_energy(handle_, packed.data(), 1, i, &e));
        return e;
    }
    double StateEnergy(const State& s) { return profiles({s}, false).total[0]; }                       // :171-173
    bool StateIsStable(const State& s) { return profiles({s}, false).stable[0] != 0; }                 // :214-225
    bool AllStatesAreStable(const std::vector<State>& v) {                                              // :236-243
        for (uint8_t f : profiles(v, false).stable) if (!f) return false;
        return true;
    }

    // LearnStates (HopfieldNetwork.go:254-262): activation in place, append to the learned states, run the
    // learning method with the selected rule (noise + update orders drawn from the network RNG), normalise.
    std::vector<LearnStateData> LearnStates(std::vector<State>& states) {
        const int n = (int)states.size(), N = cfg_.dimension;
        for (State& s : states) for (double& x : s) x = (x <= 0.0) ? low() : 1.0;
        std::vector<uint64_t> packed = pack(states);
        const int cap = epochs_ * n * (method_ == IterativeBatchMethod ? std::max(1, n / 5) : 1);
        std::vector<int32_t> ep((size_t)cap), ix((size_t)cap);
        std::vector<uint8_t> st((size_t)cap);
        std::vector<double> en((size_t)cap * N);
        int32_t rows = 0;
        check(hn_learn_states(handle_, packed.data(), n, rule_, method_, epochs_, noise_, noise_scale_, rng_, cap,
                              ep.data(), ix.data(), st.data(), en.data(), &rows));
        std::vector<LearnStateData> out;
        for (int r = 0; r < rows && r < cap; r++)
            out.push_back({ep[r], ix[r], std::vector<double>(en.begin() + (size_t)r * N, en.begin() + (size_t)(r + 1) * N), st[r] != 0});
        return out;
    }

    // ConcurrentRelaxStates (HopfieldNetwork.go:461-490); numThreads accepted and ignored for compute.
    // States are relaxed in place.  The schedule is what one relaxation goroutine would draw.
    std::vector<RelaxationResult> ConcurrentRelaxStates(std::vector<State>& states, int /*numThreads*/) {
        const int b = (int)states.size(), N = cfg_.dimension, w = hn_words(N);
        int32_t p = 0;
        check(hn_num_targets(handle_, &p));
        std::vector<uint64_t> packed = pack(states), finals((size_t)b * w);
        std::vector<uint16_t> list((size_t)N), pool((size_t)cfg_.max_relax_iterations * N);
        check(hn_rng_perm_pool(rng_, N, cfg_.max_relax_iterations, 1, list.data(), pool.data()));
        std::vector<int32_t> steps((size_t)b), dist((size_t)b * p + 1);
        std::vector<uint8_t> stable((size_t)b);
        std::vector<double> en((size_t)b * N);
        hn_relax_out out{};
        out.final_states = finals.data(); out.num_steps = steps.data(); out.stable = stable.data();
        out.distances = p ? dist.data() : nullptr; out.energies = en.data();
        check(hn_relax_batch(handle_, packed.data(), b, pool.data(), cfg_.max_relax_iterations, nullptr, 0, &out));
        std::vector<RelaxationResult> res((size_t)b);
        for (int i = 0; i < b; i++) {
            unpack(finals.data() + (size_t)i * w, states[(size_t)i]);
            res[i].Stable = stable[i] != 0; res[i].NumSteps = steps[i]; res[i].FinalState = states[(size_t)i];
            res[i].DistancesToTargets.assign(dist.begin() + (size_t)i * p, dist.begin() + (size_t)(i + 1) * p);
            res[i].EnergyProfile.assign(en.begin() + (size_t)i * N, en.begin() + (size_t)(i + 1) * N);
        }
        return res;
    }
    RelaxationResult RelaxState(State& s) {  // :302-359
        std::vector<State> v{s};
        RelaxationResult r = ConcurrentRelaxStates(v, 1)[0];
        s = v[0];
        return r;
    }
    void UpdateState(State& s) {  // :280-291
        const int N = cfg_.dimension;
        std::vector<uint16_t> perm((size_t)N);
        for (int i = 0; i < N; i++) perm[i] = (uint16_t)i;
        check(hn_rng_shuffle_u16(rng_, perm.data(), N));
        std::vector<State> v{s};
        std::vector<uint64_t> packed = pack(v);
        check(hn_update_states(handle_, packed.data(), 1, perm.data()));
        unpack(packed.data(), s);
    }

private:
    friend class HopfieldNetworkBuilder;
    HopfieldNetwork() = default;
    struct Profiles { std::vector<std::vector<double>> energies; std::vector<uint8_t> stable; std::vector<double> total; };
    double low() const { return cfg_.domain == BinaryDomain ? 0.0 : -1.0; }
    std::vector<uint64_t> pack(const std::vector<State>& v) const {
        const int N = cfg_.dimension, w = hn_words(N);
        std::vector<uint64_t> out((size_t)v.size() * w, 0);
        for (size_t s = 0; s < v.size(); s++) {
            if ((int)v[s].size() != N) throw std::invalid_argument("state length != network dimension");
            for (int i = 0; i < N; i++) if (v[s][(size_t)i] > 0.0) out[s * w + (size_t)(i >> 6)] |= 1ULL << (i & 63);
        }
        return out;
    }
    void unpack(const uint64_t* p, State& s) const {
        s.resize((size_t)cfg_.dimension);
        for (int i = 0; i < cfg_.dimension; i++) s[(size_t)i] = ((p[i >> 6] >> (i & 63)) & 1ULL) ? 1.0 : low();
    }
    Profiles profiles(const std::vector<State>& v, bool want) {
        const int n = (int)v.size(), N = cfg_.dimension;
        std::vector<uint64_t> packed = pack(v);
        std::vector<double> en(want ? (size_t)n * N : 0);
        Profiles p; p.stable.resize((size_t)n); p.total.resize((size_t)n);
        check(hn_energy_profiles(handle_, packed.data(), n, want ? en.data() : nullptr, nullptr, p.stable.data(), p.total.data()));
        if (want) for (int i = 0; i < n; i++) p.energies.emplace_back(en.begin() + (size_t)i * N, en.begin() + (size_t)(i + 1) * N);
        return p;
    }
    hn_config cfg_{};
    hn_handle handle_ = nullptr;
    hn_rng rng_ = nullptr;
    int epochs_ = 0, rule_ = 0, method_ = 0, noise_ = 0, units_updated_ = 1;
    double noise_scale_ = 0.0;
};

class HopfieldNetworkBuilder {  // HopfieldNetworkBuilder.go:16-279
public:
    HopfieldNetworkBuilder() { hn_default_config(&cfg_); }  // NewHopfieldNetworkBuilder :40-55
    HopfieldNetworkBuilder& SetRandMatrixInit(bool f) { rand_init_ = f; return *this; }
    HopfieldNetworkBuilder& SetNetworkDimension(int d) { cfg_.dimension = d; return *this; }
    HopfieldNetworkBuilder& SetNetworkDomain(DomainEnum d) { cfg_.domain = d; return *this; }
    HopfieldNetworkBuilder& SetForceSymmetric(bool f) { cfg_.force_symmetric = f; return *this; }
    HopfieldNetworkBuilder& SetForceZeroDiagonal(bool f) { cfg_.force_zero_diagonal = f; return *this; }
    HopfieldNetworkBuilder& SetNetworkLearningMethod(LearningMethodEnum m) { method_ = m; return *this; }
    HopfieldNetworkBuilder& SetNetworkLearningRule(LearningRuleEnum r) { rule_ = r; return *this; }
    HopfieldNetworkBuilder& SetEpochs(int e) { epochs_ = e; return *this; }
    HopfieldNetworkBuilder& SetMaximumRelaxationUnstableUnits(int u) { cfg_.max_relax_unstable_units = u; return *this; }
    HopfieldNetworkBuilder& SetMaximumRelaxationIterations(int i) { cfg_.max_relax_iterations = i; return *this; }
    HopfieldNetworkBuilder& SetLearningRate(double r) { cfg_.learning_rate = r; return *this; }
    HopfieldNetworkBuilder& SetLearningNoiseMethod(NoiseApplicationEnum m) { noise_ = m; return *this; }
    HopfieldNetworkBuilder& SetLearningNoiseRatio(double r) { noise_scale_ = r; return *this; }
    HopfieldNetworkBuilder& SetUnitsUpdatedPerStep(int u) { units_ = u; return *this; }
    HopfieldNetworkBuilder& SetAllowIntensiveDataCollection(bool) { return *this; }
    HopfieldNetworkBuilder& SetSeed(uint64_t s) { seed_ = s; return *this; }   // additive: reproducible runs
    HopfieldNetworkBuilder& SetDevice(int d) { cfg_.device = d; return *this; } // additive
    // additive: which of gonum's two Dlassq variants the Frobenius norm of LearnStates reproduces (hn_set_norm_algorithm)
    HopfieldNetworkBuilder& SetNormAlgorithm(int a) { norm_algorithm_ = a; return *this; }

    std::unique_ptr<HopfieldNetwork> Build() {
        const std::string pre = "HopfieldNetworkBuilder encountered an error during build! ";
        if (cfg_.dimension <= 0) throw std::invalid_argument(pre + "Dimension must be explicitly set to a positive integer!");
        if (epochs_ <= 0) throw std::invalid_argument(pre + "Epochs must be a positive integer!");
        if (noise_scale_ < 0.0 || noise_scale_ > 1.0) throw std::invalid_argument(pre + "learningNoiseRatio must be in range [0.0, 1.0]!");
        if (units_ < 0 || units_ > cfg_.dimension)
            throw std::invalid_argument(pre + "unitsUpdatedPerStep must be a positive integer that is smaller than the network dimension!");
        std::unique_ptr<HopfieldNetwork> net(new HopfieldNetwork());
        net->cfg_ = cfg_; net->epochs_ = epochs_; net->rule_ = rule_; net->method_ = method_; net->noise_ = noise_;
        net->noise_scale_ = noise_scale_; net->units_updated_ = units_;
        check(hn_rng_create(seed_, &net->rng_));
        check(hn_create(&cfg_, &net->handle_));
        if (norm_algorithm_ != HN_NORM_DLASSQ_LAPACK310) check(hn_set_norm_algorithm(net->handle_, norm_algorithm_));
        if (rand_init_) {  // :235-246 distuv.Normal{0, 0.01, randSrc}
            std::vector<double> w((size_t)cfg_.dimension * cfg_.dimension);
            for (double& x : w) x = hn_rng_norm_float64(net->rng_) * 0.01 + 0.0;
            net->SetMatrix(w);
        }
        return net;
    }

private:
    hn_config cfg_{};
    bool rand_init_ = false;
    int epochs_ = 0, rule_ = HebbianLearningRule, method_ = FullSetMethod, noise_ = None, units_ = 1;
    int norm_algorithm_ = HN_NORM_DLASSQ_LAPACK310;
    double noise_scale_ = 0.0;
    uint64_t seed_ = 0x5EED;
};

inline HopfieldNetworkBuilder NewHopfieldNetworkBuilder() { return HopfieldNetworkBuilder(); }

}  // namespace hopfieldnetwork
