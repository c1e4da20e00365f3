// Exercises include/hopfieldnetwork.hpp (the C++ host mirror of the Go API) against libhopfield_b200.so.
//   host_mirror_test cpu : builder validation + "no CPU fallback" (no GPU needed)
//   host_mirror_test gpu : P=1 Hebbian known answer through LearnStates / ConcurrentRelaxStates
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "hopfieldnetwork.hpp"

using namespace hopfieldnetwork;

static int fail(const char* what) { std::printf("FAIL: %s\n", what); return 1; }

template <class F> static bool throws_with(F f, const char* needle) {
    try { f(); } catch (const std::exception& e) { return std::strstr(e.what(), needle) != nullptr; }
    return false;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    if (!throws_with([] { NewHopfieldNetworkBuilder().SetEpochs(1).Build(); }, "Dimension must be explicitly set")) return fail("dimension check");
    if (!throws_with([] { NewHopfieldNetworkBuilder().SetNetworkDimension(8).Build(); }, "Epochs must be a positive integer")) return fail("epochs check");
    if (!throws_with([] { NewHopfieldNetworkBuilder().SetNetworkDimension(8).SetEpochs(1).SetLearningNoiseRatio(2.0).Build(); }, "learningNoiseRatio")) return fail("noise check");
    if (!throws_with([] { NewHopfieldNetworkBuilder().SetNetworkDimension(8).SetEpochs(1).SetUnitsUpdatedPerStep(9).Build(); }, "unitsUpdatedPerStep")) return fail("units check");
    if (mode == "cpu") {
        if (!throws_with([] { NewHopfieldNetworkBuilder().SetNetworkDimension(8).SetEpochs(1).Build(); }, "no CPU fallback")) return fail("no-device check");
        std::printf("OK cpu\n");
        return 0;
    }
    const int n = 96;
    auto net = NewHopfieldNetworkBuilder().SetNetworkDimension(n).SetEpochs(1).SetNetworkLearningRule(HebbianLearningRule)
                   .SetNetworkLearningMethod(FullSetMethod).SetSeed(42).Build();
    State x((size_t)n);
    for (int i = 0; i < n; i++) x[(size_t)i] = (i * 7 % 5 < 2) ? 1.0 : -1.0;
    std::vector<State> targets{x};
    auto lsd = net->LearnStates(targets);
    if (lsd.size() != 1 || !lsd[0].Stable || lsd[0].Epoch != 0) return fail("LearnStateData");
    auto w = net->GetMatrix();
    double ss = 0;
    for (double v : w) ss += v * v;
    if (std::fabs(std::sqrt(ss) - 1.0) > 1e-12) return fail("Frobenius norm");
    for (int i = 0; i < n; i++) if (w[(size_t)i * n + i] != 0.0) return fail("diagonal");
    if (!net->StateIsStable(x)) return fail("target stable");
    std::vector<State> probes(8, x);
    for (int p = 0; p < 8; p++) for (int i = 0; i < 20 + p; i++) probes[(size_t)p][(size_t)((i * 13 + p) % n)] *= -1.0;
    auto res = net->ConcurrentRelaxStates(probes, 4);
    for (auto& r : res) {
        if (!r.Stable || r.NumSteps != 2 || r.DistancesToTargets.size() != 1 || r.DistancesToTargets[0] != 0.0) return fail("relaxation result");
        if (r.FinalState != x) return fail("final state");
        for (double e : r.EnergyProfile) if (!(e < 0.0)) return fail("energy profile");
    }
    // UnitEnergy is the reference's scalar loop: for a bipolar network it agrees with AllUnitEnergies to rounding
    if (std::fabs(net->UnitEnergy(x, 5) - net->AllUnitEnergies(x)[5]) > 1e-12) return fail("UnitEnergy");
    auto summary = net->GetNetworkSummary();
    if (summary.Dimension != n || summary.Epochs != 1 || !summary.ForceSymmetric || summary.Matrix.size() != (size_t)n * n) return fail("summary");
    std::printf("OK gpu\n");
    return 0;
}
