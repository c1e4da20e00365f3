// Package hopfieldnetwork — cgo drop-in for hmcalister/hopfield/hopfieldnetwork.
//
// This file replaces HopfieldNetwork.go, LearningRule.go and LearningMethod.go of the reference (the enum
// declarations of LearningRule.go:32-39 / LearningMethod.go:22-25 are kept as they are); the edited
// HopfieldNetworkBuilder.go next to it keeps every setter and the four Build panics and calls newDeviceNetwork
// below instead of allocating a gonum matrix.  Every method forwards to libhopfield_b200.so
// (include/hopfield_b200.h); the sub-packages domain, noiseapplication, distancemeasure, datacollector,
// states and hopfieldutils are used unchanged.
//
// UNVERIFIED: NOT COMPILED in the build environment of this repository (no Go toolchain there, and the gonum /
// x-exp module sources are absent).  It documents the binding a maintainer adds; every behaviour lives in the C
// library, where the ctypes tests reach it.  INTEGRATION.md walks through it.
package hopfieldnetwork

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../lib -lhopfield_b200 -Wl,-rpath,${SRCDIR}/../../lib
#include <stdlib.h>
#include "hopfield_b200.h"
*/
import "C"

import (
	"fmt"
	"log"
	"runtime"
	"unsafe"

	"hmcalister/hopfield/hopfieldnetwork/datacollector"
	"hmcalister/hopfield/hopfieldnetwork/domain"
	"hmcalister/hopfield/hopfieldnetwork/noiseapplication"
	"hmcalister/hopfield/hopfieldutils"

	"golang.org/x/exp/rand"
	"gonum.org/v1/gonum/mat"
)

// HopfieldNetwork keeps the reference's fields that are host-side state; the weight matrix and the
// learned states live in HBM behind `handle`.
type HopfieldNetwork struct {
	handle                         C.hn_handle
	dimension                      int
	domain                         domain.DomainEnum
	domainManager                  domain.DomainManager
	forceSymmetric                 bool
	forceZeroDiagonal              bool
	learningMethodEnum             LearningMethodEnum
	learningRuleEnum               LearningRuleEnum
	learningNoiseMethod            noiseapplication.NoiseApplicationMethod
	epochs                         int
	maximumRelaxationUnstableUnits int
	maximumRelaxationIterations    int
	learningRate                   float64
	learningNoiseScale             float64
	unitsUpdatedPerStep            int
	randomGenerator                *rand.Rand
	targetStates                   []*mat.VecDense
	dataCollector                  *datacollector.DataCollector
	logger                         *log.Logger
	allowIntensiveDataCollection   bool
}

func check(rc C.int) {
	if rc != C.HN_OK {
		panic(fmt.Sprintf("libhopfield_b200: [%d] %s", int(rc), C.GoString(C.hn_last_error())))
	}
}

// newDeviceNetwork is called by HopfieldNetworkBuilder.Build after its four validation panics
// (HopfieldNetworkBuilder.go:215-229).  initial is nil for the zero matrix or the N*N Gaussian(0,0.01)
// draw of :235-246.
func newDeviceNetwork(net *HopfieldNetwork, device int, initial []float64) *HopfieldNetwork {
	var cfg C.hn_config
	C.hn_default_config(&cfg)
	// cfg.column_stream: C.HN_COLUMNS_AUTO (default), C.HN_COLUMNS_F64, C.HN_COLUMNS_F32 or C.HN_COLUMNS_I16 — results are identical,
	// see include/hopfield_b200.h; a SetColumnStream builder option can expose it.
	cfg.dimension = C.int32_t(net.dimension)
	cfg.domain = C.int32_t(net.domain)
	cfg.force_symmetric = boolToC(net.forceSymmetric)
	cfg.force_zero_diagonal = boolToC(net.forceZeroDiagonal)
	cfg.max_relax_unstable_units = C.int32_t(net.maximumRelaxationUnstableUnits)
	cfg.max_relax_iterations = C.int32_t(net.maximumRelaxationIterations)
	cfg.learning_rate = C.double(net.learningRate)
	cfg.device = C.int32_t(device)
	check(C.hn_create(&cfg, &net.handle))
	// ||W||_F of LearnStates (:260): gonum's Dlassq exists in two variants across releases; the library reproduces both bit
	// for bit for Hebbian / Delta weights.  A maintainer checks once which one their gonum has (mat.Norm of a saved Hebbian
	// matrix against both) and, if it is the classic recurrence, selects it here:
	//   check(C.hn_set_norm_algorithm(net.handle, C.HN_NORM_DLASSQ_CLASSIC))
	if initial != nil {
		check(C.hn_set_weights(net.handle, (*C.double)(unsafe.Pointer(&initial[0]))))
	}
	runtime.SetFinalizer(net, func(n *HopfieldNetwork) { C.hn_destroy(n.handle) })
	return net
}

func boolToC(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

func (network *HopfieldNetwork) words() int { return (network.dimension + 63) / 64 }

// pack flattens []*mat.VecDense into one contiguous packed buffer (cgo: no Go pointer to Go pointer may
// cross the boundary).  The activation function is applied while packing: x <= 0 -> 0 bit.
func (network *HopfieldNetwork) pack(states []*mat.VecDense) []uint64 {
	w := network.words()
	out := make([]uint64, w*len(states))
	for s, v := range states {
		raw := v.RawVector().Data
		for i := 0; i < network.dimension; i++ {
			if raw[i] > 0 {
				out[s*w+i/64] |= 1 << uint(i%64)
			}
		}
	}
	return out
}

// unpackInto writes packed state s back into the caller's vector (the reference relaxes in place).
func (network *HopfieldNetwork) unpackInto(packed []uint64, s int, v *mat.VecDense) {
	w := network.words()
	low := -1.0
	if network.domain == domain.BinaryDomain {
		low = 0.0
	}
	for i := 0; i < network.dimension; i++ {
		if packed[s*w+i/64]>>uint(i%64)&1 == 1 {
			v.SetVec(i, 1.0)
		} else {
			v.SetVec(i, low)
		}
	}
}

// ---- getters (HopfieldNetwork.go:93-149) ----

// GetMatrix returns a COPY of the device-resident matrix (the reference returns the live matrix;
// caller-side mutation cannot propagate — use SetMatrix).
func (network *HopfieldNetwork) GetMatrix() *mat.Dense {
	data := make([]float64, network.dimension*network.dimension)
	check(C.hn_get_weights(network.handle, (*C.double)(unsafe.Pointer(&data[0]))))
	return mat.NewDense(network.dimension, network.dimension, data)
}

// SetMatrix uploads a matrix (additive extension: resume from matrix.bin).
func (network *HopfieldNetwork) SetMatrix(m *mat.Dense) {
	data := mat.DenseCopyOf(m).RawMatrix().Data
	check(C.hn_set_weights(network.handle, (*C.double)(unsafe.Pointer(&data[0]))))
}

func (network *HopfieldNetwork) GetDimension() int                 { return network.dimension }
func (network *HopfieldNetwork) GetLearnedStates() []*mat.VecDense { return network.targetStates }

// HopfieldNetworkSummary / GetNetworkSummary: HopfieldNetwork.go:124-149, consumed by main.go:262-279.
type HopfieldNetworkSummary struct {
	Matrix                         *mat.Dense
	Dimension                      int
	ForceSymmetric                 bool
	ForceZeroDiagonal              bool
	Epochs                         int
	MaximumRelaxationUnstableUnits int
	MaximumRelaxationIterations    int
	LearningNoiseScale             float64
	UnitsUpdatedPerStep            int
}

func (network *HopfieldNetwork) GetNetworkSummary() *HopfieldNetworkSummary {
	return &HopfieldNetworkSummary{
		Matrix: network.GetMatrix(), Dimension: network.dimension,
		ForceSymmetric: network.forceSymmetric, ForceZeroDiagonal: network.forceZeroDiagonal,
		Epochs: network.epochs, MaximumRelaxationUnstableUnits: network.maximumRelaxationUnstableUnits,
		MaximumRelaxationIterations: network.maximumRelaxationIterations,
		LearningNoiseScale: network.learningNoiseScale, UnitsUpdatedPerStep: network.unitsUpdatedPerStep,
	}
}

func (network *HopfieldNetwork) String() string {
	return fmt.Sprintf("Hopfield Network\n\tDomain: %s\n\tDimension: %d\n", network.domain, network.dimension)
}

// ---- energies / stability (HopfieldNetwork.go:171-243) ----

func (network *HopfieldNetwork) energyProfiles(states []*mat.VecDense, wantEnergies bool) ([][]float64, []bool, []float64) {
	n := len(states)
	packed := network.pack(states)
	var energies []float64
	var ep *C.double
	if wantEnergies {
		energies = make([]float64, n*network.dimension)
		ep = (*C.double)(unsafe.Pointer(&energies[0]))
	}
	stable := make([]C.uint8_t, n)
	total := make([]float64, n)
	check(C.hn_energy_profiles(network.handle, (*C.uint64_t)(unsafe.Pointer(&packed[0])), C.int32_t(n), ep, nil,
		&stable[0], (*C.double)(unsafe.Pointer(&total[0]))))
	profiles := make([][]float64, n)
	flags := make([]bool, n)
	for i := range profiles {
		if wantEnergies {
			profiles[i] = energies[i*network.dimension : (i+1)*network.dimension]
		}
		flags[i] = stable[i] != 0
	}
	return profiles, flags, total
}

func (network *HopfieldNetwork) AllUnitEnergies(state *mat.VecDense) []float64 {
	p, _, _ := network.energyProfiles([]*mat.VecDense{state}, true)
	return p[0]
}

// UnitEnergy is the reference's SCALAR loop (BipolarDomainManager.go:29-37; BinaryDomainManager.go:43-52 with its
// "-1" per term), not AllUnitEnergies(state)[unitIndex]: hn_unit_energy reproduces both.
func (network *HopfieldNetwork) UnitEnergy(state *mat.VecDense, unitIndex int) float64 {
	packed := network.pack([]*mat.VecDense{state})
	var e C.double
	check(C.hn_unit_energy(network.handle, (*C.uint64_t)(unsafe.Pointer(&packed[0])), 1, C.int32_t(unitIndex), &e))
	return float64(e)
}
func (network *HopfieldNetwork) StateEnergy(state *mat.VecDense) float64 {
	_, _, t := network.energyProfiles([]*mat.VecDense{state}, false)
	return t[0]
}
func (network *HopfieldNetwork) StateIsStable(state *mat.VecDense) bool {
	_, f, _ := network.energyProfiles([]*mat.VecDense{state}, false)
	return f[0]
}
func (network *HopfieldNetwork) AllStatesAreStable(states []*mat.VecDense) bool {
	_, f, _ := network.energyProfiles(states, false)
	for _, ok := range f {
		if !ok {
			return false
		}
	}
	return true
}

// ---- learning (HopfieldNetwork.go:254-262, LearningMethod.go:38-89, LearningRule.go:64-126) ----

// learnEpoch applies the selected rule once (all six of LearningRule.go:32-39) through hn_rule_epoch.  Noise and
// update orders are drawn from the network RNG — the reference's own golang.org/x/exp/rand stream — in the reference's
// order (per target: the noise draws, then the N-1 shuffle draws of UpdateState, LearningRule.go:98-104) and uploaded.
// The bipolar-mapped rules re-activate copies of the targets with the manager the reference names before the noise is
// applied (:78-87, :117-126, :161-170); the library applies the same value map on its side.
func (network *HopfieldNetwork) learnEpoch(states []*mat.VecDense) {
	packed := network.pack(states)
	sp := (*C.uint64_t)(unsafe.Pointer(&packed[0]))
	rule := C.int32_t(network.learningRuleEnum)
	switch network.learningRuleEnum {
	case HebbianLearningRule, BipolarMappedHebbianLearningRule:
		check(C.hn_rule_epoch(network.handle, rule, sp, nil, nil, C.int32_t(len(states)), nil))
		return
	}
	var mapper domain.DomainManager // nil: the rule sees the targets as they are
	switch network.learningRuleEnum {
	case BipolarMappedDeltaLearningRule, BipolarMappedThermalDeltaLearningRule:
		mapper = &domain.BipolarDomainManager{}
	}
	noised := make([]*mat.VecDense, len(states))
	perms := make([]uint16, len(states)*network.dimension)
	idx := make([]int, network.dimension)
	for i, s := range states {
		noised[i] = mat.VecDenseCopyOf(s)
		if mapper != nil {
			mapper.ActivationFunction(noised[i])
		}
		network.learningNoiseMethod(network.randomGenerator, noised[i], network.learningNoiseScale)
		network.domainManager.ActivationFunction(noised[i])
		for k := range idx { // UpdateState rebuilds the identity list every call (HopfieldNetwork.go:281)
			idx[k] = k
		}
		hopfieldutils.ShuffleList(network.randomGenerator, idx)
		for k, u := range idx {
			perms[i*network.dimension+k] = uint16(u)
		}
	}
	np := network.pack(noised)
	check(C.hn_rule_epoch(network.handle, rule, sp, (*C.uint64_t)(unsafe.Pointer(&np[0])),
		(*C.uint16_t)(unsafe.Pointer(&perms[0])), C.int32_t(len(states)), nil))
}

func (network *HopfieldNetwork) fullSet(states []*mat.VecDense) []*datacollector.LearnStateData {
	out := []*datacollector.LearnStateData{}
	for epoch := 0; epoch < network.epochs; epoch++ {
		network.learnEpoch(states)
		profiles, stable, _ := network.energyProfiles(states, true)
		all := true
		for i := range states {
			out = append(out, &datacollector.LearnStateData{Epoch: epoch, TargetStateIndex: i,
				EnergyProfile: profiles[i], Stable: stable[i]})
			all = all && stable[i]
		}
		if all {
			break
		}
	}
	return out
}

func (network *HopfieldNetwork) LearnStates(states []*mat.VecDense) []*datacollector.LearnStateData {
	for _, state := range states {
		network.domainManager.ActivationFunction(state)
	}
	network.targetStates = append(network.targetStates, states...)
	all := network.pack(network.targetStates)
	check(C.hn_set_targets(network.handle, (*C.uint64_t)(unsafe.Pointer(&all[0])), C.int32_t(len(network.targetStates))))
	var data []*datacollector.LearnStateData
	if network.learningMethodEnum == FullSetMethod {
		data = network.fullSet(states)
	} else { // IterativeBatchMethod, LearningMethod.go:70-89
		total := 0
		for it := 1; it <= len(states)/5; it++ {
			d := network.fullSet(states[:5*it])
			for _, row := range d {
				row.Epoch += total
			}
			total = d[len(d)-1].Epoch
			data = append(data, d...)
		}
	}
	check(C.hn_normalize_frobenius(network.handle, nil))
	return data
}

// ---- relaxation (HopfieldNetwork.go:268-490) ----

type RelaxationResult struct {
	Stable             bool
	DistancesToTargets []float64
	StateHistory       []*mat.VecDense
	EnergyHistory      [][]float64
}

// UpdateState: one sweep in a fresh shuffled order (HopfieldNetwork.go:280-291).
func (network *HopfieldNetwork) UpdateState(state *mat.VecDense) {
	idx := make([]int, network.dimension)
	for k := range idx {
		idx[k] = k
	}
	hopfieldutils.ShuffleList(network.randomGenerator, idx)
	perm := make([]uint16, network.dimension)
	for k, u := range idx {
		perm[k] = uint16(u)
	}
	packed := network.pack([]*mat.VecDense{state})
	check(C.hn_update_states(network.handle, (*C.uint64_t)(unsafe.Pointer(&packed[0])), 1, (*C.uint16_t)(unsafe.Pointer(&perm[0]))))
	network.unpackInto(packed, 0, state)
}

// schedule draws the unit orders one relaxation goroutine would see: one persistent list, shuffled in
// place once per sweep (HopfieldNetwork.go:374,393).
func (network *HopfieldNetwork) schedule(rows int) []uint16 {
	idx := make([]int, network.dimension)
	for k := range idx {
		idx[k] = k
	}
	pool := make([]uint16, rows*network.dimension)
	for r := 0; r < rows; r++ {
		hopfieldutils.ShuffleList(network.randomGenerator, idx)
		for k, u := range idx {
			pool[r*network.dimension+k] = uint16(u)
		}
	}
	return pool
}

// ConcurrentRelaxStates relaxes every state on the GPU.  numThreads is accepted for API compatibility
// and ignored for compute.  States are relaxed in place, as in the reference.
func (network *HopfieldNetwork) ConcurrentRelaxStates(states []*mat.VecDense, numThreads int) []*RelaxationResult {
	b := len(states)
	w := network.words()
	p := len(network.targetStates)
	packed := network.pack(states)
	pool := network.schedule(network.maximumRelaxationIterations)
	finals := make([]uint64, b*w)
	steps := make([]C.int32_t, b)
	stable := make([]C.uint8_t, b)
	dist := make([]C.int32_t, b*p+1)
	energies := make([]float64, b*network.dimension)
	var out C.hn_relax_out
	out.final_states = (*C.uint64_t)(unsafe.Pointer(&finals[0]))
	out.num_steps = &steps[0]
	out.stable = &stable[0]
	if p > 0 {
		out.distances = &dist[0]
	}
	out.energies = (*C.double)(unsafe.Pointer(&energies[0]))
	var flags C.uint32_t
	var history []uint64
	if network.allowIntensiveDataCollection {
		flags |= C.HN_RELAX_HISTORY
		history = make([]uint64, b*(network.maximumRelaxationIterations+1)*w)
		out.history = (*C.uint64_t)(unsafe.Pointer(&history[0]))
	}
	check(C.hn_relax_batch(network.handle, (*C.uint64_t)(unsafe.Pointer(&packed[0])), C.int32_t(b),
		(*C.uint16_t)(unsafe.Pointer(&pool[0])), C.int32_t(network.maximumRelaxationIterations), nil, flags, &out))
	results := make([]*RelaxationResult, b)
	for i := range results {
		n := int(steps[i])
		r := &RelaxationResult{Stable: stable[i] != 0, DistancesToTargets: make([]float64, p),
			StateHistory: make([]*mat.VecDense, n), EnergyHistory: make([][]float64, n)}
		for t := 0; t < p; t++ {
			r.DistancesToTargets[t] = float64(dist[i*p+t])
		}
		network.unpackInto(finals, i, states[i])
		r.StateHistory[n-1] = mat.VecDenseCopyOf(states[i])
		r.EnergyHistory[n-1] = energies[i*network.dimension : (i+1)*network.dimension]
		if network.allowIntensiveDataCollection {
			hs := make([]*mat.VecDense, n)
			for s := 0; s < n; s++ {
				hs[s] = mat.NewVecDense(network.dimension, nil)
				network.unpackInto(history[(i*(network.maximumRelaxationIterations+1)+s)*w:], 0, hs[s])
			}
			r.StateHistory = hs
			prof, _, _ := network.energyProfiles(hs, true)
			r.EnergyHistory = prof
		}
		results[i] = r
	}
	return results
}

// RelaxState relaxes a single state (HopfieldNetwork.go:302-359).
func (network *HopfieldNetwork) RelaxState(state *mat.VecDense) *RelaxationResult {
	return network.ConcurrentRelaxStates([]*mat.VecDense{state}, 1)[0]
}
