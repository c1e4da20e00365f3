// HopfieldNetworkBuilder.go — the reference's builder (hopfieldnetwork/HopfieldNetworkBuilder.go:16-279) edited for
// the device-resident network: every setter and the four Build() panics are kept (same names, same messages), the
// network's rule and method are remembered as their enums (the device library runs them), and Build() hands the
// configuration to newDeviceNetwork (hopfield_b200.go) instead of allocating a gonum matrix.  Additive: SetSeed,
// SetDevice.
//
// UNVERIFIED: not compiled in this repository's build environment (no Go toolchain; see hopfield_b200.go).
package hopfieldnetwork

import (
	"log"
	"time"

	"hmcalister/hopfield/hopfieldnetwork/datacollector"
	"hmcalister/hopfield/hopfieldnetwork/domain"
	"hmcalister/hopfield/hopfieldnetwork/noiseapplication"

	"golang.org/x/exp/rand"
	"gonum.org/v1/gonum/stat/distuv"
)

type HopfieldNetworkBuilder struct {
	randMatrixInit                 bool
	dimension                      int
	domain                         domain.DomainEnum
	forceSymmetric                 bool
	forceZeroDiagonal              bool
	learningMethod                 LearningMethodEnum
	learningRule                   LearningRuleEnum
	epochs                         int
	maximumRelaxationUnstableUnits int
	maximumRelaxationIterations    int
	learningRate                   float64
	learningNoiseMethod            noiseapplication.NoiseApplicationMethod
	learningNoiseScale             float64
	unitsUpdatedPerStep            int
	dataCollector                  *datacollector.DataCollector
	logger                         *log.Logger
	allowIntensiveDataCollection   bool
	seed                           uint64 // additive; 0 = time-seeded as the reference (:231)
	device                         int    // additive; CUDA device ordinal
}

// NewHopfieldNetworkBuilder: the reference's defaults (:40-55).
func NewHopfieldNetworkBuilder() *HopfieldNetworkBuilder {
	return &HopfieldNetworkBuilder{
		domain:                      domain.BipolarDomain,
		forceSymmetric:              true,
		forceZeroDiagonal:           true,
		maximumRelaxationIterations: 100,
		learningRate:                1.0,
		unitsUpdatedPerStep:         1,
		dataCollector:               datacollector.NewDataCollector(),
		logger:                      log.Default(),
	}
}

func (b *HopfieldNetworkBuilder) SetRandMatrixInit(f bool) *HopfieldNetworkBuilder { b.randMatrixInit = f; return b }
func (b *HopfieldNetworkBuilder) SetNetworkDimension(d int) *HopfieldNetworkBuilder { b.dimension = d; return b }
func (b *HopfieldNetworkBuilder) SetNetworkDomain(d domain.DomainEnum) *HopfieldNetworkBuilder {
	b.domain = d
	return b
}
func (b *HopfieldNetworkBuilder) SetForceSymmetric(f bool) *HopfieldNetworkBuilder    { b.forceSymmetric = f; return b }
func (b *HopfieldNetworkBuilder) SetForceZeroDiagonal(f bool) *HopfieldNetworkBuilder { b.forceZeroDiagonal = f; return b }
func (b *HopfieldNetworkBuilder) SetNetworkLearningMethod(m LearningMethodEnum) *HopfieldNetworkBuilder {
	b.learningMethod = m
	return b
}
func (b *HopfieldNetworkBuilder) SetNetworkLearningRule(r LearningRuleEnum) *HopfieldNetworkBuilder {
	b.learningRule = r
	return b
}
func (b *HopfieldNetworkBuilder) SetEpochs(e int) *HopfieldNetworkBuilder { b.epochs = e; return b }
func (b *HopfieldNetworkBuilder) SetMaximumRelaxationUnstableUnits(u int) *HopfieldNetworkBuilder {
	b.maximumRelaxationUnstableUnits = u
	return b
}
func (b *HopfieldNetworkBuilder) SetMaximumRelaxationIterations(i int) *HopfieldNetworkBuilder {
	b.maximumRelaxationIterations = i
	return b
}
func (b *HopfieldNetworkBuilder) SetLearningRate(r float64) *HopfieldNetworkBuilder { b.learningRate = r; return b }
func (b *HopfieldNetworkBuilder) SetLearningNoiseMethod(m noiseapplication.NoiseApplicationEnum) *HopfieldNetworkBuilder {
	b.learningNoiseMethod = noiseapplication.GetNoiseApplicationMethod(m)
	return b
}
func (b *HopfieldNetworkBuilder) SetLearningNoiseRatio(r float64) *HopfieldNetworkBuilder {
	b.learningNoiseScale = r
	return b
}
func (b *HopfieldNetworkBuilder) SetUnitsUpdatedPerStep(u int) *HopfieldNetworkBuilder {
	b.unitsUpdatedPerStep = u
	return b
}
func (b *HopfieldNetworkBuilder) SetDataCollector(c *datacollector.DataCollector) *HopfieldNetworkBuilder {
	b.dataCollector = c
	return b
}
func (b *HopfieldNetworkBuilder) SetLogger(l *log.Logger) *HopfieldNetworkBuilder { b.logger = l; return b }
func (b *HopfieldNetworkBuilder) SetAllowIntensiveDataCollection(f bool) *HopfieldNetworkBuilder {
	b.allowIntensiveDataCollection = f
	return b
}

// SetSeed (additive): seed of the network generator; the reference seeds from the clock (:231), which makes its
// runs irreproducible (SURVEY F1).  SetDevice (additive): CUDA device ordinal of the replica.
func (b *HopfieldNetworkBuilder) SetSeed(seed uint64) *HopfieldNetworkBuilder { b.seed = seed; return b }
func (b *HopfieldNetworkBuilder) SetDevice(device int) *HopfieldNetworkBuilder { b.device = device; return b }

// Build: the four validation panics of :215-229 with the reference's texts, the generator of :231-232, the optional
// Gaussian(0, 0.01) initial matrix of :235-246 drawn from the same source, then the device network.
func (b *HopfieldNetworkBuilder) Build() *HopfieldNetwork {
	if b.dimension <= 0 {
		panic("HopfieldNetworkBuilder encountered an error during build! Dimension must be explicitly set to a positive integer!")
	}
	if b.epochs <= 0 {
		panic("HopfieldNetworkBuilder encountered an error during build! Epochs must be a positive integer!")
	}
	if b.learningNoiseScale < 0.0 || b.learningNoiseScale > 1.0 {
		panic("HopfieldNetworkBuilder encountered an error during build! learningNoiseRatio must be in range [0.0, 1.0]!")
	}
	if b.unitsUpdatedPerStep < 0 || b.unitsUpdatedPerStep > b.dimension {
		panic("HopfieldNetworkBuilder encountered an error during build! unitsUpdatedPerStep must be a positive integer that is smaller than the network dimension!")
	}
	seed := b.seed
	if seed == 0 {
		seed = uint64(time.Now().UnixNano())
	}
	randSrc := rand.NewSource(seed)
	var initial []float64
	if b.randMatrixInit {
		normal := distuv.Normal{Mu: 0, Sigma: 0.01, Src: randSrc}
		initial = make([]float64, b.dimension*b.dimension)
		for i := range initial {
			initial[i] = normal.Rand()
		}
	}
	noise := b.learningNoiseMethod
	if noise == nil {
		noise = noiseapplication.GetNoiseApplicationMethod(noiseapplication.None)
	}
	net := &HopfieldNetwork{
		dimension:                      b.dimension,
		domain:                         b.domain,
		domainManager:                  domain.GetDomainManager(b.domain),
		forceSymmetric:                 b.forceSymmetric,
		forceZeroDiagonal:              b.forceZeroDiagonal,
		learningMethodEnum:             b.learningMethod,
		learningRuleEnum:               b.learningRule,
		learningNoiseMethod:            noise,
		epochs:                         b.epochs,
		maximumRelaxationUnstableUnits: b.maximumRelaxationUnstableUnits,
		maximumRelaxationIterations:    b.maximumRelaxationIterations,
		learningRate:                   b.learningRate,
		learningNoiseScale:             b.learningNoiseScale,
		unitsUpdatedPerStep:            b.unitsUpdatedPerStep,
		randomGenerator:                rand.New(randSrc),
		dataCollector:                  b.dataCollector,
		logger:                         b.logger,
		allowIntensiveDataCollection:   b.allowIntensiveDataCollection,
	}
	return newDeviceNetwork(net, b.device, initial)
}
