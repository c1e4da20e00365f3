/*
 * oracle/xrand.h — TEST INFRASTRUCTURE (see oracle/README.md).
 *
 * CPU restatement of the random stream the reference draws from:
 * golang.org/x/exp/rand @ v0.0.0-20230315142452-642cacee5cc0 (go.mod:12).
 * That module's source is NOT under /root/reference and there is no Go
 * toolchain here, so this is restated from the published PCG XSL-RR 128/64
 * algorithm (O'Neill 2014) and the package's documented behaviour; call sites
 * in the reference: hopfieldutils/HopfieldUtils.go:35-39 (Shuffle),
 * noiseapplication/NoiseApplication.go:88 (Float64), :103 (NormFloat64),
 * HopfieldNetworkBuilder.go:231-232 and states/StateGeneratorBuilder.go:32-33,123
 * (NewSource/New), states/StateGenerator.go:46 (distuv.Uniform.Rand).
 *
 * PARITY UNPINNED against Go itself; the core generator is pinned against
 * numpy's independent PCG64 (same algorithm) in tests/test_oracle_rng.py.
 */
#ifndef ORACLE_XRAND_H
#define ORACLE_XRAND_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t low, high; } orc_rng;

void     orc_rng_seed(orc_rng* r, uint64_t seed);          /* PCGSource.Seed       */
uint64_t orc_rng_uint64(orc_rng* r);                       /* PCGSource.Uint64     */
uint64_t orc_rng_uint64n(orc_rng* r, uint64_t n);          /* Rand.Uint64n         */
uint32_t orc_rng_uint32(orc_rng* r);                       /* Rand.Uint32          */
double   orc_rng_float64(orc_rng* r);                      /* Rand.Float64         */
double   orc_rng_norm_float64(orc_rng* r);                 /* Rand.NormFloat64     */
/* Rand.Shuffle driving hopfieldutils.ShuffleList on an int list */
void     orc_rng_shuffle_i32(orc_rng* r, int32_t* list, int32_t n);
void     orc_rng_shuffle_u16(orc_rng* r, uint16_t* list, int32_t n);
/* distuv.Uniform{Min,Max,Src}.Rand() */
double   orc_rng_uniform(orc_rng* r, double mn, double mx);
/* distuv.Normal{Mu,Sigma,Src}.Rand() */
double   orc_rng_normal(orc_rng* r, double mu, double sigma);

#ifdef __cplusplus
}
#endif
#endif
