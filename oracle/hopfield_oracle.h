/*
 * oracle/hopfield_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference algorithm on the hot path
 * (hmcalister/Hopfield-Network-Go).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this; the product
 * library (libhopfield_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and cannot be built here (no Go toolchain; gonum and
 * x/exp/rand sources absent).  The oracle is pinned only by hand-derived
 * known-answer cases and an independent numpy restatement (tests/).
 *
 * Arithmetic that lives in gonum v0.12.0 (go.mod:5, source absent) is restated
 * from its published algorithm:
 *   mat.Dot -> blas64.Ddot -> f64.DotUnitary (amd64 SSE2): four independent
 *     partial sums (lane l sums indices == l mod 4, multiply then add, no FMA),
 *     tail elements added to lane 0, result (a0+a2)+(a1+a3).
 *   VecDense.MulVec -> Dgemv(NoTrans) -> f64.GemvN: y_i = sum_j A_ij x_j; the
 *     assembly's summation order is not recoverable here; the oracle uses the
 *     pure-Go fallback's order, y_i = DotUnitary(row_i, x).
 *   Dense.RankOne -> Dger: a_ij += (alpha*x_i)*y_j, unfused.
 *   Dense.Norm(2) = Frobenius via Dlange -> per-row Dlassq (scaled sum of squares).
 */
#ifndef HOPFIELD_ORACLE_H
#define HOPFIELD_ORACLE_H
#include <stdint.h>
#include "xrand.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BIPOLAR = 0, ORC_BINARY = 1 };

/* HopfieldNetwork struct (HopfieldNetwork.go:25-47), numeric fields only. */
typedef struct {
    int32_t n;                 /* dimension                           */
    int32_t domain;            /* domain                              */
    int32_t force_symmetric;   /* forceSymmetric                      */
    int32_t force_zero_diag;   /* forceZeroDiagonal                   */
    int32_t max_unstable;      /* maximumRelaxationUnstableUnits      */
    int32_t max_iter;          /* maximumRelaxationIterations         */
    double  lr;                /* learningRate                        */
    double* w;                 /* matrix, row-major n*n (caller owned)*/
    int32_t n_targets;         /* len(targetStates)                   */
    const double* targets;     /* targetStates, row-major [n_targets][n] (caller owned) */
} orc_net;

/* gonum f64.DotUnitary order (see header comment). */
double orc_dot_unitary(const double* x, const double* y, int32_t n);

/* domain/BipolarDomainManager.go:10-22, BinaryDomainManager.go:23-35 */
double orc_activation_unit(int32_t domain, double x);
void   orc_activation(int32_t domain, double* v, int32_t n);

/* AllUnitEnergies: HopfieldNetwork.go:198-200 -> BipolarDomainManager.go:39-46 /
 * BinaryDomainManager.go:54-63.  e must hold n doubles. */
void   orc_all_unit_energies(const orc_net* net, const double* state, double* e);
/* UnitEnergy: BipolarDomainManager.go:29-37 / BinaryDomainManager.go:43-52 */
double orc_unit_energy(const orc_net* net, const double* state, int32_t i);
/* StateEnergy: BipolarDomainManager.go:48-55 */
double orc_state_energy(const orc_net* net, const double* state);
/* StateIsStable: HopfieldNetwork.go:214-225.  *unstable_count optional. */
int    orc_state_is_stable(const orc_net* net, const double* state, int32_t* unstable_count);

/* distancemeasure/DistanceMeasure.go:12-18 with :28-42 (Manhattan with inversion) */
void   orc_distances_to_targets(const orc_net* net, const double* state, double* out);

/* decision margin used for reporting (DESIGN.md "margin"):
 * tau = tau_base + tau_per_flip * flips */
void   orc_margin(const orc_net* net, double* tau_base, double* tau_per_flip);

/* UpdateState (HopfieldNetwork.go:280-291) with the shuffled order given. */
void   orc_update_state(const orc_net* net, double* state, const uint16_t* perm);

typedef struct {
    int32_t stable;        /* RelaxationResult.Stable                           */
    int32_t num_steps;     /* len(StateHistory) = stepIndex+1 (main.go:231)     */
    int32_t flips;         /* counter                                           */
    int32_t margin_hits;   /* decisions (incl. stability tests) with |h| < tau  */
    double  min_abs_h;     /* smallest |h| seen at any sweep decision           */
} orc_relax_info;

/* concurrentRelaxStateRoutine body for one state (HopfieldNetwork.go:382-441)
 * == RelaxState (:302-359) without the logging.
 *   state       : n doubles, relaxed in place
 *   perm_pool   : [n_pool][n]; sweep t (0-based) uses row offset+t
 *   distances   : optional [n_targets]
 *   energies    : optional [n] final EnergyProfile
 *   history     : optional [(max_iter+1)][n] state after each step (step 0 = input);
 *                 only rows < num_steps are written
 *   faithful_cost: nonzero -> recompute the stability gemv as often as the
 *                 reference does (:403,:405,:408) — used by the CPU baseline timing
 * returns 0, or -1 when the pool is exhausted. */
int orc_relax_state(const orc_net* net, double* state, const uint16_t* perm_pool, int32_t n_pool,
                    int64_t offset, orc_relax_info* info, double* distances, double* energies,
                    double* history, int faithful_cost);

/* ConcurrentRelaxStates (HopfieldNetwork.go:461-490) over B states.
 *   offsets NULL -> all 0; serial_stream != 0 -> offsets ignored, probe p starts
 *   at row sum_{q<p} sweeps(q) (threads=1 semantics; forces 1 thread)
 *   n_threads worker threads pull states from a shared queue.
 * outputs (each optional): info[B], distances[B][n_targets], energies[B][n]. */
int orc_relax_batch(const orc_net* net, double* states, int32_t n_states, const uint16_t* perm_pool,
                    int32_t n_pool, const int32_t* offsets, int serial_stream, int n_threads,
                    orc_relax_info* info, double* distances, double* energies, int faithful_cost);

/* enforceConstraints: HopfieldNetwork.go:55-67 */
void orc_enforce_constraints(orc_net* net);
/* hebbian: LearningRule.go:64-75 (states already activated) */
void orc_hebbian(orc_net* net, const double* states, int32_t p);
/* delta: LearningRule.go:90-114 with the RNG-derived inputs explicit:
 *   noised : [p][n] copies after noise + activation (:99-102)
 *   perms  : [p][n] fresh shuffled order for each UpdateState (:103)
 *   relaxed_out: optional [p][n] */
void orc_delta(orc_net* net, const double* states, const double* noised, const uint16_t* perms,
               int32_t p, double* relaxed_out);
/* thermalDelta: LearningRule.go:129-158 with the RNG-derived inputs explicit (as orc_delta).
 * weightFactor = 1/(1+||W||_F); per state f = exp(-weightFactor*||W t||_2 / 1.0); dW += f*(t-r) t^T (no 0.5). */
void orc_thermal_delta(orc_net* net, const double* states, const double* noised, const uint16_t* perms,
                       int32_t p, double* relaxed_out);
/* Dense.Norm(2): Frobenius via Dlange/Dlassq */
double orc_frobenius(const double* w, int32_t n);
/* matrix.Scale(1/matrix.Norm(2), matrix): HopfieldNetwork.go:260; returns the norm */
double orc_normalize(orc_net* net);
/* the same with the Dlassq variant chosen: 0 classic scaled recurrence, 1 the LAPACK-3.10 rewrite (three accumulators;
 * what orc_frobenius / orc_normalize use); see hopfield_oracle.c */
double orc_frobenius_variant(const double* w, int32_t n, int32_t variant);
double orc_normalize_variant(orc_net* net, int32_t variant);

/* noiseapplication/NoiseApplication.go:51-105 on one f64 state, in place */
void orc_apply_noise(orc_rng* rng, int32_t method, double scale, double* state, int32_t n);

/* LearnStateData row (datacollector/LearnStateData.go:14-19) */
typedef struct { int32_t epoch, index, stable; } orc_learn_row;

/* LearnStates: HopfieldNetwork.go:254-262 + LearningMethod.go:38-89.
 *   states [p][n] activated in place; rule/method/noise enums as the reference.
 *   rows/energies optional with capacity cap; returns number of rows produced.
 *   NOTE: does not touch net->targets (caller appends). */
int32_t orc_learn_states(orc_net* net, double* states, int32_t p, int32_t rule, int32_t method,
                         int32_t epochs, int32_t noise_method, double noise_scale, orc_rng* rng,
                         int32_t cap, orc_learn_row* rows, double* energies);

/* handleUniqueRelaxedState: datacollector/UniqueRelaxedStateHandler.go:41-73.
 * key = fmt.Sprint(EnergyProfile) <=> bit pattern of the n doubles.
 * Processes probes in index order; writes attractor_id[B], and the table
 * (first index, hits) in first-seen order. returns n_unique. */
int32_t orc_unique_table(const double* energies, int32_t n_states, int32_t n, int32_t* attractor_id,
                         int32_t* first, int32_t* hits);

/* states/StateGenerator.go:44-52 + :65-74: Uniform(min,max) then activation */
void orc_generate_states(orc_rng* rng, int32_t domain, int32_t n, int32_t count, double mn, double mx,
                         double* out);

/* ---- helpers shared with the tests ---- */
void orc_pack(int32_t n, int32_t count, const double* f64, uint64_t* packed);
void orc_unpack(int32_t n, int32_t domain, int32_t count, const uint64_t* packed, double* f64);

#ifdef __cplusplus
}
#endif
#endif
