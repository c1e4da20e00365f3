/*
 * hopfield_b200.h — C ABI of libhopfield_b200.so
 *
 * The drop-in boundary for the one data-parallel hot path of
 * hmcalister/Hopfield-Network-Go: batched relaxation of probe states, the
 * per-state energy profile + stability check, and the Hebbian / Delta
 * learning-rule weight updates, implemented as hand-written sm_100a CUDA.
 *
 * The reference has no FFI seam today (pure Go calling gonum); the seam this
 * header introduces sits directly underneath the exported Go API of the
 * `hopfieldnetwork` package.  Each entry point below cites the reference
 * function (file:line under the reference tree) whose body it replaces; the
 * cgo binding a maintainer would add is shown in INTEGRATION.md and in
 * hopfield_network_go_b200/go/.
 *
 * Conventions
 *  - plain C: pointers + sizes only, no C++/torch types.
 *  - every function returns an int status: HN_OK (0) or a negative HN_E_*.
 *    hn_last_error() returns a human readable message for the last failure.
 *  - the caller owns every host buffer; the library copies in/out during the
 *    call and never retains a host pointer after returning (cgo rule).
 *  - there is NO CPU fallback: hn_create fails with HN_E_NO_DEVICE when no
 *    sm_100 device is usable.
 *  - a handle is single-caller (thread-compatible, not thread-safe).
 *
 * State formats
 *  - "packed": one bit per unit, 64 units per little-endian uint64 word,
 *    unit j of a state lives in bit (j % 64) of word (j / 64); a state is
 *    hn_words(N) = ceil(N/64) words; padding bits must be 0.
 *    bit 1 <-> +1 (bipolar) or 1 (binary); bit 0 <-> -1 (bipolar) or 0 (binary).
 *    This is exactly the image of the reference activation function
 *    (domain/BipolarDomainManager.go:16-22, domain/BinaryDomainManager.go:29-35):
 *    x <= 0 -> low value, else high value.
 *  - "f64": row-major float64, N per state, as in gonum mat.VecDense; on input
 *    the activation function is applied (that is what the reference does
 *    before any use of a state: HopfieldNetwork.go:255-257, states/StateGenerator.go:50).
 *
 * Randomness
 *  The reference draws unit-update orders and learning noise from a
 *  time-seeded RNG (HopfieldNetworkBuilder.go:231-232).  Here every RNG-derived
 *  quantity is an explicit INPUT generated host-side (permutation pools,
 *  noised target copies), so trajectories are reproducible.  hn_rng_* below is
 *  a host-side restatement of the reference's generator (golang.org/x/exp/rand
 *  PCG source + Shuffle/Float64/NormFloat64) for hosts that are not Go.
 */
#ifndef HOPFIELD_B200_H
#define HOPFIELD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------ */
#define HN_OK              0
#define HN_E_INVALID      -1  /* bad argument / configuration              */
#define HN_E_NO_DEVICE    -2  /* no usable sm_100 CUDA device              */
#define HN_E_CUDA         -3  /* CUDA runtime error                        */
#define HN_E_OOM          -4  /* device or host allocation failed          */
#define HN_E_POOL         -5  /* permutation pool exhausted                */
#define HN_E_NCCL         -6  /* NCCL error / NCCL not loadable            */
#define HN_E_STATE        -7  /* call not valid in the handle's state      */

/* ---- enums: numeric values equal the reference's iota values ----------- */
/* domain/DomainEnum.go:6-9 */
#define HN_DOMAIN_BIPOLAR 0
#define HN_DOMAIN_BINARY  1
/* LearningRule.go:32-39 */
#define HN_RULE_HEBBIAN                      0
#define HN_RULE_BIPOLAR_MAPPED_HEBBIAN       1
#define HN_RULE_DELTA                        2
#define HN_RULE_BIPOLAR_MAPPED_DELTA         3
#define HN_RULE_THERMAL_DELTA                4
#define HN_RULE_BIPOLAR_MAPPED_THERMAL_DELTA 5
/* LearningMethod.go:22-25 */
#define HN_METHOD_FULL_SET        0
#define HN_METHOD_ITERATIVE_BATCH 1
/* noiseapplication/NoiseApplication.go:24-36 */
#define HN_NOISE_NONE                        0
#define HN_NOISE_MAXIMAL_INVERSION           1
#define HN_NOISE_RANDOM_SUBMAXIMAL_INVERSION 2
#define HN_NOISE_GAUSSIAN                    3

/* ---- configuration: mirrors the HopfieldNetworkBuilder fields ----------
 * (HopfieldNetworkBuilder.go:16-34; defaults :40-55; Build checks :214-229) */
typedef struct hn_config {
    int32_t dimension;                 /* SetNetworkDimension; must be 1..65535              */
    int32_t domain;                    /* SetNetworkDomain; HN_DOMAIN_*                       */
    int32_t force_symmetric;           /* SetForceSymmetric (default 1)                       */
    int32_t force_zero_diagonal;       /* SetForceZeroDiagonal (default 1)                    */
    int32_t max_relax_unstable_units;  /* SetMaximumRelaxationUnstableUnits (default 0)       */
    int32_t max_relax_iterations;      /* SetMaximumRelaxationIterations (default 100)        */
    double  learning_rate;             /* SetLearningRate (default 1.0)                       */
    int32_t device;                    /* CUDA device ordinal (additive extension)            */
    int32_t column_stream;             /* HN_COLUMNS_AUTO (default), _F64, _F32 or _I16, see below */
} hn_config;

/* Column stream of the relaxation kernel (additive extension; results are bit-identical in every mode).
 * A flip of unit k adds +-delta*W[:,k] to the resident float64 fields.
 *   HN_COLUMNS_F64  the column is read from the float64 matrix (8N bytes per flip).
 *   HN_COLUMNS_F32  the column is read from a float32-rounded shadow copy of W (4N bytes per flip; the shadow of
 *                   a 4096-unit network fits the L2).  The fields stay float64; the extra rounding error is
 *                   bounded by flips * delta * max|W| * 2^-24 and added to the decision margin tau, and every
 *                   decision inside the margin is re-evaluated exactly on the float64 matrix in the reference's
 *                   summation order, so final states, step counts, flags and distances do not change.
 *   HN_COLUMNS_I16  exact integer stream for weights learned in this handle (Hebbian / Delta from the zero matrix,
 *                   learning rate 1): then W = c * Wraw with K = m * Wraw integer for an m in {0.5, 1, 2, 4}, and the
 *                   kernel keeps the EXACT int32 fields g = K s, adding int16 columns (2N bytes per flip).  sign(g) is the sign of the
 *                   reference's field wherever g != 0 (its rounding error is orders of magnitude below c); g == 0 is
 *                   an exact tie of the true field and is re-evaluated on the float64 matrix in the reference's
 *                   summation order.  Falls back to HN_COLUMNS_F64 whenever the integer structure is not known
 *                   (hn_set_weights, further learning after normalisation, |K| > 32767).                     */
#define HN_COLUMNS_F64 0
#define HN_COLUMNS_F32 1
#define HN_COLUMNS_I16 2
/*   HN_COLUMNS_AUTO the default: HN_COLUMNS_I16 whenever the integer structure of the current weights is known and
 *                   passes its exactness checks, HN_COLUMNS_F64 otherwise (random initial matrix, hn_set_weights,
 *                   thermal rules, non-dyadic learning rate).  Results are bit-identical either way;
 *                   hn_integer_stream_active tells which one the next relaxation uses.                           */
#define HN_COLUMNS_AUTO 3

typedef struct hn_handle_s* hn_handle;

/* Words per packed state. */
static inline int32_t hn_words(int32_t dimension) { return (dimension + 63) / 64; }

/* Fill `cfg` with the builder defaults (HopfieldNetworkBuilder.go:40-55). */
void hn_default_config(hn_config* cfg);

int  hn_abi_version(void);

/* Message for the most recent failing call on this thread ("" if none). */
const char* hn_last_error(void);

/* ---- lifetime ----------------------------------------------------------
 * Replaces HopfieldNetworkBuilder.Build (HopfieldNetworkBuilder.go:214-279):
 * allocates the zero N x N float64 weight matrix resident in HBM.  Random
 * Gaussian init (:235-246) is done by the host through hn_set_weights. */
int hn_create(const hn_config* cfg, hn_handle* out);
int hn_destroy(hn_handle h);

/* ---- weights -----------------------------------------------------------
 * GetMatrix (HopfieldNetwork.go:93-95) hands out the live gonum matrix; with
 * W resident in HBM the binding copies.  Row-major float64, N*N.            */
int hn_set_weights(hn_handle h, const double* w_host);
int hn_get_weights(hn_handle h, double* w_host);
/* Declares the integer structure of the resident weights for HN_COLUMNS_I16: the current W equals c * w_unnormalized
 * for some c > 0 (e.g. a host that saved both the normalised matrix and the pre-normalisation matrix and resumes).
 * Checked: 0.5, 1, 2 or 4 times w_unnormalized is integer valued within int16, and it is proportional to W.  Row-major float64, N*N.      */
int hn_set_integer_structure(hn_handle h, const double* w_unnormalized_host);
/* How hn_normalize_frobenius / hn_learn_states compute ||W||_F for quarter-integer weights (every Hebbian / Delta matrix
 * learned from zero).  mat.Norm(W, 2) is Dlange('F') over gonum's Dlassq (HopfieldNetwork.go:260), and gonum replaced
 * its classic Dlassq by a port of the LAPACK-3.10 routine (believed: before v0.11, so in the pinned v0.12.0); this
 * cannot be checked without the module sources, and the two differ in the last digits of the norm (3e-14 relative at
 * N=100) and hence in the last bits of W.  Both are reproduced BIT FOR BIT for such weights (oracle:
 * orc_frobenius_variant):
 *   HN_NORM_DLASSQ_LAPACK310 (default, what oracle/ assumes) three accumulators: for these weights an exact sum of
 *                            squares, computed on the device
 *   HN_NORM_DLASSQ_CLASSIC   the sequential scale/ssq recurrence (every addition rounds), run on the host
 * Weights that are not quarter-integers use a deterministic parallel sum under both settings (<= 1e-9 relative). */
#define HN_NORM_DLASSQ_CLASSIC   0
#define HN_NORM_DLASSQ_LAPACK310 1
int hn_set_norm_algorithm(hn_handle h, int32_t algorithm);
/* 1 if the next relaxation will use the exact integer stream, else 0 */
int hn_integer_stream_active(hn_handle h, int32_t* active);

/* ---- target (learned) states -------------------------------------------
 * network.targetStates (HopfieldNetwork.go:43, appended at :258).  Distances
 * in relaxation results are measured against this list (:411,:432).
 * hn_learn_* append automatically; hn_set_targets replaces the list.        */
int hn_set_targets(hn_handle h, const uint64_t* targets_packed, int32_t n_targets);
int hn_num_targets(hn_handle h, int32_t* n_targets);
int hn_get_targets(hn_handle h, uint64_t* targets_packed);

/* ---- learning ----------------------------------------------------------
 * One application of the Hebbian rule over `n` states
 * (LearningRule.go:64-75): W += lr * sum_mu x x^T, then enforceConstraints
 * (HopfieldNetwork.go:55-67).  Does NOT append to the target list and does
 * NOT normalise; hn_learn_states does the whole LearnStates.                */
int hn_hebbian_epoch(hn_handle h, const uint64_t* states_packed, int32_t n);

/* One application of the Delta rule (LearningRule.go:90-114).
 *   states_packed : the n target states t_mu
 *   noised_packed : the n copies after learning noise + activation
 *                   (LearningRule.go:99-102), generated host-side
 *   perms         : n permutations of 0..N-1 (uint16), the fresh shuffled unit
 *                   order UpdateState draws for each copy (HopfieldNetwork.go:281-282)
 *   relaxed_out   : optional (may be NULL) n packed states after the one sweep
 * Device work: fields of the noised copies (FP64 tensor-core GEMM), one
 * asynchronous sweep per copy, dW = sum_mu 0.5 (t - r) t^T (GEMM),
 * [NCCL allreduce of dW when a communicator is attached], W += lr dW,
 * enforceConstraints.                                                       */
int hn_delta_epoch(hn_handle h, const uint64_t* states_packed, const uint64_t* noised_packed,
                   const uint16_t* perms, int32_t n, uint64_t* relaxed_out);

/* One application of the thermal Delta rule (LearningRule.go:129-158): as hn_delta_epoch, but every target's
 * rank-one term is weighted by exp(-||W t||_2 / (1 + ||W||_F)) and the 0.5 factor is absent.  The weights are
 * no longer integers: dW matches the reference within rounding of the summation order (1e-9 relative).       */
int hn_thermal_delta_epoch(hn_handle h, const uint64_t* states_packed, const uint64_t* noised_packed,
                           const uint16_t* perms, int32_t n, uint64_t* relaxed_out);

/* One application of ANY of the six learning rules (LearningRule.go:32-39, getLearningRule :50-61) with the RNG-derived
 * inputs explicit — what a host that owns the reference's own generator (the Go package: golang.org/x/exp/rand) calls
 * once per epoch so that the random stream stays its own.  The bipolar-mapped rules (:78-87, :117-126, :161-170) see
 * the targets re-activated by the manager the reference names (mapped Hebbian: the BINARY manager, as written there);
 * noised copies and relaxed copies always carry the network domain's values.  noised_packed / perms are ignored by
 * the Hebbian rules (may be NULL).  Same device work and all-reduce behaviour as hn_hebbian_epoch / hn_delta_epoch /
 * hn_thermal_delta_epoch, which are the rule-0 / 2 / 4 cases.                                                     */
int hn_rule_epoch(hn_handle h, int32_t rule, const uint64_t* states_packed, const uint64_t* noised_packed,
                  const uint16_t* perms, int32_t n, uint64_t* relaxed_out);

/* enforceConstraints (HopfieldNetwork.go:55-67) on the resident W. */
int hn_enforce_constraints(hn_handle h);

/* W <- W * (1 / ||W||_F)  (HopfieldNetwork.go:260; gonum Norm(2) of a matrix
 * is the Frobenius norm).  Returns the norm that was divided out.           */
int hn_normalize_frobenius(hn_handle h, double* norm_out);

/* Per-state diagnostics used by the learning methods (LearningMethod.go:45-58)
 * and by main.go:191-204:  AllUnitEnergies (HopfieldNetwork.go:198-200 ->
 * domain/BipolarDomainManager.go:39-46, BinaryDomainManager.go:54-63),
 * StateIsStable (:214-225), StateEnergy (:171-173).
 *   energies_out   : optional [n][N] float64
 *   unstable_out   : optional [n] int32, #units with energy > 0
 *   stable_out     : optional [n] uint8, unstable <= max_relax_unstable_units
 *   state_energy_out: optional [n] float64                                   */
int hn_energy_profiles(hn_handle h, const uint64_t* states_packed, int32_t n,
                       double* energies_out, int32_t* unstable_out, uint8_t* stable_out,
                       double* state_energy_out);

/* UnitEnergy (HopfieldNetwork.go:185-187) for unit `unit_index` of each of n states: the reference's SCALAR loop,
 * which is not AllUnitEnergies(state)[i]:
 *   bipolar (domain/BipolarDomainManager.go:29-37): sum_j ((-0.5 * W_ij) * s_i) * s_j, accumulated in index order;
 *   binary  (domain/BinaryDomainManager.go:43-52):  sum_j (((-0.5 * W_ij) * s_i) * (2 s_j - 1) - 1)  — the reference
 *   subtracts 1 PER TERM (so the value is offset by -N); reproduced as written.
 *   energy_out : [n] float64                                                                                      */
int hn_unit_energy(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t unit_index, double* energy_out);

/* Whole LearnStates (HopfieldNetwork.go:254-262) with the learning-method
 * loop of LearningMethod.go:38-89 run by the library, drawing noise and
 * permutations from `rng` (see hn_rng_*) in the reference's draw order.
 *   learn_epoch_out / learn_index_out / learn_stable_out : optional arrays of
 *     capacity `learn_cap` receiving one LearnStateData row each
 *     (datacollector/LearnStateData.go:14-19) ; learn_energy_out optional
 *     [learn_cap][N].  *n_learn_rows receives the number of rows produced
 *     (rows beyond learn_cap are counted but not stored).
 * Appends `states` to the target list and normalises W.                      */
struct hn_rng_s;
typedef struct hn_rng_s* hn_rng;
int hn_learn_states(hn_handle h, const uint64_t* states_packed, int32_t n,
                    int32_t rule, int32_t method, int32_t epochs,
                    int32_t noise_method, double noise_scale, hn_rng rng,
                    int32_t learn_cap, int32_t* learn_epoch_out, int32_t* learn_index_out,
                    uint8_t* learn_stable_out, double* learn_energy_out,
                    int32_t* n_learn_rows);

/* LearnStates with the TARGETS SHARDED over the ranks of the attached communicator (config 5).
 * Every rank passes the same n states and an identically seeded rng and names its shard [shard_begin, shard_end):
 *   - noise and update orders are drawn for ALL n targets in the reference's order (LearningRule.go:98-104), so the
 *     random stream — and therefore W — is bit-identical to the single-process hn_learn_states;
 *   - each epoch's dW is all-reduced inside the library (exact integers: order independent) and every rank applies the
 *     same update, so the replicas of W stay bitwise equal;
 *   - the all-stable early exit (LearningMethod.go:56-58) is ncclAllReduce(min) of the per-rank flags: every rank runs
 *     the same number of epochs (a rank whose shard of an iterative-batch prefix is empty contributes a zero dW);
 *   - the LearnStateData rows returned are those of this rank's targets (TargetStateIndex is global); all n states
 *     are appended to the target list of every replica.
 * Without a communicator only the whole range [0, n) is accepted (then it IS hn_learn_states).
 * hn_learn_states itself returns HN_E_STATE while a communicator is attached.                                    */
int hn_learn_states_sharded(hn_handle h, const uint64_t* states_packed, int32_t n,
                            int32_t shard_begin, int32_t shard_end,
                            int32_t rule, int32_t method, int32_t epochs,
                            int32_t noise_method, double noise_scale, hn_rng rng,
                            int32_t learn_cap, int32_t* learn_epoch_out, int32_t* learn_index_out,
                            uint8_t* learn_stable_out, double* learn_energy_out,
                            int32_t* n_learn_rows);

/* ---- relaxation --------------------------------------------------------
 * flags for hn_relax_batch */
#define HN_RELAX_SERIAL_STREAM 0x1  /* emulate threads=1: probe p starts at pool row
                                       K_p = sum_{q<p} sweeps(q); offsets ignored */
#define HN_RELAX_DEDUP         0x2  /* fill attractor ids + unique table                 */
#define HN_RELAX_HISTORY       0x4  /* capture per-step packed state history             */

typedef struct hn_relax_out {
    /* all pointers are host buffers owned by the caller; any may be NULL */
    uint64_t* final_states;   /* [B][hn_words(N)] packed final state (StateHistory last, main.go:232) */
    int32_t*  num_steps;      /* [B] NumSteps = len(StateHistory) (main.go:231)                      */
    uint8_t*  stable;         /* [B] RelaxationResult.Stable                                         */
    int32_t*  distances;      /* [B][P] Manhattan-with-inversion distance to every target
                                 (distancemeasure/DistanceMeasure.go:28-42); exact small integers   */
    double*   energies;       /* [B][N] EnergyProfile of the final state (main.go:234)               */
    int32_t*  flips;          /* [B] counter: units flipped                                          */
    int32_t*  margin_hits;    /* [B] counter: decisions taken by exact-order recomputation because
                                 |h| fell below the margin tau                                       */
    /* unique-attractor table (datacollector/UniqueRelaxedStateHandler.go:41-73), HN_RELAX_DEDUP */
    int32_t*  attractor_id;   /* [B] index into the unique table                                     */
    int32_t*  unique_first;   /* [unique_cap] StateIndex of first occurrence, in first-seen order    */
    int32_t*  unique_hits;    /* [unique_cap] Hits                                                   */
    int32_t   unique_cap;
    int32_t   n_unique;       /* out */
    /* history (HN_RELAX_HISTORY): packed state after every step incl. step 0 */
    uint64_t* history;        /* [B][max_relax_iterations+1][hn_words(N)]                            */
    /* 128-bit key the unique table was grouped on (HN_RELAX_DEDUP): the exact-profile key where the library knows the
       profile exactly, else the hash of the sign-canonical final state; lets a host merge per-GPU tables without
       shipping the states (hn_merge_unique; hn_merge_unique_device compares the states too)            */
    uint64_t* state_hash;     /* [B][2]                                                              */
    /* totals (out) */
    int64_t   total_flips;
    int64_t   total_sweeps;
    int64_t   total_margin_hits;
    double    device_ms;      /* device time of the relaxation kernels, CUDA events                 */
} hn_relax_out;

/* ConcurrentRelaxStates / RelaxState (HopfieldNetwork.go:302-359, 371-442,
 * 461-490): relax B probe states until stable or max_relax_iterations sweeps.
 *   probes_packed : [B][hn_words(N)]
 *   perm_pool     : [n_pool][N] uint16; row r is the unit visiting order of one
 *                   sweep (the state of the reference's shuffled unitIndices
 *                   list, HopfieldNetwork.go:374,393)
 *   offsets       : [B] or NULL(=all 0); probe p uses pool rows offsets[p]+t
 *                   for sweep t = 0,1,...                                      */
int hn_relax_batch(hn_handle h, const uint64_t* probes_packed, int32_t n_probes,
                   const uint16_t* perm_pool, int32_t n_pool, const int32_t* offsets,
                   uint32_t flags, hn_relax_out* out);

/* UpdateState (HopfieldNetwork.go:280-291): one sweep per state in the given
 * fresh order, no stability test; states updated in place.                   */
int hn_update_states(hn_handle h, uint64_t* states_packed, int32_t n, const uint16_t* perms);

/* ---- device-resident bulk path (throughput / benchmarking) -------------
 * Upload once, relax many times without host traffic; see bench.py.          */
int hn_upload_perm_pool(hn_handle h, const uint16_t* perm_pool, int32_t n_pool);
int hn_upload_probes(hn_handle h, const uint64_t* probes_packed, int32_t n_probes);
/* states.StateGenerator on the device (states/StateGenerator.go:44-52, :65-74): `count` probe states are
 * generated straight into the resident probe buffer — unit i of state s from draw s*N+i of the generator's stream
 * (Uniform{min,max}.Rand() then the domain activation), bit for bit what hn_rng_generate_states returns — and the host
 * generator is advanced by count*N draws, so the stream continues as if the host had drawn them.  No probe upload. */
int hn_generate_probes(hn_handle h, hn_rng r, int32_t count, double rand_min, double rand_max);
/* Copy the first n resident probes (uploaded or generated) back to the host, packed. */
int hn_download_probes(hn_handle h, uint64_t* probes_packed_out, int32_t n_probes);
/* Relax the uploaded probes with the uploaded pool (offsets all 0). Results
 * stay on the device; *device_ms gets the CUDA-event time of the whole call. */
int hn_relax_resident(hn_handle h, uint32_t flags, double* device_ms);
int hn_download_results(hn_handle h, hn_relax_out* out);

/* Timing breakdown of the last relaxation (CUDA events, ms) and kernel count */
typedef struct hn_relax_timing {
    double fields_ms;   /* initial fields GEMM (0 when the dense sweeps supplied the fields) */
    double relax_ms;    /* relaxation kernel    */
    double post_ms;     /* energies / dedup     */
    int32_t launches;   /* kernels launched     */
    int32_t dense_sweeps; /* sweeps taken by the dense lockstep kernel (tcgen05 path) before the hand-over */
    double dense_ms;    /* those sweeps, with the fields contraction and stability scan after each of them */
    double dense_kernel_ms;  /* ... of which the lockstep sweep kernel itself (tcgen05 block contractions + decisions) */
    double dense_gemm_ms;    /* ... the exact fields contraction after every sweep (igemm_tc05)                        */
    double dense_scan_ms;    /* ... the stability scan after every sweep                                               */
    int64_t dense_flips;     /* units flipped by the dense sweeps (the rest of total_flips belongs to the sparse kernel) */
} hn_relax_timing;
int hn_last_relax_timing(hn_handle h, hn_relax_timing* t);

/* Unique-attractor table, exactness evidence (lifetime totals of the handle).  Equal sign-canonical states are merged
 * by comparing the packed states; DIFFERENT states are merged only when their exact-profile keys coincide AND their
 * float64 energy profiles, recomputed in the reference's summation order, are equal bit for bit
 * (UniqueRelaxedStateHandler.go:44 keys on the printed profile).  verified_merges counts those comparisons,
 * refuted_merges the ones that failed (a 128-bit key collision, or integer profiles that coincide where the rounded
 * float profiles do not) and were separated again.                                                              */
int hn_unique_table_stats(hn_handle h, int64_t* verified_merges, int64_t* refuted_merges);

/* Diagnostic: the exact integer fields (M s) of n states as the integer stream's contraction kernel computes them
 * (tcgen05 / TMEM / TMA kernel on the byte image: *kernel_used = 1; mma.sync kernel: 0).  M is the int16 image of the
 * weights (2K with the self term on the diagonal).  fields_out: [n][N] int32.  HN_E_STATE unless the integer stream is
 * active and W is symmetric.                                                                                    */
int hn_debug_integer_fields(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t* fields_out, int32_t* kernel_used);

/* Device-time breakdown (CUDA events, ms) of the last hn_delta_epoch / hn_thermal_delta_epoch (first five fields)
 * and of the last hn_energy_profiles (last two): where a Delta-rule epoch spends its time.                        */
typedef struct hn_learn_timing {
    double sweep_fields_ms;     /* fields of the noised copies: FP64 tensor-core GEMM, 2 N^2 P flop          */
    double sweep_ms;            /* the P one-sweep relaxations (UpdateState)                                  */
    double dw_gemm_ms;          /* dW = sum 0.5 (t - r) t^T: FP64 tensor-core GEMM, 2 N^2 P flop              */
    double allreduce_ms;        /* NCCL all-reduce of dW (0 without a communicator)                           */
    double update_ms;           /* W += lr dW, zero diagonal, symmetrise                                      */
    double energy_gemm_ms;      /* fields of the states given to hn_energy_profiles: 2 N^2 n flop             */
    double energy_epilogue_ms;  /* energies, unstable counts, stability flags                                 */
    int32_t targets;            /* states in the epoch                                                        */
    int32_t reserved;
    int64_t sweep_flips;        /* units flipped by the P sweeps (one weight column streamed per flip)        */
} hn_learn_timing;
int hn_last_learn_timing(hn_handle h, hn_learn_timing* t);

/* Decision margin tau = tau_base + tau_per_flip * flips currently in force
 * (DESIGN.md "margin").                                                      */
int hn_get_margin(hn_handle h, double* tau_base, double* tau_per_flip);

/* ---- conversions -------------------------------------------------------
 * f64 rows -> packed with the domain activation function applied. Host-side
 * helpers for hosts that hold gonum-style float64 state vectors.             */
/* Merge of per-GPU unique-attractor tables on the host (probes sharded over GPUs): `n` entries, each the
 * first occurrence of an attractor on some GPU with its GLOBAL probe index `first`, its hit count and the 128-bit hash
 * of its canonical final state (hn_relax_out.state_hash of that probe).  Entries with equal hashes are one attractor:
 * first = minimum, hits = sum; the merged table comes back in ascending first index — the order and the counts
 * handleUniqueRelaxedState (datacollector/UniqueRelaxedStateHandler.go:41-66) produces over the whole probe set.
 * rep_out[j] = index of the entry that supplied first_out[j].  Outputs hold up to n entries.  Pure host code.   */
int hn_merge_unique(int64_t n, const int64_t* first, const int32_t* hits, const uint64_t* hash /* [n][2] */,
                    int64_t* first_out, int64_t* hits_out, int64_t* rep_out, int64_t* n_groups_out);

/* Merge of per-GPU unique tables ON A GPU, exact (the probes are sharded, the tables meet on one rank).
 * hn_export_unique_device gathers this handle's table (after a relaxation with HN_RELAX_DEDUP) into caller-owned
 * DEVICE arrays — per unique attractor: global first index (index_offset + local), hits, the four key words, the packed
 * final state and its zero-field mark — so that the entries can travel GPU to GPU (NCCL gather) instead of through the
 * host.  Called with d_first == NULL it only reports *n_unique / *profile_keyed.
 * hn_merge_unique_device takes the concatenation of such exports in ascending rank order (DEVICE pointers; d_keys is
 * scratch and may be modified) and runs the same two-level table over it as a single-GPU batch does: equal
 * sign-canonical states are merged by comparing the states, different states only after the float-profile comparison.
 * Outputs (HOST, capacity n): the merged table in first-seen order — first index, summed hits, and rep_out[j] = which
 * input entry supplied first_out[j] (may be NULL).  The handle's own unique table is consumed by the call.        */
int hn_export_unique_device(hn_handle h, int64_t index_offset, int64_t* d_first, int32_t* d_hits, uint64_t* d_keys,
                            uint64_t* d_states, uint8_t* d_zero_field, int32_t* n_unique, int32_t* profile_keyed);
int hn_merge_unique_device(hn_handle h, int64_t n, const int64_t* d_first, const int32_t* d_hits, uint64_t* d_keys,
                           const uint64_t* d_states, const uint8_t* d_zero_field, int32_t profile_keyed,
                           int64_t* first_out, int64_t* hits_out, int64_t* rep_out, int64_t* n_groups_out);

/* Page-locked host staging buffers (additive extension): every host pointer of this API may be pageable, but
 * probes / results in buffers from hn_host_alloc move at PCIe speed and without a staging copy (a Go binding
 * keeps its flattened []float64 / []uint64 slices in such buffers; reference: the [] *mat.VecDense results of
 * ConcurrentRelaxStates, HopfieldNetwork.go:461-490).  Freed with hn_host_free only.
 * hn_relax_batch goes further with page-locked arrays (from hn_host_alloc, cudaHostAlloc or cudaHostRegister; detected per
 * call): final_states and distances are WRITTEN BY THE RELAXATION KERNEL into the caller's memory as each probe
 * finishes, and — when the batch starts with dense lockstep sweeps — the probes are read where they lie, so the two
 * large transfers ride under the kernels instead of following them (config 3: end-to-end 0.993 of the device rate on
 * one GPU, 0.973 on eight).  The arrays are complete when the call returns, as with pageable memory.  HN_ZERO_COPY=0
 * (environment) selects the copy pipeline for measurements.                                                     */
int hn_host_alloc(void** ptr_out, size_t bytes);
int hn_host_free(void* ptr);

int hn_pack_states_f64(int32_t dimension, int32_t n, const double* states_f64, uint64_t* packed);
int hn_unpack_states_f64(int32_t dimension, int32_t domain, int32_t n, const uint64_t* packed,
                         double* states_f64);

/* ---- multi-GPU (targets sharded for learning) --------------------------
 * One handle per GPU.  When a communicator is attached, hn_hebbian_epoch /
 * hn_delta_epoch treat their `states` as this rank's shard and allreduce dW
 * (ncclAllReduce, sum, float64, N*N) before the update, so every replica of W
 * stays bitwise identical (dW is integer valued: order independent).
 * NCCL is dlopen()ed on first use.                                           */
#define HN_NCCL_UNIQUE_ID_BYTES 128
int hn_comm_unique_id(uint8_t id[HN_NCCL_UNIQUE_ID_BYTES]);
int hn_comm_attach(hn_handle h, const uint8_t id[HN_NCCL_UNIQUE_ID_BYTES], int32_t rank, int32_t n_ranks);
int hn_comm_detach(hn_handle h);
/* broadcast W from `root` to all ranks of the attached communicator */
int hn_broadcast_weights(hn_handle h, int32_t root);

/* ---- host-side RNG: restatement of golang.org/x/exp/rand ----------------
 * (module golang.org/x/exp @ 642cacee5cc0, go.mod:12 — source absent from the
 * reference tree; algorithm restated from the published PCG XSL-RR 128/64
 * generator and the package's documented Shuffle/Float64/NormFloat64.)        */
int      hn_rng_create(uint64_t seed, hn_rng* out);        /* rand.New(rand.NewSource(seed)) */
int      hn_rng_destroy(hn_rng r);
uint64_t hn_rng_uint64(hn_rng r);                          /* Rand.Uint64   */
uint64_t hn_rng_uint64n(hn_rng r, uint64_t n);             /* Rand.Uint64n  */
double   hn_rng_float64(hn_rng r);                         /* Rand.Float64  */
double   hn_rng_norm_float64(hn_rng r);                    /* Rand.NormFloat64 */
/* hopfieldutils.ShuffleList (hopfieldutils/HopfieldUtils.go:35-39) on a uint16 list */
int      hn_rng_shuffle_u16(hn_rng r, uint16_t* list, int32_t n);
/* n_rows successive in-place shuffles of one persistent list that starts as
 * `list_io` (identity when start_identity != 0); row r of pool = list after
 * shuffle r+1.  This is the stream of orders one relaxation goroutine sees
 * (HopfieldNetwork.go:374,393).                                              */
int      hn_rng_perm_pool(hn_rng r, int32_t dimension, int32_t n_rows, int32_t start_identity,
                          uint16_t* list_io, uint16_t* pool_out);
/* learning noise on one f64 state copy, in place, drawing from r exactly as
 * noiseapplication/NoiseApplication.go:51-105 does.                          */
int      hn_rng_apply_noise(hn_rng r, int32_t noise_method, double noise_scale,
                            double* state_f64, int32_t dimension);

/* Skip `draws` Uint64 values of the stream in O(log draws) (128-bit LCG jump-ahead): the state after the call is the
 * state `draws` calls of hn_rng_uint64 would leave.  Used by hn_generate_probes; lets ranks of a multi-GPU job take
 * disjoint sections of ONE reference stream (states/StateGeneratorBuilder.go:115-123 seeds one generator).        */
int      hn_rng_advance(hn_rng r, uint64_t draws);

/* states.StateGenerator (states/StateGenerator.go:44-52, :65-74): `count` states whose elements are
 * distuv.Uniform{min,max}.Rand() = Float64()*(max-min)+min followed by the domain activation, packed. */
int      hn_rng_generate_states(hn_rng r, int32_t dimension, int32_t count, double rand_min, double rand_max,
                                uint64_t* packed_out);

#ifdef __cplusplus
}
#endif
#endif /* HOPFIELD_B200_H */
