/*
 * oracle/hopfield_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of the reference hot path; see hopfield_oracle.h.
 * Compile with -O2 -ffp-contract=off (no FMA contraction, no fast-math): every
 * floating-point operation below is meant to round exactly where the
 * reference's does.  PARITY UNPINNED (no reference tests / golden vectors).
 */
#include "hopfield_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- gonum f64.DotUnitary (amd64 SSE2) ---------------------------------
 * X7 = (a0,a1), X8 = (a2,a3) accumulate 4 elements per iteration with
 * MULPD/ADDPD; the 1-3 tail elements go through MULSD/ADDSD into X7's low
 * lane (a0); epilogue ADDPD X8,X7 then low+high.  Used by mat.Dot at
 * HopfieldNetwork.go:286,320,396. */
double orc_dot_unitary(const double* x, const double* y, int32_t n) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int32_t i = 0;
    for (; i + 4 <= n; i += 4) {
        a0 = a0 + x[i] * y[i];
        a1 = a1 + x[i + 1] * y[i + 1];
        a2 = a2 + x[i + 2] * y[i + 2];
        a3 = a3 + x[i + 3] * y[i + 3];
    }
    for (; i < n; i++) a0 = a0 + x[i] * y[i];
    return (a0 + a2) + (a1 + a3);
}

/* domain/BipolarDomainManager.go:16-22, BinaryDomainManager.go:29-35 */
double orc_activation_unit(int32_t domain, double x) {
    if (x <= 0.0) return domain == ORC_BINARY ? 0.0 : -1.0;
    return 1.0;
}
void orc_activation(int32_t domain, double* v, int32_t n) {
    for (int32_t i = 0; i < n; i++) v[i] = orc_activation_unit(domain, v[i]);
}

/* BipolarDomainManager.go:39-46: e = MulVec(W,s); e = e .* s; e = -0.5 * e
 * BinaryDomainManager.go:54-63: same with the second factor mapped 2s-1
 * (mapVectorToBipolar :8-13: mapped = s + 2*... AddScaledVec(neg1, 2.0, s) = -1 + 2*s). */
void orc_all_unit_energies(const orc_net* net, const double* state, double* e) {
    int32_t n = net->n;
    for (int32_t i = 0; i < n; i++) {
        double h = orc_dot_unitary(net->w + (size_t)i * n, state, n);
        double v = state[i];
        if (net->domain == ORC_BINARY) v = -1.0 + 2.0 * state[i];
        double t = h * v;
        e[i] = -0.5 * t;
    }
}

/* UnitEnergy scalar loops: BipolarDomainManager.go:29-37, BinaryDomainManager.go:43-52
 * (the Binary version subtracts 1 per term, as written in the reference). */
double orc_unit_energy(const orc_net* net, const double* state, int32_t i) {
    int32_t n = net->n;
    double energy = 0.0;
    const double* row = net->w + (size_t)i * n;
    if (net->domain == ORC_BINARY) {
        for (int32_t j = 0; j < n; j++) {
            double mj = -1.0 + 2.0 * state[j];
            energy += -0.5 * row[j] * state[i] * mj - 1;
        }
    } else {
        for (int32_t j = 0; j < n; j++) energy += -0.5 * row[j] * state[i] * state[j];
    }
    return energy;
}

double orc_state_energy(const orc_net* net, const double* state) {
    int32_t n = net->n;
    double* e = (double*)malloc(sizeof(double) * n);
    orc_all_unit_energies(net, state, e);
    double s = 0.0;
    for (int32_t i = 0; i < n; i++) s += e[i];
    free(e);
    return s;
}

/* HopfieldNetwork.go:214-225 */
int orc_state_is_stable(const orc_net* net, const double* state, int32_t* unstable_count) {
    int32_t n = net->n;
    double* e = (double*)malloc(sizeof(double) * n);
    orc_all_unit_energies(net, state, e);
    int32_t c = 0;
    for (int32_t i = 0; i < n; i++)
        if (e[i] > 0) c++;
    free(e);
    if (unstable_count) *unstable_count = c;
    return c <= net->max_unstable;
}

/* InvertState: Bipolar :24-27 (scale -1 then activation), Binary :37-41 (1 - v then activation) */
static void invert_state(int32_t domain, const double* a, double* out, int32_t n) {
    for (int32_t i = 0; i < n; i++) out[i] = domain == ORC_BINARY ? 1.0 + -1.0 * a[i] : -1.0 * a[i];
    orc_activation(domain, out, n);
}

/* distancemeasure/DistanceMeasure.go:28-42 for every target (:12-18):
 * d1 = ||t - s||_1, d2 = ||inv(t) - s||_1, min. (Dasum: sequential sum of |x|;
 * all terms are small integers so the order is immaterial.) */
void orc_distances_to_targets(const orc_net* net, const double* state, double* out) {
    int32_t n = net->n;
    double* inv = (double*)malloc(sizeof(double) * n);
    for (int32_t t = 0; t < net->n_targets; t++) {
        const double* a = net->targets + (size_t)t * n;
        double d1 = 0.0, d2 = 0.0;
        for (int32_t i = 0; i < n; i++) d1 += fabs(a[i] - state[i]);
        invert_state(net->domain, a, inv, n);
        for (int32_t i = 0; i < n; i++) d2 += fabs(inv[i] - state[i]);
        out[t] = d1 < d2 ? d1 : d2;
    }
    free(inv);
}

/* Margin: bound on |incremental field - fresh DotUnitary field|.
 * R = max_i sum_k |W_ik|, u = 2^-53:
 *   tau_base = 2*(1.25 n + 4)*u*R, tau_per_flip = 2*u*R.    */
void orc_margin(const orc_net* net, double* tau_base, double* tau_per_flip) {
    int32_t n = net->n;
    double r = 0.0;
    for (int32_t i = 0; i < n; i++) {
        double s = 0.0;
        for (int32_t k = 0; k < n; k++) s += fabs(net->w[(size_t)i * n + k]);
        if (s > r) r = s;
    }
    const double u = 1.1102230246251565e-16;
    *tau_base = 2.0 * (1.25 * n + 4.0) * u * r;
    *tau_per_flip = 2.0 * u * r;
}

/* HopfieldNetwork.go:280-291 */
void orc_update_state(const orc_net* net, double* state, const uint16_t* perm) {
    int32_t n = net->n;
    for (int32_t i = 0; i < n; i++) {
        int32_t k = perm[i];
        double h = orc_dot_unitary(net->w + (size_t)k * n, state, n);
        state[k] = orc_activation_unit(net->domain, h);
    }
    orc_activation(net->domain, state, n);
}

/* One pass of AllUnitEnergies + count, with margin accounting. */
static int stable_check(const orc_net* net, const double* state, double* e, double tau,
                        int32_t* margin_hits) {
    int32_t n = net->n, c = 0;
    orc_all_unit_energies(net, state, e);
    for (int32_t i = 0; i < n; i++) {
        if (e[i] > 0) c++;
        if (margin_hits && fabs(e[i]) * 2.0 < tau) (*margin_hits)++;
    }
    return c <= net->max_unstable;
}

int orc_relax_state(const orc_net* net, double* state, const uint16_t* perm_pool, int32_t n_pool,
                    int64_t offset, orc_relax_info* info, double* distances, double* energies,
                    double* history, int faithful_cost) {
    int32_t n = net->n;
    double tau_base, tau_flip;
    orc_margin(net, &tau_base, &tau_flip);
    double* e = (double*)malloc(sizeof(double) * n);
    orc_relax_info li;
    li.stable = 0; li.num_steps = 0; li.flips = 0; li.margin_hits = 0; li.min_abs_h = INFINITY;
    if (history) memcpy(history, state, sizeof(double) * n); /* stateHistory[0], :387-390 */
    int rc = 0, done = 0;
    for (int32_t step = 1; step <= net->max_iter && !done; step++) { /* :392 */
        int64_t row = offset + (step - 1);
        if (row < 0 || row >= n_pool) { rc = -1; break; }
        const uint16_t* perm = perm_pool + (size_t)row * n; /* list after ShuffleList, :393 */
        for (int32_t i = 0; i < n; i++) { /* :394-399 */
            int32_t k = perm[i];
            double h = orc_dot_unitary(net->w + (size_t)k * n, state, n);
            double ah = fabs(h);
            if (ah < li.min_abs_h) li.min_abs_h = ah;
            if (ah < tau_base + tau_flip * li.flips) li.margin_hits++;
            double v = orc_activation_unit(net->domain, h);
            if (v != state[k]) li.flips++;
            state[k] = v;
        }
        orc_activation(net->domain, state, n); /* :400 */
        double tau = tau_base + tau_flip * li.flips;
        int st = stable_check(net, state, e, tau, &li.margin_hits); /* :403 */
        if (faithful_cost) {
            if (st) stable_check(net, state, e, tau, NULL);  /* :405 AllUnitEnergies */
            st = stable_check(net, state, e, tau, NULL);     /* :408 second StateIsStable */
        }
        if (history) memcpy(history + (size_t)step * n, state, sizeof(double) * n);
        li.num_steps = step + 1; /* len(stateHistory[:stepIndex+1]) */
        if (st) { li.stable = 1; done = 1; }
    }
    if (rc == 0 && !done) {
        /* :425-435 not stable after max_iter sweeps: history is the full maxIter+1 slots */
        li.stable = 0;
        li.num_steps = net->max_iter + 1;
    }
    if (rc == 0) {
        if (energies) orc_all_unit_energies(net, state, energies); /* energyHistory last */
        if (distances) orc_distances_to_targets(net, state, distances); /* :411,:432 */
    }
    if (info) *info = li;
    free(e);
    return rc;
}

typedef struct {
    const orc_net* net; double* states; int32_t n_states; const uint16_t* pool; int32_t n_pool;
    const int32_t* offsets; orc_relax_info* info; double* distances; double* energies;
    int faithful; int32_t next; int rc; pthread_mutex_t mu;
} batch_ctx;

static void* batch_worker(void* arg) {
    batch_ctx* c = (batch_ctx*)arg;
    int32_t n = c->net->n, p = c->net->n_targets;
    for (;;) {
        pthread_mutex_lock(&c->mu);
        int32_t i = c->next++;
        pthread_mutex_unlock(&c->mu);
        if (i >= c->n_states) break;
        int r = orc_relax_state(c->net, c->states + (size_t)i * n, c->pool, c->n_pool,
                                c->offsets ? c->offsets[i] : 0, c->info ? c->info + i : NULL,
                                c->distances ? c->distances + (size_t)i * p : NULL,
                                c->energies ? c->energies + (size_t)i * n : NULL, NULL, c->faithful);
        if (r) { pthread_mutex_lock(&c->mu); c->rc = r; pthread_mutex_unlock(&c->mu); }
    }
    return NULL;
}

int orc_relax_batch(const orc_net* net, double* states, int32_t n_states, const uint16_t* perm_pool,
                    int32_t n_pool, const int32_t* offsets, int serial_stream, int n_threads,
                    orc_relax_info* info, double* distances, double* energies, int faithful_cost) {
    int32_t n = net->n, p = net->n_targets;
    if (serial_stream) {
        /* threads=1: one goroutine, one persistent list, one RNG stream */
        int64_t k = 0;
        for (int32_t i = 0; i < n_states; i++) {
            orc_relax_info li;
            int r = orc_relax_state(net, states + (size_t)i * n, perm_pool, n_pool, k, &li,
                                    distances ? distances + (size_t)i * p : NULL,
                                    energies ? energies + (size_t)i * n : NULL, NULL, faithful_cost);
            if (r) return r;
            if (info) info[i] = li;
            k += li.num_steps - 1;
        }
        return 0;
    }
    batch_ctx c;
    c.net = net; c.states = states; c.n_states = n_states; c.pool = perm_pool; c.n_pool = n_pool;
    c.offsets = offsets; c.info = info; c.distances = distances; c.energies = energies;
    c.faithful = faithful_cost; c.next = 0; c.rc = 0;
    pthread_mutex_init(&c.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads == 1) {
        batch_worker(&c);
    } else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &c);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&c.mu);
    return c.rc;
}

/* HopfieldNetwork.go:55-67.  matrix.Add(matrix, matrix.T()) goes through
 * gonum's aliasing workspace, so every new_ij = W_ij + W_ji uses the old
 * values; then Scale(0.5). */
void orc_enforce_constraints(orc_net* net) {
    int32_t n = net->n;
    double* w = net->w;
    if (net->force_zero_diag)
        for (int32_t i = 0; i < n; i++) w[(size_t)i * n + i] = 0.0;
    if (net->force_symmetric) {
        for (int32_t i = 0; i < n; i++)
            for (int32_t j = i; j < n; j++) {
                double a = w[(size_t)i * n + j], b = w[(size_t)j * n + i];
                double sij = a + b, sji = b + a;
                w[(size_t)i * n + j] = 0.5 * sij;
                w[(size_t)j * n + i] = 0.5 * sji;
            }
    }
}

/* Dger: a_ij += (alpha*x_i)*y_j (multiply then add, unfused) */
static void rank_one(double* a, int32_t n, double alpha, const double* x, const double* y) {
    for (int32_t i = 0; i < n; i++) {
        double t = alpha * x[i];
        double* row = a + (size_t)i * n;
        for (int32_t j = 0; j < n; j++) row[j] = row[j] + t * y[j];
    }
}

/* updatedMatrix.Scale(lr) ; matrix.Add(matrix, updatedMatrix) ; enforceConstraints */
static void apply_update(orc_net* net, double* upd) {
    size_t nn = (size_t)net->n * net->n;
    for (size_t i = 0; i < nn; i++) upd[i] = net->lr * upd[i];
    for (size_t i = 0; i < nn; i++) net->w[i] = net->w[i] + upd[i];
    orc_enforce_constraints(net);
}

/* LearningRule.go:64-75 */
void orc_hebbian(orc_net* net, const double* states, int32_t p) {
    int32_t n = net->n;
    double* upd = (double*)calloc((size_t)n * n, sizeof(double));
    for (int32_t s = 0; s < p; s++) rank_one(upd, n, 1.0, states + (size_t)s * n, states + (size_t)s * n);
    apply_update(net, upd);
    free(upd);
}

/* LearningRule.go:90-114 */
void orc_delta(orc_net* net, const double* states, const double* noised, const uint16_t* perms,
               int32_t p, double* relaxed_out) {
    int32_t n = net->n;
    double* upd = (double*)calloc((size_t)n * n, sizeof(double));
    double* relaxed = (double*)malloc(sizeof(double) * (size_t)p * n);
    double* diff = (double*)malloc(sizeof(double) * n);
    memcpy(relaxed, noised, sizeof(double) * (size_t)p * n);
    for (int32_t s = 0; s < p; s++) /* :98-104 (noise + activation done by caller) */
        orc_update_state(net, relaxed + (size_t)s * n, perms + (size_t)s * n);
    for (int32_t s = 0; s < p; s++) { /* :106-109 */
        const double* t = states + (size_t)s * n;
        for (int32_t i = 0; i < n; i++) diff[i] = t[i] - relaxed[(size_t)s * n + i];
        rank_one(upd, n, 0.5, diff, t);
    }
    apply_update(net, upd); /* :111-113 */
    if (relaxed_out) memcpy(relaxed_out, relaxed, sizeof(double) * (size_t)p * n);
    free(diff); free(relaxed); free(upd);
}

static void dlassq(int32_t n, const double* x, double* scale, double* sumsq);

/* LearningRule.go:129-158.  mat.Norm(matrix, 2.0) = Frobenius (Dlange); MulVec = per-row DotUnitary (see header);
 * mat.Norm(vector, 2) = Dnrm2 = scaled sum of squares. */
void orc_thermal_delta(orc_net* net, const double* states, const double* noised, const uint16_t* perms,
                       int32_t p, double* relaxed_out) {
    int32_t n = net->n;
    double* upd = (double*)calloc((size_t)n * n, sizeof(double));
    double* relaxed = (double*)malloc(sizeof(double) * (size_t)p * n);
    double* diff = (double*)malloc(sizeof(double) * n);
    double* tv = (double*)malloc(sizeof(double) * n);
    double weight_factor = 1 / (1 + orc_frobenius(net->w, n)); /* :134 */
    memcpy(relaxed, noised, sizeof(double) * (size_t)p * n);
    for (int32_t s = 0; s < p; s++) /* :138-144 */
        orc_update_state(net, relaxed + (size_t)s * n, perms + (size_t)s * n);
    for (int32_t s = 0; s < p; s++) { /* :146-153 */
        const double* t = states + (size_t)s * n;
        for (int32_t i = 0; i < n; i++) diff[i] = t[i] - relaxed[(size_t)s * n + i];
        for (int32_t i = 0; i < n; i++) tv[i] = orc_dot_unitary(net->w + (size_t)i * n, t, n);
        double scale = 0.0, ssq = 1.0;
        dlassq(n, tv, &scale, &ssq);
        double nrm = scale * sqrt(ssq);
        double f = exp(-1.0 * weight_factor * nrm / (1.0));
        rank_one(upd, n, f, diff, t);
    }
    apply_update(net, upd); /* :155-157 */
    if (relaxed_out) memcpy(relaxed_out, relaxed, sizeof(double) * (size_t)p * n);
    free(tv); free(diff); free(relaxed); free(upd);
}

/* gonum Dlassq (classic scaled sum of squares) */
static void dlassq(int32_t n, const double* x, double* scale, double* sumsq) {
    for (int32_t i = 0; i < n; i++) {
        double absxi = fabs(x[i]);
        if (absxi > 0 || isnan(absxi)) {
            if (*scale < absxi) {
                double ratio = *scale / absxi;
                *sumsq = 1 + *sumsq * ratio * ratio;
                *scale = absxi;
            } else {
                double ratio = absxi / *scale;
                *sumsq += ratio * ratio;
            }
        }
    }
}
/* Dlassq as rewritten after LAPACK 3.10 (Blue's algorithm: three accumulators for big / medium / small magnitudes,
 * thresholds 2^486 / 2^-511, scalings 2^-538 / 2^537).  gonum replaced its classic Dlassq by a port of this routine at some
 * release (believed: before v0.11, so the pinned v0.12.0 of go.mod:5 has it); that cannot be checked here (no module
 * sources), so both variants are restated and the library implements both (hn_set_norm_algorithm).  For the medium range — every weight
 * matrix of this path — it is a plain sequential sum of squares per row plus the running total, scale = 1. */
static void dlassq_310(int32_t n, const double* x, double* scale, double* sumsq) {
    const double tsml = ldexp(1.0, -511), tbig = ldexp(1.0, 486), ssml = ldexp(1.0, 537), sbig = ldexp(1.0, -538);
    if (isnan(*scale) || isnan(*sumsq)) return;
    if (*sumsq == 0.0) *scale = 1.0;
    if (*scale == 0.0) { *scale = 1.0; *sumsq = 0.0; }
    if (n <= 0) return;
    int notbig = 1;
    double asml = 0.0, amed = 0.0, abig = 0.0;
    for (int32_t i = 0; i < n; i++) {
        double ax = fabs(x[i]);
        if (ax > tbig) { double t = ax * sbig; abig += t * t; notbig = 0; }
        else if (ax < tsml) { if (notbig) { double t = ax * ssml; asml += t * t; } }
        else amed += ax * ax;
    }
    if (*sumsq > 0.0) {
        double ax = *scale * sqrt(*sumsq);
        if (ax > tbig) { double t = *scale * sbig; abig += t * t * *sumsq; notbig = 0; }
        else if (ax < tsml) { if (notbig) { double t = *scale * ssml; asml += t * t * *sumsq; } }
        else amed += *scale * *scale * *sumsq;
    }
    if (abig > 0.0) {
        if (amed > 0.0 || isnan(amed)) abig += (amed * sbig) * sbig;
        *scale = 1.0 / sbig; *sumsq = abig;
    } else if (asml > 0.0) {
        if (amed > 0.0 || isnan(amed)) {
            double a = sqrt(amed), b = sqrt(asml) / ssml;
            double ymin = b > a ? a : b, ymax = b > a ? b : a;
            *scale = 1.0; *sumsq = ymax * ymax * (1.0 + (ymin / ymax) * (ymin / ymax));
        } else { *scale = 1.0 / ssml; *sumsq = asml; }
    } else { *scale = 1.0; *sumsq = amed; }
}
/* Dlange(Frobenius): per-row Dlassq, scale*sqrt(sum).  variant 0: classic Dlassq, 1: the LAPACK-3.10 rewrite */
double orc_frobenius_variant(const double* w, int32_t n, int32_t variant) {
    double scale = 0.0, sum = 1.0;
    for (int32_t i = 0; i < n; i++) {
        if (variant) dlassq_310(n, w + (size_t)i * n, &scale, &sum);
        else dlassq(n, w + (size_t)i * n, &scale, &sum);
    }
    return scale * sqrt(sum);
}
/* default: the LAPACK-3.10 rewrite — gonum's dlassq.go cites "Algorithm 978: Safe Scaling in the Level 1 BLAS" (Anderson
 * 2017) and was ported from LAPACK 3.10 (2021) before v0.11; recalled, not verifiable here: parity stays UNPINNED and
 * the classic variant remains selectable on both sides */
double orc_frobenius(const double* w, int32_t n) { return orc_frobenius_variant(w, n, 1); }
double orc_normalize_variant(orc_net* net, int32_t variant) {
    double nrm = orc_frobenius_variant(net->w, net->n, variant);
    double f = 1 / nrm;
    size_t nn = (size_t)net->n * net->n;
    for (size_t i = 0; i < nn; i++) net->w[i] = f * net->w[i];
    return nrm;
}
double orc_normalize(orc_net* net) {
    double nrm = orc_frobenius(net->w, net->n);
    double f = 1 / nrm;
    size_t nn = (size_t)net->n * net->n;
    for (size_t i = 0; i < nn; i++) net->w[i] = f * net->w[i];
    return nrm;
}

/* noiseapplication/NoiseApplication.go */
static void maximal_invert(orc_rng* rng, double* state, int32_t n, double ratio) { /* :64-75 */
    int32_t num = (int32_t)((double)n * ratio);
    int32_t m = n - 1; /* sliceIndices has length Len()-1: last unit never inverted */
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (m > 0 ? m : 1));
    for (int32_t i = 0; i < m; i++) idx[i] = i;
    orc_rng_shuffle_i32(rng, idx, m);
    for (int32_t i = 0; i < num && i < m; i++) state[idx[i]] = -1 * state[idx[i]];
    free(idx);
}
void orc_apply_noise(orc_rng* rng, int32_t method, double scale, double* state, int32_t n) {
    switch (method) {
    case 0: break; /* :51 */
    case 1: maximal_invert(rng, state, n, scale); break;
    case 2: { /* :87-90 */
        double sel = fmod(orc_rng_float64(rng), scale);
        maximal_invert(rng, state, n, sel);
        break;
    }
    case 3: /* :101-105 */
        for (int32_t i = 0; i < n; i++) state[i] += orc_rng_norm_float64(rng) * scale;
        break;
    default: break;
    }
}

/* One application of the chosen rule drawing from rng in the reference's
 * order: per target [noise draws][fresh identity list shuffled] (LearningRule.go:98-104). */
static void apply_rule(orc_net* net, const double* states, int32_t p, int32_t rule,
                       int32_t noise_method, double noise_scale, orc_rng* rng) {
    int32_t n = net->n;
    double* mapped = NULL;
    const double* st = states;
    if (rule == 1 || rule == 3 || rule == 5) {
        /* bipolarMappedHebbian re-activates with the BINARY manager (sic, LearningRule.go:80),
         * bipolarMappedDelta / bipolarMappedThermalDelta with the Bipolar manager (:119, :163) */
        mapped = (double*)malloc(sizeof(double) * (size_t)p * n);
        memcpy(mapped, states, sizeof(double) * (size_t)p * n);
        orc_activation(rule == 1 ? ORC_BINARY : ORC_BIPOLAR, mapped, p * n);
        st = mapped;
    }
    if (rule == 0 || rule == 1) {
        orc_hebbian(net, st, p);
    } else {
        double* noised = (double*)malloc(sizeof(double) * (size_t)p * n);
        uint16_t* perms = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)p * n);
        memcpy(noised, st, sizeof(double) * (size_t)p * n);
        for (int32_t s = 0; s < p; s++) {
            orc_apply_noise(rng, noise_method, noise_scale, noised + (size_t)s * n, n);
            orc_activation(net->domain, noised + (size_t)s * n, n);
            uint16_t* pr = perms + (size_t)s * n;
            for (int32_t i = 0; i < n; i++) pr[i] = (uint16_t)i; /* getUnitIndices :71-77 */
            orc_rng_shuffle_u16(rng, pr, n);
        }
        if (rule >= 4) orc_thermal_delta(net, st, noised, perms, p, NULL);
        else orc_delta(net, st, noised, perms, p, NULL);
        free(perms); free(noised);
    }
    free(mapped);
}

/* fullSetLearningMethod: LearningMethod.go:38-61 */
static int32_t full_set(orc_net* net, const double* states, int32_t p, int32_t rule, int32_t epochs,
                        int32_t noise_method, double noise_scale, orc_rng* rng, int32_t cap,
                        orc_learn_row* rows, double* energies, int32_t produced, int32_t epoch_offset,
                        int32_t* last_epoch) {
    int32_t n = net->n;
    double* e = (double*)malloc(sizeof(double) * n);
    for (int32_t epoch = 0; epoch < epochs; epoch++) {
        apply_rule(net, states, p, rule, noise_method, noise_scale, rng);
        int all_stable = 1;
        for (int32_t s = 0; s < p; s++) {
            const double* st = states + (size_t)s * n;
            orc_all_unit_energies(net, st, e);
            int32_t c = 0;
            for (int32_t i = 0; i < n; i++) if (e[i] > 0) c++;
            int stable = c <= net->max_unstable;
            if (!stable) all_stable = 0;
            if (produced < cap) {
                if (rows) { rows[produced].epoch = epoch + epoch_offset; rows[produced].index = s; rows[produced].stable = stable; }
                if (energies) memcpy(energies + (size_t)produced * n, e, sizeof(double) * n);
            }
            produced++;
            *last_epoch = epoch + epoch_offset;
        }
        if (all_stable) break; /* AllStatesAreStable :56-58 */
    }
    free(e);
    return produced;
}

int32_t orc_learn_states(orc_net* net, double* states, int32_t p, int32_t rule, int32_t method,
                         int32_t epochs, int32_t noise_method, double noise_scale, orc_rng* rng,
                         int32_t cap, orc_learn_row* rows, double* energies) {
    orc_activation(net->domain, states, p * net->n); /* HopfieldNetwork.go:255-257 */
    int32_t produced = 0, last = 0;
    if (method == 0) {
        produced = full_set(net, states, p, rule, epochs, noise_method, noise_scale, rng, cap, rows,
                            energies, 0, 0, &last);
    } else {
        /* iterativeBatchLearningMethod: LearningMethod.go:70-89 (BATCHSIZE 5) */
        int32_t batches = p / 5, total = 0;
        for (int32_t it = 1; it <= batches; it++) {
            produced = full_set(net, states, 5 * it, rule, epochs, noise_method, noise_scale, rng, cap,
                                rows, energies, produced, total, &last);
            total = last; /* totalEpochsPassed = last row's Epoch (:84) */
        }
    }
    orc_normalize(net); /* :260 */
    return produced;
}

/* datacollector/UniqueRelaxedStateHandler.go:41-66 */
int32_t orc_unique_table(const double* energies, int32_t n_states, int32_t n, int32_t* attractor_id,
                         int32_t* first, int32_t* hits) {
    int32_t cap = 16;
    while (cap < 2 * n_states) cap <<= 1;
    int32_t* slot = (int32_t*)malloc(sizeof(int32_t) * cap);
    for (int32_t i = 0; i < cap; i++) slot[i] = -1;
    int32_t nu = 0;
    for (int32_t s = 0; s < n_states; s++) {
        const unsigned char* key = (const unsigned char*)(energies + (size_t)s * n);
        uint64_t hsh = 1469598103934665603ULL;
        for (size_t b = 0; b < sizeof(double) * (size_t)n; b++) { hsh ^= key[b]; hsh *= 1099511628211ULL; }
        int32_t pos = (int32_t)(hsh & (uint64_t)(cap - 1));
        for (;;) {
            int32_t u = slot[pos];
            if (u < 0) {
                slot[pos] = nu; first[nu] = s; hits[nu] = 1;
                if (attractor_id) attractor_id[s] = nu;
                nu++;
                break;
            }
            if (memcmp(energies + (size_t)first[u] * n, key, sizeof(double) * (size_t)n) == 0) {
                hits[u]++;
                if (attractor_id) attractor_id[s] = u;
                break;
            }
            pos = (pos + 1) & (cap - 1);
        }
    }
    free(slot);
    return nu;
}

/* states/StateGenerator.go:44-52, :65-74 */
void orc_generate_states(orc_rng* rng, int32_t domain, int32_t n, int32_t count, double mn, double mx,
                         double* out) {
    for (int32_t s = 0; s < count; s++) {
        double* v = out + (size_t)s * n;
        for (int32_t i = 0; i < n; i++) v[i] = orc_rng_uniform(rng, mn, mx);
        orc_activation(domain, v, n);
    }
}

void orc_pack(int32_t n, int32_t count, const double* f64, uint64_t* packed) {
    int32_t w = (n + 63) / 64;
    memset(packed, 0, sizeof(uint64_t) * (size_t)w * count);
    for (int32_t s = 0; s < count; s++)
        for (int32_t i = 0; i < n; i++)
            if (f64[(size_t)s * n + i] > 0.0) packed[(size_t)s * w + i / 64] |= 1ULL << (i % 64);
}
void orc_unpack(int32_t n, int32_t domain, int32_t count, const uint64_t* packed, double* f64) {
    int32_t w = (n + 63) / 64;
    double lo = domain == ORC_BINARY ? 0.0 : -1.0;
    for (int32_t s = 0; s < count; s++)
        for (int32_t i = 0; i < n; i++)
            f64[(size_t)s * n + i] = (packed[(size_t)s * w + i / 64] >> (i % 64)) & 1 ? 1.0 : lo;
}
