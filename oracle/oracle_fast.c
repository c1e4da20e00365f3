/*
 * oracle/oracle_fast.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Exact-integer accelerator for the oracle's relaxation, valid when the weight
 * matrix is a positive scalar multiple of an integer matrix, W_ij = fl(c * K2_ij)
 * (always true for Hebbian / Delta learning from a zero matrix: SURVEY F3).
 * It reproduces orc_relax_state (hopfield_oracle.c, itself restating
 * HopfieldNetwork.go:382-441) decision for decision:
 *   - the exact field g_k = sum_j K2_kj s_j is maintained incrementally in
 *     int64; where g_k != 0 the reference's rounded DotUnitary field has the
 *     sign of g_k (its rounding error is orders of magnitude below c);
 *   - where g_k == 0 (an exact tie of the true field) the reference's result is
 *     whatever DotUnitary's rounding produces, so the oracle's DotUnitary is
 *     evaluated on the float64 row, exactly as the slow oracle does.
 * tests/test_oracle_fast.py checks it against the slow oracle bit for bit.
 * It exists so that parity can be checked at BASELINE sizes (N=4096), where the
 * slow oracle needs seconds per probe.
 */
#include "hopfield_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const orc_net* net;
    const int32_t* k2;   /* [n][n] integer matrix, W = c*K2           */
    const int32_t* k2t;  /* its transpose (== k2 when symmetric)      */
} fast_net;

static int fast_relax_one(const fast_net* fn, double* state, const uint16_t* perm_pool, int32_t n_pool,
                          int64_t offset, orc_relax_info* info, double* distances, double* energies,
                          int64_t* g) {
    const orc_net* net = fn->net;
    int32_t n = net->n;
    double lo = net->domain == ORC_BINARY ? 0.0 : -1.0;
    int64_t dscale = net->domain == ORC_BINARY ? 1 : 2;
    /* exact initial fields */
    for (int32_t i = 0; i < n; i++) {
        const int32_t* row = fn->k2 + (size_t)i * n;
        int64_t s = 0;
        for (int32_t j = 0; j < n; j++) s += (int64_t)row[j] * (int64_t)state[j];
        g[i] = s;
    }
    orc_relax_info li;
    li.stable = 0; li.num_steps = 0; li.flips = 0; li.margin_hits = 0; li.min_abs_h = INFINITY;
    int rc = 0, done = 0;
    for (int32_t step = 1; step <= net->max_iter && !done; step++) {
        int64_t row = offset + (step - 1);
        if (row < 0 || row >= n_pool) { rc = -1; break; }
        const uint16_t* perm = perm_pool + (size_t)row * n;
        for (int32_t i = 0; i < n; i++) {
            int32_t k = perm[i];
            double v;
            if (g[k] != 0) {
                v = g[k] > 0 ? 1.0 : lo;
            } else {
                double h = orc_dot_unitary(net->w + (size_t)k * n, state, n);
                li.margin_hits++;
                v = orc_activation_unit(net->domain, h);
            }
            if (v != state[k]) {
                int64_t d = v > state[k] ? dscale : -dscale;
                const int32_t* col = fn->k2t + (size_t)k * n; /* column k of K2 */
                for (int32_t j = 0; j < n; j++) g[j] += d * col[j];
                state[k] = v;
                li.flips++;
            }
        }
        /* StateIsStable: #{e_i > 0} = #{h_i * v_i < 0} */
        int32_t c = 0;
        for (int32_t i = 0; i < n; i++) {
            double vi = net->domain == ORC_BINARY ? -1.0 + 2.0 * state[i] : state[i];
            if (g[i] != 0) {
                if ((g[i] > 0) != (vi > 0)) c++;
            } else {
                double h = orc_dot_unitary(net->w + (size_t)i * n, state, n);
                li.margin_hits++;
                if (-0.5 * (h * vi) > 0) c++;
            }
        }
        li.num_steps = step + 1;
        if (c <= net->max_unstable) { li.stable = 1; done = 1; }
    }
    if (rc == 0 && !done) { li.stable = 0; li.num_steps = net->max_iter + 1; }
    if (rc == 0) {
        if (energies) orc_all_unit_energies(net, state, energies);
        if (distances) orc_distances_to_targets(net, state, distances);
    }
    if (info) *info = li;
    return rc;
}

typedef struct {
    fast_net fn; double* states; int32_t n_states; const uint16_t* pool; int32_t n_pool;
    const int32_t* offsets; orc_relax_info* info; double* distances; double* energies;
    int32_t next; int rc; pthread_mutex_t mu;
} fast_ctx;

static void* fast_worker(void* arg) {
    fast_ctx* c = (fast_ctx*)arg;
    int32_t n = c->fn.net->n, p = c->fn.net->n_targets;
    int64_t* g = (int64_t*)malloc(sizeof(int64_t) * n);
    for (;;) {
        pthread_mutex_lock(&c->mu);
        int32_t i = c->next++;
        pthread_mutex_unlock(&c->mu);
        if (i >= c->n_states) break;
        int r = fast_relax_one(&c->fn, c->states + (size_t)i * n, c->pool, c->n_pool,
                               c->offsets ? c->offsets[i] : 0, c->info ? c->info + i : NULL,
                               c->distances ? c->distances + (size_t)i * p : NULL,
                               c->energies ? c->energies + (size_t)i * n : NULL, g);
        if (r) { pthread_mutex_lock(&c->mu); c->rc = r; pthread_mutex_unlock(&c->mu); }
    }
    free(g);
    return NULL;
}

/* Same contract as orc_relax_batch; k2 is the integer matrix with W = c*K2
 * (c > 0), k2t its transpose (pass k2 again when symmetric).  Returns -2 if the
 * exactness precondition n^2 * max|K2| * 2^-50 < 0.5 fails. */
int orc_relax_batch_fast(const orc_net* net, const int32_t* k2, const int32_t* k2t, double* states,
                         int32_t n_states, const uint16_t* perm_pool, int32_t n_pool,
                         const int32_t* offsets, int serial_stream, int n_threads,
                         orc_relax_info* info, double* distances, double* energies) {
    int32_t n = net->n, p = net->n_targets;
    int64_t mx = 0;
    for (size_t i = 0; i < (size_t)n * n; i++) {
        int64_t a = k2[i] < 0 ? -(int64_t)k2[i] : k2[i];
        if (a > mx) mx = a;
    }
    if ((double)n * (double)n * (double)mx * 8.881784197001252e-16 >= 0.5) return -2;
    fast_ctx c;
    c.fn.net = net; c.fn.k2 = k2; c.fn.k2t = k2t;
    c.states = states; c.n_states = n_states; c.pool = perm_pool; c.n_pool = n_pool;
    c.offsets = offsets; c.info = info; c.distances = distances; c.energies = energies;
    c.next = 0; c.rc = 0;
    if (serial_stream) {
        int64_t* g = (int64_t*)malloc(sizeof(int64_t) * n);
        int64_t k = 0;
        for (int32_t i = 0; i < n_states; i++) {
            orc_relax_info li;
            int r = fast_relax_one(&c.fn, states + (size_t)i * n, perm_pool, n_pool, k, &li,
                                   distances ? distances + (size_t)i * p : NULL,
                                   energies ? energies + (size_t)i * n : NULL, g);
            if (r) { free(g); return r; }
            if (info) info[i] = li;
            k += li.num_steps - 1;
        }
        free(g);
        return 0;
    }
    pthread_mutex_init(&c.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, fast_worker, &c);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&c.mu);
    return c.rc;
}
