// relax.cuh — asynchronous relaxation of probe states (the headline kernel).
//
// Replaces the body of concurrentRelaxStateRoutine / RelaxState / UpdateState
// (hopfieldnetwork/HopfieldNetwork.go:371-442, 302-359, 280-291): up to maxIter
// sweeps; each sweep visits the units in a given random order k = perm[i] and
// sets s_k = (h_k <= 0 ? low : +1) with h_k = W[k,:] . s; after the sweep the
// state is stable iff #{i : e_i > 0} <= maxUnstable, e = -0.5 (W s) (.) v.
//
// B200 design (DESIGN.md "K2"):
//  * the local fields h = W s of a probe live in registers (UPT doubles per
//    thread) and are updated INCREMENTALLY: a flip of unit k adds +-delta*W[:,k]
//    (one coalesced, 128-bit vectorised pass over a 8N-byte column that is
//    L2/HBM resident) instead of recomputing N dot products per sweep;
//  * between two flips h does not change, so the CTA scans ahead with one
//    warp-shuffle/`redux` min-reduction to the next position of the sweep order
//    whose unit wants to flip: the sequential chain is flips long, not N long;
//  * energies and the stability test are O(N) reductions over the h already in
//    registers (no gemv);
//  * a decision whose incremental |h_k| is below the margin tau (rounding noise
//    could change its sign with respect to the reference's fresh dot product)
//    is re-evaluated EXACTLY in the reference's own summation order (gonum
//    f64.DotUnitary: 4 lanes, multiply then add, unfused), so results stay
//    bit-identical to the oracle even at exact ties of the true field.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace hn {

struct RelaxParams {
    int N;               // dimension
    int ldw;             // leading dimension of W / WT (multiple of 16, zero padded)
    int words;           // 64-bit words per packed state
    const double* Wcol;  // row k of this matrix == column k of W  (W if symmetric else W^T)
    const double* Wrow;  // W: row k . s gives h_k (exact recomputation)
    const double* H0;    // [B][ldh] initial fields (from the GEMM)
    int ldh;
    const uint64_t* probes;     // [B][words]
    const uint16_t* perm_pool;  // [n_pool][ldp]
    const uint16_t* inv_pool;   // [n_pool][ldp] inverse permutations (position of each unit)
    int ldp;
    int n_pool;
    const int32_t* offsets;  // [B] or nullptr
    int B;
    int max_iter;
    int max_unstable;
    int domain;
    int check_stability;  // 0: UpdateState semantics (exactly max_iter sweeps, no test)
    int serial_stream;    // 1: one CTA walks the probes in order, offsets = running sweep count
    double tau_base, tau_flip;
    // outputs (device)
    uint64_t* final_states;  // [B][words]
    int32_t* num_steps;      // [B]
    uint8_t* stable;         // [B]
    int32_t* flips;          // [B]
    int32_t* margin_hits;    // [B]
    int32_t* distances;      // [B][P] or nullptr
    const uint64_t* targets; // [P][words]
    int P;
    uint64_t* hash;          // [B][2] canonical-state hash or nullptr
    uint64_t* history;       // [B][max_iter+1][words] or nullptr
    int32_t* work_counter;   // dynamic probe scheduler
    int32_t* status;         // [0]: set to 1 when the pool was exhausted; [1]: rows needed
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31; return x;
}


// Exact field of one unit in the reference's summation order (gonum f64.DotUnitary,
// the routine behind mat.Dot at HopfieldNetwork.go:286,320,396): lane l of 4
// accumulates indices == l (mod 4), multiply then add (never fused); the n%4 tail
// elements go to lane 0; result (a0+a2)+(a1+a3).  Executed by ONE full warp; `bits`
// is the packed state (32-bit words, shared or global).  Valid in lane 0.
__device__ __forceinline__ double exact_field_warp(const double* __restrict__ row, const uint32_t* bits,
                                                   int N, int domain, int lane) {
    const double lowv = domain == HN_DOMAIN_BINARY ? 0.0 : -1.0;
    double acc = 0.0;
    if (lane < 4) {
        const int n4 = N & ~3;
        for (int j = lane; j < n4; j += 4) {
            const uint32_t b = (bits[j >> 5] >> (j & 31)) & 1u;
            acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), b ? 1.0 : lowv));
        }
        if (lane == 0) {
            for (int j = n4; j < N; j++) {
                const uint32_t b = (bits[j >> 5] >> (j & 31)) & 1u;
                acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), b ? 1.0 : lowv));
            }
        }
    }
    const double other = __shfl_down_sync(0xFFFFFFFFu, acc, 2);  // lane0 <- a2, lane1 <- a3
    const double pair = __dadd_rn(acc, other);                    // lane0: a0+a2, lane1: a1+a3
    const double hi = __shfl_down_sync(0xFFFFFFFFu, pair, 1);
    return __dadd_rn(pair, hi);
}

constexpr int RELAX_THREADS = 256;
constexpr int RELAX_WARPS = RELAX_THREADS / 32;
constexpr uint32_t RELAX_NONE = 0xFFFFFFFFu;

// dynamic shared memory layout: [state bits: ldbits uint32][perm row: ldp uint16]
template <int UPT>
__global__ void __launch_bounds__(RELAX_THREADS)
relax_kernel(RelaxParams p) {
    constexpr int VEC = (UPT >= 2) ? 2 : 1;  // units per vector load
    constexpr int NV = UPT / VEC;            // vector loads per thread per column
    using mask_t = typename std::conditional<(UPT > 32), uint64_t, uint32_t>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbw = (p.N + 31) / 32;         // 32-bit state words
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(smem_raw + (size_t)((nbw + 3) / 4 * 4) * 4);
    __shared__ uint32_t s_red[2][RELAX_WARPS];
    __shared__ double s_exact[4];
    __shared__ int s_probe;
    __shared__ long long s_rowbase;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N;
    const double dscale = p.domain == HN_DOMAIN_BINARY ? 1.0 : 2.0;
    int red_buf = 0;

    auto unit_of = [&](int v, int e) -> int { return (v * RELAX_THREADS + tid) * VEC + e; };

    auto block_min = [&](uint32_t x) -> uint32_t {
        x = __reduce_min_sync(0xFFFFFFFFu, x);
        if (lane == 0) s_red[red_buf][warp] = x;
        __syncthreads();
        uint32_t m = s_red[red_buf][0];
#pragma unroll
        for (int w = 1; w < RELAX_WARPS; w++) m = min(m, s_red[red_buf][w]);
        red_buf ^= 1;
        return m;
    };
    auto block_sum = [&](uint32_t x) -> uint32_t {
        x = __reduce_add_sync(0xFFFFFFFFu, x);
        if (lane == 0) s_red[red_buf][warp] = x;
        __syncthreads();
        uint32_t m = 0;
#pragma unroll
        for (int w = 0; w < RELAX_WARPS; w++) m += s_red[red_buf][w];
        red_buf ^= 1;
        return m;
    };

    // exact re-evaluation of one field, block-uniform call; all threads get the value
    auto exact_field = [&](int k) -> double {
        __syncthreads();  // s_bits complete and visible
        if (warp == 0) {
            const double v = exact_field_warp(p.Wrow + (size_t)k * p.ldw, s_bits, N, p.domain, lane);
            if (lane == 0) s_exact[0] = v;
        }
        __syncthreads();
        return s_exact[0];
    };

    if (tid == 0) { s_probe = (p.serial_stream && blockIdx.x != 0) ? p.B : 0; s_rowbase = 0; }
    for (;;) {
        // ---- fetch next probe ----
        __syncthreads();
        if (tid == 0 && !p.serial_stream) s_probe = atomicAdd(p.work_counter, 1);
        // (serial stream: s_probe / s_rowbase were advanced at the end of the previous probe)
        __syncthreads();
        const int pi = s_probe;
        if (pi >= p.B) break;
        long long rowbase = p.serial_stream ? s_rowbase : (p.offsets ? (long long)p.offsets[pi] : 0LL);

        // ---- load state bits and initial fields ----
        const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(p.probes + (size_t)pi * p.words);
        for (int w = tid; w < nbw; w += RELAX_THREADS) s_bits[w] = pw32[w];
        __syncthreads();
        double h[UPT];
        mask_t sb = 0, valid = 0;
#pragma unroll
        for (int v = 0; v < NV; v++) {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const int u = unit_of(v, e);
                const int r = v * VEC + e;
                if (u < N) {
                    valid |= (mask_t)1 << r;
                    sb |= (mask_t)((s_bits[u >> 5] >> (u & 31)) & 1u) << r;
                    h[r] = p.H0[(size_t)pi * p.ldh + u];
                } else {
                    h[r] = 0.0;
                }
            }
        }
        if (p.history) {
            uint32_t* hw = reinterpret_cast<uint32_t*>(p.history + ((size_t)pi * (p.max_iter + 1)) * p.words);
            for (int w = tid; w < p.words * 2; w += RELAX_THREADS) hw[w] = w < nbw ? s_bits[w] : 0u;
        }

        int flips = 0, margin = 0, sweeps = 0, is_stable = 0, pool_fail = 0;
        for (int sweep = 0; sweep < p.max_iter; sweep++) {
            const long long row = rowbase + sweep;
            if (row >= p.n_pool || row < 0) { pool_fail = 1; break; }
            // positions of my units in this sweep's order; the order itself in smem
            uint32_t posw[NV];  // VEC==2: two 16-bit positions per word
            const uint16_t* invrow = p.inv_pool + (size_t)row * p.ldp;
            const uint16_t* permrow = p.perm_pool + (size_t)row * p.ldp;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int u0 = unit_of(v, 0);
                if (VEC == 2) posw[v] = (u0 < N) ? *reinterpret_cast<const uint32_t*>(invrow + u0) : 0u;
                else posw[v] = (u0 < N) ? (uint32_t)invrow[u0] : 0u;
            }
            __syncthreads();  // previous users of s_perm are done
            for (int i = tid; i < N; i += RELAX_THREADS) s_perm[i] = permrow[i];
            __syncthreads();

            int cur = -1;
            for (;;) {
                const double tau = p.tau_base + p.tau_flip * (double)flips;
                uint32_t best = RELAX_NONE;
#pragma unroll
                for (int r = 0; r < UPT; r++) {
                    const double hv = h[r];
                    const uint32_t nb = hv > 0.0 ? 1u : 0u;
                    const uint32_t ob = (uint32_t)((sb >> r) & 1u);
                    const uint32_t unc = fabs(hv) < tau ? 1u : 0u;
                    const uint32_t ps = (VEC == 2) ? ((posw[r >> 1] >> ((r & 1) * 16)) & 0xFFFFu) : posw[r];
                    const bool cand = ((valid >> r) & 1u) && ((int)ps > cur) && (unc | (nb ^ ob));
                    const uint32_t key = (ps << 3) | (unc << 2) | (ob << 1) | nb;
                    if (cand) best = min(best, key);
                }
                best = block_min(best);
                if (best == RELAX_NONE) break;
                cur = (int)(best >> 3);
                const int k = s_perm[cur];
                uint32_t nb = best & 1u;
                const uint32_t ob = (best >> 1) & 1u;
                if (best & 4u) {
                    const double hx = exact_field(k);
                    margin++;
                    nb = hx > 0.0 ? 1u : 0u;
                }
                if (nb != ob) {
                    const double d = nb ? dscale : -dscale;
                    const double* col = p.Wcol + (size_t)k * p.ldw;
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        const int u0 = unit_of(v, 0);
                        if (u0 < p.ldw) {
                            if (VEC == 2) {
                                const double2 w2 = __ldg(reinterpret_cast<const double2*>(col + u0));
                                h[v * 2] = fma(d, w2.x, h[v * 2]);
                                h[v * 2 + 1] = fma(d, w2.y, h[v * 2 + 1]);
                            } else {
                                h[v] = fma(d, __ldg(col + u0), h[v]);
                            }
                        }
                    }
                    // owner flips its register bit and the shared copy
                    const int kv = k / VEC, ke = k % VEC;
                    if ((kv % RELAX_THREADS) == tid) {
                        sb ^= (mask_t)1 << ((kv / RELAX_THREADS) * VEC + ke);
                        s_bits[k >> 5] ^= 1u << (k & 31);
                    }
                    flips++;
                }
            }
            sweeps++;

            if (p.history) {
                __syncthreads();
                uint32_t* hw = reinterpret_cast<uint32_t*>(p.history + ((size_t)pi * (p.max_iter + 1) + sweeps) * p.words);
                for (int w = tid; w < p.words * 2; w += RELAX_THREADS) hw[w] = w < nbw ? s_bits[w] : 0u;
            }
            if (!p.check_stability) continue;

            // ---- stability: #{i : e_i > 0}, e_i = -0.5 * (h_i * v_i) ----
            const double tau = p.tau_base + p.tau_flip * (double)flips;
            uint32_t cnt = 0;
#pragma unroll
            for (int r = 0; r < UPT; r++) {
                const uint32_t ob = (uint32_t)((sb >> r) & 1u);
                const bool ok = ((valid >> r) & 1u) && !(fabs(h[r]) < tau);
                // v_i = +1 when the bit is set, -1 otherwise (Binary: 2s-1, same sign)
                if (ok && (ob ? (h[r] < 0.0) : (h[r] > 0.0))) cnt++;
            }
            cnt = block_sum(cnt);
            // units inside the margin: exact recomputation, in increasing unit index
            int last = -1;
            for (;;) {
                uint32_t c = RELAX_NONE;
#pragma unroll
                for (int v = 0; v < NV; v++) {
#pragma unroll
                    for (int e = 0; e < VEC; e++) {
                        const int r = v * VEC + e;
                        const int u = unit_of(v, e);
                        if (((valid >> r) & 1u) && fabs(h[r]) < tau && u > last) c = min(c, (uint32_t)u);
                    }
                }
                c = block_min(c);
                if (c == RELAX_NONE) break;
                last = (int)c;
                const double hx = exact_field(last);
                margin++;
                const uint32_t ob = (s_bits[last >> 5] >> (last & 31)) & 1u;
                const double vi = ob ? 1.0 : -1.0;
                const double en = -0.5 * (hx * vi);
                if (en > 0.0) cnt++;
            }
            if ((int)cnt <= p.max_unstable) { is_stable = 1; break; }
        }

        // ---- results ----
        __syncthreads();
        if (pool_fail) {
            if (tid == 0) { atomicExch(&p.status[0], 1); atomicMax(&p.status[1], (int)(rowbase + p.max_iter)); }
        }
        if (p.final_states) {
            uint32_t* fw = reinterpret_cast<uint32_t*>(p.final_states + (size_t)pi * p.words);
            for (int w = tid; w < p.words * 2; w += RELAX_THREADS) fw[w] = w < nbw ? s_bits[w] : 0u;
        }
        if (tid == 0) {
            const int steps = is_stable ? sweeps + 1 : p.max_iter + 1;  // len(StateHistory), main.go:231
            if (p.num_steps) p.num_steps[pi] = p.check_stability ? steps : sweeps + 1;
            if (p.stable) p.stable[pi] = (uint8_t)is_stable;
            if (p.flips) p.flips[pi] = flips;
            if (p.margin_hits) p.margin_hits[pi] = margin;
            if (p.serial_stream) { s_probe = pi + 1; s_rowbase = rowbase + sweeps; }
        }
        // Manhattan distance with inversion to every target (DistanceMeasure.go:28-42):
        // bipolar 2*min(dH, N-dH), binary min(dH, N-dH); exact integers.
        if (p.distances) {
            for (int t = warp; t < p.P; t += RELAX_WARPS) {
                const uint32_t* tw = reinterpret_cast<const uint32_t*>(p.targets + (size_t)t * p.words);
                int dh = 0;
                for (int w = lane; w < nbw; w += 32) dh += __popc(s_bits[w] ^ tw[w]);
                dh = __reduce_add_sync(0xFFFFFFFFu, dh);
                if (lane == 0) {
                    const int m = min(dh, N - dh);
                    p.distances[(size_t)pi * p.P + t] = p.domain == HN_DOMAIN_BINARY ? m : 2 * m;
                }
            }
        }
        // 128-bit hash of the sign-canonical final state (bipolar: s and -s share an
        // energy profile bit for bit, which is the reference's dedup key).
        if (p.hash) {
            const bool inv = (p.domain == HN_DOMAIN_BIPOLAR) && (s_bits[0] & 1u);
            uint64_t h0 = 0, h1 = 0;
            for (int w = tid; w < nbw; w += RELAX_THREADS) {
                uint32_t x = s_bits[w];
                if (inv) {
                    x = ~x;
                    if (w == nbw - 1 && (N & 31)) x &= (1u << (N & 31)) - 1u;
                }
                const uint64_t y = ((uint64_t)(w + 1) << 32) | x;
                h0 ^= mix64(y ^ 0x9E3779B97F4A7C15ULL);
                h1 ^= mix64(y * 0xD6E8FEB86659FD93ULL + 0x2545F4914F6CDD1DULL);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                h0 ^= __shfl_xor_sync(0xFFFFFFFFu, h0, o);
                h1 ^= __shfl_xor_sync(0xFFFFFFFFu, h1, o);
            }
            if (lane == 0) {
                atomicXor(reinterpret_cast<unsigned long long*>(&p.hash[(size_t)pi * 2]), (unsigned long long)h0);
                atomicXor(reinterpret_cast<unsigned long long*>(&p.hash[(size_t)pi * 2 + 1]), (unsigned long long)h1);
            }
        }
    }
}

}  // namespace hn
