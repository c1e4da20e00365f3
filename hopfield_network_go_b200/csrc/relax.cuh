// relax.cuh — asynchronous relaxation of probe states (the headline kernel).
//
// Replaces the body of concurrentRelaxStateRoutine / RelaxState / UpdateState
// (hopfieldnetwork/HopfieldNetwork.go:371-442, 302-359, 280-291): up to maxIter
// sweeps; each sweep visits the units in a given random order k = perm[i] and
// sets s_k = (h_k <= 0 ? low : +1) with h_k = W[k,:] . s; after the sweep the
// state is stable iff #{i : e_i > 0} <= maxUnstable, e = -0.5 (W s) (.) v.
//
// B200 design (DESIGN.md "K2"):
//  * the local fields h = W s of a probe live in registers (UPT per thread:
//    float64, or exact int32 on the integer stream) and are updated
//    INCREMENTALLY: a flip of unit k adds +-delta*W[:,k] (one coalesced,
//    128/256-bit vectorised pass over the column, L2 resident) instead of
//    recomputing N dot products per sweep;
//  * between two flips h does not change, so the CTA scans ahead with one
//    `redux` min-reduction per 1024 positions to the next position of the sweep
//    order whose unit wants to flip: the sequential chain is flips long, not N long;
//  * few, fat threads: 32 units per thread (one warp per probe up to N = 1024,
//    4 warps at N = 4096) because what every warp repeats per flip (scan, two
//    barriers, flag, addresses) costs more than the update itself;
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
    const void* Wcol;    // row k of this matrix == column k of W  (W if symmetric else W^T); float64 or float32 shadow
    const double* Wrow;  // W: row k . s gives h_k (exact recomputation)
    const double* H0;    // [B][ldh] initial fields (from the GEMM)
    const int32_t* H0i;  // integer stream: [B][ldh] exact int32 fields g = scale*Wraw*s from the integer GEMM (or nullptr)
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
    double int_scale;    // exact integer stream: fields g = int_scale * (Wraw s)
    double tau_base, tau_flip;  // margin tau(F) = tau_base + tau_flip * F (tau_flip includes the float32 stream bound)
    double key_scale;    // != 0: weights have a known integer structure, rint(key_scale * h) is the exact integer field (unique-table key)
    int exact_arith;     // float64 stream on quarter-integer weights: every field is exact, tau = 0, a zero field is a zero
    // outputs (device)
    uint64_t* final_states;  // [B][words]
    uint64_t* final_states_host;  // [B][words] or nullptr: second copy, written straight into mapped page-locked host memory
    int32_t* num_steps;      // [B]
    uint8_t* stable;         // [B]
    int32_t* flips;          // [B]
    int32_t* margin_hits;    // [B]
    int32_t* distances;      // [B][P] or nullptr (device memory, or mapped page-locked host memory: rows are written coalesced)
    const uint64_t* targets; // [P][words]
    int P;
    uint64_t* hash;          // [B][4] unique-table keys (zeroed by the host): [0..1] canonical-state hash, [2..3] exact-profile key; or nullptr
    uint8_t* zero_field;     // [B] 1 iff the final state has a unit with an exactly zero field (dedup key, see same_canonical)
    uint64_t* history;       // [B][max_iter+1][words] or nullptr
    const int32_t* sweeps_done;   // [B] or nullptr: sweeps already done by the dense lockstep kernel (dense.cuh)
    const int32_t* flips_done;    // [B] or nullptr: flips counted so far
    const int32_t* margin_done;   // [B] or nullptr: exact ties counted so far
    int32_t* work_counter;   // dynamic probe scheduler
    int32_t* status;         // [0]: set to 1 when the pool was exhausted; [1]: rows needed
};



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

constexpr uint32_t RELAX_NONE = 0xFFFFFFFFu;
constexpr int RELAX_TAU_STEP = 256;
constexpr int RELAX_DIST_STAGE = 512;   // distances staged in shared memory per probe before one coalesced row write

// Dynamic shared memory layout (see relax_smem_bytes):
//   s_bits  [nbw, padded to 4]   uint32  packed state (exact path, outputs)
//   s_perm  [ldp]                uint16  this sweep's visiting order        (position -> unit)
//   s_inv   [ldp]                uint16  its inverse                        (unit -> position)
//   s_cand  [ncw, padded to 4]   uint32  POSITION-indexed candidate bitmap: bit i is set iff the unit at
//           position i of this sweep has aligned field h*v < tau, i.e. wants to flip or sits inside the margin.
//
// Per flip: every thread updates its UPT fields in registers (one coalesced pass over the column), re-derives the
// candidacy of its units (float: hi-word integer compare; integer stream: sign bits by funnel shifts) and toggles
// the bitmap only where it changed; __syncthreads; every warp finds the next candidate position with one coalesced
// read of the bitmap (32 words = 1024 positions per step) + redux.min; the unit's owner publishes (inside-margin?,
// old bit); __syncthreads.  The L1/LSU data pipe carries the column and almost nothing else.
template <int UPT, typename CT> struct RelaxCfg {
    static constexpr int VMAX = std::is_same<CT, int16_t>::value ? 8 : 4;   // one 128-bit load of int16 / float32, one 256-bit load of float64
    static constexpr int VEC = UPT >= VMAX ? VMAX : UPT;   // contiguous units per ownership chunk
    static constexpr int NCH = UPT / VEC;            // chunks per thread
};
static inline int relax_cand_words(int nbw) { return (nbw + 31) / 32 * 32 + 32; }   // scan reads 32 words from any word < nbw
static inline size_t relax_smem_bytes(int N, int ldp, int upt) {  // independent of the thread count
    const int nbw = (N + 31) / 32;
    (void)upt;
    return (size_t)((nbw + 3) / 4 * 4 + relax_cand_words(nbw)) * 4 + (size_t)((ldp + 7) / 8 * 8) * 2 * 2;
}

// v[idx] for a register-resident array and a run-time index: a binary tree of selects (N-1 selects, log2 N deep)
// instead of the N-long dependent compare/select chain a loop would give.
template <int N_>
__device__ __forceinline__ int tree_select(const int (&v)[N_], int idx) {
    int t[N_];
#pragma unroll
    for (int i = 0; i < N_; i++) t[i] = v[i];
#pragma unroll
    for (int s = 1; s < N_; s <<= 1) {
        const bool up = (idx & s) != 0;
#pragma unroll
        for (int i = 0; i + s < N_; i += 2 * s) t[i] = up ? t[i + s] : t[i];
    }
    return t[0];
}

// UPT units per thread, THREADS threads per probe, MINB = CTAs per SM the register allocation must allow (chosen per
// geometry in api.cu: launch_relax_c / launch_relax_int)
template <int UPT, typename CT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
relax_kernel(RelaxParams p) {
    constexpr int RELAX_THREADS = THREADS;   // threads per probe: 256 up to N=4096, 512 / 1024 for N=8192 / 16384
    constexpr int RELAX_WARPS = THREADS / 32;
    constexpr int VEC = RelaxCfg<UPT, CT>::VEC;
    constexpr int NCH = RelaxCfg<UPT, CT>::NCH;
    constexpr int NU = RELAX_THREADS * UPT;  // padded unit count
    using mask_t = typename std::conditional<(UPT > 32), uint64_t, uint32_t>::type;
    // Field type: float64 local fields, or — CT = int16_t, the exact-integer stream — int32 fields g = 2*(Wraw s) with
    // Wraw the un-normalised dyadic matrix (W = c*Wraw, c > 0): sign(g) is the sign of the reference's field wherever
    // g != 0, and g == 0 is an exact tie that is re-evaluated on the float64 matrix (as oracle/oracle_fast.c does).
    constexpr bool INTF = std::is_same<CT, int16_t>::value;
    using FT = typename std::conditional<INTF, int, double>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbw = (p.N + 31) / 32;
    const int nbw4 = (nbw + 3) / 4 * 4, ldp8 = (p.ldp + 7) / 8 * 8;
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem_raw);
    const int ncw = ((nbw + 31) / 32) * 32 + 32;   // relax_cand_words(nbw); words >= nbw stay zero
    uint32_t* s_cand = s_bits + nbw4;
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(s_cand + ncw);
    uint16_t* s_inv = s_perm + ldp8;
    __shared__ __align__(16) uint32_t s_red[2][32];
    __shared__ double s_exact[4];
    __shared__ int s_probe;
    __shared__ int s_zero_field;                 // the state after the LAST sweep has a unit whose reference-order field is exactly zero
    __shared__ unsigned long long s_tz[2];       // hash marks of those units, by the sign of v (unique-table key, epilogue)
    __shared__ int s_flag;
    __shared__ long long s_rowbase;
    __shared__ int s_dist[RELAX_DIST_STAGE];

    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));   // keep the thread index in a register (the compiler would re-read SR_TID every flip)
    const int lane = tid & 31, warp = tid >> 5;
    const bool full_rows = p.ldw == NU;   // every ownership chunk lies inside the padded row: unpredicated column pass
    const int N = p.N;
    const double dscale = p.domain == HN_DOMAIN_BINARY ? 1.0 : 2.0;
    // Margin threshold on the hi word, valid for the next RELAX_TAU_STEP flips: T(F) = hi32(tau(F + STEP)) + 1.
    // It is refreshed every RELAX_TAU_STEP flips (tau grows with the flip count).
    auto thr = [&](int f) -> int { return __double2hiint(p.tau_base + p.tau_flip * (double)(f + RELAX_TAU_STEP)) + 1; };
    int T = thr(0);
    int red_buf = 0;

    auto unit_of = [&](int c, int e) -> int { return (c * RELAX_THREADS + tid) * VEC + e; };
    auto unit_of_r = [&](int r) -> int { return ((r / VEC) * RELAX_THREADS + tid) * VEC + (r % VEC); };

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
        T = thr(0);
        FT h[UPT];         // float64: local fields h of my units.  Integer stream: U = 2g - 2b with the EXACT field
                           // g = K s (K = scale*Wraw) and b the unit's current bit: the unit wants to flip or sits on an
                           // exact tie (v*g <= 0) exactly when  b = 1: g <= 0 <=> U < 0;   b = 0: g >= 0 <=> U >= 0,
                           // so the candidacy mask is signmask(U) ^ ~bits: one funnel shift per unit and ONE xor per flip.
                           // A flip adds the same multiple (+-delta) of the column of M = 2K to every field — including the
                           // flipped unit's own, whose -+2 (its bit changed) sits on M's diagonal (to_i16_kernel): no
                           // per-unit sign registers and no owner fix-up of a register chosen at run time.
        mask_t sbm = 0;    // my units' current bits; SIGNW(r) = 0x80000000 when the bit is clear (v = -1): hi32(h)^SIGNW = hi32(h*v)
#define SIGNW(r) ((int)((uint32_t)((~sbm >> (r)) & 1u) << 31))
#pragma unroll
        for (int c = 0; c < NCH; c++) {
#pragma unroll
            for (int e = 0; e < VEC; e++) {
                const int u = unit_of(c, e), r = c * VEC + e;
                if (u < N) {
                    const uint32_t b = (s_bits[u >> 5] >> (u & 31)) & 1u;
                    sbm |= (mask_t)b << r;
                    if constexpr (INTF) {   // H0 = S * Wraw^T is exact in float64
                                                // integer GEMM: (M s)_u = 2 g_u + diag * s_u, i.e. U_u up to -1 (bipolar) / 0 (binary);  float64 GEMM
                        // (asymmetric weights): K s = int_scale * H0 exactly
                        if (p.H0i) h[r] = p.H0i[(size_t)pi * p.ldh + u] - (p.domain == HN_DOMAIN_BINARY ? 0 : 1);
                        else h[r] = 2 * ((int)llrint(p.int_scale * p.H0[(size_t)pi * p.ldh + u]) - (int)b);
                    } else h[r] = p.H0[(size_t)pi * p.ldh + u];
                } else {  // padding: never a candidate, never visited
                    sbm |= (mask_t)1 << r;
                    if constexpr (INTF) h[r] = 0x3FFFFFFE; else h[r] = __longlong_as_double(0x7FF0000000000000LL);   // (bit 1, U >= 0; its column entries are 0)
                }
            }
        }
        // candidacy of my units: bit r set iff hi32(h*v) < T   (aligned field below tau)
        auto cand_mask = [&]() -> mask_t {
            mask_t m = 0;
            if constexpr (INTF) {   // signmask(u) ^ ~bits: wants to flip (v*g < 0) or exact tie (g == 0)
                if constexpr (UPT >= 8) {   // two independent chains halve the dependent latency
                    uint32_t hi = 0, lo = 0;
#pragma unroll
                    for (int r = UPT / 2 - 1; r >= 0; r--) {
                        hi = __funnelshift_l((uint32_t)h[r + UPT / 2], hi, 1);
                        lo = __funnelshift_l((uint32_t)h[r], lo, 1);
                    }
                    m = ((mask_t)hi << (UPT / 2)) | (mask_t)lo;
                } else {
#pragma unroll
                    for (int r = UPT - 1; r >= 0; r--) m = (mask_t)__funnelshift_l((uint32_t)h[r], (uint32_t)m, 1);
                }
                m ^= ~sbm;
                if constexpr (UPT < 8 * (int)sizeof(mask_t)) m &= ((mask_t)1 << UPT) - 1;
            } else {
#pragma unroll
                for (int r = 0; r < UPT; r++)
                    if ((__double2hiint(h[r]) ^ SIGNW(r)) < T) m |= (mask_t)1 << r;
            }
            return m;
        };
        // owner of a flipping unit: new bit into its bit mask and the packed state (integer stream: its own field moves
        // through the diagonal of M with everyone else's)
        auto owner_flip = [&](int rr, int k) {
            sbm ^= (mask_t)1 << rr;
            s_bits[k >> 5] ^= 1u << (k & 31);
        };
        mask_t cm = cand_mask();
        if (p.history) {
            uint32_t* hw = reinterpret_cast<uint32_t*>(p.history + ((size_t)pi * (p.max_iter + 1)) * p.words);
            for (int w = tid; w < p.words * 2; w += RELAX_THREADS) hw[w] = w < nbw ? s_bits[w] : 0u;
        }

        int flips = p.flips_done ? p.flips_done[pi] : 0, margin = p.margin_done ? p.margin_done[pi] : 0;
        int sweeps = p.sweeps_done ? p.sweeps_done[pi] : 0, is_stable = 0, pool_fail = 0;
        if (tid == 0) { s_zero_field = 0; s_tz[0] = s_tz[1] = 0ULL; }   // rare-path state lives in shared memory, not registers
        // A probe handed over by the dense lockstep sweeps (dense.cuh) arrives with `sweeps` sweeps done and the fields of
        // its current state: the stability test that follows its last sweep comes first (its exact ties were already
        // counted by the scan that handed it over), then the sweeps continue at pool row rowbase + sweeps.
        bool precheck = sweeps > 0 && p.check_stability;
        for (;;) {
            const bool count_margin = !precheck;
            if (!precheck) {
            if (sweeps >= p.max_iter) break;
            const long long row = rowbase + sweeps;
            if (row >= p.n_pool || row < 0) { pool_fail = 1; break; }
            const uint16_t* permrow = p.perm_pool + (size_t)row * p.ldp;
            const uint16_t* invrow = p.inv_pool + (size_t)row * p.ldp;
            __syncthreads();  // previous readers of s_perm / s_inv / s_cand are done
#pragma unroll 1
            for (int i = tid; i < p.ldp / 4; i += RELAX_THREADS) {   // (128-bit copies make the integer kernel spill at 56 registers: +48% time)
                const uint2 a = __ldg(reinterpret_cast<const uint2*>(permrow) + i);
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(invrow) + i);
                reinterpret_cast<uint2*>(s_perm)[i] = a;
                reinterpret_cast<uint2*>(s_inv)[i] = b;
            }
            for (int w = tid; w < ncw; w += RELAX_THREADS) s_cand[w] = 0u;
            __syncthreads();
            {   // bitmap of this sweep's order from the candidacy masks held in registers
                mask_t m = cm;
                while (m) {
                    const int r = (sizeof(mask_t) == 8) ? (__ffsll((long long)m) - 1) : (__ffs((int)m) - 1);
                    m &= m - 1;
                    const int pos = s_inv[unit_of_r(r)];
                    atomicOr(&s_cand[pos >> 5], 1u << (pos & 31));
                }
            }
            __syncthreads();

            int base = 0;  // next position of the sweep order to examine
            for (;;) {
                // ---- scan ahead (every warp, redundantly): first set bit of the bitmap at position >= base ----
                int found = -1;
                {
                    uint32_t first = lane == 0 ? 0xFFFFFFFFu << (base & 31) : 0xFFFFFFFFu;   // positions < base are done
                    for (int w0 = base >> 5; w0 < nbw; w0 += 32) {
                        const int wi = w0 + lane;
                        const uint32_t word = s_cand[wi] & first;            // padded with zero words: no bounds check
                        uint32_t pos = word ? (uint32_t)(wi * 32 + __ffs((int)word) - 1) : RELAX_NONE;
                        pos = __reduce_min_sync(0xFFFFFFFFu, pos);
                        if (pos != RELAX_NONE) { found = (int)pos; break; }
                        first = 0xFFFFFFFFu;
                    }
                }
                if (found < 0) break;  // sweep finished
                base = found + 1;
                const int k = s_perm[found];
                const int kch = k / VEC;
                const bool owner = (kch % RELAX_THREADS) == tid;
                if (owner) {   // publishes: bit1 = inside the margin (integer stream: exact tie), bit0 = current (old) bit
                    const int rr = (kch / RELAX_THREADS) * VEC + k % VEC;   // register slot of unit k
                    int hw[UPT];
#pragma unroll
                    for (int r = 0; r < UPT; r++) { if constexpr (INTF) hw[r] = h[r]; else hw[r] = __double2hiint(h[r]); }
                    const int a = tree_select(hw, rr);
                    const uint32_t obit = (s_bits[k >> 5] >> (k & 31)) & 1u;
                    const bool inside = INTF ? (a + 2 * (int)obit == 0) : ((a & 0x7FFFFFFF) < T);   // integer stream: g == 0
                    s_flag = (inside ? 2 : 0) | (int)obit;
                }
                __syncthreads();
                const int fl = s_flag;
                const uint32_t ob = (uint32_t)fl & 1u;
                uint32_t nb = ob ^ 1u;
                if (fl & 2) {  // inside the margin: decide exactly, in the reference's summation order
                    margin++;
                    if (p.exact_arith) {
                        nb = 0u;   // the reference's field is this exact zero too: h <= 0 selects the low value
                        // no flip: every thread must have read s_flag before the next owner may overwrite it (the
                        // other tie path passes exact_field's two barriers, a flip passes the barrier after the update)
                        if (RELAX_WARPS > 1 && nb == ob) __syncthreads();
                    } else nb = exact_field(k) > 0.0 ? 1u : 0u;
                }
                if (nb == ob) continue;  // exact decision: no flip; nothing was written, keep scanning

                // ---- flip unit k: h += d * W[:,k] over my units; toggle the bitmap where candidacy changed ----
                {   // one coalesced pass over column k (row k of Wcol), fused into the register-resident fields
                    const CT* mycol = reinterpret_cast<const CT*>(p.Wcol) + ((unsigned)k * (unsigned)p.ldw + (unsigned)(tid * VEC));   // my chunk 0 of column k (N*ldw < 2^32)
                    const double d = nb ? dscale : -dscale;
                    if constexpr (INTF) {
                        // exact int16 column stream, 2N bytes per flip: u += (+-delta) * K[:,k] by dp2a — two int16 column entries
                        // per word against the multiplier in byte 0 (dp2a.lo: even unit) / byte 3 (dp2a.hi: odd unit)
                        const int ds = nb ? (int)dscale : -(int)dscale;
                        const int q = (ds & 0xFF) | ((ds & 0xFF) << 24);
                        auto ipass = [&](auto full_tag) {
                        constexpr bool FULL = decltype(full_tag)::value;   // every chunk lies inside the padded row
                        constexpr int XW = (VEC + 1) / 2;
                        int x[NCH][XW];
#pragma unroll
                        for (int c = 0; c < NCH; c++) {
#pragma unroll
                            for (int e = 0; e < XW; e++) x[c][e] = 0;
                            if (FULL || unit_of(c, 0) < p.ldw) {
                                const CT* src = mycol + c * (RELAX_THREADS * VEC);
                                if constexpr (VEC == 8) {
                                    const int4 t = __ldg(reinterpret_cast<const int4*>(src));
                                    x[c][0] = t.x; x[c][1 % XW] = t.y; x[c][2 % XW] = t.z; x[c][3 % XW] = t.w;
                                } else if constexpr (VEC == 4) {
                                    const int2 t = __ldg(reinterpret_cast<const int2*>(src));
                                    x[c][0] = t.x; x[c][1 % XW] = t.y;
                                } else if constexpr (VEC == 2) {
                                    x[c][0] = __ldg(reinterpret_cast<const int*>(src));
                                } else {
                                    x[c][0] = (int)__ldg(src);   // hi half = sign bits, multiplied by byte 1 = 0
                                }
                            }
                        }
                        // owner, in the shadow of the loads: its own sign changes (commutes with the column update)
                        if (owner) owner_flip((kch / RELAX_THREADS) * VEC + k % VEC, k);
#pragma unroll
                        for (int c = 0; c < NCH; c++) {
#pragma unroll
                            for (int e = 0; e < VEC; e++) {
                                const int r = c * VEC + e;
                                h[r] = (r & 1) ? __dp2a_hi(x[c][e >> 1], q, h[r]) : __dp2a_lo(x[c][e >> 1], q, h[r]);
                            }
                        }
                        };
                        if (full_rows) ipass(std::true_type{}); else ipass(std::false_type{});
                    } else {
                    if (owner) owner_flip((kch / RELAX_THREADS) * VEC + k % VEC, k);
                    auto pass = [&](auto full_tag) {
                        constexpr bool FULL = decltype(full_tag)::value;   // every chunk lies inside the padded row
#pragma unroll
                        for (int c = 0; c < NCH; c++) {
                            const int u0 = unit_of(c, 0);
                            {
                            double w[VEC];
#pragma unroll
                            for (int e = 0; e < VEC; e++) w[e] = 0.0;
                            if (FULL || u0 < p.ldw) {
                                if constexpr (sizeof(CT) == 4) {   // float32 column stream: 4N bytes per flip
                                    if (VEC == 4) {
                                        const float4 x = __ldg(reinterpret_cast<const float4*>(mycol + c * (RELAX_THREADS * VEC)));
                                        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
                                    } else if (VEC == 2) {
                                        const float2 x = __ldg(reinterpret_cast<const float2*>(mycol + c * (RELAX_THREADS * VEC)));
                                        w[0] = x.x; w[1] = x.y;
                                    } else {
                                        w[0] = __ldg(mycol + c * (RELAX_THREADS * VEC));
                                    }
                                } else if (VEC == 4) {   // 256-bit loads (8% faster than two 128-bit loads here)
                                    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                                                 : "=d"(w[0]), "=d"(w[1]), "=d"(w[2]), "=d"(w[3]) : "l"(mycol + c * (RELAX_THREADS * VEC)));
                                } else if (VEC == 2) {
                                    const double2 x = __ldg(reinterpret_cast<const double2*>(mycol + c * (RELAX_THREADS * VEC)));
                                    w[0] = x.x; w[1] = x.y;
                                } else {
                                    w[0] = __ldg(mycol + c * (RELAX_THREADS * VEC));
                                }
                            }
#pragma unroll
                            for (int e = 0; e < VEC; e++) {
                                const int r = c * VEC + e;
                                h[r] = fma(d, w[e], h[r]);
                            }
                            }
                        }
                    };
                    if (full_rows) pass(std::true_type{}); else pass(std::false_type{});
                    }
                }
                if (!INTF && ((flips + 1) & (RELAX_TAU_STEP - 1)) == 0) T = thr(flips + 1);  // candidacy below is re-derived with it
                {
                    const mask_t nm = cand_mask();
                    mask_t ch = nm ^ cm;
                    cm = nm;
                    while (ch) {
                        const int r = (sizeof(mask_t) == 8) ? (__ffsll((long long)ch) - 1) : (__ffs((int)ch) - 1);
                        ch &= ch - 1;
                        const int pos = s_inv[unit_of_r(r)];
                        atomicXor(&s_cand[pos >> 5], 1u << (pos & 31));
                    }
                }
                flips++;
                __syncthreads();  // bitmap complete before anyone scans it
            }
            sweeps++;

            if (p.history) {
                __syncthreads();
                uint32_t* hw = reinterpret_cast<uint32_t*>(p.history + ((size_t)pi * (p.max_iter + 1) + sweeps) * p.words);
                for (int w = tid; w < p.words * 2; w += RELAX_THREADS) hw[w] = w < nbw ? s_bits[w] : 0u;
            }
            if (!p.check_stability) continue;
            }
            precheck = false;

            // ---- stability: #{i : e_i > 0}, e_i = -0.5 h_i v_i  (HopfieldNetwork.go:214-225) ----
            uint32_t cnt = 0;
#pragma unroll
            for (int r = 0; r < UPT; r++) {
                if constexpr (INTF) {
                    if (((sbm >> r) & 1u) ? (h[r] < -2) : (h[r] > 0)) cnt++;   // v*g < 0: certainly unstable
                } else {
                    const int a = __double2hiint(h[r]) ^ SIGNW(r);
                    if (a < 0 && !((a & 0x7FFFFFFF) < T)) cnt++;
                }
            }
            cnt = block_sum(cnt);
            if (!INTF && p.exact_arith) {   // exact zeros: e = -0.5 * 0 * v is not > 0; only counted as margin decisions
                uint32_t z = 0;
#pragma unroll
                for (int r = 0; r < UPT; r++)
                    if ((__double2hiint(h[r]) & 0x7FFFFFFF) < T) z++;
                const int nz = (int)block_sum(z);
                if (count_margin) margin += nz;
                if (tid == 0) s_zero_field = nz > 0;
                if ((int)cnt <= p.max_unstable) { is_stable = 1; break; }
                continue;
            }
            // units inside the margin: exact recomputation, in increasing unit index
            int last = -1;
            if (tid == 0) { s_zero_field = 0; s_tz[0] = s_tz[1] = 0ULL; }
            for (;;) {
                uint32_t cmin = RELAX_NONE;
#pragma unroll
                for (int r = 0; r < UPT; r++) {
                    const int u = unit_of_r(r);
                    bool inside;
                    if constexpr (INTF) inside = h[r] + 2 * (int)((sbm >> r) & 1u) == 0; else inside = (__double2hiint(h[r]) & 0x7FFFFFFF) < T;
                    if (inside && u > last) cmin = min(cmin, (uint32_t)u);
                }
                cmin = block_min(cmin);
                if (cmin == RELAX_NONE) break;
                last = (int)cmin;
                const double hx = exact_field(last);
                if (count_margin) margin++;
                const uint32_t ob = (s_bits[last >> 5] >> (last & 31)) & 1u;
                if (hx == 0.0 && tid == 0) {   // unit energy -0.5*(+0*v): a zero whose sign follows v
                    s_zero_field = 1;
                    s_tz[0] ^= mix64(((uint64_t)(last + 1) << 1 | ob) ^ 0xA24BAED4963EE407ULL);
                    s_tz[1] ^= mix64(((uint64_t)(last + 1) << 1 | ob) * 0x9FB21C651E98DF25ULL + 0x632BE59BD9B4E019ULL);
                }
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
            if (p.final_states_host) {
                uint32_t* fh = reinterpret_cast<uint32_t*>(p.final_states_host + (size_t)pi * p.words);
                for (int w = tid; w < p.words * 2; w += RELAX_THREADS) fh[w] = w < nbw ? s_bits[w] : 0u;
            }
        }
        if (tid == 0) {
            const int steps = is_stable ? sweeps + 1 : p.max_iter + 1;  // len(StateHistory), main.go:231
            if (p.num_steps) p.num_steps[pi] = p.check_stability ? steps : sweeps + 1;
            if (p.stable) p.stable[pi] = (uint8_t)is_stable;
            if (p.flips) p.flips[pi] = flips;
            if (p.margin_hits) p.margin_hits[pi] = margin;
            if (p.zero_field) p.zero_field[pi] = (uint8_t)s_zero_field;
            if (p.serial_stream) { s_probe = pi + 1; s_rowbase = rowbase + sweeps; }
        }
        // Manhattan distance with inversion to every target (DistanceMeasure.go:28-42):
        // bipolar 2*min(dH, N-dH), binary min(dH, N-dH); exact integers.
        // The row of a probe is staged in shared memory and written with consecutive threads on consecutive words, so
        // that it also goes out in full segments when the destination is mapped host memory (no device copy, no
        // device -> host transfer after the kernel).
        if (p.distances) {
            int32_t* drow = p.distances + (size_t)pi * p.P;
            for (int t0 = 0; t0 < p.P; t0 += RELAX_DIST_STAGE) {
                const int tn = min(RELAX_DIST_STAGE, p.P - t0);
                for (int t = warp; t < tn; t += RELAX_WARPS) {
                    const uint32_t* tw = reinterpret_cast<const uint32_t*>(p.targets + (size_t)(t0 + t) * p.words);
                    int dh = 0;
                    for (int w = lane; w < nbw; w += 32) dh += __popc(s_bits[w] ^ tw[w]);
                    dh = __reduce_add_sync(0xFFFFFFFFu, dh);
                    if (lane == 0) {
                        const int m = min(dh, N - dh);
                        s_dist[t] = p.domain == HN_DOMAIN_BINARY ? m : 2 * m;
                    }
                }
                __syncthreads();
                for (int i = tid; i < tn; i += RELAX_THREADS) drow[t0 + i] = s_dist[i];
                if (t0 + RELAX_DIST_STAGE < p.P) __syncthreads();
            }
        }
        // Unique-table keys (UniqueRelaxedStateHandler.go:44 keys on the printed energy profile), four words per probe:
        //  [0..1] hash of the sign-canonical final state (s and -s share a profile bit for bit unless a field is an
        //         exact zero, whose energy is a zero with the sign of v): level 1 of the table, equality decided on
        //         the full packed state;
        //  [2..3] where the profile is known EXACTLY — integer stream: v*g per unit; exact float arithmetic: the float64
        //         values themselves; known integer structure on a float stream: rint(key_scale*h) — a hash of the
        //         profile itself plus a mark for every unit whose re-evaluated float field is +0.  Equal profiles imply
        //         equal keys, so DISTINCT states with identical profiles (the two stored patterns of a P=2 network)
        //         meet at level 2 of the table, where the candidates' float profiles are compared bit for bit in the
        //         reference's summation order (dedup_verify_kernel).
        if (p.hash) {
            uint64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
            if (INTF || p.exact_arith || p.key_scale != 0.0) {
                // 128 key bits: four multilinear hashes  sum_u value_u * c_j(u)  mod 2^32  with unit-dependent odd constants
                // (a universal family: distinct profiles collide with probability ~2^-32 per hash), per thread; threads are
                // combined by XOR.  No 64-bit arithmetic in this part of the epilogue.
                uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;
                auto mix_in = [&](uint32_t uk, uint32_t v) {
                    k0 += v * (uk | 1u);
                    k1 += v * (__funnelshift_l(uk, uk, 11) * 0x85EBCA6Bu | 1u);
                    k2 += v * ((uk ^ 0xC2B2AE35u) * 0x27D4EB2Fu | 1u);
                    k3 += (v ^ (v >> 15)) * (__funnelshift_l(uk, uk, 19) * 0x165667B1u | 1u);
                };
#pragma unroll
                for (int r = 0; r < UPT; r++) {
                    const int u = unit_of_r(r);
                    if (u < N) {
                        const uint32_t b = (uint32_t)((sbm >> r) & 1u);
                        const uint32_t uk = (uint32_t)(u + 1) * 0x9E3779B1u;
                        if constexpr (INTF) {
                            mix_in(uk, (uint32_t)(b ? ((int)h[r] >> 1) + 1 : -((int)h[r] >> 1)));   // v*g from U = 2g - 2b (even)
                        } else if (p.exact_arith) {
                            const double hv = (double)h[r];
                            const long long t = __double_as_longlong((hv == 0.0 ? 0.0 : hv) * (b ? 1.0 : -1.0));   // h*v, the bits behind e
                            mix_in(uk, (uint32_t)t);
                            mix_in(uk ^ 0x5BD1E995u, (uint32_t)(t >> 32));
                        } else {
                            const long long g = __double2ll_rn((double)h[r] * p.key_scale);   // exact: |error| << 1/2
                            mix_in(uk, (uint32_t)(int)(b ? g : -g));
                        }
                    }
                }
                // final avalanche per thread so that the XOR across threads does not cancel structure
                k0 ^= k0 >> 16; k0 *= 0x7FEB352Du; k0 ^= k0 >> 15;
                k1 ^= k1 >> 15; k1 *= 0x846CA68Bu; k1 ^= k1 >> 16;
                k2 ^= k2 >> 17; k2 *= 0xC2B2AE35u; k2 ^= k2 >> 13;
                k3 ^= k3 >> 14; k3 *= 0x85EBCA6Bu; k3 ^= k3 >> 16;
                h2 = ((uint64_t)k1 << 32) | k0;
                h3 = ((uint64_t)k3 << 32) | k2;
                if ((INTF || !p.exact_arith) && tid == 0) { h2 ^= s_tz[0]; h3 ^= s_tz[1]; }   // units whose re-evaluated float field is +0
            }
            {
                const bool inv = (p.domain == HN_DOMAIN_BIPOLAR) && (s_bits[0] & 1u) && !s_zero_field;
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
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                h0 ^= __shfl_xor_sync(0xFFFFFFFFu, h0, o);
                h1 ^= __shfl_xor_sync(0xFFFFFFFFu, h1, o);
                h2 ^= __shfl_xor_sync(0xFFFFFFFFu, h2, o);
                h3 ^= __shfl_xor_sync(0xFFFFFFFFu, h3, o);
            }
            if (lane == 0) {
                unsigned long long* hp = reinterpret_cast<unsigned long long*>(&p.hash[(size_t)pi * 4]);
                if (RELAX_WARPS == 1) { hp[0] = h0; hp[1] = h1; hp[2] = h2; hp[3] = h3; }
                else { atomicXor(hp, h0); atomicXor(hp + 1, h1); atomicXor(hp + 2, h2); atomicXor(hp + 3, h3); }
            }
        }
    }
}

}  // namespace hn
