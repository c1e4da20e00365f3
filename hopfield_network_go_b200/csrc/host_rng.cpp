// host_rng.cpp — host-side restatement of the reference's random stream, so that
// hosts which are not Go can generate the update orders and learning noise the
// reference would draw and upload them (BASELINE north_star: "generated
// host-side from the reference's seeded RNG and uploaded").
//
// Restates golang.org/x/exp/rand @ 642cacee5cc0 (go.mod:12; source not in the
// reference tree): PCG XSL-RR 128/64 source, Rand.Uint64n / Shuffle / Float64 /
// NormFloat64.  Reference call sites: hopfieldutils/HopfieldUtils.go:35-39,
// noiseapplication/NoiseApplication.go:64-105, HopfieldNetworkBuilder.go:231-232.
// A Go host does not need this file: it calls x/exp/rand directly.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/hopfield_b200.h"

namespace hn { void set_error(const char* fmt, ...); }

namespace {

using u128 = unsigned __int128;

struct Pcg {
    u128 state;
    static constexpr u128 kMul = ((u128)2549297995355413924ULL << 64) | 4865540595714422341ULL;
    static constexpr u128 kInc = ((u128)6364136223846793005ULL << 64) | 1442695040888963407ULL;
    explicit Pcg(uint64_t seed) : state(((u128)seed << 64) | seed) {}  // Seed: low = high = seed
    uint64_t next() {
        state = state * kMul + kInc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const unsigned rot = (unsigned)(hi >> 58);
        const uint64_t x = hi ^ lo;
        return (x >> rot) | (x << ((64u - rot) & 63u));
    }
    uint64_t below(uint64_t n) {  // Rand.Uint64n
        if ((n & (n - 1)) == 0) return next() & (n - 1);
        uint64_t v = next();
        if (v > UINT64_MAX - n) {
            const uint64_t ceiling = UINT64_MAX - UINT64_MAX % n;
            while (v >= ceiling) v = next();
        }
        return v % n;
    }
    double unit() {  // Rand.Float64
        for (;;) {
            const double f = (double)below(1ULL << 53) / 9007199254740992.0;
            if (f != 1.0) return f;
        }
    }
    template <class T> void shuffle(T* list, int n) {  // Rand.Shuffle (Fisher-Yates from the top)
        for (int i = n - 1; i > 0; i--) {
            const int j = (int)below((uint64_t)i + 1);
            T t = list[i]; list[i] = list[j]; list[j] = t;
        }
    }
};

// Ziggurat tables for NormFloat64 (Marsaglia & Tsang 2000, 128 strips); Go ships
// them as float32/uint32 literals of the same construction.
struct Zig {
    uint32_t kn[128]; float wn[128], fn[128];
    static constexpr double rn = 3.442619855899;
    Zig() {
        const double m1 = 2147483648.0;
        double dn = rn, tn = dn, vn = 9.91256303526217e-3;
        const double q = vn / std::exp(-.5 * dn * dn);
        kn[0] = (uint32_t)((dn / q) * m1); kn[1] = 0;
        wn[0] = (float)(q / m1); wn[127] = (float)(dn / m1);
        fn[0] = 1.0f; fn[127] = (float)std::exp(-.5 * dn * dn);
        for (int i = 126; i >= 1; i--) {
            dn = std::sqrt(-2. * std::log(vn / dn + std::exp(-.5 * dn * dn)));
            kn[i + 1] = (uint32_t)((dn / tn) * m1);
            tn = dn;
            fn[i] = (float)std::exp(-.5 * dn * dn);
            wn[i] = (float)(dn / m1);
        }
    }
};
const Zig& zig() { static Zig z; return z; }

double norm_float64(Pcg& g) {
    const Zig& z = zig();
    for (;;) {
        const int32_t j = (int32_t)(uint32_t)(g.next() >> 32);
        const int32_t i = j & 0x7F;
        double x = (double)j * (double)z.wn[i];
        const uint32_t aj = j < 0 ? (uint32_t)(-(int64_t)j) : (uint32_t)j;
        if (aj < z.kn[i]) return x;
        if (i == 0) {
            for (;;) {
                x = -std::log(g.unit()) * (1.0 / Zig::rn);
                const double y = -std::log(g.unit());
                if (y + y >= x * x) break;
            }
            return j > 0 ? Zig::rn + x : -Zig::rn - x;
        }
        if (z.fn[i] + (float)g.unit() * (z.fn[i - 1] - z.fn[i]) < (float)std::exp(-.5 * x * x)) return x;
    }
}

// maximalRatioInvertSliceElements (NoiseApplication.go:64-75): the index list has
// length Len()-1, so the last unit is never inverted; count = int(N*ratio).
void maximal_invert(Pcg& g, double* state, int n, double ratio) {
    const int num = (int)((double)n * ratio);
    const int m = n - 1;
    std::vector<int32_t> idx((size_t)(m > 0 ? m : 0));
    for (int i = 0; i < m; i++) idx[i] = i;
    g.shuffle(idx.data(), m);
    for (int i = 0; i < num && i < m; i++) state[idx[i]] = -1 * state[idx[i]];
}

}  // namespace

struct hn_rng_s { Pcg g; explicit hn_rng_s(uint64_t s) : g(s) {} };

namespace hn {
// Jump-ahead of the 128-bit LCG under the PCG source: entry j holds (M^(2^j), I*(M^(2^j)-1)/(M-1)) as
// {mult lo, mult hi, plus lo, plus hi}; advancing by d composes the entries of the set bits of d.
void rng_jump_table(uint64_t table[64][4]) {
    u128 m = Pcg::kMul, c = Pcg::kInc;
    for (int j = 0; j < 64; j++) {
        table[j][0] = (uint64_t)m; table[j][1] = (uint64_t)(m >> 64);
        table[j][2] = (uint64_t)c; table[j][3] = (uint64_t)(c >> 64);
        c = (m + 1) * c;
        m = m * m;
    }
}
void rng_get_state(hn_rng r, uint64_t out[2]) { out[0] = (uint64_t)r->g.state; out[1] = (uint64_t)(r->g.state >> 64); }
void rng_advance(hn_rng r, uint64_t draws) {   // as if `draws` Uint64 values had been drawn
    uint64_t t[64][4];
    rng_jump_table(t);
    u128 s = r->g.state;
    for (int j = 0; j < 64; j++)
        if ((draws >> j) & 1ULL) s = s * (((u128)t[j][1] << 64) | t[j][0]) + (((u128)t[j][3] << 64) | t[j][2]);
    r->g.state = s;
}
}  // namespace hn

extern "C" {

// Host-side merge of per-GPU unique-attractor tables (multi-GPU probe sharding): groups entries by their
// 128-bit canonical-state hash; a group's first index is the minimum, its hits the sum; groups are returned in
// ascending first index = the first-seen order of handleUniqueRelaxedState (UniqueRelaxedStateHandler.go:41-66)
// over the whole probe set.  O(n) open addressing; pure host code.
int hn_merge_unique(int64_t n, const int64_t* first, const int32_t* hits, const uint64_t* hash,
                    int64_t* first_out, int64_t* hits_out, int64_t* rep_out, int64_t* n_groups_out) {
    if (n < 0 || !n_groups_out || (n > 0 && (!first || !hits || !hash || !first_out || !hits_out || !rep_out))) {
        hn::set_error("hn_merge_unique: bad argument");
        return HN_E_INVALID;
    }
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int64_t> slot, gfirst, ghits, grep;
    try {
        slot.assign(cap, -1);
        gfirst.reserve((size_t)n); ghits.reserve((size_t)n); grep.reserve((size_t)n);
    } catch (const std::bad_alloc&) { hn::set_error("hn_merge_unique: out of host memory"); return HN_E_OOM; }
    auto home = [&](int64_t i) { return (size_t)((hash[2 * i] ^ (hash[2 * i + 1] * 0x9E3779B97F4A7C15ULL)) & (cap - 1)); };
    constexpr int64_t kAhead = 16;   // the table is far larger than the caches: fetch the home slots ahead
    for (int64_t i = 0; i < n; i++) {
        if (i + kAhead < n) __builtin_prefetch(&slot[home(i + kAhead)], 1, 0);
        const uint64_t h0 = hash[2 * i], h1 = hash[2 * i + 1];
        size_t s = home(i);
        for (;;) {
            const int64_t g = slot[s];
            if (g < 0) {
                slot[s] = (int64_t)gfirst.size();
                gfirst.push_back(first[i]); ghits.push_back(hits[i]); grep.push_back(i);
                break;
            }
            const int64_t r = grep[(size_t)g];
            if (hash[2 * r] == h0 && hash[2 * r + 1] == h1) {
                ghits[(size_t)g] += hits[i];
                if (first[i] < gfirst[(size_t)g]) { gfirst[(size_t)g] = first[i]; grep[(size_t)g] = i; }
                break;
            }
            s = (s + 1) & (cap - 1);
        }
    }
    // NOTE: a group's representative may move to a later entry (smaller first index), but the hash it is compared by is
    // the same for every member, so the probe loop above stays correct.
    const size_t ng = gfirst.size();
    std::vector<int64_t> order(ng);
    for (size_t g = 0; g < ng; g++) order[g] = (int64_t)g;
    bool ascending = true;
    for (size_t g = 1; g < ng; g++) if (gfirst[g] < gfirst[g - 1]) { ascending = false; break; }
    if (!ascending)
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return gfirst[(size_t)a] < gfirst[(size_t)b]; });
    for (size_t j = 0; j < ng; j++) {
        const size_t g = (size_t)order[j];
        first_out[j] = gfirst[g]; hits_out[j] = ghits[g]; rep_out[j] = grep[g];
    }
    *n_groups_out = (int64_t)ng;
    return HN_OK;
}

int hn_rng_create(uint64_t seed, hn_rng* out) {
    if (!out) { hn::set_error("hn_rng_create: out is NULL"); return HN_E_INVALID; }
    *out = new (std::nothrow) hn_rng_s(seed);
    return *out ? HN_OK : HN_E_OOM;
}
int hn_rng_destroy(hn_rng r) { delete r; return HN_OK; }
uint64_t hn_rng_uint64(hn_rng r) { return r->g.next(); }
uint64_t hn_rng_uint64n(hn_rng r, uint64_t n) { return r->g.below(n); }
double hn_rng_float64(hn_rng r) { return r->g.unit(); }
double hn_rng_norm_float64(hn_rng r) { return norm_float64(r->g); }

int hn_rng_advance(hn_rng r, uint64_t draws) {
    if (!r) { hn::set_error("hn_rng_advance: rng is NULL"); return HN_E_INVALID; }
    hn::rng_advance(r, draws);
    return HN_OK;
}

int hn_rng_shuffle_u16(hn_rng r, uint16_t* list, int32_t n) {
    if (!r || !list || n < 0) { hn::set_error("hn_rng_shuffle_u16: bad argument"); return HN_E_INVALID; }
    r->g.shuffle(list, n);
    return HN_OK;
}

int hn_rng_perm_pool(hn_rng r, int32_t dimension, int32_t n_rows, int32_t start_identity,
                     uint16_t* list_io, uint16_t* pool_out) {
    if (!r || !list_io || !pool_out || dimension <= 0 || dimension > 65535 || n_rows < 0) {
        hn::set_error("hn_rng_perm_pool: bad argument");
        return HN_E_INVALID;
    }
    if (start_identity)
        for (int i = 0; i < dimension; i++) list_io[i] = (uint16_t)i;  // getUnitIndices, HopfieldNetwork.go:71-77
    for (int row = 0; row < n_rows; row++) {
        r->g.shuffle(list_io, dimension);  // ShuffleList on the persistent list, :393
        std::memcpy(pool_out + (size_t)row * dimension, list_io, sizeof(uint16_t) * (size_t)dimension);
    }
    return HN_OK;
}

int hn_rng_apply_noise(hn_rng r, int32_t noise_method, double noise_scale, double* state, int32_t n) {
    if (!r || !state || n <= 0) { hn::set_error("hn_rng_apply_noise: bad argument"); return HN_E_INVALID; }
    switch (noise_method) {
    case HN_NOISE_NONE: break;  // noNoiseApplication :51
    case HN_NOISE_MAXIMAL_INVERSION: maximal_invert(r->g, state, n, noise_scale); break;
    case HN_NOISE_RANDOM_SUBMAXIMAL_INVERSION: {  // :87-90
        const double sel = std::fmod(r->g.unit(), noise_scale);
        maximal_invert(r->g, state, n, sel);
        break;
    }
    case HN_NOISE_GAUSSIAN:  // :101-105
        for (int i = 0; i < n; i++) state[i] += norm_float64(r->g) * noise_scale;
        break;
    default:
        hn::set_error("hn_rng_apply_noise: unknown noise method %d", noise_method);
        return HN_E_INVALID;
    }
    return HN_OK;
}

int hn_rng_generate_states(hn_rng r, int32_t dimension, int32_t count, double rand_min, double rand_max,
                           uint64_t* packed_out) {
    if (!r || !packed_out || dimension <= 0 || count < 0 || !(rand_min < rand_max)) {
        hn::set_error("hn_rng_generate_states: bad argument");
        return HN_E_INVALID;
    }
    const int words = (dimension + 63) / 64;
    std::memset(packed_out, 0, sizeof(uint64_t) * (size_t)count * words);
    for (int s = 0; s < count; s++)
        for (int i = 0; i < dimension; i++) {
            const double x = r->g.unit() * (rand_max - rand_min) + rand_min;  // distuv.Uniform.Rand
            if (x > 0.0) packed_out[(size_t)s * words + (i >> 6)] |= 1ULL << (i & 63);  // activation: x <= 0 -> low
        }
    return HN_OK;
}

}  // extern "C"
