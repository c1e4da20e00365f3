/*
 * oracle/xrand.c — TEST INFRASTRUCTURE.  See xrand.h for provenance.
 * Restates golang.org/x/exp/rand (PCG source, Uint64n, Shuffle, Float64,
 * NormFloat64) and the two gonum distuv samplers the reference uses.
 */
#include "xrand.h"
#include <math.h>

typedef unsigned __int128 u128;

/* PCG default 128-bit LCG constants (x/exp/rand rng.go: multiplier / increment):
 *   multiplier = 47026247687942121848144207491837523525
 *              = (2549297995355413924 << 64) | 4865540595714422341
 *   increment  = 117397592171526113268558934119004209487
 *              = (6364136223846793005 << 64) | 1442695040888963407            */
#define PCG_MUL_HI 2549297995355413924ULL
#define PCG_MUL_LO 4865540595714422341ULL
#define PCG_INC_HI 6364136223846793005ULL
#define PCG_INC_LO 1442695040888963407ULL

/* PCGSource.Seed: low = seed; high = seed. */
void orc_rng_seed(orc_rng* r, uint64_t seed) { r->low = seed; r->high = seed; }

/* PCGSource.Uint64: state = state*mul + inc; return rotr64(high ^ low, high >> 58). */
uint64_t orc_rng_uint64(orc_rng* r) {
    u128 s   = ((u128)r->high << 64) | r->low;
    u128 mul = ((u128)PCG_MUL_HI << 64) | PCG_MUL_LO;
    u128 inc = ((u128)PCG_INC_HI << 64) | PCG_INC_LO;
    s = s * mul + inc;
    r->low  = (uint64_t)s;
    r->high = (uint64_t)(s >> 64);
    uint64_t x = r->high ^ r->low;
    unsigned rot = (unsigned)(r->high >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

/* Rand.Uint64n: mask for powers of two, else rejection of the biased top range. */
uint64_t orc_rng_uint64n(orc_rng* r, uint64_t n) {
    if ((n & (n - 1)) == 0) { /* power of two (n == 0 panics in Go) */
        return orc_rng_uint64(r) & (n - 1);
    }
    uint64_t v = orc_rng_uint64(r);
    if (v > UINT64_MAX - n) {
        uint64_t ceiling = UINT64_MAX - UINT64_MAX % n;
        while (v >= ceiling) v = orc_rng_uint64(r);
    }
    return v % n;
}

/* Rand.Uint32 = uint32(Uint64() >> 32). */
uint32_t orc_rng_uint32(orc_rng* r) { return (uint32_t)(orc_rng_uint64(r) >> 32); }

/* Rand.Float64 = float64(Uint64n(1<<53)) / (1<<53), redrawn if it rounds to 1. */
double orc_rng_float64(orc_rng* r) {
    for (;;) {
        double f = (double)orc_rng_uint64n(r, 1ULL << 53) / (double)(1ULL << 53);
        if (f != 1.0) return f;
    }
}

/* Rand.Shuffle: Fisher-Yates from the top, j = Int31n(i+1) = Uint64n(i+1). */
void orc_rng_shuffle_i32(orc_rng* r, int32_t* list, int32_t n) {
    for (int32_t i = n - 1; i > 0; i--) {
        int32_t j = (int32_t)orc_rng_uint64n(r, (uint64_t)(i + 1));
        int32_t t = list[i]; list[i] = list[j]; list[j] = t;
    }
}
void orc_rng_shuffle_u16(orc_rng* r, uint16_t* list, int32_t n) {
    for (int32_t i = n - 1; i > 0; i--) {
        int32_t j = (int32_t)orc_rng_uint64n(r, (uint64_t)(i + 1));
        uint16_t t = list[i]; list[i] = list[j]; list[j] = t;
    }
}

/* ---- NormFloat64: Marsaglia-Tsang ziggurat, 128 strips -----------------
 * Go ships the kn/wn/fn tables as literals; they are the float32/uint32
 * roundings of the tables produced by the published generator below
 * (Marsaglia & Tsang 2000, zigset), which we run once.                      */
#define ZIG_RN 3.442619855899
static uint32_t zig_kn[128];
static float    zig_wn[128], zig_fn[128];
static int      zig_ready = 0;

static void zig_init(void) {
    const double m1 = 2147483648.0;
    double dn = ZIG_RN, tn = dn, vn = 9.91256303526217e-3;
    double q = vn / exp(-.5 * dn * dn);
    zig_kn[0] = (uint32_t)((dn / q) * m1);
    zig_kn[1] = 0;
    zig_wn[0] = (float)(q / m1);
    zig_wn[127] = (float)(dn / m1);
    zig_fn[0] = 1.0f;
    zig_fn[127] = (float)exp(-.5 * dn * dn);
    for (int i = 126; i >= 1; i--) {
        dn = sqrt(-2. * log(vn / dn + exp(-.5 * dn * dn)));
        zig_kn[i + 1] = (uint32_t)((dn / tn) * m1);
        tn = dn;
        zig_fn[i] = (float)exp(-.5 * dn * dn);
        zig_wn[i] = (float)(dn / m1);
    }
    zig_ready = 1;
}

static uint32_t abs_i32(int32_t j) { return j < 0 ? (uint32_t)(-(int64_t)j) : (uint32_t)j; }

double orc_rng_norm_float64(orc_rng* r) {
    if (!zig_ready) zig_init();
    for (;;) {
        int32_t j = (int32_t)orc_rng_uint32(r);
        int32_t i = j & 0x7F;
        double x = (double)j * (double)zig_wn[i];
        if (abs_i32(j) < zig_kn[i]) return x;
        if (i == 0) {
            for (;;) {
                x = -log(orc_rng_float64(r)) * (1.0 / ZIG_RN);
                double y = -log(orc_rng_float64(r));
                if (y + y >= x * x) break;
            }
            return j > 0 ? ZIG_RN + x : -ZIG_RN - x;
        }
        if (zig_fn[i] + (float)orc_rng_float64(r) * (zig_fn[i - 1] - zig_fn[i]) <
            (float)exp(-.5 * x * x))
            return x;
    }
}

/* distuv.Uniform.Rand: rand.New(Src).Float64()*(Max-Min) + Min */
double orc_rng_uniform(orc_rng* r, double mn, double mx) {
    return orc_rng_float64(r) * (mx - mn) + mn;
}
/* distuv.Normal.Rand: rand.New(Src).NormFloat64()*Sigma + Mu */
double orc_rng_normal(orc_rng* r, double mu, double sigma) {
    return orc_rng_norm_float64(r) * sigma + mu;
}
