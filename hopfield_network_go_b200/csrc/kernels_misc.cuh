// kernels_misc.cuh — HBM-bound elementwise / reduction kernels around the contractions:
// weight update + constraints + Frobenius normalisation, energy epilogue,
// permutation inversion, bit-matrix transpose, unique-attractor table.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "relax.cuh"

namespace hn {

// ---- W <- W + lr*dW  (LearningRule.go:72-73 / :111-112) --------------------
// updatedMatrix.Scale(lr) rounds once, matrix.Add rounds once: keep both
// roundings (no FMA) so non-dyadic learning rates stay bit-identical.
__global__ void apply_update_kernel(double* __restrict__ w, const double* __restrict__ dw, int n, int ld,
                                    double lr) {
    const int64_t total = (int64_t)n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % ld);
        if (c < n) w[i] = __dadd_rn(w[i], __dmul_rn(lr, dw[i]));
    }
}

// ---- int8 operands of the learning contractions (tcgen05 path) ---------------------------------------------------
// out[unit][mu] = va(bitA[unit][mu]) - (bitB ? vb(bitB[unit][mu]) : 0) for mu < P, 0 up to the pitch.  The bit matrices
// are unit-major ([N][pwords], transpose_bits_kernel); the value maps are integers in {-1, 0, 1}, so the entries are
// integers in [-2, 2]: t (Hebbian / the t^T factor) or t - r (twice the Delta rule's 0.5 (t - r)).
__global__ void expand_units_i8_kernel(const uint64_t* __restrict__ bitsA, const uint64_t* __restrict__ bitsB, int n, int pwords,
                                       int P, int pitch, int alo, int ahi, int blo, int bhi, int8_t* __restrict__ out) {
    const int64_t total = (int64_t)n * (pitch / 8);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int unit = (int)(t / (pitch / 8)), m0 = (int)(t % (pitch / 8)) * 8;
        uint64_t pack = 0;
        if (m0 < P) {
            const unsigned a = (unsigned)((bitsA[(size_t)unit * pwords + (m0 >> 6)] >> (m0 & 63)) & 0xFFu);
            const unsigned b = bitsB ? (unsigned)((bitsB[(size_t)unit * pwords + (m0 >> 6)] >> (m0 & 63)) & 0xFFu) : 0u;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                int v = 0;
                if (m0 + e < P) v = (((a >> e) & 1u) ? ahi : alo) - (bitsB ? (((b >> e) & 1u) ? bhi : blo) : 0);
                pack |= (uint64_t)(uint8_t)(int8_t)v << (8 * e);
            }
        }
        reinterpret_cast<uint64_t*>(out)[t] = pack;
    }
}

// which fractions occur: off <- v is not a multiple of 1/4; fr |= 1 (not an even integer) | 2 (not an integer) | 4 (not a
// multiple of 1/2).  One conversion each way instead of four rint (FP64 pipe): q = 4v as an integer when it is one.
__device__ __forceinline__ void frac_scan(double v, bool& off, int& fr) {
    const double v4 = 4.0 * v;
    if (fabs(v4) < 4503599627370496.0) {
        const long long q = __double2ll_rn(v4);
        if ((double)q != v4) { off = true; fr |= 7; }
        else fr |= ((q & 7) ? 1 : 0) | ((q & 3) ? 2 : 0) | ((q & 1) ? 4 : 0);
    } else { off = true; fr |= 7; }   // beyond 2^50: no exact-arithmetic claim
}

// ---- W <- W + lr*(s*dWi), then enforceConstraints, in ONE pass over tile pairs (LearningRule.go:72-74 / :111-113) ----
// dWi: the exact integer contraction (all-reduced over the ranks), s = 1 (Hebbian) or 0.5 (Delta): s*dWi is the
// reference's updatedMatrix entry exactly, fl(lr * .) its Scale, fl(W + .) its Add; zero diagonal, then 0.5*(W + W^T)
// from the updated values.  Also the max |W| and the "every entry a multiple of 1/4" flag ensure_margin needs.
__global__ void update_constraints_i32_kernel(double* __restrict__ w, const int32_t* __restrict__ dwi, int n, int ld, double lr,
                                              double s, int zero_diag, int symmetric, unsigned long long* maxabs_bits,
                                              int32_t* not_quarters, int32_t* frac_flags) {
    __shared__ double ta[32][33], tb[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        double a = 0.0;
        if (i < n && j < n) a = __dadd_rn(w[(size_t)i * ld + j], __dmul_rn(lr, s * (double)dwi[(size_t)i * ld + j]));
        if (zero_diag && i == j) a = 0.0;
        ta[r][tx] = a;
        const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
        double b = 0.0;
        if (i2 < n && j2 < n) b = __dadd_rn(w[(size_t)i2 * ld + j2], __dmul_rn(lr, s * (double)dwi[(size_t)i2 * ld + j2]));
        if (zero_diag && i2 == j2) b = 0.0;
        tb[r][tx] = b;
    }
    __syncthreads();
    double m = 0.0;
    bool off = false;
    int fr = 0;   // 1: some entry is not an even integer, 2: not an integer, 4: not a multiple of 1/2
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        if (i < n && j < n) {
            const double a = ta[r][tx];
            const double v = symmetric ? __dmul_rn(0.5, __dadd_rn(a, tb[tx][r])) : a;
            w[(size_t)i * ld + j] = v;
            m = fmax(m, fabs(v)); frac_scan(v, off, fr);
        }
        if (bi != bj) {
            const int i2 = bj * 32 + r, j2 = bi * 32 + tx;
            if (i2 < n && j2 < n) {
                const double b = tb[r][tx];
                const double v = symmetric ? __dmul_rn(0.5, __dadd_rn(b, ta[tx][r])) : b;
                w[(size_t)i2 * ld + j2] = v;
                m = fmax(m, fabs(v)); frac_scan(v, off, fr);
            }
        }
    }
    if (fr) atomicOr(frac_flags, fr);
    if (off) *not_quarters = 1;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if (tx == 0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(m));
}

// ---- enforceConstraints (HopfieldNetwork.go:55-67) --------------------------
// zero diagonal first, then W <- 0.5*(W + W^T) from the OLD values (gonum's Add
// against the transposed view goes through an aliasing workspace).
// One CTA per tile pair (bi <= bj); both tiles staged in shared memory so every
// global access is coalesced.
constexpr int SYM_T = 32;
__global__ void constraints_kernel(double* __restrict__ w, int n, int ld, int zero_diag, int symmetric) {
    __shared__ double ta[SYM_T][SYM_T + 1], tb[SYM_T][SYM_T + 1];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bi * SYM_T + r, j = bj * SYM_T + tx;
        double a = (i < n && j < n) ? w[(size_t)i * ld + j] : 0.0;
        if (zero_diag && i == j) a = 0.0;
        ta[r][tx] = a;
        const int i2 = bj * SYM_T + r, j2 = bi * SYM_T + tx;
        double b = (i2 < n && j2 < n) ? w[(size_t)i2 * ld + j2] : 0.0;
        if (zero_diag && i2 == j2) b = 0.0;
        tb[r][tx] = b;
    }
    __syncthreads();
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bi * SYM_T + r, j = bj * SYM_T + tx;
        if (i < n && j < n) {
            const double a = ta[r][tx];
            w[(size_t)i * ld + j] = symmetric ? __dmul_rn(0.5, __dadd_rn(a, tb[tx][r])) : a;
        }
        if (bi != bj) {
            const int i2 = bj * SYM_T + r, j2 = bi * SYM_T + tx;
            if (i2 < n && j2 < n) {
                const double b = tb[r][tx];
                w[(size_t)i2 * ld + j2] = symmetric ? __dmul_rn(0.5, __dadd_rn(b, ta[tx][r])) : b;
            }
        }
    }
}

// flag <- 1 if W_ij != W_ji anywhere (exact comparison)
__global__ void asymmetry_kernel(const double* __restrict__ w, int n, int ld, int32_t* flag) {
    __shared__ double t[SYM_T][SYM_T + 1];
    const int bi = blockIdx.y, bj = blockIdx.x, tx = threadIdx.x, ty = threadIdx.y;
    if (bj < bi) return;
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bj * SYM_T + r, j = bi * SYM_T + tx;
        t[r][tx] = (i < n && j < n) ? w[(size_t)i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bi * SYM_T + r, j = bj * SYM_T + tx;
        if (i < n && j < n && w[(size_t)i * ld + j] != t[tx][r]) *flag = 1;
    }
}

// WT <- W^T (kept resident when W is not symmetric: a flip of unit k needs column k)
__global__ void transpose_kernel(const double* __restrict__ w, double* __restrict__ wt, int n, int ld) {
    __shared__ double t[SYM_T][SYM_T + 1];
    const int bi = blockIdx.y, bj = blockIdx.x, tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bi * SYM_T + r, j = bj * SYM_T + tx;
        t[r][tx] = (i < n && j < n) ? w[(size_t)i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < SYM_T; r += blockDim.y) {
        const int i = bj * SYM_T + r, j = bi * SYM_T + tx;
        if (i < n && j < n) wt[(size_t)i * ld + j] = t[tx][r];
    }
}

// ---- Frobenius norm: deterministic two-pass reduction -----------------------
constexpr int RED_THREADS = 256;
__device__ __forceinline__ double block_reduce_sum_det(double v, double* sm) {
    // fixed-order tree: identical result for identical launch geometry
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)blockDim.x / 32; w++) r += sm[w];
    __syncthreads();
    return r;  // valid in thread 0
}
__global__ void sumsq_partial_kernel(const double* __restrict__ w, int n, int ld, double* __restrict__ partial) {
    __shared__ double sm[RED_THREADS / 32];
    double acc = 0.0;
    const int64_t total = (int64_t)n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double x = w[i];  // padding columns are zero
        acc = fma(x, x, acc);
    }
    const double r = block_reduce_sum_det(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
// out[0] = sqrt(sum partial), out[1] = 1/out[0]
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int np, double* __restrict__ out) {
    __shared__ double sm[RED_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += partial[i];
    const double r = block_reduce_sum_det(acc, sm);
    if (threadIdx.x == 0) { const double nrm = sqrt(r); out[0] = nrm; out[1] = 1.0 / nrm; }
}
// W <- fl(f * W), f = 1/||W||_F  (matrix.Scale(1/matrix.Norm(2), matrix), HopfieldNetwork.go:260)
__global__ void scale_kernel(double* __restrict__ w, int n, int ld, const double* __restrict__ f) {
    const double s = f[1];
    const int64_t total = (int64_t)n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x)
        w[i] = __dmul_rn(s, w[i]);
}

// Dense [n][n] int16 image of 4*W for the host's exact Frobenius norm (api.cu: exact_frobenius_quarters); bad <- 1 if some
// entry is not a multiple of 0.25 or 4|w| leaves int16.
__global__ void quarters_i16_kernel(const double* __restrict__ w, int n, int ld, int16_t* __restrict__ out, int32_t* bad) {
    const int64_t total = (int64_t)n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = 4.0 * w[(i / n) * ld + (i % n)];
        const double r = rint(x);
        if (r != x || fabs(r) > 32767.0) *bad = 1;
        out[i] = (int16_t)r;
    }
}

// max_i sum_k |W_ik|  (the R of the decision margin); one CTA per row, atomicMax on the bits
__global__ void rowabs_max_kernel(const double* __restrict__ w, int n, int ld, unsigned long long* out) {
    __shared__ double sm[RED_THREADS / 32];
    const double* row = w + (size_t)blockIdx.x * ld;
    double acc = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) acc += fabs(row[j]);
    const double r = block_reduce_sum_det(acc, sm);
    if (threadIdx.x == 0) atomicMax(out, (unsigned long long)__double_as_longlong(r));
}

// max_ij |W_ij| (per-flip rounding bound of the float32 column stream); atomicMax on the bits.
// not_quarters (optional) <- 1 if some entry is not an integer multiple of 0.25 (exact-arithmetic test, ensure_margin)
__global__ void maxabs_kernel(const double* __restrict__ w, int n, int ld, unsigned long long* out, int32_t* not_quarters) {
    double m = 0.0;
    bool off = false;
    const int64_t total = (int64_t)n * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = w[i];
        m = fmax(m, fabs(x));
        off |= rint(4.0 * x) != 4.0 * x;
    }
    if (not_quarters && off) *not_quarters = 1;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}
// float32 (round-to-nearest) shadow of a padded float64 matrix, same leading dimension
__global__ void to_f32_kernel(const double* __restrict__ w, float* __restrict__ w32, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        w32[i] = __double2float_rn(w[i]);
}

// thermal Delta factors (LearningRule.go:134,149-150): f[s] = exp(-wf * ||H[s,:]||_2), wf = 1/(1 + ||W||_F)
// nrm[0] holds ||W||_F (sumsq_final_kernel). One CTA per state.
__global__ void __launch_bounds__(RED_THREADS)
thermal_factor_kernel(const double* __restrict__ H, int ldh, int n, const double* __restrict__ nrm, double* __restrict__ f) {
    __shared__ double sm[RED_THREADS / 32];
    const int s = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += RED_THREADS) { const double x = H[(size_t)s * ldh + i]; acc = fma(x, x, acc); }
    const double r = block_reduce_sum_det(acc, sm);
    if (threadIdx.x == 0) f[s] = exp(-1.0 * (1.0 / (1.0 + nrm[0])) * sqrt(r) / 1.0);
}

// int16 image M of the weights for the exact integer stream: M = 2*K with K = scale*Wraw integer valued, and
// M_kk = 2*K_kk + diag (diag = -2/delta: -1 bipolar, -2 binary) so that a flip of unit k moves its own tracked field
// U_k = 2 g_k - 2 b_k by exactly Delta_s*M_kk = 2 Delta_s K_kk - 2 Delta_b: the column update needs no owner fix-up.
// flag <- 1 if some scale*w is not an integer or M leaves int16.  Row-major [n][ld] in both layouts (diagonal = [k][k]).
__global__ void to_i16_kernel(const double* __restrict__ w, int16_t* __restrict__ k16, int64_t total, int ld, double scale,
                              int diag, int32_t* flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = scale * w[i];
        const double r = rint(x);
        double m = 2.0 * r;
        if (i / ld == i % ld) m += (double)diag;
        if (r != x || fabs(m) > 32767.0) *flag = 1;
        k16[i] = (int16_t)(long long)m;
    }
}

// byte image of the int16 image (every |M| <= 127), four entries per thread
__global__ void i16_to_i8_kernel(const int16_t* __restrict__ k16, int8_t* __restrict__ k8, int64_t quads) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (int64_t)gridDim.x * blockDim.x) {
        const short4 v = reinterpret_cast<const short4*>(k16)[i];
        reinterpret_cast<char4*>(k8)[i] = make_char4((signed char)v.x, (signed char)v.y, (signed char)v.z, (signed char)v.w);
    }
}

// ---- integer structure of caller-supplied weights (hn_set_weights: resume from matrix.bin, SetMatrix) ----
// smallest non-zero |W_ij| (positive doubles order like their bit patterns); out starts at +inf bits
__global__ void min_nonzero_abs_kernel(const double* __restrict__ w, int64_t total, unsigned long long* out) {
    unsigned long long m = 0x7FF0000000000000ULL;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = fabs(w[i]);
        if (x > 0.0) m = min(m, (unsigned long long)__double_as_longlong(x));
    }
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out, m);
}
// hypothesis W = fl(c * K) with K integer: K = rint(W / c) into kout (as float64); flag <- 1 where fl(c*K) != W bit for
// bit or |K| leaves the int16 image's range.  One rounding per element is exactly what Scale(1/norm) produced.
__global__ void recover_integer_kernel(const double* __restrict__ w, double* __restrict__ kout, int64_t total, double c,
                                       int32_t* flag) {
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = w[i];
        const double k = rint(x / c);
        if (__dmul_rn(c, k) != x || fabs(k) > 16383.0) bad = true;
        kout[i] = k;
    }
    if (bad) *flag = 1;
}

// ---- inverse permutations: inv[row][perm[row][i]] = i -----------------------
__global__ void invert_perm_kernel(const uint16_t* __restrict__ perm, uint16_t* __restrict__ inv, int n,
                                   int ldp, int rows) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
        const int u = perm[(size_t)r * ldp + i];
        if (u < n) inv[(size_t)r * ldp + u] = (uint16_t)i;   // entries >= n are reported by check_perm_kernel
    }
}
// every row must be a permutation of 0..n-1 (a bad row would send the relaxation kernel out of bounds): after the
// inversion, perm[inv[u]] == u for every unit u exactly when the row is one.  bad <- 1 + index of an offending row.
__global__ void check_perm_kernel(const uint16_t* __restrict__ perm, const uint16_t* __restrict__ inv, int n, int ldp, int rows,
                                  int32_t* bad) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    for (int u = blockIdx.y * blockDim.x + threadIdx.x; u < n; u += gridDim.y * blockDim.x) {
        const int pos = inv[(size_t)r * ldp + u];
        if (perm[(size_t)r * ldp + u] >= n || pos >= n || perm[(size_t)r * ldp + pos] != u) atomicMax(bad, r + 1);
    }
}

// ---- bit-matrix transpose: X[P][words] (state-major) -> XT[N][pwords] (unit-major)
__global__ void transpose_bits_kernel(const uint64_t* __restrict__ x, uint64_t* __restrict__ xt, int n,
                                      int words, int p, int pwords) {
    const int64_t total = (int64_t)n * pwords;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
         o += (int64_t)gridDim.x * blockDim.x) {
        const int unit = (int)(o / pwords), q = (int)(o % pwords);
        uint64_t out = 0;
        const int m0 = q * 64;
        for (int b = 0; b < 64 && m0 + b < p; b++)
            out |= ((x[(size_t)(m0 + b) * words + (unit >> 6)] >> (unit & 63)) & 1ULL) << b;
        xt[o] = out;
    }
}

// ---- energy profile epilogue (AllUnitEnergies / StateIsStable / StateEnergy) ----
// BipolarDomainManager.go:39-55, BinaryDomainManager.go:54-72, HopfieldNetwork.go:214-225.
// H holds the fields from the GEMM; e = -0.5 * (h * v), v = s (bipolar) or 2s-1 (binary).
// A unit whose |h| is inside the margin is re-evaluated in the reference's order.
// One CTA per state.
__global__ void __launch_bounds__(RED_THREADS)
energy_epilogue_kernel(const double* __restrict__ H, int ldh, const uint64_t* __restrict__ states, int words,
                       const double* __restrict__ W, int ldw, int n, int domain, double tau,
                       int max_unstable, double* __restrict__ energies, int32_t* __restrict__ unstable,
                       uint8_t* __restrict__ stable, double* __restrict__ state_energy,
                       int32_t* __restrict__ margin_hits) {
    __shared__ double sm[RED_THREADS / 32];
    __shared__ int s_cnt, s_margin;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(states + (size_t)s * words);
    if (tid == 0) { s_cnt = 0; s_margin = 0; }
    __syncthreads();
    double esum = 0.0;
    int cnt = 0;
    // pass 1: units outside the margin, strided
    for (int i = tid; i < n; i += RED_THREADS) {
        const double h = H[(size_t)s * ldh + i];
        const uint32_t b = (bits[i >> 5] >> (i & 31)) & 1u;
        if (!(fabs(h) < tau)) {
            const double v = b ? 1.0 : -1.0;
            const double e = -0.5 * (h * v);
            if (energies) energies[(size_t)s * n + i] = e;
            esum += e;
            if (e > 0.0) cnt++;
        }
    }
    // pass 2: units inside the margin, one warp per unit, exact order
    int mcount = 0;
    for (int i0 = warp; i0 < n; i0 += RED_THREADS / 32) {
        // warp-uniform test
        const double h = H[(size_t)s * ldh + i0];
        if (fabs(h) < tau) {
            const double hx = exact_field_warp(W + (size_t)i0 * ldw, bits, n, domain, lane);
            if (lane == 0) {
                const uint32_t b = (bits[i0 >> 5] >> (i0 & 31)) & 1u;
                const double v = b ? 1.0 : -1.0;
                const double e = -0.5 * (hx * v);
                if (energies) energies[(size_t)s * n + i0] = e;
                esum += e;
                if (e > 0.0) cnt++;
                mcount++;
            }
        }
    }
    atomicAdd(&s_cnt, cnt);
    atomicAdd(&s_margin, mcount);
    const double r = block_reduce_sum_det(esum, sm);
    __syncthreads();
    if (tid == 0) {
        if (unstable) unstable[s] = s_cnt;
        if (stable) stable[s] = (uint8_t)(s_cnt <= max_unstable);
        if (state_energy) state_energy[s] = r;
        if (margin_hits) margin_hits[s] = s_margin;
    }
}

// The same from the EXACT integer fields G = M s of the integer tensor path (exact-arithmetic weights, i.e. inside a
// learning run): (M s)_u = 2 g_u + diag * s_u with g = K s and K = scale * W, so the float64 field is h = g / scale
// exactly and e = -0.5 * (h * v) is the reference's value bit for bit; a zero field is an exact zero (energy not > 0).
__global__ void __launch_bounds__(RED_THREADS)
energy_epilogue_i32_kernel(const int32_t* __restrict__ G, int ldg, const uint64_t* __restrict__ states, int words, int n, int domain,
                           double inv_scale, int max_unstable, double* __restrict__ energies, int32_t* __restrict__ unstable,
                           uint8_t* __restrict__ stable, double* __restrict__ state_energy) {
    __shared__ double sm[RED_THREADS / 32];
    __shared__ int s_cnt;
    const int s = blockIdx.x, tid = threadIdx.x;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(states + (size_t)s * words);
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    double esum = 0.0;
    int cnt = 0;
    for (int i = tid; i < n; i += RED_THREADS) {
        const int b = (int)((bits[i >> 5] >> (i & 31)) & 1u);
        // diag * s_u: bipolar -1 * (+-1); binary -2 * (0 / 1)
        const int self = domain == HN_DOMAIN_BINARY ? -2 * b : (b ? -1 : 1);
        const int g = (G[(size_t)s * ldg + i] - self) / 2;   // exact: the difference is even
        const double h = (double)g * inv_scale;
        const double v = b ? 1.0 : -1.0;
        const double e = -0.5 * (h * v);
        if (energies) energies[(size_t)s * n + i] = e;
        esum += e;
        if (e > 0.0) cnt++;
    }
    atomicAdd(&s_cnt, cnt);
    const double r = block_reduce_sum_det(esum, sm);
    __syncthreads();
    if (tid == 0) {
        if (unstable) unstable[s] = s_cnt;
        if (stable) stable[s] = (uint8_t)(s_cnt <= max_unstable);
        if (state_energy) state_energy[s] = r;
    }
}

// ---- UnitEnergy: the reference's scalar loop (API surface, not used by its main) ----
// BipolarDomainManager.go:29-37:  energy += -0.5 * W_ij * s_i * s_j            (left to right, sequential over j)
// BinaryDomainManager.go:43-52:   energy += -0.5 * W_ij * s_i * (2 s_j - 1) - 1 (sic: one -1 per term)
// One thread per state: the summation order IS the contract (no FMA, no tree).
__global__ void unit_energy_kernel(const double* __restrict__ W, int ldw, const uint64_t* __restrict__ states, int words,
                                   int n, int count, int domain, int unit, double* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const uint64_t* bits = states + (size_t)s * words;
    const double* row = W + (size_t)unit * ldw;
    const bool binary = domain == HN_DOMAIN_BINARY;
    const double si = ((bits[unit >> 6] >> (unit & 63)) & 1ULL) ? 1.0 : (binary ? 0.0 : -1.0);
    double energy = 0.0;
    for (int j = 0; j < n; j++) {
        const double sj = ((bits[j >> 6] >> (j & 63)) & 1ULL) ? 1.0 : -1.0;   // binary: the mapped vector 2 s_j - 1
        double term = __dmul_rn(__dmul_rn(__dmul_rn(-0.5, row[j]), si), sj);
        if (binary) term = __dsub_rn(term, 1.0);
        energy = __dadd_rn(energy, term);
    }
    out[s] = energy;
}

// ---- unique-attractor table (datacollector/UniqueRelaxedStateHandler.go:41-73) ----
// key: sign-canonical packed final state (== identical energy profile, the
// reference's key); 128-bit hash picks the bucket, the full state decides equality.
// za / zb: the state has a unit whose field is an exact zero.  Its unit energy is then -0.5*(+0*v) = -+0, whose SIGN
// follows v, so s and -s print different energy profiles (the reference's key is fmt.Sprint of the profile,
// UniqueRelaxedStateHandler.go:44) and must not be merged; without such a unit the profiles are bit-identical.
__device__ __forceinline__ bool same_canonical(const uint64_t* a, const uint64_t* b, int words, int n,
                                               int domain, bool za, bool zb) {
    const bool ia = domain == HN_DOMAIN_BIPOLAR && (a[0] & 1ULL) && !za;
    const bool ib = domain == HN_DOMAIN_BIPOLAR && (b[0] & 1ULL) && !zb;
    for (int w = 0; w < words; w++) {
        uint64_t x = a[w], y = b[w];
        uint64_t mask = ~0ULL;
        if (w == words - 1 && (n & 63)) mask = (1ULL << (n & 63)) - 1ULL;
        if (ia) x = ~x & mask;
        if (ib) y = ~y & mask;
        if (x != y) return false;
    }
    return true;
}
// Level 1: table[slot] = representative probe index or -1; rep[p] = representative of p.  Key: the canonical-state hash
// (keys[p][0..1]) picks the bucket, the sign-canonical packed state decides — exact by construction.
__global__ void dedup_insert_kernel(const uint64_t* __restrict__ finals, const uint64_t* __restrict__ keys,
                                    int B, int words, int n, int domain, int32_t* table, int cap_mask,
                                    int32_t* __restrict__ rep, const uint8_t* __restrict__ zero_field) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const uint64_t k0 = keys[(size_t)p * 4], k1 = keys[(size_t)p * 4 + 1];
    uint32_t slot = (uint32_t)(k0 ^ (k1 >> 17)) & cap_mask;
    for (;;) {
        int32_t cur = atomicCAS(&table[slot], -1, p);
        if (cur == -1) { rep[p] = p; return; }
        if (keys[(size_t)cur * 4] == k0 && keys[(size_t)cur * 4 + 1] == k1 &&
            same_canonical(finals + (size_t)cur * words, finals + (size_t)p * words, words, n, domain,
                           zero_field[cur] != 0, zero_field[p] != 0)) {
            rep[p] = cur;
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
}
// Level 2 (only where the exact-profile key exists): the level-1 representatives are grouped by their profile key
// (keys[p][2..3]).  Equal profiles always give equal keys; every representative that joins an existing group is a
// DIFFERENT state with the same key, so the pair is recorded and its float profiles are compared by dedup_verify_kernel.
__global__ void dedup_insert_profile_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rep1, int B,
                                            int32_t* table, int cap_mask, int32_t* __restrict__ rep2,
                                            int2* __restrict__ pairs, int32_t* n_pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B || rep1[p] != p) return;
    const uint64_t k2 = keys[(size_t)p * 4 + 2], k3 = keys[(size_t)p * 4 + 3];
    uint32_t slot = (uint32_t)(k2 ^ (k3 >> 17)) & cap_mask;
    for (;;) {
        int32_t cur = atomicCAS(&table[slot], -1, p);
        if (cur == -1) { rep2[p] = p; return; }
        if (keys[(size_t)cur * 4 + 2] == k2 && keys[(size_t)cur * 4 + 3] == k3) {
            rep2[p] = cur;
            pairs[atomicAdd(n_pairs, 1)] = make_int2(p, cur);
            return;
        }
        slot = (slot + 1) & cap_mask;
    }
}
// One CTA per recorded pair (a, b): the reference's key is the printed energy profile (UniqueRelaxedStateHandler.go:44),
// e_i = -0.5 * (h_i * v_i) with h_i the float64 row dot product in gonum's summation order.  Both profiles are computed
// in that order and compared bit for bit; if they differ (a 128-bit key collision, or integer profiles that coincide
// while the rounded float profiles do not) the key of `a` is salted so that the next insertion pass separates it.
__global__ void __launch_bounds__(256)
dedup_verify_kernel(const int2* __restrict__ pairs, const uint64_t* __restrict__ finals, int words,
                    const double* __restrict__ W, int ldw, int n, int domain, uint64_t* __restrict__ keys,
                    uint64_t salt, int32_t* n_bad) {
    __shared__ int s_bad;
    const int2 pr = pairs[blockIdx.x];
    const uint32_t* ba = reinterpret_cast<const uint32_t*>(finals + (size_t)pr.x * words);
    const uint32_t* bb = reinterpret_cast<const uint32_t*>(finals + (size_t)pr.y * words);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int i = warp; i < n; i += 8) {
        const double ha = exact_field_warp(W + (size_t)i * ldw, ba, n, domain, lane);
        const double hb = exact_field_warp(W + (size_t)i * ldw, bb, n, domain, lane);
        if (lane == 0) {
            const double va = ((ba[i >> 5] >> (i & 31)) & 1u) ? 1.0 : -1.0, vb = ((bb[i >> 5] >> (i & 31)) & 1u) ? 1.0 : -1.0;
            const double ea = -0.5 * (ha * va), eb = -0.5 * (hb * vb);
            if (__double_as_longlong(ea) != __double_as_longlong(eb)) s_bad = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_bad) {
        keys[(size_t)pr.x * 4 + 2] ^= mix64(salt);
        keys[(size_t)pr.x * 4 + 3] += mix64(salt ^ 0x9E3779B97F4A7C15ULL);
        atomicAdd(n_bad, 1);
    }
}
// rep[p] = representative of p's final group: level 2 of its level-1 representative where level 2 ran
__global__ void dedup_compose_kernel(const int32_t* __restrict__ rep1, const int32_t* __restrict__ rep2, int B,
                                     int32_t* __restrict__ rep) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    rep[p] = rep2 ? rep2[rep1[p]] : rep1[p];
}
// first[rep] = min member index, hits[rep] = #members (or the sum of the members' own hit counts: merge of tables)
__global__ void dedup_count_kernel(const int32_t* __restrict__ rep, int B, int32_t* first, int32_t* hits,
                                   const int32_t* __restrict__ weight) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const int r = rep[p];
    atomicMin(&first[r], p);
    atomicAdd(&hits[r], weight ? weight[p] : 1);
}
// flag[p] = 1 iff p is the first-seen member of its group
__global__ void dedup_flag_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ first, int B,
                                  int32_t* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    flag[p] = (first[rep[p]] == p) ? 1 : 0;
}
// single-CTA exclusive scan of flag -> rank (first-seen order == ascending first index)
__global__ void __launch_bounds__(1024) scan_flags_kernel(const int32_t* __restrict__ flag, int B,
                                                         int32_t* __restrict__ rank, int32_t* __restrict__ total) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < B; base += 1024) {
        const int i = base + threadIdx.x;
        const int32_t v = i < B ? flag[i] : 0;
        int32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int32_t s = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += y; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int32_t prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + x - v;
        if (i < B) rank[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// attractor_id[p] = rank[first[rep[p]]]; table rows in first-seen order
__global__ void dedup_finalize_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ first,
                                      const int32_t* __restrict__ hits, const int32_t* __restrict__ rank, int B,
                                      int32_t* __restrict__ attractor_id, int32_t* __restrict__ ufirst,
                                      int32_t* __restrict__ uhits) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const int r = rep[p];
    const int f = first[r];
    const int id = rank[f];
    attractor_id[p] = id;
    if (f == p) { ufirst[id] = f; uhits[id] = hits[r]; }
}

// unique entries of a batch gathered into caller-owned DEVICE arrays (cross-GPU merge of tables: the entries travel
// between GPUs, not through the host): entry j = the first occurrence ufirst[j] of attractor j
__global__ void export_unique_kernel(const int32_t* __restrict__ ufirst, const int32_t* __restrict__ uhits, int nu,
                                     const uint64_t* __restrict__ finals, int words, const uint64_t* __restrict__ keys,
                                     const uint8_t* __restrict__ zf, long long offset, long long* __restrict__ first_out,
                                     int32_t* __restrict__ hits_out, uint64_t* __restrict__ keys_out,
                                     uint64_t* __restrict__ states_out, uint8_t* __restrict__ zf_out) {
    const int per = words + 4;
    const long long total = (long long)nu * per;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t / per), w = (int)(t % per);
        const int p = ufirst[j];
        if (w < words) states_out[(size_t)j * words + w] = finals[(size_t)p * words + w];
        else keys_out[(size_t)j * 4 + (w - words)] = keys[(size_t)p * 4 + (w - words)];
        if (w == 0) { first_out[j] = offset + p; hits_out[j] = uhits[j]; zf_out[j] = zf[p]; }
    }
}
// merged table rows: first = the global index carried by the group's earliest entry, hits = summed
__global__ void merge_rows_kernel(const int32_t* __restrict__ ufirst, const int32_t* __restrict__ uhits, int nu,
                                  const long long* __restrict__ first_in, long long* __restrict__ first_out,
                                  long long* __restrict__ hits_out, long long* __restrict__ rep_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nu) return;
    first_out[j] = first_in[ufirst[j]];
    hits_out[j] = uhits[j];
    rep_out[j] = ufirst[j];
}

// totals of the per-probe counters of a batch: flips, sweeps (= NumSteps - 1), margin hits
__global__ void sum_counters_kernel(const int32_t* __restrict__ flips, const int32_t* __restrict__ steps,
                                    const int32_t* __restrict__ margin, int B, unsigned long long* __restrict__ tot) {
    unsigned long long f = 0, s = 0, m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        f += (unsigned)flips[i];
        if (steps) s += (unsigned)(steps[i] - 1);
        if (margin) m += (unsigned)margin[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        f += __shfl_xor_sync(0xFFFFFFFFu, f, o); s += __shfl_xor_sync(0xFFFFFFFFu, s, o); m += __shfl_xor_sync(0xFFFFFFFFu, m, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&tot[0], f); atomicAdd(&tot[1], s); atomicAdd(&tot[2], m); }
}

// ---- f64 -> packed with activation (x <= 0 -> 0 bit) -------------------------
__global__ void pack_f64_kernel(const double* __restrict__ x, uint64_t* __restrict__ out, int n, int words,
                                int count) {
    const int64_t total = (int64_t)count * words;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
         o += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(o / words), w = (int)(o % words);
        uint64_t v = 0;
        for (int b = 0; b < 64 && w * 64 + b < n; b++)
            if (x[(size_t)s * n + w * 64 + b] > 0.0) v |= 1ULL << b;
        out[o] = v;
    }
}


// states.StateGenerator (states/StateGenerator.go:44-52, :65-74) on the device: unit i of state s takes draw number
// s*N + i of the PCG XSL-RR 128/64 stream (one Uint64 per unit: Float64() = (Uint64() & (2^53-1)) / 2^53),
// x = f*(max-min)+min, bit = x > 0 (domain activation).  One thread per output word: it jumps the 128-bit LCG to
// its first draw with the power table (rng_jump_table) and then steps it 64 times.
__global__ void generate_states_kernel(uint64_t state_lo, uint64_t state_hi, const uint64_t* __restrict__ jump,  // [64][4]
                                       int N, int words, int count, double rmin, double rmax, uint64_t* __restrict__ out) {
    using u128 = unsigned __int128;
    const long long total = (long long)count * words;
    const u128 kMul = ((u128)2549297995355413924ULL << 64) | 4865540595714422341ULL;
    const u128 kInc = ((u128)6364136223846793005ULL << 64) | 1442695040888963407ULL;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(t / words), w = (int)(t % words);
        const unsigned long long first = (unsigned long long)s * (unsigned long long)N + (unsigned long long)w * 64ULL;
        u128 st = ((u128)state_hi << 64) | state_lo;
        for (int j = 0; j < 64; j++)
            if ((first >> j) & 1ULL)
                st = st * (((u128)jump[j * 4 + 1] << 64) | jump[j * 4 + 0]) + (((u128)jump[j * 4 + 3] << 64) | jump[j * 4 + 2]);
        uint64_t bits = 0;
        const int n = min(64, N - w * 64);
        for (int i = 0; i < n; i++) {
            st = st * kMul + kInc;
            const uint64_t hi = (uint64_t)(st >> 64), lo = (uint64_t)st;
            const unsigned rot = (unsigned)(hi >> 58);
            const uint64_t x = hi ^ lo;
            const uint64_t r = (x >> rot) | (x << ((64u - rot) & 63u));
            const double f = (double)(r & ((1ULL << 53) - 1)) / 9007199254740992.0;
            const double v = __dadd_rn(__dmul_rn(f, rmax - rmin), rmin);   // Uniform.Rand: rnd*(Max-Min) + Min, unfused
            if (v > 0.0) bits |= 1ULL << i;
        }
        out[t] = bits;
    }
}


// Relaxation on the all-zero matrix in closed form (a network that has learned nothing): every field is an exact zero,
// so the first sweep sets every unit to the low value (h <= 0, BipolarDomainManager.go:10-22), no unit energy is
// positive and the state is stable after that sweep: NumSteps = 2, flips = units that were high, and every probe ends
// in the one all-low attractor.  One warp per probe.
__global__ void zero_matrix_relax_kernel(const uint64_t* __restrict__ probes, int B, int words, int n, int domain, int max_iter,
                                         uint64_t* __restrict__ finals, int32_t* __restrict__ steps, uint8_t* __restrict__ stable,
                                         int32_t* __restrict__ flips, int32_t* __restrict__ margin, int32_t* __restrict__ dist,
                                         const uint64_t* __restrict__ targets, int P, uint64_t* __restrict__ hash,
                                         uint8_t* __restrict__ zero_field, uint64_t* __restrict__ history) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= B) return;
    int f = 0;
    for (int w = lane; w < words; w += 32) {
        const uint64_t s = probes[(size_t)p * words + w];
        f += __popcll(s);
        finals[(size_t)p * words + w] = 0ULL;
        if (history) {
            history[((size_t)p * (max_iter + 1)) * words + w] = s;
            history[((size_t)p * (max_iter + 1) + 1) * words + w] = 0ULL;
        }
    }
    f = __reduce_add_sync(0xFFFFFFFFu, f);
    if (dist)
        for (int t = 0; t < P; t++) {
            int dh = 0;
            for (int w = lane; w < words; w += 32) dh += __popcll(targets[(size_t)t * words + w]);
            dh = __reduce_add_sync(0xFFFFFFFFu, dh);
            if (lane == 0) { const int m = min(dh, n - dh); dist[(size_t)p * P + t] = domain == HN_DOMAIN_BINARY ? m : 2 * m; }
        }
    if (lane == 0) {
        steps[p] = 2; stable[p] = 1; flips[p] = f; margin[p] = 0;
        if (hash) { hash[(size_t)p * 4] = 0x5A45524F4D415452ULL; hash[(size_t)p * 4 + 1] = 0x4958414C4C4C4F57ULL; }   // one attractor (equal states)
        if (zero_field) zero_field[p] = 1;
    }
}

}  // namespace hn
