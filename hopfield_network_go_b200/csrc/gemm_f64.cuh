// gemm_f64.cuh — FP64 tensor-core contraction  C[M x N] = A[M x K] * B[N x K]^T
//
// Used only where the work is a real dense contraction (BASELINE north_star):
//   * initial local fields of a probe batch, H = S * W^T       (replaces the N
//     mat.Dot calls per sweep of HopfieldNetwork.go:394-399 for the first fields,
//     and the MulVec of domain/BipolarDomainManager.go:41 for energy profiles)
//   * Hebbian dW = X^T X                                        (P RankOne calls, LearningRule.go:68-70)
//   * Delta   dW = sum_mu 0.5 (t - r) t^T                       (LearningRule.go:106-109)
//
// Blackwell's tcgen05 has no f64 kind, so the FP64 tensor pipe is driven with
// mma.sync.m8n8k4.f64 (SASS DMMA).  Both operands are K-contiguous; bipolar /
// binary states are expanded from their packed bit form while staging to shared
// memory, so a state costs 1 bit of HBM traffic per unit, not 8 bytes.
#pragma once
#include "common.cuh"

namespace hn {

// ---- operand loaders: load8(row, k0, v) yields the 8 values (row, k0..k0+7) ----

// float64 row-major, zero padded so that ld >= round_up(K,16).
struct OpF64 {
    const double* p; int ld; int rows;
    __device__ __forceinline__ void load8(int row, int k0, double (&v)[8]) const {
        if (row < rows) {
            const double2* q = reinterpret_cast<const double2*>(p + (size_t)row * ld + k0);
#pragma unroll
            for (int e = 0; e < 4; e++) { double2 t = __ldg(q + e); v[2 * e] = t.x; v[2 * e + 1] = t.y; }
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) v[e] = 0.0;
        }
    }
};

// packed states: bit -> hi, no bit -> lo, beyond K -> 0.
struct OpBits {
    const uint64_t* p; int words; int rows; int K; double lo, hi;
    __device__ __forceinline__ void load8(int row, int k0, double (&v)[8]) const {
        unsigned bits = 0;
        if (row < rows && k0 < K) bits = (unsigned)((p[(size_t)row * words + (k0 >> 6)] >> (k0 & 63)) & 0xFFu);
        const bool rv = row < rows;
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = (rv && (k0 + e) < K) ? (((bits >> e) & 1u) ? hi : lo) : 0.0;
    }
};

// scale * (bitA - bitB): the 0.5*(t - r) factor of the Delta rule (LearningRule.go:107-108).
struct OpBitsDiff {
    const uint64_t* a; const uint64_t* b; int words; int rows; int K; double scale;
    __device__ __forceinline__ void load8(int row, int k0, double (&v)[8]) const {
        unsigned ba = 0, bb = 0;
        if (row < rows && k0 < K) {
            size_t o = (size_t)row * words + (k0 >> 6);
            ba = (unsigned)((a[o] >> (k0 & 63)) & 0xFFu);
            bb = (unsigned)((b[o] >> (k0 & 63)) & 0xFFu);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
            int d = (int)((ba >> e) & 1u) - (int)((bb >> e) & 1u);
            v[e] = ((k0 + e) < K) ? scale * (double)d : 0.0;
        }
    }
};

// ---- epilogues: called with two adjacent columns (col even) ----
struct EpiStore {
    double* c; int ld; int M; int N;
    __device__ __forceinline__ void operator()(int row, int col, double v0, double v1) const {
        if (row < M) {
            double* q = c + (size_t)row * ld + col;
            if (col + 1 < N) { *reinterpret_cast<double2*>(q) = make_double2(v0, v1); }
            else if (col < N) { q[0] = v0; }
        }
    }
};

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int GEMM_BK = 16;
constexpr int GEMM_LDS = GEMM_BK + 4;  // padded row stride (doubles): conflict-free fragment loads

template <int BM, int BN>
constexpr size_t gemm_smem_bytes() { return (size_t)2 * (BM + BN) * GEMM_LDS * sizeof(double); }

template <class OpA, class OpB, class Epi, int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_f64_kernel(OpA opa, OpB opb, Epi epi, int K) {
    constexpr int THREADS = (BM / WM) * (BN / WN) * 32;
    constexpr int TM = WM / 8, TN = WN / 8;
    constexpr int CHA = (BM * 2 + THREADS - 1) / THREADS;  // 8-wide chunks per thread
    constexpr int CHB = (BN * 2 + THREADS - 1) / THREADS;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;                          // [2][BM][LDS]
    double* Bs = smem + 2 * BM * GEMM_LDS;      // [2][BN][LDS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    double sa[CHA][8], sb[CHB][8];
    auto gload = [&](int k0) {
#pragma unroll
        for (int c = 0; c < CHA; c++) {
            int ch = tid + c * THREADS;
            if (ch < BM * 2) opa.load8(m0 + (ch >> 1), k0 + (ch & 1) * 8, sa[c]);
        }
#pragma unroll
        for (int c = 0; c < CHB; c++) {
            int ch = tid + c * THREADS;
            if (ch < BN * 2) opb.load8(n0 + (ch >> 1), k0 + (ch & 1) * 8, sb[c]);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int c = 0; c < CHA; c++) {
            int ch = tid + c * THREADS;
            if (ch < BM * 2) {
                double2* q = reinterpret_cast<double2*>(As + ((size_t)buf * BM + (ch >> 1)) * GEMM_LDS + (ch & 1) * 8);
#pragma unroll
                for (int e = 0; e < 4; e++) q[e] = make_double2(sa[c][2 * e], sa[c][2 * e + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < CHB; c++) {
            int ch = tid + c * THREADS;
            if (ch < BN * 2) {
                double2* q = reinterpret_cast<double2*>(Bs + ((size_t)buf * BN + (ch >> 1)) * GEMM_LDS + (ch & 1) * 8);
#pragma unroll
                for (int e = 0; e < 4; e++) q[e] = make_double2(sb[c][2 * e], sb[c][2 * e + 1]);
            }
        }
    };

    const int ktiles = (K + GEMM_BK - 1) / GEMM_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < ktiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) gload((kt + 1) * GEMM_BK);
        const double* Ab = As + (size_t)buf * BM * GEMM_LDS + (wm0 + (lane >> 2)) * GEMM_LDS + (lane & 3);
        const double* Bb = Bs + (size_t)buf * BN * GEMM_LDS + (wn0 + (lane >> 2)) * GEMM_LDS + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; kk++) {
            double af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) af[i] = Ab[i * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < TN; j++) bf[j] = Bb[j * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (kt + 1 < ktiles) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++)
            epi(m0 + wm0 + i * 8 + (lane >> 2), n0 + wn0 + j * 8 + 2 * (lane & 3), acc[i][j][0], acc[i][j][1]);
}

// Host launcher. M x N output tiles; picks a tile shape by problem size.
template <class OpA, class OpB, class Epi>
static int launch_gemm_f64(const OpA& a, const OpB& b, const Epi& e, int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || N <= 0) return HN_OK;
    if ((int64_t)M * N >= (int64_t)128 * 128 * 296) {
        constexpr int BM = 128, BN = 128;
        auto kern = gemm_f64_kernel<OpA, OpB, Epi, BM, BN, 64, 32>;
        HN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<BM, BN>()));
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
        kern<<<grid, 256, gemm_smem_bytes<BM, BN>(), st>>>(a, b, e, K);
    } else {
        constexpr int BM = 64, BN = 64;
        auto kern = gemm_f64_kernel<OpA, OpB, Epi, BM, BN, 32, 32>;
        HN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<BM, BN>()));
        dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
        kern<<<grid, 128, gemm_smem_bytes<BM, BN>(), st>>>(a, b, e, K);
    }
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}

}  // namespace hn
