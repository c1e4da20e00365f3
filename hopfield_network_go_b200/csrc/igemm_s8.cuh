// igemm_s8.cuh — exact integer contraction for the integer column stream (HN_COLUMNS_I16):
//   G[B x N] (int32) = S[B x N] (states as +-1 / 0-1 int8, expanded from packed bits) * K[N x N]^T (int16)
// i.e. the initial exact integer fields g = scale * Wraw * s of a probe batch, on the integer tensor-core path
// (mma.sync.m16n8k32.s8: SASS IMMA).  K is split into a signed high byte plane and an unsigned low byte plane,
// K = 256*Khi + Klo, so two s8 x {s8,u8} MMAs per tile give the exact int32 result: G = 256*(S Khi^T) + S Klo^T;
// while every |K| <= 127 the low byte IS the entry as a signed byte and one s8 x s8 MMA does (SINGLE).
// Replaces the float64 fields GEMM (gemm_f64.cuh) when the integer structure of W is known and W is symmetric
// (K is stored column-major == row-major then); everything is exact, there is no rounding anywhere.
#pragma once
#include "common.cuh"

namespace hn {

constexpr int IG_BM = 128, IG_BN = 64, IG_BK = 64;   // CTA tile (probes x units x k-bytes)
constexpr int IG_LD = IG_BK + 16;                    // padded smem row stride in bytes (conflict-free fragments)

__device__ __forceinline__ void imma_s8s8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void imma_s8u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 4 state bits -> 4 int8 values in one word: bipolar +1 / -1 (0x01 / 0xFF), binary 1 / 0
__device__ __forceinline__ uint32_t expand4(uint32_t nib, bool bipolar) {
    const uint32_t spread = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);   // bytes 0/1
    return bipolar ? (0xFFFFFFFFu ^ (spread * 0xFEu)) : spread;
}

// grid: (ceil(N/IG_BN), ceil(B/IG_BM)); 256 threads = 8 warps as 4 (M) x 2 (N), warp tile 32 x 32
// SINGLE: every |K| <= 127, the low byte of an entry is the entry as a signed byte: one s8 x s8 MMA per tile, no high plane
template <bool SINGLE>
__global__ void __launch_bounds__(256)
igemm_fields_kernel(const uint64_t* __restrict__ states, int words, int B, const int16_t* __restrict__ K, int ldk, int N,
                    int bipolar, int32_t* __restrict__ G, int ldg) {
    __shared__ __align__(16) unsigned char sA[2][IG_BM * IG_LD];
    __shared__ __align__(16) unsigned char sBh[2][IG_BN * IG_LD];
    __shared__ __align__(16) unsigned char sBl[2][IG_BN * IG_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;
    const int m0 = blockIdx.y * IG_BM, n0 = blockIdx.x * IG_BN;

    int acc_h[2][4][4], acc_l[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc_h[i][j][e] = acc_l[i][j][e] = 0;

    // staging registers: A: half a row (32 units = 32 bits) per thread; B: 16 int16 of one row per thread
    uint32_t a_bits = 0;
    uint4 b_lo = make_uint4(0, 0, 0, 0), b_hi = make_uint4(0, 0, 0, 0);
    const int a_row = tid >> 1, a_half = tid & 1;
    const int b_row = tid >> 2, b_chunk = tid & 3;
    auto gload = [&](int k0) {
        const int row = m0 + a_row;
        a_bits = 0;
        if (row < B && k0 + a_half * 32 < N)
            a_bits = (uint32_t)(states[(size_t)row * words + (k0 >> 6)] >> (a_half * 32));
        const int n = n0 + b_row, kk = k0 + b_chunk * 16;
        b_lo = b_hi = make_uint4(0, 0, 0, 0);
        if (n < N && kk < ldk) {
            const uint4* q = reinterpret_cast<const uint4*>(K + (size_t)n * ldk + kk);
            b_lo = __ldg(q);
            b_hi = __ldg(q + 1);
        }
    };
    auto sstore = [&](int buf, int k0) {
        // A: 32 bits -> 8 words of 4 int8 (units beyond N are masked to 0 so they contribute nothing)
        uint32_t bits = a_bits;
        const int base = k0 + a_half * 32;
        uint32_t* qa = reinterpret_cast<uint32_t*>(&sA[buf][a_row * IG_LD + a_half * 32]);
#pragma unroll
        for (int wq = 0; wq < 8; wq++) {
            uint32_t v = expand4((bits >> (4 * wq)) & 0xFu, bipolar != 0);
            const int u = base + 4 * wq;
            if (u + 3 >= N) {   // ragged tail: zero the bytes of units >= N
                uint32_t mask = 0;
#pragma unroll
                for (int e = 0; e < 4; e++) if (u + e < N) mask |= 0xFFu << (8 * e);
                v &= mask;
            }
            qa[wq] = v;
        }
        // B: 16 int16 (two uint4) -> 16 low bytes + 16 high bytes
        const uint32_t w[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
        uint32_t* ql = reinterpret_cast<uint32_t*>(&sBl[buf][b_row * IG_LD + b_chunk * 16]);
        uint32_t* qh = reinterpret_cast<uint32_t*>(&sBh[buf][b_row * IG_LD + b_chunk * 16]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            ql[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x6420);
            if constexpr (!SINGLE) qh[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x7531);
        }
    };

    const int ktiles = (N + IG_BK - 1) / IG_BK;
    gload(0);
    sstore(0, 0);
    __syncthreads();
    for (int kt = 0; kt < ktiles; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) gload((kt + 1) * IG_BK);
#pragma unroll
        for (int ks = 0; ks < IG_BK / 32; ks++) {
            uint32_t af[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const unsigned char* base = &sA[buf][(wm0 + i * 16 + g) * IG_LD + ks * 32 + 4 * t];
                af[i][0] = *reinterpret_cast<const uint32_t*>(base);
                af[i][1] = *reinterpret_cast<const uint32_t*>(base + 8 * IG_LD);
                af[i][2] = *reinterpret_cast<const uint32_t*>(base + 16);
                af[i][3] = *reinterpret_cast<const uint32_t*>(base + 8 * IG_LD + 16);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int off = (wn0 + j * 8 + g) * IG_LD + ks * 32 + 4 * t;
                if constexpr (!SINGLE) {
                    bh[j][0] = *reinterpret_cast<const uint32_t*>(&sBh[buf][off]);
                    bh[j][1] = *reinterpret_cast<const uint32_t*>(&sBh[buf][off + 16]);
                }
                bl[j][0] = *reinterpret_cast<const uint32_t*>(&sBl[buf][off]);
                bl[j][1] = *reinterpret_cast<const uint32_t*>(&sBl[buf][off + 16]);
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if constexpr (SINGLE) imma_s8s8(acc_l[i][j], af[i], bl[j]);
                    else { imma_s8s8(acc_h[i][j], af[i], bh[j]); imma_s8u8(acc_l[i][j], af[i], bl[j]); }
                }
        }
        if (kt + 1 < ktiles) sstore(buf ^ 1, (kt + 1) * IG_BK);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = n0 + wn0 + j * 8 + 2 * t;
#pragma unroll
            for (int hrow = 0; hrow < 2; hrow++) {
                const int row = m0 + wm0 + i * 16 + g + 8 * hrow;
                if (row < B && col < ldg) {   // ldg is even and >= N; padding columns get zeros (K rows beyond N are zero)
                    const int v0 = (SINGLE ? 0 : 256 * acc_h[i][j][2 * hrow]) + acc_l[i][j][2 * hrow];
                    const int v1 = (SINGLE ? 0 : 256 * acc_h[i][j][2 * hrow + 1]) + acc_l[i][j][2 * hrow + 1];
                    *reinterpret_cast<int2*>(G + (size_t)row * ldg + col) = make_int2(v0, v1);
                }
            }
        }
}

static inline int launch_igemm_fields(const uint64_t* states, int words, int B, const int16_t* K, int ldk, int N, bool bipolar,
                                      int32_t* G, int ldg, bool fits_s8, cudaStream_t st) {
    if (B <= 0) return HN_OK;
    dim3 grid((N + IG_BN - 1) / IG_BN, (B + IG_BM - 1) / IG_BM);
    if (fits_s8) igemm_fields_kernel<true><<<grid, 256, 0, st>>>(states, words, B, K, ldk, N, bipolar ? 1 : 0, G, ldg);
    else igemm_fields_kernel<false><<<grid, 256, 0, st>>>(states, words, B, K, ldk, N, bipolar ? 1 : 0, G, ldg);
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}

}  // namespace hn
