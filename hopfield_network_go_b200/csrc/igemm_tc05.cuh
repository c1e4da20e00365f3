// igemm_tc05.cuh — the exact integer fields contraction on the Blackwell tensor path (tcgen05 + TMEM + TMA):
//   G[B x N] (int32) = S[B x N] (states as +-1 / 0-1 int8, expanded from packed bits) * M8[N x N]^T (int8)
// i.e. the initial exact integer fields of a probe batch on the integer stream (HN_COLUMNS_I16 / AUTO) whenever the
// int16 image of the weights fits a signed byte (|M| <= 127: every Hebbian / Delta matrix of the BASELINE configs).
// Replaces the mma.sync kernel of igemm_s8.cuh on that path (which stays for images that need two byte planes).
//
// One persistent CTA per SM, 128 probes x 256 units per tile, K streamed in 128-byte steps through a 4-stage ring:
//   warp 0      TMA producer: the M8 tile (256 rows x 128 B, 128-byte swizzle) of every K step;
//   warp 1      allocates TMEM (2 accumulators of 256 columns) and issues tcgen05.mma.kind::i8 (M=128, N=256, K=32);
//   warps 2-9   expand the probes' bits into the int8 A tile in the same swizzled layout (generic proxy + proxy fence);
//               eight warps because this expansion, not the tensor pipe or the weight stream, bounds a K step otherwise;
//   warps 10-13 epilogue: tcgen05.ld of the finished accumulator (one TMEM sub-partition per warp) -> global int32,
//               overlapped with the next tile's MMAs through the second accumulator.
// Everything is exact integer arithmetic: the result is bit-identical to igemm_fields_kernel and to the float64 GEMM.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace hn {

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 128, TG_STAGES = 4;
constexpr int TG_A_BYTES = TG_BM * TG_BK, TG_B_BYTES = TG_BN * TG_BK;
constexpr int TG_THREADS = 448;   // warp 0 TMA, warp 1 MMA, warps 2-9 A producers, warps 10-13 epilogue
constexpr int TG_PRODUCER_WARPS = 8;
constexpr int TG_GROUPS_DEFAULT = 4;   // HN_IGEMM_GROUPS picks 1 / 2 / 4 at run time (measured per config-3 step: 8.98 / 7.74 / 7.10 ms)
// rows per TMA box: an operand tile is fetched as several boxes of this many 128-byte rows (same shared-memory layout:
// the 128-byte swizzle repeats every 8 rows); HN_TMA_BOX picks it at run time (measured at config 3: 128 and 32 rows alike, 8 rows 1.8x slower).
constexpr int TG_BOX_DEFAULT = 128;
static inline int tg_box_rows() {
    static const int v = [] {
        const char* e = getenv("HN_TMA_BOX");
        const int x = e ? atoi(e) : TG_BOX_DEFAULT;
        return (x == 8 || x == 16 || x == 32 || x == 64 || x == 128 || x == 256) ? x : TG_BOX_DEFAULT;
    }();
    return v;
}
constexpr size_t TG_SMEM = (size_t)TG_STAGES * (TG_A_BYTES + TG_B_BYTES) + 256 + 1024;   // + barriers + alignment slack

// 4 state bits -> 4 int8 in one word: bipolar +1 / -1 (0x01 / 0xFF), binary 1 / 0
// (nib < 16: the four shifted copies nib << 0 / 7 / 14 / 21 do not overlap, bit 8j of their sum is bit j of nib; and
// 0xFF - 0xFE * bit never borrows across bytes: one multiply-add instead of a multiply and an xor)
__device__ __forceinline__ uint32_t tg_expand4(uint32_t nib, bool bipolar) {
    const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;
    return bipolar ? 0xFFFFFFFFu - spread * 0xFEu : spread;
}

// A_BITS: the A operand (rows = probes) is expanded from packed state bits by warps 2-5 (the fields contraction);
// otherwise both operands are int8 matrices fetched by TMA (the learning contractions dW = X^T X, D^T T) and K = Kdim.
template <bool A_BITS, int GROUPS>
__global__ void __launch_bounds__(TG_THREADS, 1)
igemm_tc05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const uint64_t* __restrict__ states, int words, int B, int N, int Kdim,
                  int bipolar, int32_t* __restrict__ G, int ldg, int32_t* __restrict__ scan_unstable, int32_t* __restrict__ scan_zero,
                  int box_rows) {
    using namespace tc05;
    extern __shared__ unsigned char tg_smem_raw[];
    const uint32_t raw = smem_u32(tg_smem_raw);
    unsigned char* smem = tg_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* sA = smem;
    unsigned char* sB = smem + TG_STAGES * TG_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TG_STAGES * (TG_A_BYTES + TG_B_BYTES));
    uint64_t* full = bars;                    // [stages] 1 TMA arrive(+tx) + 8 producer warps
    uint64_t* empty = bars + TG_STAGES;       // [stages] tcgen05.commit
    uint64_t* acc_full = bars + 2 * TG_STAGES;       // [2] tcgen05.commit after the last K step of a tile
    uint64_t* acc_empty = bars + 2 * TG_STAGES + 2;  // [2] 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TG_STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_tiles = (B + TG_BM - 1) / TG_BM, n_tiles = (N + TG_BN - 1) / TG_BN;
    const int tiles = m_tiles * n_tiles;
    const int kblocks = (Kdim + TG_BK - 1) / TG_BK;

    if (tid == 0) {
        for (int s = 0; s < TG_STAGES; s++) { mbar_init(&full[s], A_BITS ? 1 + TG_PRODUCER_WARPS / GROUPS : 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (the whole warp runs the loop converged, one elected lane issues: tc05.cuh) =====
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int n0 = (tile % n_tiles) * TG_BN, m0 = (tile / n_tiles) * TG_BM;
                for (int kb = 0; kb < kblocks; kb++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx_elect(&full[stage], A_BITS ? TG_B_BYTES : TG_A_BYTES + TG_B_BYTES);
                    for (int r = 0; r < TG_BN; r += box_rows)
                        tma_load_2d_elect(sB + (size_t)stage * TG_B_BYTES + (size_t)r * TG_BK, &map_b, kb * TG_BK, n0 + r, &full[stage]);
                    if constexpr (!A_BITS)
                        for (int r = 0; r < TG_BM; r += box_rows)
                            tma_load_2d_elect(sA + (size_t)stage * TG_A_BYTES + (size_t)r * TG_BK, &map_a, kb * TG_BK, m0 + r, &full[stage]);
                    if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (converged warp, elected issue) =====
        constexpr uint32_t idesc = idesc_i8(TG_BM, TG_BN, true, true);
        const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);   // warp-uniform by construction
        const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
        int stage = 0; uint32_t phase = 0; int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t use = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&acc_empty[acc], use ^ 1u);      // the epilogue has drained this accumulator
            tc_fence_after_sync();
            for (int kb = 0; kb < kblocks; kb++) {
                mbar_wait(&full[stage], phase);
                tc_fence_after_sync();
                const uint64_t da = smem_desc_k128(sA_addr + (uint32_t)stage * (uint32_t)TG_A_BYTES);
                const uint64_t db = smem_desc_k128(sB_addr + (uint32_t)stage * (uint32_t)TG_B_BYTES);
#pragma unroll
                for (int j = 0; j < TG_BK / 32; j++)   // K = 32 bytes per instruction: +32 B on the start address (>>4)
                    mma_i8_elect(tmem_u + (uint32_t)acc * TG_BN, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc,
                                 (uint32_t)((kb | j) != 0));
                mma_commit_elect(&empty[stage]);             // frees the stage when these MMAs have read it
                if (kb == kblocks - 1) mma_commit_elect(&acc_full[acc]);
                if (++stage == TG_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < 2 + TG_PRODUCER_WARPS) {
        if constexpr (A_BITS) {
        // ===== A producers: bits -> int8, 128-byte swizzled K-major tile =====
        // The eight warps form GROUPS groups; group g fills the ring slots of the K steps i = g (mod GROUPS), i counted
        // across tiles (slot i mod 4).  What a warp spends per slot is mostly latency (slot wait, proxy fence, barrier
        // arrival), so slots are filled side by side rather than by all eight warps one slot after the other.
        static_assert(TG_STAGES == 4, "slot = step & 3");
        constexpr int PT = 32 * TG_PRODUCER_WARPS / GROUPS, PI = TG_BM * 4 / PT;   // threads per group, (row, quarter) items per thread
        const int g = (tid - 64) / PT, a = (tid - 64) % PT;
        auto fetch = [&](int m0, int kb, uint32_t (&out)[PI]) {   // requested one of this group's steps ahead of the slot wait
#pragma unroll
            for (int j = 0; j < PI; j++) {
                const int q = a + PT * j;
                const int probe = m0 + (q >> 2), unit0 = kb * TG_BK + (q & 3) * 32;
                out[j] = 0;
                if (probe < B && unit0 < N)
                    out[j] = (uint32_t)(__ldg(states + (size_t)probe * words + (unit0 >> 6)) >> (unit0 & 63));
            }
        };
        int tile = blockIdx.x, kb = 0;
        uint32_t step = 0;
        auto advance = [&](int n) {
            for (int t = 0; t < n; t++) {
                if (++kb == kblocks) { kb = 0; tile += gridDim.x; }
                step++;
            }
        };
        advance(g);
        uint32_t nxt[PI] = {};
        if (tile < tiles) fetch((tile / n_tiles) * TG_BM, kb, nxt);
        while (tile < tiles) {
            const int ckb = kb;
            const uint32_t stage = step & 3u, phase = (step >> 2) & 1u;
            uint32_t cur[PI];
#pragma unroll
            for (int j = 0; j < PI; j++) cur[j] = nxt[j];
            advance(GROUPS);
            if (tile < tiles) fetch((tile / n_tiles) * TG_BM, kb, nxt);
            mbar_wait(&empty[stage], phase ^ 1u);
            unsigned char* dst = sA + (size_t)stage * TG_A_BYTES;
#pragma unroll
            for (int j = 0; j < PI; j++) {
                const int q = a + PT * j;
                const int row = q >> 2, quarter = q & 3;
                const int unit0 = ckb * TG_BK + quarter * 32;
                const uint32_t bits = cur[j];
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; e++) w[e] = tg_expand4((bits >> (4 * e)) & 0xFu, bipolar != 0);
                if (unit0 + 32 > N) {   // ragged tail: units >= N contribute nothing
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int bb = 0; bb < 4; bb++) if (unit0 + 4 * e + bb < N) mask |= 0xFFu << (8 * bb);
                        w[e] &= mask;
                    }
                }
                const uint32_t rbase = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
                const uint32_t c0 = (uint32_t)quarter * 2u, sw = (uint32_t)(row & 7);
                *reinterpret_cast<uint4*>(dst + rbase + (((c0) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + rbase + (((c0 + 1u) ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int sub = warp & 3;   // TMEM sub-partition this warp may read
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t use = (uint32_t)(it >> 1) & 1u;
            const int m0 = (tile / n_tiles) * TG_BM, n0 = (tile % n_tiles) * TG_BN;
            mbar_wait(&acc_full[acc], use);
            tc_fence_after_sync();
            const int row = m0 + sub * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < TG_BN / 32; c++) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * TG_BN + c * 32), r);
                tmem_ld_wait();
                const int col = n0 + c * 32;
                if constexpr (A_BITS) {
                    // stability scan fused into the fields contraction (the test that follows a dense lockstep sweep,
                    // HopfieldNetwork.go:214-225): per probe, the units with v*g < 0 and the units with an exactly zero field.
                    // The tile holds (M s)_u = 2 g_u + diag s_u; with the unit's bit b:  2 g = value - bias + 2b.
                    if (scan_unstable && row < B && col < N) {
                        const uint32_t bw = reinterpret_cast<const uint32_t*>(states + (size_t)row * words)[col >> 5];
                        const int bias = bipolar ? 1 : 0;
                        int cu = 0, cz = 0;
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            const int b = (int)((bw >> i) & 1u);
                            const int g2 = (int)r[i] - bias + 2 * b;
                            const bool in = col + i < N;
                            cu += (in && (b ? g2 < 0 : g2 > 0)) ? 1 : 0;
                            cz += (in && g2 == 0) ? 1 : 0;
                        }
                        if (cu) atomicAdd(scan_unstable + row, cu);
                        if (cz) atomicAdd(scan_zero + row, cz);
                    }
                }
                if (row < B && G) {
                    int32_t* out = G + (size_t)row * ldg + col;
#pragma unroll
                    for (int v = 0; v < 8; v++)
                        if (col + 4 * v < ldg)   // ldg is a multiple of 16: whole vectors
                            *reinterpret_cast<int4*>(out + 4 * v) = make_int4((int)r[4 * v], (int)r[4 * v + 1], (int)r[4 * v + 2], (int)r[4 * v + 3]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// M8: row-major int8 [N][ld8] (ld8 multiple of 16), row k = column k of the integer image (symmetric weights)
// scan_unstable / scan_zero (optional, [B], zeroed by the caller): per-probe counts of units with v*g < 0 / g == 0;
// G may be nullptr then (fields not stored)
static inline int launch_igemm_tc05(const uint64_t* states, int words, int B, const int8_t* M8, int ld8, int N, bool bipolar,
                                    int32_t* G, int ldg, int sms, cudaStream_t st, int32_t* scan_unstable = nullptr,
                                    int32_t* scan_zero = nullptr) {
    if (B <= 0) return HN_OK;
    CUtensorMap map;
    const int box = tg_box_rows();
    const int rc = tc05::make_map_bytes_2d(&map, M8, (uint64_t)ld8, (uint64_t)N, (uint64_t)ld8, box);
    if (rc != 0) { set_error("cuTensorMapEncodeTiled failed (%d)", rc); return HN_E_CUDA; }
    const int tiles = ((B + TG_BM - 1) / TG_BM) * ((N + TG_BN - 1) / TG_BN);
    static const int groups = [] { const char* e = getenv("HN_IGEMM_GROUPS"); const int x = e ? atoi(e) : TG_GROUPS_DEFAULT; return (x == 1 || x == 2 || x == 4) ? x : TG_GROUPS_DEFAULT; }();
    auto go = [&](auto kern) -> int {
        HN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
        kern<<<std::min(tiles, sms), TG_THREADS, TG_SMEM, st>>>(map, map, states, words, B, N, N, bipolar ? 1 : 0, G, ldg, scan_unstable, scan_zero, box);
        return HN_OK;
    };
    if (groups == 1) HN_TRY(go(igemm_tc05_kernel<true, 1>));
    else if (groups == 2) HN_TRY(go(igemm_tc05_kernel<true, 2>));
    else HN_TRY(go(igemm_tc05_kernel<true, 4>));
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}

// C[M x N] (int32, leading dimension ldc) = A8[M x K] * B8[N x K]^T, both operands row-major int8 with pitch `pitch`
// (a multiple of 16 bytes, K <= pitch, bytes k >= K zero).  Exact: the learning contractions' operands are in
// {-2, .., 2}, so |C| <= 4K.
static inline int launch_igemm_tc05_mats(const int8_t* A8, int M, const int8_t* B8, int N, int K, int pitch, int32_t* C, int ldc,
                                         int sms, cudaStream_t st) {
    if (M <= 0 || N <= 0) return HN_OK;
    CUtensorMap ma, mb;
    const int box = std::min(tg_box_rows(), TG_BM);   // one box size for both operands: it must divide the smaller tile
    int rc = tc05::make_map_bytes_2d(&ma, A8, (uint64_t)pitch, (uint64_t)M, (uint64_t)pitch, box);
    if (rc == 0) rc = tc05::make_map_bytes_2d(&mb, B8, (uint64_t)pitch, (uint64_t)N, (uint64_t)pitch, box);
    if (rc != 0) { set_error("cuTensorMapEncodeTiled failed (%d)", rc); return HN_E_CUDA; }
    HN_CUDA_TRY(cudaFuncSetAttribute(igemm_tc05_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
    const int tiles = ((M + TG_BM - 1) / TG_BM) * ((N + TG_BN - 1) / TG_BN);
    igemm_tc05_kernel<false, 1><<<std::min(tiles, sms), TG_THREADS, TG_SMEM, st>>>(ma, mb, nullptr, 0, M, N, K, 0, C, ldc, nullptr, nullptr, box);
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}

}  // namespace hn
