// dense.cuh — the first, flip-dense sweeps of a relaxation on the Blackwell tensor path (tcgen05 + TMEM + TMA).
//
// The sparse kernel (relax.cuh) pays per FLIP (one weight column per flip); in the first sweeps of a random probe almost
// every second unit flips (N=4096, P=400: 1764, 428, 228, 175 flips in sweeps 1-4 of ~3900 in all), so those sweeps are
// done per POSITION instead, for 32 probes in lockstep (the shared schedule gives every probe the same visiting order):
//
//   * a sweep visits the units in blocks of 32 positions.  For a block, the exact integer fields of its 32 units are
//     computed FRESH for all 32 probes by the tensor core:  D = S8 * Mperm^T  with S8 the probes' current states as int8
//     (resident in shared memory, 4 KiB per probe) and Mperm the block's 32 rows of the byte image of the weights in
//     visiting order (streamed by TMA from a copy of the image whose rows are permuted once per schedule row).
//     To fill the 128-row MMA with 32 probes, K is cut in four quarters: row (q, p) of A holds quarter q of probe p's
//     state, row (q', i) of B quarter q' of position i's weight row; the four diagonal 32x32 blocks of the 128x128
//     result are the partial fields and are summed after tcgen05.ld (one TMEM sub-partition per quarter).
//   * one warp then takes the 32 decisions of the block in order, one probe per lane: a flip at position i corrects the
//     fields of the later positions of the block with the 32x32 in-block matrix (dp4a against a byte selector), is
//     written into S8 (so the next block's contraction sees it) and into the packed state in global memory.
//     An exact tie (integer field 0) is decided as the sparse kernel decides it: the float64 row in gonum's order.
//   * after the sweep the exact fields of all probes are recomputed by the integer fields GEMM (igemm_tc05.cuh) and a scan
//     (dense_scan_kernel) takes the reference's stability test, counts its ties, and freezes the probes that are stable.
// After a few such sweeps the probes are handed to the sparse kernel with their sweep / flip / tie counters
// (RelaxParams::sweeps_done ...).  Every quantity is an exact integer: results are bit-identical to the sparse path.
#pragma once
#include "common.cuh"
#include "igemm_tc05.cuh"
#include "relax.cuh"
#include "tc05.cuh"

namespace hn {

constexpr int DS_T = 32;            // probes per tile (x 4 K-quarters = 128 MMA rows)
constexpr int DS_BLK = 32;          // positions per block
constexpr int DS_STAGES = 4;        // ring of weight tiles: 4 quarters x 32 rows x 128 B = 16 KiB per stage
constexpr int DS_STAGE_BYTES = 16384;
constexpr int DS_THREADS = 256;
constexpr uint32_t DS_NULL = 0x80000000u;

static inline int ds_np(int n) { return (n + 511) / 512 * 512; }        // padded K: four quarters of whole 128-byte chunks
static inline int ds_rows(int n) { return (n + 31) / 32 * 32; }         // padded positions
static inline size_t ds_smem_bytes(int n) {
    return (size_t)DS_T * ds_np(n) + (size_t)DS_STAGES * DS_STAGE_BYTES + 3 * 32 * 33 * 4 + 1024 + 256 + 1024;
}

// Per schedule row: the byte image with its rows in visiting order (zero padded to [rowsP][Np]), the in-block 32x32
// matrices, and per position the shared-memory offset of its unit's state byte (uniform part) with the unit index.
//   info[pos] = unit | chunk16 << 16 ... see ds_pos_info
struct DsPos { uint32_t off; uint32_t unit; };   // off: DS_NULL for padding positions
__global__ void dense_prep_rows_kernel(const int8_t* __restrict__ k8, int ld8, const uint16_t* __restrict__ perm, int n, int np,
                                       int8_t* __restrict__ mperm) {
    const int pos = blockIdx.x;   // one CTA per position
    const int8_t* src = pos < n ? k8 + (size_t)perm[pos] * ld8 : nullptr;
    uint32_t* dst = reinterpret_cast<uint32_t*>(mperm + (size_t)pos * np);
    for (int w = threadIdx.x; w < np / 4; w += blockDim.x) {
        uint32_t v = 0;
        if (src && 4 * w < ld8) v = reinterpret_cast<const uint32_t*>(src)[w];   // ld8 is a multiple of 16, entries beyond n are 0
        dst[w] = v;
    }
}
__global__ void dense_prep_blocks_kernel(const int8_t* __restrict__ k8, int ld8, const uint16_t* __restrict__ perm, int n, int np,
                                         int rows_p, int8_t* __restrict__ mbb, DsPos* __restrict__ info) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_p * DS_BLK) return;
    const int pos = t / DS_BLK, j = t % DS_BLK, b = pos / DS_BLK;
    const int other = b * DS_BLK + j;
    int8_t v = 0;
    if (pos < n && other < n) v = k8[(size_t)perm[pos] * ld8 + perm[other]];
    mbb[t] = v;   // [b][i][j] = M[unit(b,i)][unit(b,j)]
    if (j == 0) {
        DsPos d{DS_NULL, 0xFFFFu};
        if (pos < n) {
            const int k = perm[pos], kq = np / 4;
            const int q = k / kq, kk = k - q * kq;
            // byte (row 32q + p, column kk & 127) of panel kk >> 7: uniform part of the swizzled offset + the 16-byte chunk
            d.off = (uint32_t)(kk >> 7) * 16384u + (uint32_t)q * 4096u + (uint32_t)(kk & 15) + ((uint32_t)((kk >> 4) & 7) << 24);
            d.unit = (uint32_t)k;
        }
        info[pos] = d;
    }
}

struct DenseParams {
    int N, np, rows_p, words, B, domain, ldw;
    const int8_t* mbb;         // [rows_p/32][32][32] of this schedule row
    const DsPos* info;         // [rows_p]
    int row_base;              // first row of this schedule row inside the tensor map (r * rows_p)
    uint64_t* states;          // [B][words] current packed states (updated in place)
    const uint8_t* finished;   // [B]
    int32_t* sweeps_done;      // [B]
    int32_t* flips_done;       // [B]
    int32_t* margin_done;      // [B]
    const double* Wrow;        // float64 W (exact ties)
    int sweep_index;           // this is sweep number sweep_index + 1
    unsigned long long* flips_total;   // sum of this sweep's flips
};

// exact field of unit k for probe row p of the tile (DotUnitary order, see exact_field_warp) with the state read from S8
__device__ __forceinline__ double ds_exact_field(const double* __restrict__ row, const unsigned char* s8, int p, int N, int kq, int lane) {
    double acc = 0.0;
    const uint32_t prow = (uint32_t)(p >> 3) * 1024u + (uint32_t)(p & 7) * 128u;
    auto sval = [&](int j) -> double {
        const int q = j / kq, kk = j - q * kq;
        const uint32_t off = (uint32_t)(kk >> 7) * 16384u + (uint32_t)q * 4096u + prow + ((((uint32_t)(kk >> 4) ^ (uint32_t)p) & 7u) << 4) + (uint32_t)(kk & 15);
        return (double)(int)(signed char)s8[off];
    };
    if (lane < 4) {
        const int n4 = N & ~3;
        for (int j = lane; j < n4; j += 4) acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), sval(j)));
        if (lane == 0) for (int j = n4; j < N; j++) acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), sval(j)));
    }
    const double other = __shfl_down_sync(0xFFFFFFFFu, acc, 2);
    const double pair = __dadd_rn(acc, other);
    const double hi = __shfl_down_sync(0xFFFFFFFFu, pair, 1);
    return __shfl_sync(0xFFFFFFFFu, __dadd_rn(pair, hi), 0);
}

__global__ void __launch_bounds__(DS_THREADS, 1)
dense_sweep_kernel(const __grid_constant__ CUtensorMap map_w, DenseParams p) {
    using namespace tc05;
    extern __shared__ unsigned char ds_smem_raw[];
    const uint32_t raw = smem_u32(ds_smem_raw);
    unsigned char* smem = ds_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int kq = p.np / 4, KC = kq / 128, nblk = p.rows_p / DS_BLK;
    unsigned char* s8 = smem;                                              // KC panels of 128 rows x 128 B
    unsigned char* ring = smem + (size_t)DS_T * p.np;                     // == KC * 16384
    int* part = reinterpret_cast<int*>(ring + DS_STAGES * DS_STAGE_BYTES); // [3][32][33]
    uint32_t* mbb_s = reinterpret_cast<uint32_t*>(part + 3 * 32 * 33);     // [32][8] words: in-block matrix of the current block
    uint64_t* bars = reinterpret_cast<uint64_t*>(mbb_s + 256);
    uint64_t* full = bars;                  // [stages] TMA
    uint64_t* empty = bars + DS_STAGES;     // [stages] tcgen05.commit
    uint64_t* d_full = bars + 2 * DS_STAGES;        // block's fields are in TMEM
    uint64_t* s8_ready = bars + 2 * DS_STAGES + 1;  // S8 carries every flip up to the previous block
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DS_STAGES + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles = (p.B + DS_T - 1) / DS_T;
    const bool bipolar = p.domain == HN_DOMAIN_BIPOLAR;

    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(d_full, 1);
        mbar_init(s8_ready, 1);
        mbar_fence_init();
        tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    int stage = 0; uint32_t phase = 0;      // ring position (TMA producer and MMA issuer each keep their own copy)
    uint32_t s8_par = 0;                    // MMA issuer: parity of the next s8_ready completion (nblk + 1 per tile)
    uint32_t d_par = 0;                     // warps 4-7: parity of the next d_full completion (nblk per tile)

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int p0 = tile * DS_T;
        // ---- the tile's states: packed bits -> int8, K-quartered 128-byte-swizzled panels ----
        {
            const int words32 = p.words * 2, nw = p.np / 32;   // 32-unit groups per probe (padded)
            for (int t = tid; t < DS_T * nw; t += DS_THREADS) {
                const int pp = t / nw, w = t % nw;
                uint32_t bits = 0;
                const int unit0 = 32 * w;
                if (p0 + pp < p.B && w < words32 && unit0 < p.N)
                    bits = reinterpret_cast<const uint32_t*>(p.states + (size_t)(p0 + pp) * p.words)[w];
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) v[e] = tg_expand4((bits >> (4 * e)) & 0xFu, bipolar);
                if (unit0 + 32 > p.N || p0 + pp >= p.B) {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int b = 0; b < 4; b++) if (p0 + pp < p.B && unit0 + 4 * e + b < p.N) mask |= 0xFFu << (8 * b);
                        v[e] &= mask;
                    }
                }
                const int q = unit0 / kq, kk = unit0 - q * kq, row = 32 * q + pp;
                unsigned char* base = s8 + (size_t)(kk >> 7) * 16384 + (size_t)(row >> 3) * 1024 + (size_t)(row & 7) * 128;
                const uint32_t c0 = (uint32_t)((kk & 127) >> 4), sw = (uint32_t)(row & 7);
                *reinterpret_cast<uint4*>(base + ((c0 ^ sw) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<uint4*>(base + (((c0 + 1u) ^ sw) << 4)) = make_uint4(v[4], v[5], v[6], v[7]);
            }
            fence_proxy_async_smem();
        }
        __syncthreads();
        if (tid == 128) mbar_arrive(s8_ready);   // block 0 may start (warp 4, lane 0: the thread that arrives for every block)

        if (warp == 0) {
            // ===== TMA producer: block by block, the 4 K-quarters of the block's 32 weight rows, chunk by chunk =====
            if (lane == 0) {
                for (int b = 0; b < nblk; b++)
                    for (int kc = 0; kc < KC; kc++) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        mbar_arrive_expect_tx(&full[stage], DS_STAGE_BYTES);
                        unsigned char* dst = ring + (size_t)stage * DS_STAGE_BYTES;
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            tma_load_2d(dst + q * 4096, &map_w, q * kq + kc * 128, p.row_base + b * DS_BLK, &full[stage]);
                        if (++stage == DS_STAGES) { stage = 0; phase ^= 1u; }
                    }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = idesc_i8(128, 128, true, true);
            for (int b = 0; b < nblk; b++) {
                mbar_wait(s8_ready, s8_par);
                s8_par ^= 1u;
                tc_fence_after_sync();
                for (int kc = 0; kc < KC; kc++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint64_t da = smem_desc_k128(smem_u32(s8 + (size_t)kc * 16384));
                        const uint64_t db = smem_desc_k128(smem_u32(ring + (size_t)stage * DS_STAGE_BYTES));
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            mma_i8(tmem_base, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, (uint32_t)((kc | j) != 0));
                        mma_commit(&empty[stage]);
                        if (kc == KC - 1) mma_commit(d_full);
                    }
                    __syncwarp();
                    if (++stage == DS_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            mbar_wait(s8_ready, s8_par);   // the arrival that follows the last block's decisions: keeps the parity in step
            s8_par ^= 1u;
        } else if (warp >= 4) {
            // ===== partial fields out of TMEM (warp 4+q: quarter q), decisions (warp 4) =====
            const int q = warp - 4;
            const int pg = p0 + lane;                                  // this lane's probe (warp 4)
            bool active = false;
            int flips = 0, ties = 0;
            uint32_t* gbits = nullptr;
            if (q == 0) {
                active = pg < p.B && !p.finished[pg];
                gbits = reinterpret_cast<uint32_t*>(p.states + (size_t)(pg < p.B ? pg : 0) * p.words);
            }
            const uint32_t prow = (uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u;
            const int delta = bipolar ? 2 : 1, bias = bipolar ? 1 : 0;
            const int low_byte = bipolar ? 0xFF : 0x00;
            // next block's in-block matrix row and position info, fetched one block ahead (warp 4 only)
            uint4 mrow0 = make_uint4(0, 0, 0, 0), mrow1 = mrow0;
            DsPos pinfo{DS_NULL, 0xFFFFu};
            if (q == 0) {
                const uint4* src = reinterpret_cast<const uint4*>(p.mbb + (size_t)lane * 32);
                mrow0 = __ldg(src); mrow1 = __ldg(src + 1);
                pinfo = p.info[lane];
            }
            for (int b = 0; b < nblk; b++) {
                mbar_wait(d_full, d_par);
                d_par ^= 1u;
                tc_fence_after_sync();
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * q), r);
                tmem_ld_wait();
                tc_fence_before_sync();
                if (q > 0) {
                    int* dst = part + ((q - 1) * 32 + lane) * 33;
#pragma unroll
                    for (int i = 0; i < 32; i++) dst[i] = (int)r[i];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // warps 4-7: the partials are in shared memory
                if (q > 0) continue;

                // ---- warp 4: one probe per lane, the block's 32 decisions in order ----
                int u[32];
#pragma unroll
                for (int i = 0; i < 32; i++)
                    u[i] = (int)r[i] + part[lane * 33 + i] + part[(32 + lane) * 33 + i] + part[(64 + lane) * 33 + i] - bias;
                mbb_s[lane * 8 + 0] = mrow0.x; mbb_s[lane * 8 + 1] = mrow0.y; mbb_s[lane * 8 + 2] = mrow0.z; mbb_s[lane * 8 + 3] = mrow0.w;
                mbb_s[lane * 8 + 4] = mrow1.x; mbb_s[lane * 8 + 5] = mrow1.y; mbb_s[lane * 8 + 6] = mrow1.z; mbb_s[lane * 8 + 7] = mrow1.w;
                const DsPos cur = pinfo;
                if (b + 1 < nblk) {   // prefetch the next block's row / info; lands while the tensor core works on it
                    const uint4* src = reinterpret_cast<const uint4*>(p.mbb + ((size_t)(b + 1) * 32 + lane) * 32);
                    mrow0 = __ldg(src); mrow1 = __ldg(src + 1);
                    pinfo = p.info[(b + 1) * DS_BLK + lane];
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const uint32_t off_i = __shfl_sync(0xFFFFFFFFu, cur.off, i);
                    if (off_i == DS_NULL) continue;   // padding position (uniform)
                    const uint32_t chunk = off_i >> 24;
                    unsigned char* sp = s8 + (off_i & 0xFFFFFFu) + prow + (((chunk ^ (uint32_t)lane) & 7u) << 4);
                    const int sv = (int)(signed char)*sp;
                    const int bit = sv > 0 ? 1 : 0;
                    const int U = u[i];
                    bool cand = active && (bit ? (U < 0) : (U >= 0));
                    const bool tie = active && (U + 2 * bit == 0);
                    int nb = bit ^ 1;
                    const uint32_t tmask = __ballot_sync(0xFFFFFFFFu, tie);
                    if (tmask) {   // exact tie of the true field: the reference's float64 dot product decides (rare)
                        const int k = (int)__shfl_sync(0xFFFFFFFFu, cur.unit, i);
                        uint32_t m = tmask;
                        __syncwarp();   // the lanes' earlier flips in S8 are read by lanes 0-3 below
                        while (m) {
                            const int pl = __ffs((int)m) - 1;
                            m &= m - 1;
                            const double hx = ds_exact_field(p.Wrow + (size_t)k * p.ldw, s8, pl, p.N, kq, lane);
                            if (lane == pl) { nb = hx > 0.0 ? 1 : 0; ties++; }
                        }
                    }
                    const bool flip = cand && (nb != bit);
                    const int ds = flip ? (bit ? -delta : delta) : 0;
                    if (flip) *sp = (unsigned char)(bit ? low_byte : 0x01);
                    if (__ballot_sync(0xFFFFFFFFu, flip)) {
                        const uint4 w0 = *reinterpret_cast<const uint4*>(mbb_s + i * 8);
                        const uint4 w1 = *reinterpret_cast<const uint4*>(mbb_s + i * 8 + 4);
                        const uint32_t wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                        const uint32_t d8 = (uint32_t)ds & 0xFFu;
                        const uint32_t sel[4] = {d8, d8 << 8, d8 << 16, d8 << 24};
#pragma unroll
                        for (int j = i + 1; j < 32; j++) u[j] = __dp4a((int)wv[j >> 2], (int)sel[j & 3], u[j]);
                    }
                    if (flip) flips++;
                    // packed state in global memory: fire-and-forget toggle of this probe's bit
                    {
                        const uint32_t k = __shfl_sync(0xFFFFFFFFu, cur.unit, i);
                        if (flip) atomicXor(gbits + (k >> 5), 1u << (k & 31));
                    }
                }
                fence_proxy_async_smem();   // the flips written into S8 must be visible to the next block's MMAs
                __syncwarp();
                if (lane == 0) mbar_arrive(s8_ready);
            }
            if (q == 0 && pg < p.B && active) {
                p.sweeps_done[pg] = p.sweep_index + 1;
                p.flips_done[pg] += flips;
                p.margin_done[pg] += ties;
            }
            if (q == 0) {
                unsigned long long f = (unsigned long long)__reduce_add_sync(0xFFFFFFFFu, (unsigned)flips);
                if (lane == 0 && f) atomicAdd(p.flips_total, f);
            }
        }
        __syncthreads();   // every role has finished the tile: S8 may be rewritten
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// Stability test after a dense sweep (HopfieldNetwork.go:214-225) on the exact fields G = M s of the current states
// (integer fields GEMM): one warp per probe.  U = G - bias, g2 = U + 2b = 2g; unit unstable iff v*g < 0; g == 0 is an
// exact tie, decided on the float64 row in the reference's order and counted.  Stable probes are frozen.
__global__ void __launch_bounds__(256)
dense_scan_kernel(const int32_t* __restrict__ G, int ldg, const uint64_t* __restrict__ states, int words, int N, int B, int domain,
                  const double* __restrict__ W, int ldw, int max_unstable, uint8_t* __restrict__ finished,
                  int32_t* __restrict__ margin_done, int32_t* __restrict__ unfinished_total) {
    const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (pi >= B || finished[pi]) return;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(states + (size_t)pi * words);
    const int bias = domain == HN_DOMAIN_BIPOLAR ? 1 : 0;
    int cnt = 0, ties = 0;
    for (int base = 0; base < N; base += 32) {
        const int u = base + lane;
        bool tie = false;
        int bit = 0;
        if (u < N) {
            bit = (int)((bits[u >> 5] >> (u & 31)) & 1u);
            const int g2 = G[(size_t)pi * ldg + u] - bias + 2 * bit;
            if (bit ? (g2 < 0) : (g2 > 0)) cnt++;
            tie = g2 == 0;
        }
        uint32_t m = __ballot_sync(0xFFFFFFFFu, tie);
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            const int uu = base + l;
            const double hx = exact_field_warp(W + (size_t)uu * ldw, bits, N, domain, lane);
            if (lane == 0) {
                ties++;
                const double vi = ((bits[uu >> 5] >> (uu & 31)) & 1u) ? 1.0 : -1.0;
                if (-0.5 * (hx * vi) > 0.0) cnt++;
            }
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) {
        margin_done[pi] += ties;
        if (cnt <= max_unstable) finished[pi] = 1;
        else atomicAdd(unfinished_total, 1);
    }
}

}  // namespace hn
