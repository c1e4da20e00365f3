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
//   * four warps then take the 32 decisions of the block in order, four threads per probe (each owns eight positions):
//     a flip at position i is told to the probe's threads by one width-4 shuffle and corrects the fields of the later
//     positions of the block with the 32x32 in-block matrix (dp4a against a byte selector); it is written into S8 (so
//     the next block's contraction sees it) and into the packed state in global memory.  An exact tie (integer field
//     0) stops the probe; the warp then decides it as the sparse kernel does — the float64 row in gonum's order — and
//     the probe resumes from that position.
//   * after the sweep the exact fields of all probes are recomputed by the integer fields GEMM (igemm_tc05.cuh) and a scan
//     (dense_scan_kernel) takes the reference's stability test, counts its ties, and freezes the probes that are stable.
//   * the contraction of block b+1 overlaps the decisions of block b: it reads S8 without block b's flips, whose effect on
//     block b+1's fields is added afterwards from a precomputed 32x32 cross matrix (two TMEM accumulators alternate).
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
constexpr int DS_STAGES = 5;        // ring of weight tiles: 4 quarters x 32 rows x 128 B = 16 KiB per stage
constexpr int DS_STAGE_BYTES = 16384;
constexpr int DS_THREADS = 256;
constexpr uint32_t DS_NULL = 0x80000000u;

static inline int ds_np(int n) { return (n + 511) / 512 * 512; }        // padded K: four quarters of whole 128-byte chunks
static inline int ds_rows(int n) { return (n + 31) / 32 * 32; }         // padded positions
static inline size_t ds_smem_bytes(int n) {
    return (size_t)DS_T * ds_np(n) + (size_t)DS_STAGES * DS_STAGE_BYTES + 32 * 33 * 4 + 2 * 1024 + 128 + 256 + 256 + 1024;
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
                                         int rows_p, int8_t* __restrict__ mbb, int8_t* __restrict__ mx, DsPos* __restrict__ info) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_p * DS_BLK) return;
    const int pos = t / DS_BLK, j = t % DS_BLK, b = pos / DS_BLK;
    const int other = b * DS_BLK + j, prev = (b - 1) * DS_BLK + j;
    int8_t v = 0, x = 0;
    if (pos < n && other < n) v = k8[(size_t)perm[pos] * ld8 + perm[other]];
    if (pos < n && b > 0 && prev < n) x = k8[(size_t)perm[pos] * ld8 + perm[prev]];
    mbb[t] = v;   // [b][i][j] = M[unit(b,i)][unit(b,j)]
    mx[t] = x;    // [b][i][j] = M[unit(b,i)][unit(b-1,j)]: what the previous block's flips do to this block's fields
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
    int box_rows;              // rows per TMA box of the tensor map (a quarter's 32 rows arrive as 32 / box_rows boxes)
    int map3d;                 // the tensor map is the 3-D view [quarter][row][kq bytes]: one box per ring slot
    const int8_t* mbb;         // [rows_p/32][32][32] of this schedule row
    const int8_t* mx;          // [rows_p/32][32][32] cross matrices block b <- block b-1
    const DsPos* info;         // [rows_p]
    int row_base;              // first row of this schedule row inside the tensor map (r * rows_p)
    uint64_t* states;          // [B][words] current packed states (updated in place)
    const uint64_t* states_src;  // first sweep of a batch whose probes are still in mapped page-locked host memory: every tile
    uint64_t* states_copy;       // reads its probes from there and leaves them in `states` and in the resident probe buffer
    const uint8_t* finished;   // [B]
    int32_t* sweeps_done;      // [B]
    int32_t* flips_done;       // [B]
    int32_t* margin_done;      // [B]
    const double* Wrow;        // float64 W (exact ties)
    int sweep_index;           // this is sweep number sweep_index + 1
    unsigned long long* flips_total;   // sum of this sweep's flips
    unsigned long long* prof;          // optional [8]: clock cycles of CTA 0 by phase (HN_DENSE_PROF=1), else nullptr
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

// exact_field_warp (relax.cuh) on the packed state in GLOBAL memory, read past L1 (the state is toggled by atomics)
__device__ __forceinline__ double exact_field_warp_cg(const double* __restrict__ row, const uint32_t* bits, int N, int domain, int lane) {
    const double lowv = domain == HN_DOMAIN_BINARY ? 0.0 : -1.0;
    double acc = 0.0;
    if (lane < 4) {
        const int n4 = N & ~3;
        for (int j = lane; j < n4; j += 4) {
            const uint32_t b = (__ldcg(bits + (j >> 5)) >> (j & 31)) & 1u;
            acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), b ? 1.0 : lowv));
        }
        if (lane == 0)
            for (int j = n4; j < N; j++) {
                const uint32_t b = (__ldcg(bits + (j >> 5)) >> (j & 31)) & 1u;
                acc = __dadd_rn(acc, __dmul_rn(__ldg(row + j), b ? 1.0 : lowv));
            }
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
    int* part = reinterpret_cast<int*>(ring + DS_STAGES * DS_STAGE_BYTES); // [32][33] fields [probe][position]: the four K-quarters are
                                                                           // ADDED here by their warps (shared-memory atomics; integers, any
                                                                           // order) and every entry is zeroed again by the thread that reads it
                                                                           // — one buffer instead of four buys the ring its fifth slot
    uint32_t* mbb_s = reinterpret_cast<uint32_t*>(part + 32 * 33);         // [32][8] words: in-block matrix of the current block
    uint32_t* mx_s = mbb_s + 256;                                          // [32] rows of 8 words: cross matrix (previous block's flips); row i at
                                                                           // word 8 i + 4 (i >> 3): the four rows 8 sl + e that one 128-bit load of a
                                                                           // warp touches then sit in four different bank groups (a plain [32][8]
                                                                           // layout put them all on the same four banks: 16 wavefronts per load)
    uint32_t* pos_off = mx_s + 288;                                        // [32] state-byte offsets of the block's positions
    uint32_t* pos_unit = pos_off + 32;                                     // [32] their units
    uint64_t* bars = reinterpret_cast<uint64_t*>(pos_unit + 32);
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
    for (int i = tid; i < 32 * 33; i += DS_THREADS) part[i] = 0;
    if (warp == 1) tmem_alloc(tmem_slot, 256);   // two accumulators: block b+1 is contracted while block b is decided
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    int stage = 0; uint32_t phase = 0;      // ring position (TMA producer and MMA issuer each keep their own copy)
    uint32_t s8_par = 0;                    // MMA issuer: parity of the next s8_ready completion (nblk per tile)
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
                if (p0 + pp < p.B && w < words32) {
                    const size_t at = (size_t)(p0 + pp) * p.words * 2 + w;
                    if (p.states_src) {
                        bits = reinterpret_cast<const uint32_t*>(p.states_src)[at];
                        reinterpret_cast<uint32_t*>(p.states)[at] = bits;
                        reinterpret_cast<uint32_t*>(p.states_copy)[at] = bits;
                    } else bits = reinterpret_cast<const uint32_t*>(p.states)[at];
                    if (unit0 >= p.N) bits = 0;
                }
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
            // (the whole warp runs the loop converged; one elected lane issues, see tc05.cuh "warp-convergent variants")
            for (int b = 0; b < nblk; b++)
                for (int kc = 0; kc < KC; kc++) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx_elect(&full[stage], DS_STAGE_BYTES);
                    unsigned char* dst = ring + (size_t)stage * DS_STAGE_BYTES;
                    if (p.map3d) tma_load_3d_elect(dst, &map_w, kc * 128, p.row_base + b * DS_BLK, 0, &full[stage]);
                    else
                        for (int q = 0; q < 4; q++)
                            for (int r = 0; r < DS_BLK; r += p.box_rows)
                                tma_load_2d_elect(dst + q * 4096 + r * 128, &map_w, q * kq + kc * 128, p.row_base + b * DS_BLK + r, &full[stage]);
                    if (++stage == DS_STAGES) { stage = 0; phase ^= 1u; }
                }
        } else if (warp == 1) {
            // ===== MMA issuer (converged warp, elected issue) =====
            constexpr uint32_t idesc = idesc_i8(128, 128, true, true);
            const bool prof = p.prof && blockIdx.x == 0;
            long long pm_wait = 0, pm_issue = 0;
            const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);   // warp-uniform by construction
            const uint32_t s8_addr = smem_u32(s8), ring_addr = smem_u32(ring);
            for (int b = 0; b < nblk; b++) {
                const long long m0 = prof ? clock64() : 0;
                mbar_wait(s8_ready, s8_par);
                s8_par ^= 1u;
                tc_fence_after_sync();
                const long long m1 = prof ? clock64() : 0;
                pm_wait += m1 - m0;
                for (int kc = 0; kc < KC; kc++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after_sync();
                    const uint64_t da = smem_desc_k128(s8_addr + (uint32_t)kc * 16384u);
                    const uint64_t db = smem_desc_k128(ring_addr + (uint32_t)stage * (uint32_t)DS_STAGE_BYTES);
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        mma_i8_elect(tmem_u + (uint32_t)((b & 1) * 128), da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc,
                                     (uint32_t)((kc | j) != 0));
                    mma_commit_elect(&empty[stage]);
                    if (kc == KC - 1) mma_commit_elect(d_full);
                    if (++stage == DS_STAGES) { stage = 0; phase ^= 1u; }
                }
                if (prof) pm_issue += clock64() - m1;
            }
            if (prof && lane == 0) { atomicAdd(p.prof + 5, (unsigned long long)pm_wait); atomicAdd(p.prof + 6, (unsigned long long)pm_issue); }
        } else if (warp >= 4) {
            // ===== warps 4-7: partial fields out of TMEM (warp 4+q reads quarter q), then the block's 32 decisions =====
            // Pipeline: the contraction of block b+1 runs while block b is decided.  It reads S8 WITHOUT block b's flips;
            // their effect on block b+1's fields is added from the 32x32 cross matrix before block b+1 is decided, and
            // the flips are written into S8 right after the contraction of block b+1 has completed.
            // Decisions: 4 threads per probe (8 probes per warp), thread (probe, slice) owns positions 8*slice .. +7 of the
            // block.  Step i: the owning thread decides and tells the probe's four threads (one width-4 shuffle); every
            // thread then corrects ITS eight fields with the in-block matrix row (8 dp4a).  Fields of positions already
            // decided are dead, so nothing is masked.
            const int q = warp - 4;
            const int pl = 8 * q + (lane >> 2), sl = lane & 3;       // tile-local probe and slice of this thread
            const int pg = p0 + pl;
            const bool active = pg < p.B && !p.finished[pg];
            uint32_t* gbits = reinterpret_cast<uint32_t*>(p.states + (size_t)(pg < p.B ? pg : 0) * p.words);
            const uint32_t prow = (uint32_t)(pl >> 3) * 1024u + (uint32_t)(pl & 7) * 128u;
            const int delta = bipolar ? 2 : 1, bias = bipolar ? 1 : 0;
            const int low_byte = bipolar ? 0xFF : 0x00;
            int flips = 0, ties = 0;
            // warp 4: the next block's matrix rows and position info, fetched one block ahead (lane = row)
            uint4 mrow0 = make_uint4(0, 0, 0, 0), mrow1 = mrow0, xrow0 = mrow0, xrow1 = mrow0;
            DsPos pinfo{DS_NULL, 0xFFFFu};
            if (q == 0) {
                const uint4* src = reinterpret_cast<const uint4*>(p.mbb + (size_t)lane * 32);
                mrow0 = __ldg(src); mrow1 = __ldg(src + 1);
                pinfo = p.info[lane];
            }
            uint32_t dsw_prev[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // the previous block's flips of my probe, one signed byte per position
            uint32_t soff_prev[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fm_prev = 0, sb_prev = 0;   // ... and mine, not yet written into S8
            const bool prof = p.prof && blockIdx.x == 0 && tid == 128;
            long long pd[5] = {0, 0, 0, 0, 0};
            for (int b = 0; b < nblk; b++) {
                const long long t0 = prof ? clock64() : 0;
                mbar_wait(d_full, d_par);
                d_par ^= 1u;
                tc_fence_after_sync();
                const long long t1 = prof ? clock64() : 0;
                // the contraction of block b is complete: the previous block's flips may enter S8 now, before block b+1's starts
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if ((fm_prev >> e) & 1u) s8[soff_prev[e]] = (unsigned char)(((sb_prev >> e) & 1u) ? low_byte : 0x01);
                {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((b & 1) * 128 + 32 * q), r);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    int* dst = part + lane * 33;   // [probe][position]
#pragma unroll
                    for (int i = 0; i < 32; i++) atomicAdd(dst + i, (int)r[i]);
                }
                if (q == 0) {
                    *reinterpret_cast<uint4*>(mbb_s + lane * 8) = mrow0;
                    *reinterpret_cast<uint4*>(mbb_s + lane * 8 + 4) = mrow1;
                    *reinterpret_cast<uint4*>(mx_s + lane * 8 + (lane >> 3) * 4) = xrow0;
                    *reinterpret_cast<uint4*>(mx_s + lane * 8 + (lane >> 3) * 4 + 4) = xrow1;
                    pos_off[lane] = pinfo.off; pos_unit[lane] = pinfo.unit;
                    if (b + 1 < nblk) {   // prefetch; lands while this block is decided
                        const uint4* src = reinterpret_cast<const uint4*>(p.mbb + ((size_t)(b + 1) * 32 + lane) * 32);
                        const uint4* srx = reinterpret_cast<const uint4*>(p.mx + ((size_t)(b + 1) * 32 + lane) * 32);
                        mrow0 = __ldg(src); mrow1 = __ldg(src + 1);
                        xrow0 = __ldg(srx); xrow1 = __ldg(srx + 1);
                        pinfo = p.info[(b + 1) * DS_BLK + lane];
                    }
                }
                fence_proxy_async_smem();   // the flips written into S8 must be visible to the next block's MMAs
                asm volatile("bar.sync 1, 128;" ::: "memory");   // S8 is up to date; partials, matrices, position info are in shared memory
                if (tid == 128 && b + 1 < nblk) mbar_arrive(s8_ready);   // the contraction of block b+1 may start
                const long long t2 = prof ? clock64() : 0;

                // ---- my eight fields: partial sums of the four K-quarters + what the previous block's flips (which this
                // block's contraction did not see) add; loads first, arithmetic after (a warp issues in order) ----
                int u8[8];
                uint32_t soff[8], unit8[8], sb = 0, vm = 0;
                {
                    int pq[8];
                    uint4 xr[8][2];
                    uint32_t po[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int idx = 8 * sl + e;
                        pq[e] = part[pl * 33 + idx];
                        po[e] = pos_off[idx];
                        unit8[e] = pos_unit[idx];
                    }
#pragma unroll
                    for (int e = 0; e < 8; e++) part[pl * 33 + 8 * sl + e] = 0;   // ready for the next block's sums (after the block's closing barrier)
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const bool valid = po[e] != DS_NULL;
                        soff[e] = valid ? (po[e] & 0xFFFFFFu) + prow + ((((po[e] >> 24) ^ (uint32_t)pl) & 7u) << 4) : 0u;
                        vm |= (uint32_t)valid << e;
                    }
                    int sv[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) sv[e] = (int)(signed char)s8[soff[e]];   // a unit is visited once per sweep: still the tile's initial byte
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        xr[e][0] = *reinterpret_cast<const uint4*>(mx_s + (8 * sl + e) * 8 + sl * 4);
                        xr[e][1] = *reinterpret_cast<const uint4*>(mx_s + (8 * sl + e) * 8 + sl * 4 + 4);
                    }
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        sb |= (uint32_t)(((vm >> e) & 1u) && sv[e] > 0) << e;
                        int a0 = pq[e] - bias, a1 = 0;
                        a0 = __dp4a((int)xr[e][0].x, (int)dsw_prev[0], a0); a1 = __dp4a((int)xr[e][0].y, (int)dsw_prev[1], a1);
                        a0 = __dp4a((int)xr[e][0].z, (int)dsw_prev[2], a0); a1 = __dp4a((int)xr[e][0].w, (int)dsw_prev[3], a1);
                        a0 = __dp4a((int)xr[e][1].x, (int)dsw_prev[4], a0); a1 = __dp4a((int)xr[e][1].y, (int)dsw_prev[5], a1);
                        a0 = __dp4a((int)xr[e][1].z, (int)dsw_prev[6], a0); a1 = __dp4a((int)xr[e][1].w, (int)dsw_prev[7], a1);
                        u8[e] = a0 + a1;
                    }
                }
                if (!active) vm = 0;
                // per position of mine: xb = ~0 when the unit is low (then U >= 0 is the candidate side), the flip's signed
                // step as a byte (0: not a position to decide), and the field value that means an exact tie (U == -2b)
                uint32_t xb[8], dm[8];
                int tv[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const uint32_t bit = (sb >> e) & 1u;
                    xb[e] = bit ? 0u : 0xFFFFFFFFu;
                    dm[e] = ((vm >> e) & 1u) ? ((uint32_t)(bit ? -delta : delta) & 0xFFu) : 0u;
                    tv[e] = -2 * (int)bit;
                }
                uint32_t dsw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                uint32_t fm = 0, applied = 0;   // flips of my eight positions: decided / already sent to the packed state
                const long long t3 = prof ? clock64() : 0;

                // ---- pass 1, the common case: no memory access and no branch in the step-to-step dependency chain ----
                uint32_t live = 0xFFFFFFFFu;    // cleared when my probe stops at an exact tie
                int start = 32;                 // ... the position it stopped at
                for (int so = 0; so < 4; so++) {
                    uint2 w[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) w[e] = *reinterpret_cast<const uint2*>(mbb_s + (8 * so + e) * 8 + 2 * sl);
                    const uint32_t segm = sl == so ? 0xFFFFFFFFu : 0u;
                    uint32_t d0 = 0, d1 = 0;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int U = u8[e];
                        const uint32_t h = dm[e] & segm & live;                        // my decision to take, and its step byte
                        uint32_t msg = ((uint32_t)(U >> 31) ^ xb[e]) & h;              // candidate <=> sign(U) == bit
                        if (U == tv[e] && h != 0u) msg = 0x100u;                       // exact tie of the true field: stop here
                        fm |= (uint32_t)((msg & 0xFFu) != 0u) << e;
                        msg = __shfl_sync(0xFFFFFFFFu, msg, so, 4);                    // the owner's verdict to the probe's four threads
                        if (msg & 0x100u) { live = 0u; start = 8 * so + e; }
                        const uint32_t d8 = msg & 0xFFu;
                        if (e < 4) d0 |= d8 << (8 * (e & 3)); else d1 |= d8 << (8 * (e & 3));
                        const int s0 = (int)d8, s1 = (int)(d8 << 8), s2 = (int)(d8 << 16), s3 = (int)(d8 << 24);
                        u8[0] = __dp4a((int)w[e].x, s0, u8[0]); u8[1] = __dp4a((int)w[e].x, s1, u8[1]);
                        u8[2] = __dp4a((int)w[e].x, s2, u8[2]); u8[3] = __dp4a((int)w[e].x, s3, u8[3]);
                        u8[4] = __dp4a((int)w[e].y, s0, u8[4]); u8[5] = __dp4a((int)w[e].y, s1, u8[5]);
                        u8[6] = __dp4a((int)w[e].y, s2, u8[6]); u8[7] = __dp4a((int)w[e].y, s3, u8[7]);
                    }
#pragma unroll
                    for (int sg = 0; sg < 4; sg++) if (sg == so) { dsw[2 * sg] = d0; dsw[2 * sg + 1] = d1; }
                }
                auto send_flips = [&]() {   // this pass's flips to the packed state in global memory (fire and forget); S8 gets
                    const uint32_t pend = fm & ~applied;   // them when the next block's contraction, which must not see them, is complete
                    applied = fm;
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        if ((pend >> e) & 1u) atomicXor(gbits + (unit8[e] >> 5), 1u << (unit8[e] & 31));
                    flips += __popc(pend);
                };
                send_flips();
                // ---- exact ties (rare): decide on the probe's current state, resume the probe from that position ----
                uint32_t m = __ballot_sync(0xFFFFFFFFu, start < 32 && sl == 0);
                while (m) {
                    int res_nb = 0;
                    bool have_res = false;
                    __threadfence();   // the toggles above are performed before the packed state is read below
                    __syncwarp();
                    while (m) {   // the reference's float64 dot product in gonum's order
                        const int l0 = __ffs((int)m) - 1;
                        m &= m - 1;
                        const int pos = __shfl_sync(0xFFFFFFFFu, start, l0);
                        const int k = (int)pos_unit[pos];
                        const int pt = p0 + 8 * q + (l0 >> 2);
                        const double hx = exact_field_warp_cg(p.Wrow + (size_t)k * p.ldw,
                                                              reinterpret_cast<const uint32_t*>(p.states + (size_t)pt * p.words), p.N, p.domain, lane);
                        if ((lane & ~3) == l0) { have_res = true; res_nb = hx > 0.0 ? 1 : 0; if (sl == 0) ties++; }
                    }
                    // general pass: probes without a pending decision have start == 32 and stay idle
                    bool stalled = false;
                    int nstart = 32;
                    for (int so = 0; so < 4; so++) {
                        uint32_t d0 = 0, d1 = 0;
#pragma unroll
                        for (int sg = 0; sg < 4; sg++) if (sg == so) { d0 = dsw[2 * sg]; d1 = dsw[2 * sg + 1]; }
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const int i = 8 * so + e;
                            const uint2 w = *reinterpret_cast<const uint2*>(mbb_s + i * 8 + 2 * sl);
                            const int U = u8[e], bit = (int)((sb >> e) & 1u);
                            const bool elig = sl == so && ((vm >> e) & 1u) && !stalled && i >= start;
                            const bool cand = bit ? (U < 0) : (U >= 0);
                            const bool pre = have_res && i == start;            // the tie resolved above
                            const bool stall_now = elig && (U + 2 * bit == 0) && !pre;
                            const int nb = pre ? res_nb : (bit ^ 1);
                            const bool flip = elig && !stall_now && cand && nb != bit;
                            int msg = (flip ? (bit ? -delta : delta) : 0) & 0xFF;
                            if (stall_now) msg |= 0x100;
                            fm |= (uint32_t)flip << e;
                            msg = __shfl_sync(0xFFFFFFFFu, msg, so, 4);
                            if (msg & 0x100) { stalled = true; nstart = i; }
                            const uint32_t d8 = (uint32_t)msg & 0xFFu;
                            if (e < 4) d0 |= d8 << (8 * (e & 3)); else d1 |= d8 << (8 * (e & 3));
                            const int s0 = (int)d8, s1 = (int)(d8 << 8), s2 = (int)(d8 << 16), s3 = (int)(d8 << 24);
                            u8[0] = __dp4a((int)w.x, s0, u8[0]); u8[1] = __dp4a((int)w.x, s1, u8[1]);
                            u8[2] = __dp4a((int)w.x, s2, u8[2]); u8[3] = __dp4a((int)w.x, s3, u8[3]);
                            u8[4] = __dp4a((int)w.y, s0, u8[4]); u8[5] = __dp4a((int)w.y, s1, u8[5]);
                            u8[6] = __dp4a((int)w.y, s2, u8[6]); u8[7] = __dp4a((int)w.y, s3, u8[7]);
                        }
#pragma unroll
                        for (int sg = 0; sg < 4; sg++) if (sg == so) { dsw[2 * sg] = d0; dsw[2 * sg + 1] = d1; }
                    }
                    start = nstart;
                    send_flips();
                    m = __ballot_sync(0xFFFFFFFFu, start < 32 && sl == 0);
                }
                const long long t4 = prof ? clock64() : 0;
#pragma unroll
                for (int e = 0; e < 8; e++) { dsw_prev[e] = dsw[e]; soff_prev[e] = soff[e]; }
                fm_prev = fm; sb_prev = sb;
                asm volatile("bar.sync 1, 128;" ::: "memory");   // all four warps are done with the block's shared scratch
                if (prof) { const long long t5 = clock64(); pd[0] += t1 - t0; pd[1] += t2 - t1; pd[2] += t3 - t2; pd[3] += t4 - t3; pd[4] += t5 - t4; }
            }
            if (prof) for (int i = 0; i < 5; i++) atomicAdd(p.prof + i, (unsigned long long)pd[i]);
            flips += __shfl_xor_sync(0xFFFFFFFFu, flips, 1); flips += __shfl_xor_sync(0xFFFFFFFFu, flips, 2);
            ties += __shfl_xor_sync(0xFFFFFFFFu, ties, 1); ties += __shfl_xor_sync(0xFFFFFFFFu, ties, 2);
            if (sl == 0 && active) {
                p.sweeps_done[pg] = p.sweep_index + 1;
                p.flips_done[pg] += flips;
                p.margin_done[pg] += ties;
            }
            const unsigned wf = __reduce_add_sync(0xFFFFFFFFu, sl == 0 ? (unsigned)flips : 0u);
            if (lane == 0 && wf) atomicAdd(p.flips_total, (unsigned long long)wf);
        }
        __syncthreads();   // every role has finished the tile: S8 may be rewritten
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// Stability test after a dense sweep (HopfieldNetwork.go:214-225) on the exact fields G = M s of the current states
// (integer fields GEMM): one warp per probe.  U = G - bias, g2 = U + 2b = 2g; unit unstable iff v*g < 0; g == 0 is an
// exact tie, decided on the float64 row in the reference's order and counted.  Stable probes are frozen.
// After the fused scan of the fields contraction (igemm_tc05.cuh): per probe, unstable = #{v*g < 0}, zero = #{g == 0}.
// More strictly unstable units than the tolerance -> the probe goes on (its zero fields are the test's exact ties: counted).
// Otherwise, with no zero field the probe is stable and frozen; with zero fields the float64 re-evaluation of those units
// decides (dense_scan_kernel on the stored fields, only for such probes: only[p] = 1).
__global__ void dense_classify_kernel(const int32_t* __restrict__ unstable, const int32_t* __restrict__ zero, int B, int max_unstable,
                                      uint8_t* __restrict__ finished, int32_t* __restrict__ margin_done, uint8_t* __restrict__ only,
                                      int32_t* __restrict__ unfinished_total, int32_t* __restrict__ ambiguous_total) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= B) return;
    only[pi] = 0;
    if (finished[pi]) return;
    if (unstable[pi] > max_unstable) { margin_done[pi] += zero[pi]; atomicAdd(unfinished_total, 1); }
    else if (zero[pi] == 0) finished[pi] = 1;
    else { only[pi] = 1; atomicAdd(ambiguous_total, 1); }
}

__global__ void __launch_bounds__(256)
dense_scan_kernel(const int32_t* __restrict__ G, int ldg, const uint64_t* __restrict__ states, int words, int N, int B, int domain,
                  const double* __restrict__ W, int ldw, int max_unstable, uint8_t* __restrict__ finished,
                  int32_t* __restrict__ margin_done, int32_t* __restrict__ unfinished_total, const uint8_t* __restrict__ only) {
    const int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (pi >= B || finished[pi] || (only && !only[pi])) return;
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
