// api.cu — C ABI of libhopfield_b200.so (include/hopfield_b200.h).
// One handle = one network replica resident on one B200: W (float64, row-major,
// zero padded to a multiple of 16 columns) in HBM/L2, packed states, one stream.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "gemm_f64.cuh"
#include "igemm_s8.cuh"
#include "igemm_tc05.cuh"
#include "dense.cuh"
#include "kernels_misc.cuh"
#include "relax.cuh"

namespace hn {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- NCCL through dlopen (no link-time dependency; loaded on first use) ----
struct Id128 { char bytes[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.lib) return HN_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { set_error("NCCL not loadable: %s", dlerror()); return HN_E_NCCL; }
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclBroadcast");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.Broadcast) {
        set_error("NCCL symbols missing");
        return HN_E_NCCL;
    }
    return HN_OK;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_INT32 = 2, NCCL_SUM = 0, NCCL_MIN = 4;
#define HN_NCCL_TRY(expr)                                                                   \
    do {                                                                                    \
        int _r = (expr);                                                                    \
        if (_r != 0) {                                                                      \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return HN_E_NCCL;                                                               \
        }                                                                                   \
    } while (0)

// host_rng.cpp: the x/exp/rand PCG source behind hn_rng (state access + jump-ahead for device-side generation)
void rng_jump_table(uint64_t table[64][4]);
void rng_get_state(hn_rng r, uint64_t out[2]);
void rng_advance(hn_rng r, uint64_t draws);

}  // namespace hn

using namespace hn;

// Small device scratch block of a handle (one allocation, h->small): every kernel-side counter / flag lives here.
struct Scratch {
    int32_t work_counter;              // relax: dynamic probe scheduler
    int32_t status[2];                 // relax: [0] pool exhausted, [1] rows needed
    int32_t unique_total;              // dedup: number of unique attractors
    int32_t asym_flag;                 // detect_symmetry
    int32_t int_flag;                  // build_integer_structure: scale*Wraw not an int16 integer somewhere
    int32_t not_quarters;              // ensure_margin: some W_ij is not a multiple of 0.25
    int32_t stable_flag;               // learning: all-stable flag of this rank (all-reduced with min)
    unsigned long long rowabs_bits;    // ensure_margin: max_i sum_k |W_ik| (double bits)
    unsigned long long maxabs_bits;    // ensure_margin: max |W_ij|
    unsigned long long maxabs_raw_bits;  // build_integer_structure: max |Wraw_ij|
    int32_t dense_ambiguous;           // dense sweeps: probes whose stability hangs on exactly-zero fields (float64 re-evaluation)
    int32_t dense_unfinished;          // dense sweeps: probes not yet stable after the last scan
    int32_t frac_flags;                // fused update: 1 not even integers, 2 not integers, 4 not halves somewhere in W
    unsigned long long dense_flips;    // dense sweeps: flips of the last sweep
    int32_t pair_count[2];             // dedup level 2: [0] merges of different states recorded, [1] refuted by the float profiles
    unsigned long long totals[3];      // relax: sum of flips, sweeps, margin hits over the batch (sum_counters_kernel)
};

struct hn_handle_s {
    hn_config cfg;
    int N = 0, ldw = 0, words = 0, ldp = 0, sms = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr;               // copy stream: probes in / results out under the kernels of the next range
    std::vector<cudaEvent_t> chunk_ev;        // six events per range of a chunked batch
    std::vector<cudaEvent_t> dense_ev;        // four events per dense sweep (start, sweep kernel, fields contraction, scan)
    double acc_dense_kernel_ms = 0, acc_dense_gemm_ms = 0, acc_dense_scan_ms = 0;   // sums over the ranges of the current batch
    int64_t acc_dense_flips = 0;
    double* W = nullptr;
    double* WT = nullptr;
    double* dW = nullptr;
    float* W32 = nullptr;       // float32 shadow of the column matrix (HN_COLUMNS_F32)
    double* Wraw = nullptr;     // un-normalised W captured at normalisation (HN_COLUMNS_I16): W = c * Wraw
    int8_t* K8 = nullptr;       // the same image as signed bytes while every |M| <= 127 (tcgen05 fields GEMM); row-major [N][ldw]
    bool k8_valid = false;
    bool use_tc05 = true;       // HN_IGEMM=mma_sync in the environment selects the mma.sync kernel instead (A/B comparison)
    int16_t* K16 = nullptr;     // int16 image M = 2*int_scale*Wraw (+ self-correction on the diagonal, to_i16_kernel), column matrix layout
    bool int_ok = false;        // the integer structure of the CURRENT W is known and fits int16
    double int_maxabs = 0.0;    // max |K16 entry|
    double key_scale = 0.0;     // int_ok: (K s)_i = rint(key_scale * (W s)_i): exact energy-profile key on the float streams
    double int_scale = 2.0;     // K16 = int_scale * Wraw: the smallest of 0.5, 1, 2, 4 that makes it integer valued
    bool w32_valid = false;
    double maxabs = 0;
    bool w_symmetric = true;   // W known symmetric -> column k == row k
    bool margin_valid = false;
    double tau_base = 0, tau_flip = 0;
    bool exact_arith = false;   // every W_ij is a multiple of 0.25 and N*max|W| is small: fields are exact in float64
    int P = 0;
    DevBuf targets;
    // relaxation working set
    DevBuf jump;   // LCG jump table of hn_generate_probes
    DevBuf zf;     // per probe: final state has an exactly-zero field (unique-table key)
    DevBuf probes, pool, offsets, H, finals, steps, stable, flips, margin, dist, hash, energies, history;
    DevBuf rep, rep1, rep2, pairs, first, hits, flag, rank, table, attractor, ufirst, uhits, small, partial;
    int64_t verified_pairs = 0, refuted_pairs = 0;   // unique table, level 2: merges of different states checked / refuted
    DevBuf bitsA, bitsB, bitsC, bitsAT, bitsCT, perms, eunst, estab, esum, emargin;
    uint16_t* d_pool = nullptr;
    uint16_t* d_inv = nullptr;
    int resident_B = 0, resident_pool = 0, result_B = 0, result_P = 0, result_unique = 0;
    uint32_t result_flags = 0;
    bool result_profile_keyed = false;
    bool sweep_timed = false;
    // dense lockstep sweeps (dense.cuh)
    DevBuf d_mperm, d_mbb, d_mx, d_info, d_state, d_sweeps, d_dflips, d_dmargin, d_finished, d_scan;   // d_scan: unstable[B], zero[B], only[B]
    CUtensorMap dense_map;
    // zero-copy results (hn_relax_batch with page-locked host buffers): device-visible addresses of the caller's arrays
    bool dense_want_map3d = true, dense_map3d = false;   // HN_DENSE_MAP=2d keeps the four 2-D boxes per slot
    int dense_box_rows = 32;           // rows per TMA box of the dense sweeps weight stream (HN_TMA_BOX_DENSE; 8-row boxes measured 2.3x slower)
    bool exact_norm = false;           // hn_set_norm_algorithm: classic Dlassq recurrence on the host (true) / exact parallel sum = LAPACK-3.10 Dlassq (false); HN_EXACT_NORM presets it
    bool zero_copy = true;             // HN_ZERO_COPY=0 turns it off
    uint64_t* zc_finals = nullptr;
    int32_t* zc_dist = nullptr;
    std::vector<uint16_t> pool_host;       // the schedule last uploaded into `pool` (an identical one is not uploaded again)
    int pool_host_rows = 0;
    const uint64_t* zc_probes = nullptr;   // this range's probes in the caller's page-locked memory (read by the first dense sweep)
    int dense_rows = 0;
    uint64_t w_version = 1, pool_version = 1, dense_w_version = 0, dense_pool_version = 0;
    bool use_dense = true;
    int dense_max_rows = 6, dense_min_n = 2048;
    double dense_min_flips = 150.0;   // sweep densely again while the last sweep flipped more than this per unfinished probe (the next one
                                      // flips ~0.75x as many; a dense sweep costs what ~160 flips per probe cost the sparse kernel)
    int last_dense_sweeps = 0;
    int frac_flags = 7; uint64_t frac_version = 0;   // which fractions occur in W (fused update kernel), and for which W
    uint64_t image_tried_version = 0;   // ensure_learning_image: the W version the image was last (tried to be) built for
    bool fused_flags_valid = false;   // Scratch::maxabs_bits / not_quarters describe the current W (update_constraints_i32_kernel)
    DevBuf opA8, opB8;                // int8 operands of the learning contractions (tcgen05 path)
    cudaEvent_t ev[6] = {};
    cudaEvent_t lev[8] = {};   // learning-epoch stage events
    hn_learn_timing ltiming = {};
    hn_relax_timing timing = {};
    int64_t tot_flips = 0, tot_sweeps = 0, tot_margin = 0;
    void* comm = nullptr;
    Scratch* scratch() const { return small.as<Scratch>(); }
    int rank_id = 0, n_ranks = 1;
};

namespace {

int set_device(hn_handle h) {
    HN_CUDA_TRY(cudaSetDevice(h->cfg.device));
    return HN_OK;
}

struct ValueMap { double lo, hi; };
ValueMap domain_values(int domain) { return domain == HN_DOMAIN_BINARY ? ValueMap{0.0, 1.0} : ValueMap{-1.0, 1.0}; }

int grid_for(int64_t total, int threads, int sms) {
    int64_t g = (total + threads - 1) / threads;
    int64_t cap = (int64_t)sms * 8;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

int refresh_wt(hn_handle h) {
    if (h->w_symmetric) return HN_OK;
    if (!h->WT) HN_CUDA_TRY(cudaMalloc(&h->WT, sizeof(double) * (size_t)h->N * h->ldw));
    HN_CUDA_TRY(cudaMemsetAsync(h->WT, 0, sizeof(double) * (size_t)h->N * h->ldw, h->st));
    dim3 g((h->N + SYM_T - 1) / SYM_T, (h->N + SYM_T - 1) / SYM_T), b(SYM_T, 8);
    transpose_kernel<<<g, b, 0, h->st>>>(h->W, h->WT, h->N, h->ldw);
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}

int ensure_margin(hn_handle h) {
    if (h->margin_valid) return HN_OK;
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    Scratch* sc = h->scratch();
    if (h->fused_flags_valid) {   // the fused update kernel already scanned the new W (max |W|, quarter test)
        h->fused_flags_valid = false;
        double mx = 0; int32_t nq = 1, ff = 7;
        HN_CUDA_TRY(cudaMemcpyAsync(&mx, &sc->maxabs_bits, 8, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaMemcpyAsync(&nq, &sc->not_quarters, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaMemcpyAsync(&ff, &sc->frac_flags, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        h->frac_flags = ff; h->frac_version = h->w_version;
        if (!nq && 4.0 * (double)h->N * mx < 2251799813685248.0) {
            // exact arithmetic (see below): no margin is ever consulted; keep a valid upper bound in tau anyway
            const double u = 1.1102230246251565e-16, r = (double)h->N * mx;
            h->maxabs = mx;
            h->tau_base = 2.0 * (1.25 * h->N + 4.0) * u * r;
            h->tau_flip = 2.0 * u * r;
            h->exact_arith = true;
            h->margin_valid = true;
            return HN_OK;
        }
    }
    HN_CUDA_TRY(cudaMemsetAsync(&sc->rowabs_bits, 0, 16, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(&sc->not_quarters, 0, 4, h->st));
    rowabs_max_kernel<<<h->N, RED_THREADS, 0, h->st>>>(h->W, h->N, h->ldw, &sc->rowabs_bits);
    HN_CUDA_TRY(cudaGetLastError());
    maxabs_kernel<<<h->sms * 4, 256, 0, h->st>>>(h->W, h->N, h->ldw, &sc->maxabs_bits, &sc->not_quarters);
    HN_CUDA_TRY(cudaGetLastError());
    double rr[2] = {0, 0};
    int32_t not_quarters = 1;
    HN_CUDA_TRY(cudaMemcpyAsync(rr, &sc->rowabs_bits, 16, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaMemcpyAsync(&not_quarters, &sc->not_quarters, 4, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    const double r = rr[0];
    h->maxabs = rr[1];
    const double u = 1.1102230246251565e-16;  // 2^-53
    // |incremental field - fresh DotUnitary field| <= (N [GEMM] + N/4+3 [fresh] + flips) * u * R; x2 safety
    h->tau_base = 2.0 * (1.25 * h->N + 4.0) * u * r;
    h->tau_flip = 2.0 * u * r;
    // Exact arithmetic: all entries multiples of 1/4 (un-normalised Hebbian / Delta weights) and every partial sum of a
    // field, N*max|W| in units of 1/4, far inside 2^53: GEMM, incremental updates and the reference's own dot products
    // are all exact, so the incremental field EQUALS the reference's and no margin is needed (a zero field is a zero).
    h->exact_arith = !not_quarters && 4.0 * (double)h->N * h->maxabs < 2251799813685248.0;   // 2^51
    h->margin_valid = true;
    return HN_OK;
}

void weights_changed(hn_handle h) { h->w_version++; h->fused_flags_valid = false; h->margin_valid = false; h->w32_valid = false; h->int_ok = false; h->key_scale = 0.0; h->k8_valid = false; }

// exact symmetry test of the resident W (decides whether column k can be read as row k)
int detect_symmetry(hn_handle h) {
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    int32_t* flag = &h->scratch()->asym_flag;
    HN_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, h->st));
    dim3 g((h->N + SYM_T - 1) / SYM_T, (h->N + SYM_T - 1) / SYM_T), b(SYM_T, 8);
    asymmetry_kernel<<<g, b, 0, h->st>>>(h->W, h->N, h->ldw, flag);
    HN_CUDA_TRY(cudaGetLastError());
    int32_t f = 0;
    HN_CUDA_TRY(cudaMemcpyAsync(&f, flag, 4, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    h->w_symmetric = (f == 0);
    return HN_OK;
}

// H[count][ldw] = S * W^T
int launch_fields(hn_handle h, const uint64_t* d_states, int count, double* d_H, ValueMap vm, const double* mat = nullptr) {
    OpBits a{d_states, h->words, count, h->N, vm.lo, vm.hi};
    OpF64 b{mat ? mat : h->W, h->ldw, h->N};
    EpiStore e{d_H, h->ldw, count, h->N};
    return launch_gemm_f64(a, b, e, count, h->N, h->N, h->st);
}

// Integer structure for the exact integer stream from h->Wraw (already on the device): K16 = scale*Wraw in column-matrix
// layout with the smallest scale in {0.5, 1, 2, 4} that makes every entry an integer (Hebbian: sums of +-1 products, all
// of the parity of P; Delta: half / quarter integers); int_ok only if every entry fits int16 and the exactness bound of
// the field holds.  (An int8 image with dp4a was tried for |K| <= 127: half the column bytes, no faster — the kernel is
// bound by its dependent chain and issue slots, not by the L2->SM bytes; profiles/r01_relax_i16_final_ncu_summary.txt.)
int build_integer_structure(hn_handle h, bool symmetric, double c, const double* raw = nullptr, int first_scale = 0,
                            double known_max = -1.0) {   // c: W = fl(c * raw); raw defaults to h->Wraw
    if (!raw) raw = h->Wraw;
    const int64_t total = (int64_t)h->N * h->ldw;
    if (!h->K16) HN_CUDA_TRY(cudaMalloc(&h->K16, sizeof(int16_t) * (size_t)total));
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    int32_t* flag = &h->scratch()->int_flag;
    unsigned long long* mx = &h->scratch()->maxabs_raw_bits;
    HN_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(mx, 0, 8, h->st));
    const double* src = raw;
    if (!symmetric) {  // column k of Wraw must be contiguous: transpose into dW (free outside learning epochs)
        HN_CUDA_TRY(cudaMemsetAsync(h->dW, 0, sizeof(double) * (size_t)total, h->st));
        dim3 g((h->N + SYM_T - 1) / SYM_T, (h->N + SYM_T - 1) / SYM_T), b(SYM_T, 8);
        transpose_kernel<<<g, b, 0, h->st>>>(raw, h->dW, h->N, h->ldw);
        src = h->dW;
    }
    if (known_max < 0.0) {
        maxabs_kernel<<<h->sms * 4, 256, 0, h->st>>>(raw, h->N, h->ldw, mx, nullptr);
        HN_CUDA_TRY(cudaGetLastError());
    } else {
        HN_CUDA_TRY(cudaMemcpyAsync(mx, &known_max, 8, cudaMemcpyHostToDevice, h->st));
    }
    int32_t f = 1; double m = 0;
    const double scales[4] = {0.5, 1.0, 2.0, 4.0};
    for (int si = first_scale; si < 4; si++) {
        const double scale = scales[si];
        HN_CUDA_TRY(cudaMemsetAsync(flag, 0, 4, h->st));
        to_i16_kernel<<<grid_for(total, 256, h->sms), 256, 0, h->st>>>(src, h->K16, total, h->ldw, scale,
                                                                       h->cfg.domain == HN_DOMAIN_BINARY ? -2 : -1, flag);
        HN_CUDA_TRY(cudaGetLastError());
        HN_CUDA_TRY(cudaMemcpyAsync(&f, flag, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaMemcpyAsync(&m, mx, 8, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        h->int_scale = scale;
        if (f == 0) break;
    }
    // fresh-dot rounding error << c  <=>  N^2 * max|2 Wraw| * 2^-50 < 0.5   (same precondition as oracle_fast.c)
    const bool exact = (double)h->N * (double)h->N * (h->int_scale * m) * 8.881784197001252e-16 < 0.5;
    h->int_ok = (f == 0) && exact && m > 0.0;
    h->int_maxabs = 2.0 * h->int_scale * m + 2.0;   // bound on |M| = |2K + diagonal term|
    h->key_scale = (h->int_ok && c > 0.0) ? h->int_scale / c : 0.0;   // K s = key_scale * (W s), to the nearest integer
    h->k8_valid = false;
    if (h->int_ok && symmetric && h->int_maxabs <= 127.0) {   // byte image for the tcgen05 fields GEMM
        if (!h->K8) HN_CUDA_TRY(cudaMalloc(&h->K8, (size_t)total));
        i16_to_i8_kernel<<<grid_for(total / 4, 256, h->sms), 256, 0, h->st>>>(h->K16, h->K8, total / 4);
        HN_CUDA_TRY(cudaGetLastError());
        h->k8_valid = true;
    }
    return HN_OK;
}

// Caller-supplied weights (hn_set_weights, hn_broadcast_weights): try to RECOVER the integer structure the library would
// have captured had it learned them.  A Hebbian / Delta matrix after Scale(1/norm) is W = fl(c*K), K integer; the
// smallest non-zero |W| is then fl(c*k) for a small k.  For k = 1..8: c = w_min/k, K = rint(W/c), accepted iff
// fl(c*K_ij) == W_ij bit for bit everywhere (plus the int16 / exactness checks of build_integer_structure).  A random
// initial matrix, thermal weights or anything else fails the bit test and keeps the float64 stream.
int detect_integer_structure(hn_handle h) {
    const int64_t total = (int64_t)h->N * h->ldw;
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    Scratch* sc = h->scratch();
    const unsigned long long inf_bits = 0x7FF0000000000000ULL;
    HN_CUDA_TRY(cudaMemcpyAsync(&sc->maxabs_raw_bits, &inf_bits, 8, cudaMemcpyHostToDevice, h->st));
    min_nonzero_abs_kernel<<<h->sms * 4, 256, 0, h->st>>>(h->W, total, &sc->maxabs_raw_bits);
    HN_CUDA_TRY(cudaGetLastError());
    double wmin = 0.0;
    HN_CUDA_TRY(cudaMemcpyAsync(&wmin, &sc->maxabs_raw_bits, 8, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    if (!(wmin > 0.0) || !std::isfinite(wmin)) return HN_OK;   // the zero matrix: nothing to stream
    if (!h->Wraw) HN_CUDA_TRY(cudaMalloc(&h->Wraw, sizeof(double) * (size_t)total));
    for (int k = 1; k <= 8; k++) {
        const double c = wmin / (double)k;
        int32_t f = 1;
        HN_CUDA_TRY(cudaMemsetAsync(&sc->int_flag, 0, 4, h->st));
        recover_integer_kernel<<<grid_for(total, 256, h->sms), 256, 0, h->st>>>(h->W, h->Wraw, total, c, &sc->int_flag);
        HN_CUDA_TRY(cudaGetLastError());
        HN_CUDA_TRY(cudaMemcpyAsync(&f, &sc->int_flag, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        if (f == 0) return build_integer_structure(h, h->w_symmetric, c);   // Wraw = K, W = fl(c*K)
    }
    return HN_OK;
}

bool wants_integer_stream(hn_handle h);
bool integer_stream(hn_handle h);
// Un-normalised weights inside a learning run (quarter-integers, exact arithmetic): the integer image is rebuilt from W
// itself (W = 1 * W), so that the sweeps of the next Delta epoch run on the integer stream and the fields of the sweep
// and of the per-epoch diagnostics come from the integer tensor path instead of three FP64 GEMMs.  Symmetric W only.
int ensure_learning_image(hn_handle h) {
    if (h->int_ok || !h->use_tc05 || !wants_integer_stream(h) || !h->w_symmetric) return HN_OK;
    HN_TRY(ensure_margin(h));
    if (!h->exact_arith || h->maxabs == 0.0 || h->image_tried_version == h->w_version) return HN_OK;
    h->image_tried_version = h->w_version;
    int first = 0;   // the fused update kernel already saw which fractions occur: go straight to the scale that fits
    if (h->frac_version == h->w_version) first = (h->frac_flags & 4) ? 3 : (h->frac_flags & 2) ? 2 : (h->frac_flags & 1) ? 1 : 0;
    return build_integer_structure(h, true, 1.0, h->W, first, h->maxabs);
}

// float32 shadow of the column matrix (W if symmetric else W^T), rebuilt after any weight change
int ensure_w32(hn_handle h) {
    if (h->w32_valid) return HN_OK;
    const int64_t total = (int64_t)h->N * h->ldw;
    if (!h->W32) HN_CUDA_TRY(cudaMalloc(&h->W32, sizeof(float) * (size_t)total));
    to_f32_kernel<<<grid_for(total, 256, h->sms), 256, 0, h->st>>>(h->w_symmetric ? h->W : h->WT, h->W32, total);
    HN_CUDA_TRY(cudaGetLastError());
    h->w32_valid = true;
    return HN_OK;
}

// point p at the column stream selected by the configuration and set the matching per-flip margin growth
bool integer_stream(hn_handle h);
int select_columns(hn_handle h, RelaxParams& p) {
    const double dscale = h->cfg.domain == HN_DOMAIN_BINARY ? 1.0 : 2.0;
    p.tau_base = h->tau_base;
    p.tau_flip = h->tau_flip;
    p.exact_arith = 0;
    p.key_scale = h->int_ok ? h->key_scale : 0.0;
    if (h->exact_arith && h->cfg.column_stream != HN_COLUMNS_F32) {
        // no margin: "inside" means the field is an exact zero, decided as h <= 0 -> low without re-evaluation (on the
        // integer stream too: the sweeps of a learning epoch meet thousands of exact zeros)
        p.exact_arith = 1;
        if (!integer_stream(h)) p.tau_base = p.tau_flip = 0.0;
    }
    if (h->cfg.column_stream == HN_COLUMNS_F32) {
        HN_TRY(ensure_w32(h));
        p.Wcol = h->W32;
        // |fl32(w) - w| <= 2^-24 |w| (+ a denormal floor): each flip perturbs a field by at most delta*max|W|*2^-24
        p.tau_flip += 2.0 * dscale * (h->maxabs * 5.9604644775390625e-08 + 1.5e-45);
    } else if (integer_stream(h)) {
        p.Wcol = h->K16;
        p.int_scale = h->int_scale;
    } else {
        p.Wcol = h->w_symmetric ? (const void*)h->W : (const void*)h->WT;
    }
    return HN_OK;
}
bool wants_integer_stream(hn_handle h) { return h->cfg.column_stream == HN_COLUMNS_I16 || h->cfg.column_stream == HN_COLUMNS_AUTO; }
bool integer_stream(hn_handle h) { return wants_integer_stream(h) && h->int_ok; }

template <int UPT, typename CT, int THREADS, int MINB>
int launch_relax_t(hn_handle h, const RelaxParams& p, int grid_cap) {
    auto kern = relax_kernel<UPT, CT, THREADS, MINB>;
    const size_t smem = relax_smem_bytes(h->N, h->ldp, UPT);
    HN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    HN_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) occ = 1;
    int grid = p.serial_stream ? 1 : std::min(p.B, h->sms * occ);
    if (grid_cap > 0) grid = std::min(grid, grid_cap);
    if (grid < 1) grid = 1;
    kern<<<grid, THREADS, smem, h->st>>>(p);
    HN_CUDA_TRY(cudaGetLastError());
    return HN_OK;
}
template <typename CT>
int launch_relax_c(hn_handle h, const RelaxParams& p) {
    const int n = h->N;
    // few, fat threads per probe (what every warp repeats per flip dominates): one warp up to N = 1024 (no block barrier
    // worth the name), 32 float64 fields per thread: 2.5x at N=1024, 3.8x at N=256, +33% at N=2048 over 256 threads
    if (n <= 256) return launch_relax_t<8, CT, 32, 32>(h, p, 0);
    if (n <= 512) return launch_relax_t<16, CT, 32, 24>(h, p, 0);
    if (n <= 1024) return launch_relax_t<32, CT, 32, 16>(h, p, 0);
    if (n <= 2048) return launch_relax_t<32, CT, 64, 8>(h, p, 0);
    if (n <= 4096) return launch_relax_t<32, CT, 128, 4>(h, p, 0);
    if (n <= 8192) return launch_relax_t<32, CT, 256, 2>(h, p, 0);
    if (n <= 16384) return launch_relax_t<32, CT, 512, 1>(h, p, 0);
    set_error("relaxation kernel supports dimension <= 16384 (got %d)", n);
    return HN_E_INVALID;
}
// Integer stream: the per-flip work of a warp is dominated by what every warp repeats (scan, barriers, flag, addresses),
// so probes get FEW, FAT threads: 32 units per thread (32 int32 fields + 16 registers of column in flight: 56 registers),
// 9 CTAs of 128 threads per SM at N = 4096 (265k probes/s; 7 CTAs 236k, 8 CTAs 256k, 10 CTAs spill: 180k).
int launch_relax_int(hn_handle h, const RelaxParams& p) {
    const int n = h->N;
    if (n <= 256) return launch_relax_t<8, int16_t, 32, 32>(h, p, 0);
    if (n <= 512) return launch_relax_t<16, int16_t, 32, 32>(h, p, 0);
    if (n <= 1024) return launch_relax_t<32, int16_t, 32, 32>(h, p, 0);
    if (n <= 2048) return launch_relax_t<32, int16_t, 64, 18>(h, p, 0);
    if (n <= 4096) return launch_relax_t<32, int16_t, 128, 9>(h, p, 0);   // 64 units x 64 threads measured slower: 185-198 ms vs 170 ms at config 3 (6-10 CTAs/SM)
    if (n <= 8192) return launch_relax_t<32, int16_t, 256, 4>(h, p, 0);
    if (n <= 16384) return launch_relax_t<32, int16_t, 512, 2>(h, p, 0);
    set_error("relaxation kernel supports dimension <= 16384 (got %d)", n);
    return HN_E_INVALID;
}
int launch_relax(hn_handle h, RelaxParams& p) {
    HN_TRY(select_columns(h, p));
    if (integer_stream(h)) return launch_relax_int(h, p);
    return h->cfg.column_stream == HN_COLUMNS_F32 ? launch_relax_c<float>(h, p) : launch_relax_c<double>(h, p);
}

// upload permutations [rows][N] (host, dense) into one device buffer:
// first half = padded rows, second half = their inverses (position of each unit)
int upload_pool(hn_handle h, const uint16_t* pool, int rows, DevBuf& buf, uint16_t** d_perm, uint16_t** d_inv) {
    const size_t bytes = sizeof(uint16_t) * (size_t)rows * h->ldp;
    HN_TRY(buf.reserve(2 * bytes));
    *d_perm = buf.as<uint16_t>();
    *d_inv = buf.as<uint16_t>() + (size_t)rows * h->ldp;
    HN_CUDA_TRY(cudaMemsetAsync(buf.p, 0, 2 * bytes, h->st));
    HN_CUDA_TRY(cudaMemcpy2DAsync(*d_perm, sizeof(uint16_t) * h->ldp, pool, sizeof(uint16_t) * h->N,
                                  sizeof(uint16_t) * h->N, rows, cudaMemcpyHostToDevice, h->st));
    dim3 g(rows, std::max(1, std::min(64, (h->N + 255) / 256)));
    invert_perm_kernel<<<g, 256, 0, h->st>>>(*d_perm, *d_inv, h->N, h->ldp, rows);
    // validated on the device (the host loop over rows x N entries costs more than a learning epoch at N = 16384)
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    int32_t* bad = &h->scratch()->asym_flag;   // free here: detect_symmetry uses it only inside its own call
    HN_CUDA_TRY(cudaMemsetAsync(bad, 0, 4, h->st));
    check_perm_kernel<<<g, 256, 0, h->st>>>(*d_perm, *d_inv, h->N, h->ldp, rows, bad);
    HN_CUDA_TRY(cudaGetLastError());
    int32_t b = 0;
    HN_CUDA_TRY(cudaMemcpyAsync(&b, bad, 4, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    if (b) { set_error("perm_pool row %d is not a permutation of 0..%d", b - 1, h->N - 1); return HN_E_INVALID; }
    return HN_OK;
}

// The schedule of a relaxation call: validated and uploaded (with its inverse) unless it is the one already resident —
// callers relax batch after batch on the same pool of unit orders, and everything derived from it (the dense sweeps'
// row-permuted images, keyed by pool_version) then stays valid too.
int validate_pool_host(hn_handle h, const uint16_t* pool, int rows);
int stage_pool(hn_handle h, const uint16_t* pool, int rows) {
    const size_t count = (size_t)rows * h->N;
    if (h->d_pool && h->pool_host_rows == rows && h->pool_host.size() == count &&
        std::memcmp(h->pool_host.data(), pool, sizeof(uint16_t) * count) == 0)
        return HN_OK;
    h->pool_host_rows = 0;
    HN_TRY(validate_pool_host(h, pool, rows));
    HN_TRY(upload_pool(h, pool, rows, h->pool, &h->d_pool, &h->d_inv));
    h->pool_version++;
    h->pool_host.assign(pool, pool + count);
    h->pool_host_rows = rows;
    return HN_OK;
}

int validate_pool_host(hn_handle h, const uint16_t* pool, int rows) {
    // every row must be a permutation of 0..N-1 (a bad row would corrupt device memory); small pools are checked here,
    // before any device work, every pool again on the device when it is uploaded (upload_pool)
    if ((int64_t)rows * h->N > (1 << 20)) return HN_OK;
    std::vector<uint8_t> seen((size_t)h->N);
    for (int r = 0; r < rows; r++) {
        std::fill(seen.begin(), seen.end(), 0);
        const uint16_t* row = pool + (size_t)r * h->N;
        for (int i = 0; i < h->N; i++) {
            if (row[i] >= h->N || seen[row[i]]) { set_error("perm_pool row %d is not a permutation of 0..%d", r, h->N - 1); return HN_E_INVALID; }
            seen[row[i]] = 1;
        }
    }
    return HN_OK;
}

// G = S * M^T, exact int32 (the integer stream's initial fields; W symmetric): tcgen05 / TMEM / TMA kernel on the byte
// image, the mma.sync kernel where the image needs two byte planes
int integer_fields(hn_handle h, const uint64_t* d_states, int count, int32_t* d_G, int32_t* scan_unstable = nullptr,
                   int32_t* scan_zero = nullptr) {
    const bool bipolar = h->cfg.domain == HN_DOMAIN_BIPOLAR;
    if (h->use_tc05 && h->k8_valid)
        return launch_igemm_tc05(d_states, h->words, count, h->K8, h->ldw, h->N, bipolar, d_G, h->ldw, h->sms, h->st, scan_unstable, scan_zero);
    return launch_igemm_fields(d_states, h->words, count, h->K16, h->ldw, h->N, bipolar, d_G, h->ldw, h->int_maxabs <= 127.0, h->st);
}

// Unique-attractor table over `count` final states on the device (datacollector/UniqueRelaxedStateHandler.go:41-73).
// Level 1 groups equal sign-canonical states (exact); where the exact-profile key exists, level 2 groups the level-1
// representatives by that key and every such merge is verified on the float energy profiles in the reference's
// summation order; a refuted merge salts the key and the level is redone (never seen: it takes a 128-bit collision or
// integer profiles that coincide where the rounded float profiles do not).  weight == nullptr: every state counts once;
// else state j stands for weight[j] probes (merge of per-GPU tables).  Results: h->attractor, h->ufirst, h->uhits
// (first-seen order = ascending position) and Scratch::unique_total.
int run_dedup(hn_handle h, const uint64_t* finals, uint64_t* keys, const uint8_t* zf, int count, bool profile_keyed,
              const int32_t* weight, int* launches) {
    const int words = h->words, N = h->N;
    Scratch* sc = h->scratch();
    int cap = 1024;
    while (cap < 2 * count) cap <<= 1;
    HN_TRY(h->table.reserve(sizeof(int32_t) * (size_t)cap));
    HN_TRY(h->rep.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->rep1.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->first.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->hits.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->flag.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->rank.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->attractor.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->ufirst.reserve(sizeof(int32_t) * (size_t)count));
    HN_TRY(h->uhits.reserve(sizeof(int32_t) * (size_t)count));
    const int tb = 256, gb = (count + tb - 1) / tb;
    HN_CUDA_TRY(cudaMemsetAsync(h->table.p, 0xFF, sizeof(int32_t) * (size_t)cap, h->st));
    dedup_insert_kernel<<<gb, tb, 0, h->st>>>(finals, keys, count, words, N, h->cfg.domain, h->table.as<int32_t>(), cap - 1,
                                             h->rep1.as<int32_t>(), zf);
    HN_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    const int32_t* rep2 = nullptr;
    if (profile_keyed) {
        HN_TRY(h->rep2.reserve(sizeof(int32_t) * (size_t)count));
        HN_TRY(h->pairs.reserve(sizeof(int2) * (size_t)count));
        for (int round = 0;; round++) {
            if (round == 8) { set_error("unique table: profile keys still collide after 8 salted passes"); return HN_E_STATE; }
            HN_CUDA_TRY(cudaMemsetAsync(h->table.p, 0xFF, sizeof(int32_t) * (size_t)cap, h->st));
            HN_CUDA_TRY(cudaMemsetAsync(sc->pair_count, 0, sizeof(sc->pair_count), h->st));
            dedup_insert_profile_kernel<<<gb, tb, 0, h->st>>>(keys, h->rep1.as<int32_t>(), count, h->table.as<int32_t>(), cap - 1,
                                                             h->rep2.as<int32_t>(), h->pairs.as<int2>(), &sc->pair_count[0]);
            HN_CUDA_TRY(cudaGetLastError());
            if (launches) ++*launches;
            int32_t pc[2] = {0, 0};
            HN_CUDA_TRY(cudaMemcpyAsync(pc, sc->pair_count, sizeof(pc), cudaMemcpyDeviceToHost, h->st));
            HN_CUDA_TRY(cudaStreamSynchronize(h->st));
            if (pc[0] == 0) break;   // the usual case: no two different states share a profile key
            dedup_verify_kernel<<<pc[0], 256, 0, h->st>>>(h->pairs.as<int2>(), finals, words, h->W, h->ldw, N, h->cfg.domain, keys,
                                                         0xC0FFEE0000ULL + (uint64_t)round, &sc->pair_count[1]);
            HN_CUDA_TRY(cudaGetLastError());
            if (launches) ++*launches;
            HN_CUDA_TRY(cudaMemcpyAsync(pc, sc->pair_count, sizeof(pc), cudaMemcpyDeviceToHost, h->st));
            HN_CUDA_TRY(cudaStreamSynchronize(h->st));
            h->verified_pairs += pc[0]; h->refuted_pairs += pc[1];
            if (pc[1] == 0) break;   // every merge of different states confirmed on the float profiles
        }
        rep2 = h->rep2.as<int32_t>();
    }
    HN_CUDA_TRY(cudaMemsetAsync(h->first.p, 0x7F, sizeof(int32_t) * (size_t)count, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(h->hits.p, 0, sizeof(int32_t) * (size_t)count, h->st));
    dedup_compose_kernel<<<gb, tb, 0, h->st>>>(h->rep1.as<int32_t>(), rep2, count, h->rep.as<int32_t>());
    dedup_count_kernel<<<gb, tb, 0, h->st>>>(h->rep.as<int32_t>(), count, h->first.as<int32_t>(), h->hits.as<int32_t>(), weight);
    dedup_flag_kernel<<<gb, tb, 0, h->st>>>(h->rep.as<int32_t>(), h->first.as<int32_t>(), count, h->flag.as<int32_t>());
    scan_flags_kernel<<<1, 1024, 0, h->st>>>(h->flag.as<int32_t>(), count, h->rank.as<int32_t>(), &sc->unique_total);
    dedup_finalize_kernel<<<gb, tb, 0, h->st>>>(h->rep.as<int32_t>(), h->first.as<int32_t>(), h->hits.as<int32_t>(),
                                               h->rank.as<int32_t>(), count, h->attractor.as<int32_t>(),
                                               h->ufirst.as<int32_t>(), h->uhits.as<int32_t>());
    HN_CUDA_TRY(cudaGetLastError());
    if (launches) *launches += 5;
    return HN_OK;
}

// ---- dense lockstep sweeps (dense.cuh) ------------------------------------------------------------------------------
bool dense_applicable(hn_handle h, bool serial, bool hist, const int32_t* d_offsets, int n_pool) {
    return h->use_dense && h->use_tc05 && integer_stream(h) && h->w_symmetric && h->k8_valid && !serial && !hist && !d_offsets &&
           h->N >= h->dense_min_n && h->N <= 4096 && n_pool >= 1 && h->cfg.max_relax_iterations >= 1;
}
// images of the first `rows` schedule rows: weight rows in visiting order, in-block matrices, per-position state offsets
int prepare_dense(hn_handle h, const uint16_t* d_pool, int rows) {
    if (h->dense_rows >= rows && h->dense_w_version == h->w_version && h->dense_pool_version == h->pool_version) return HN_OK;
    const int N = h->N, np = ds_np(N), rp = ds_rows(N);
    HN_TRY(h->d_mperm.reserve((size_t)rows * rp * np));
    HN_TRY(h->d_mbb.reserve((size_t)rows * rp * DS_BLK));
    HN_TRY(h->d_mx.reserve((size_t)rows * rp * DS_BLK));
    HN_TRY(h->d_info.reserve(sizeof(DsPos) * (size_t)rows * rp));
    for (int r = 0; r < rows; r++) {
        const uint16_t* perm = d_pool + (size_t)r * h->ldp;
        dense_prep_rows_kernel<<<rp, 256, 0, h->st>>>(h->K8, h->ldw, perm, N, np, h->d_mperm.as<int8_t>() + (size_t)r * rp * np);
        dense_prep_blocks_kernel<<<(rp * DS_BLK + 255) / 256, 256, 0, h->st>>>(h->K8, h->ldw, perm, N, np, rp,
                                                                             h->d_mbb.as<int8_t>() + (size_t)r * rp * DS_BLK,
                                                                             h->d_mx.as<int8_t>() + (size_t)r * rp * DS_BLK,
                                                                             h->d_info.as<DsPos>() + (size_t)r * rp);
    }
    HN_CUDA_TRY(cudaGetLastError());
    int rc = -1;
    h->dense_map3d = false;
    if (h->dense_want_map3d && h->dense_box_rows == DS_BLK) {   // one 3-D box per ring slot; the 2-D boxes are the fallback
        rc = tc05::make_map_bytes_3d_parts(&h->dense_map, h->d_mperm.p, (uint64_t)np / 4, 4, (uint64_t)rows * rp, (uint64_t)np, DS_BLK);
        h->dense_map3d = rc == 0;
    }
    if (rc != 0) rc = tc05::make_map_bytes_2d(&h->dense_map, h->d_mperm.p, (uint64_t)np, (uint64_t)rows * rp, (uint64_t)np, h->dense_box_rows);
    if (rc != 0) { set_error("cuTensorMapEncodeTiled (dense sweeps) failed (%d)", rc); return HN_E_CUDA; }
    h->dense_rows = rows; h->dense_w_version = h->w_version; h->dense_pool_version = h->pool_version;
    return HN_OK;
}
// Up to dense_max_rows lockstep sweeps over the B probes; leaves the current states in d_state, their exact fields in
// h->H (int32) and the per-probe counters in d_sweeps / d_dflips / d_dmargin for the sparse kernel.
int run_dense_sweeps(hn_handle h, const uint64_t* d_probes, int B, const uint16_t* d_pool, int n_pool, int* launches) {
    const int N = h->N, words = h->words, np = ds_np(N), rp = ds_rows(N);
    const int rows = std::min(std::min(h->dense_max_rows, n_pool), h->cfg.max_relax_iterations);
    HN_TRY(prepare_dense(h, d_pool, rows));
    HN_TRY(h->d_state.reserve(sizeof(uint64_t) * (size_t)B * words));
    HN_TRY(h->d_sweeps.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->d_dflips.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->d_dmargin.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->d_finished.reserve((size_t)B));
    if (!h->zc_probes)
        HN_CUDA_TRY(cudaMemcpyAsync(h->d_state.p, d_probes, sizeof(uint64_t) * (size_t)B * words, cudaMemcpyDeviceToDevice, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(h->d_sweeps.p, 0, sizeof(int32_t) * (size_t)B, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(h->d_dflips.p, 0, sizeof(int32_t) * (size_t)B, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(h->d_dmargin.p, 0, sizeof(int32_t) * (size_t)B, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(h->d_finished.p, 0, (size_t)B, h->st));
    Scratch* sc = h->scratch();
    const bool prof = getenv("HN_DENSE_PROF") != nullptr;
    if (prof) { HN_TRY(h->partial.reserve(64)); HN_CUDA_TRY(cudaMemsetAsync(h->partial.p, 0, 64, h->st)); }
    const size_t smem = ds_smem_bytes(N);
    HN_CUDA_TRY(cudaFuncSetAttribute(dense_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (B + DS_T - 1) / DS_T;
    h->last_dense_sweeps = 0;
    while ((int)h->dense_ev.size() < 4 * rows) {
        cudaEvent_t e;
        HN_CUDA_TRY(cudaEventCreate(&e));
        h->dense_ev.push_back(e);
    }
    HN_TRY(h->d_scan.reserve((size_t)B * 9));
    int32_t* d_unst = h->d_scan.as<int32_t>();
    int32_t* d_zero = d_unst + B;
    uint8_t* d_only = reinterpret_cast<uint8_t*>(d_zero + B);
    int32_t* G = reinterpret_cast<int32_t*>(h->H.as<double>());
    const bool fused = h->use_tc05 && h->k8_valid;   // stability scan inside the contraction's epilogue
    int32_t unf_prev = B;
    bool fields_stored = false;
    for (int r = 0; r < rows; r++) {
        cudaEvent_t* de = &h->dense_ev[4 * (size_t)r];
        HN_CUDA_TRY(cudaEventRecord(de[0], h->st));
        HN_CUDA_TRY(cudaMemsetAsync(&sc->dense_ambiguous, 0, 8, h->st));    // ambiguous, unfinished
        HN_CUDA_TRY(cudaMemsetAsync(&sc->dense_flips, 0, 8, h->st));
        DenseParams dp{};
        dp.N = N; dp.np = np; dp.rows_p = rp; dp.words = words; dp.B = B; dp.domain = h->cfg.domain; dp.ldw = h->ldw;
        dp.box_rows = h->dense_box_rows; dp.map3d = h->dense_map3d ? 1 : 0;
        dp.mbb = h->d_mbb.as<int8_t>() + (size_t)r * rp * DS_BLK;
        dp.mx = h->d_mx.as<int8_t>() + (size_t)r * rp * DS_BLK;
        dp.info = h->d_info.as<DsPos>() + (size_t)r * rp;
        dp.row_base = r * rp;
        dp.states = h->d_state.as<uint64_t>();
        if (r == 0 && h->zc_probes) { dp.states_src = h->zc_probes; dp.states_copy = const_cast<uint64_t*>(d_probes); }
        dp.finished = h->d_finished.as<uint8_t>();
        dp.sweeps_done = h->d_sweeps.as<int32_t>(); dp.flips_done = h->d_dflips.as<int32_t>(); dp.margin_done = h->d_dmargin.as<int32_t>();
        dp.Wrow = h->W; dp.sweep_index = r; dp.flips_total = &sc->dense_flips;
        dp.prof = prof ? h->partial.as<unsigned long long>() : nullptr;
        dense_sweep_kernel<<<std::min(tiles, h->sms), DS_THREADS, smem, h->st>>>(h->dense_map, dp);
        HN_CUDA_TRY(cudaGetLastError());
        HN_CUDA_TRY(cudaEventRecord(de[1], h->st));
        // This sweep's flips decide whether another dense sweep is worth it.  The fields go to HBM (2 GiB at config 3)
        // only after the last one, where the sparse kernel starts from them, or when some probe's stability hangs on
        // exactly-zero fields.
        unsigned long long fl = 0;
        HN_CUDA_TRY(cudaMemcpyAsync(&fl, &sc->dense_flips, 8, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        const bool last = r == rows - 1 || (double)fl < h->dense_min_flips * (double)unf_prev;
        int32_t cnt[2] = {0, 0};   // ambiguous, unfinished
        if (fused) {
            HN_CUDA_TRY(cudaMemsetAsync(d_unst, 0, sizeof(int32_t) * 2 * (size_t)B, h->st));
            HN_TRY(integer_fields(h, h->d_state.as<uint64_t>(), B, last ? G : nullptr, d_unst, d_zero));
            HN_CUDA_TRY(cudaEventRecord(de[2], h->st));
            dense_classify_kernel<<<(B + 255) / 256, 256, 0, h->st>>>(d_unst, d_zero, B, h->cfg.max_relax_unstable_units,
                                                                      h->d_finished.as<uint8_t>(), h->d_dmargin.as<int32_t>(), d_only,
                                                                      &sc->dense_unfinished, &sc->dense_ambiguous);
            HN_CUDA_TRY(cudaGetLastError());
            HN_CUDA_TRY(cudaMemcpyAsync(cnt, &sc->dense_ambiguous, 8, cudaMemcpyDeviceToHost, h->st));
            HN_CUDA_TRY(cudaStreamSynchronize(h->st));
            fields_stored = last;
        } else {
            HN_TRY(integer_fields(h, h->d_state.as<uint64_t>(), B, G));
            HN_CUDA_TRY(cudaEventRecord(de[2], h->st));
            fields_stored = true;
        }
        if (!fused || cnt[0] > 0) {   // float64 re-evaluation of the exactly-zero fields (of every probe on the mma.sync path)
            if (!fields_stored) { HN_TRY(integer_fields(h, h->d_state.as<uint64_t>(), B, G)); fields_stored = true; }
            dense_scan_kernel<<<(B * 32 + 255) / 256, 256, 0, h->st>>>(G, h->ldw, h->d_state.as<uint64_t>(), words, N, B, h->cfg.domain, h->W,
                                                                       h->ldw, h->cfg.max_relax_unstable_units, h->d_finished.as<uint8_t>(),
                                                                       h->d_dmargin.as<int32_t>(), &sc->dense_unfinished,
                                                                       fused ? d_only : nullptr);
            HN_CUDA_TRY(cudaGetLastError());
        }
        HN_CUDA_TRY(cudaEventRecord(de[3], h->st));
        if (launches) *launches += 3;
        h->last_dense_sweeps = r + 1;
        int32_t unf = 0;
        HN_CUDA_TRY(cudaMemcpyAsync(&unf, &sc->dense_unfinished, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        {
            float a = 0, b2 = 0, c2 = 0;
            cudaEventElapsedTime(&a, de[0], de[1]); cudaEventElapsedTime(&b2, de[1], de[2]); cudaEventElapsedTime(&c2, de[2], de[3]);
            h->acc_dense_kernel_ms += a; h->acc_dense_gemm_ms += b2; h->acc_dense_scan_ms += c2;
            h->acc_dense_flips += (int64_t)fl;
        }
        if (prof) {   // clock cycles of CTA 0 by phase, per block of 32 positions (diagnostic)
            unsigned long long pc[8];
            HN_CUDA_TRY(cudaMemcpy(pc, h->partial.p, 64, cudaMemcpyDeviceToHost));
            HN_CUDA_TRY(cudaMemsetAsync(h->partial.p, 0, 64, h->st));
            const double nb = (double)(rp / DS_BLK) * (double)((tiles + h->sms - 1) / h->sms);
            fprintf(stderr, "[dense prof] sweep %d flips/probe %.0f | per block (clk): wait MMA %.0f, tmem->smem %.0f, gather %.0f, decide %.0f, "
                            "end barrier %.0f | MMA warp: wait decisions %.0f, issue+stream %.0f\n", r + 1, (double)fl / std::max(1, B),
                    pc[0] / nb, pc[1] / nb, pc[2] / nb, pc[3] / nb, pc[4] / nb, pc[5] / nb, pc[6] / nb);
        }
        unf_prev = std::max(unf, 1);
        if (last || unf == 0) break;   // the sparse kernel is cheaper from here on
    }
    // every probe froze before the last planned sweep: the hand-over to the sparse kernel still reads the fields
    if (!fields_stored) HN_TRY(integer_fields(h, h->d_state.as<uint64_t>(), B, G));
    return HN_OK;
}

// The whole relaxation of B probes already on the device.
// Relaxation of probes [off, off + cnt) of the batch already on the device (kernels only, stream h->st): initial fields
// (or the dense lockstep sweeps, which supply them) and the sparse kernel.  Results land at the same indices of the
// handle's result arrays.  ce: four events of this range (start, after the dense sweeps, after the fields, at the end).
int relax_range(hn_handle h, const uint64_t* d_probes, int B_total, int off, int cnt, const uint16_t* d_pool, const uint16_t* d_inv,
                int n_pool, const int32_t* d_offsets, uint32_t flags, bool want_dist, cudaEvent_t* ce, int* launches,
                bool* profile_keyed, int* dense_sweeps) {
    const int N = h->N, words = h->words;
    const bool dedup = flags & HN_RELAX_DEDUP, hist = flags & HN_RELAX_HISTORY, serial = flags & HN_RELAX_SERIAL_STREAM;
    (void)B_total;
    Scratch* sc = h->scratch();
    const uint64_t* probes = d_probes + (size_t)off * words;
    HN_CUDA_TRY(cudaMemsetAsync(&sc->work_counter, 0, 4, h->st));
    HN_CUDA_TRY(cudaEventRecord(ce[0], h->st));
    // initial fields: S * W^T (FP64 tensor pipe); exact integer stream: G = S * K^T on the integer tensor pipe when W is
    // symmetric (K's rows are its columns), else S * Wraw^T in float64 (half/quarter-integers: exact)
    const bool int_fields = integer_stream(h) && h->w_symmetric;
    const bool dense = dense_applicable(h, serial, hist, d_offsets, n_pool);
    const uint64_t* relax_probes = probes;
    h->last_dense_sweeps = 0;
    if (dense) {   // the flip-dense first sweeps in lockstep on the tensor path; leaves states, fields and counters
        HN_TRY(run_dense_sweeps(h, probes, cnt, d_pool, n_pool, launches));
        relax_probes = h->d_state.as<uint64_t>();
        HN_CUDA_TRY(cudaEventRecord(ce[1], h->st));
    } else {
        HN_CUDA_TRY(cudaEventRecord(ce[1], h->st));
        if (int_fields)
            HN_TRY(integer_fields(h, probes, cnt, reinterpret_cast<int32_t*>(h->H.as<double>())));
        else
            HN_TRY(launch_fields(h, probes, cnt, h->H.as<double>(), domain_values(h->cfg.domain), integer_stream(h) ? h->Wraw : nullptr));
        ++*launches;
    }
    *dense_sweeps = std::max(*dense_sweeps, h->last_dense_sweeps);
    HN_CUDA_TRY(cudaEventRecord(ce[2], h->st));

    RelaxParams p{};
    p.N = N; p.ldw = h->ldw; p.words = words;
    p.Wrow = h->W;
    p.H0 = h->H.as<double>(); p.ldh = h->ldw;
    p.H0i = int_fields ? reinterpret_cast<const int32_t*>(h->H.as<double>()) : nullptr;
    p.probes = relax_probes; p.perm_pool = d_pool; p.inv_pool = d_inv; p.ldp = h->ldp; p.n_pool = n_pool;
    if (dense) { p.sweeps_done = h->d_sweeps.as<int32_t>(); p.flips_done = h->d_dflips.as<int32_t>(); p.margin_done = h->d_dmargin.as<int32_t>(); }
    p.offsets = (serial || !d_offsets) ? nullptr : d_offsets + off;
    p.B = cnt; p.max_iter = h->cfg.max_relax_iterations; p.max_unstable = h->cfg.max_relax_unstable_units;
    p.domain = h->cfg.domain; p.check_stability = 1; p.serial_stream = serial ? 1 : 0;
    p.final_states = h->finals.as<uint64_t>() + (size_t)off * words; p.num_steps = h->steps.as<int32_t>() + off;
    p.stable = h->stable.as<uint8_t>() + off; p.flips = h->flips.as<int32_t>() + off; p.margin_hits = h->margin.as<int32_t>() + off;
    p.distances = (want_dist && h->P > 0) ? (h->zc_dist ? h->zc_dist : h->dist.as<int32_t>()) + (size_t)off * h->P : nullptr;
    p.final_states_host = h->zc_finals ? h->zc_finals + (size_t)off * words : nullptr;
    p.targets = h->targets.as<uint64_t>(); p.P = h->P;
    p.hash = dedup ? h->hash.as<uint64_t>() + (size_t)off * 4 : nullptr;
    p.zero_field = dedup ? h->zf.as<uint8_t>() + off : nullptr;
    p.history = hist ? h->history.as<uint64_t>() + (size_t)off * (h->cfg.max_relax_iterations + 1) * words : nullptr;
    p.work_counter = &sc->work_counter; p.status = sc->status;
    HN_TRY(launch_relax(h, p));
    ++*launches;
    *profile_keyed = p.exact_arith || integer_stream(h) || p.key_scale != 0.0;   // see the epilogue of relax_kernel
    HN_CUDA_TRY(cudaEventRecord(ce[3], h->st));
    return HN_OK;
}

// probes per range of a chunked batch: the ranges run back to back on the compute stream while the copy stream moves
// the next range's probes in and the previous range's results out (hn_relax_batch with host buffers), and the fields /
// dense-sweep working set is that of one range.  Measured at config 3 (2^17 probes per GPU): every extra range costs
// ~2 ms of kernel tails and host round trips; the copies it hides take 7 ms per step on one GPU but 27 ms when eight
// GPUs share the host's PCIe fabric (12.8 GB/s per GPU).  So: two ranges per config-3 batch when the results are
// streamed to the host, one range (up to 2^17 probes) when they stay on the device.
constexpr int RELAX_CHUNK_STREAMED = 1 << 16, RELAX_CHUNK_RESIDENT = 1 << 17;

// Device-visible address of a caller's host array if all of it is page-locked (cudaHostAlloc / cudaHostRegister: mapped
// into the unified address space), else nullptr.
static void* mapped_host(const void* ptr, size_t bytes) {
    if (!ptr || !bytes) return nullptr;
    cudaPointerAttributes a{}, b{};
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    if (cudaPointerGetAttributes(&b, static_cast<const char*>(ptr) + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (b.type != cudaMemoryTypeHost) return nullptr;
    return a.devicePointer;
}

// The whole relaxation of B probes.  d_probes: device buffer of the batch; if host_probes != nullptr the probes are
// still on the host and are uploaded range by range on the copy stream.  stream_out != nullptr: the per-probe results
// are copied to its host buffers range by range on the copy stream, overlapping the next range's kernels.
int relax_device(hn_handle h, uint64_t* d_probes, int B, const uint16_t* d_pool, const uint16_t* d_inv,
                 int n_pool, const int32_t* d_offsets, uint32_t flags, bool want_energies, bool want_dist,
                 const uint64_t* host_probes = nullptr, const hn_relax_out* stream_out = nullptr) {
    const int N = h->N, words = h->words;
    HN_TRY(ensure_margin(h));
    HN_TRY(refresh_wt(h));
    const bool dedup = flags & HN_RELAX_DEDUP, hist = flags & HN_RELAX_HISTORY, serial = flags & HN_RELAX_SERIAL_STREAM;
    const bool zero_matrix = h->maxabs == 0.0 && h->cfg.max_relax_iterations >= 1;   // closed form, zero_matrix_relax_kernel
    // Zero-copy results: when the caller's final-state and distance arrays are page-locked, the relaxation kernel writes
    // every probe's rows straight into them as the probe finishes (relax.cuh, epilogue), so the two large device -> host
    // transfers (282 MB per 2^17 probes at config 3) ride under the kernel itself instead of following it, and the
    // batch needs no splitting into ranges.
    h->zc_finals = nullptr; h->zc_dist = nullptr;
    bool zc = false;
    if (stream_out && h->zero_copy && !hist && !serial && !zero_matrix) {
        const bool dist_out = want_dist && h->P > 0 && stream_out->distances;
        uint64_t* zf = static_cast<uint64_t*>(mapped_host(stream_out->final_states, sizeof(uint64_t) * (size_t)B * words));
        int32_t* zd = dist_out ? static_cast<int32_t*>(mapped_host(stream_out->distances, sizeof(int32_t) * (size_t)B * h->P)) : nullptr;
        zc = (stream_out->final_states || dist_out) && (zf || !stream_out->final_states) && (zd || !dist_out);
        if (zc) { h->zc_finals = zf; h->zc_dist = zd; }
    }
    const uint64_t* zc_probes = (host_probes && h->zero_copy)
        ? static_cast<const uint64_t*>(mapped_host(host_probes, sizeof(uint64_t) * (size_t)B * words)) : nullptr;
    const int range = ((host_probes || stream_out) && !zc) ? RELAX_CHUNK_STREAMED : RELAX_CHUNK_RESIDENT;
    const int chunk = (serial || zero_matrix || B <= range + range / 4) ? B : range;
    const int n_chunks = (B + chunk - 1) / chunk;
    HN_TRY(h->H.reserve(sizeof(double) * (size_t)chunk * h->ldw));
    HN_TRY(h->finals.reserve(sizeof(uint64_t) * (size_t)B * words));
    HN_TRY(h->steps.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->stable.reserve((size_t)B));
    HN_TRY(h->flips.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->margin.reserve(sizeof(int32_t) * (size_t)B));
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    if (want_dist && h->P > 0 && !h->zc_dist) HN_TRY(h->dist.reserve(sizeof(int32_t) * (size_t)B * h->P));
    if (dedup) HN_TRY(h->hash.reserve(sizeof(uint64_t) * 4 * (size_t)B));
    if (dedup) HN_TRY(h->zf.reserve((size_t)B));
    if (hist) HN_TRY(h->history.reserve(sizeof(uint64_t) * (size_t)B * (h->cfg.max_relax_iterations + 1) * words));
    while ((int)h->chunk_ev.size() < 6 * n_chunks) {
        cudaEvent_t e;
        HN_CUDA_TRY(cudaEventCreate(&e));
        h->chunk_ev.push_back(e);
    }

    int launches = 0, dense_sweeps = 0;
    h->acc_dense_kernel_ms = h->acc_dense_gemm_ms = h->acc_dense_scan_ms = 0; h->acc_dense_flips = 0;
    Scratch* sc = h->scratch();
    HN_CUDA_TRY(cudaMemsetAsync(sc, 0, offsetof(Scratch, asym_flag), h->st));   // work counter, status, unique total
    if (dedup) HN_CUDA_TRY(cudaMemsetAsync(h->hash.p, 0, sizeof(uint64_t) * 4 * (size_t)B, h->st));
    if (zero_matrix && (serial ? n_pool < B : n_pool < 1)) {
        set_error("permutation pool exhausted: %d rows supplied, at least %d needed", n_pool, serial ? B : 1);
        return HN_E_POOL;
    }
    bool profile_keyed = false;
    HN_CUDA_TRY(cudaEventRecord(h->ev[0], h->st));   // the copy stream starts after everything queued so far
    HN_CUDA_TRY(cudaStreamWaitEvent(h->st2, h->ev[0], 0));
    auto d2h2 = [&](void* dst, const void* src, size_t bytes) -> int {
        if (dst && bytes) HN_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st2));
        return HN_OK;
    };
    for (int c = 0; c < n_chunks; c++) {
        const int off = c * chunk, cnt = std::min(chunk, B - off);
        cudaEvent_t* ce = &h->chunk_ev[6 * (size_t)c];
        // a range that starts with dense lockstep sweeps reads page-locked probes where they lie (dense.cuh, tile load of
        // the first sweep) instead of waiting for their transfer
        h->zc_probes = nullptr;
        if (host_probes && zc_probes && !zero_matrix && dense_applicable(h, serial, hist, d_offsets, n_pool))
            h->zc_probes = zc_probes + (size_t)off * words;
        if (host_probes && !h->zc_probes) {   // this range's probes: host -> device on the copy stream
            HN_CUDA_TRY(cudaMemcpyAsync(d_probes + (size_t)off * words, host_probes + (size_t)off * words,
                                        sizeof(uint64_t) * (size_t)cnt * words, cudaMemcpyHostToDevice, h->st2));
            HN_CUDA_TRY(cudaEventRecord(ce[4], h->st2));
            HN_CUDA_TRY(cudaStreamWaitEvent(h->st, ce[4], 0));
        }
        if (zero_matrix) {
            HN_CUDA_TRY(cudaEventRecord(ce[0], h->st));
            HN_CUDA_TRY(cudaEventRecord(ce[1], h->st));
            HN_CUDA_TRY(cudaEventRecord(ce[2], h->st));
            const int wpb = 8;
            zero_matrix_relax_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, h->st>>>(
                d_probes, B, words, N, h->cfg.domain, h->cfg.max_relax_iterations, h->finals.as<uint64_t>(), h->steps.as<int32_t>(),
                h->stable.as<uint8_t>(), h->flips.as<int32_t>(), h->margin.as<int32_t>(),
                (want_dist && h->P > 0) ? h->dist.as<int32_t>() : nullptr, h->targets.as<uint64_t>(), h->P,
                dedup ? h->hash.as<uint64_t>() : nullptr, dedup ? h->zf.as<uint8_t>() : nullptr, hist ? h->history.as<uint64_t>() : nullptr);
            HN_CUDA_TRY(cudaGetLastError());
            launches++;
            profile_keyed = false;  // every final state is the all-low state: one group at level 1
            HN_CUDA_TRY(cudaEventRecord(ce[3], h->st));
        } else {
            HN_TRY(relax_range(h, d_probes, B, off, cnt, d_pool, d_inv, n_pool, d_offsets, flags, want_dist, ce, &launches,
                               &profile_keyed, &dense_sweeps));
        }
        if (stream_out) {   // this range's results: device -> host on the copy stream, under the next range's kernels
            HN_CUDA_TRY(cudaStreamWaitEvent(h->st2, ce[3], 0));
            if (!h->zc_finals)
                HN_TRY(d2h2(stream_out->final_states ? stream_out->final_states + (size_t)off * words : nullptr,
                            h->finals.as<uint64_t>() + (size_t)off * words, sizeof(uint64_t) * (size_t)cnt * words));
            HN_TRY(d2h2(stream_out->num_steps ? stream_out->num_steps + off : nullptr, h->steps.as<int32_t>() + off, sizeof(int32_t) * (size_t)cnt));
            HN_TRY(d2h2(stream_out->stable ? stream_out->stable + off : nullptr, h->stable.as<uint8_t>() + off, (size_t)cnt));
            HN_TRY(d2h2(stream_out->flips ? stream_out->flips + off : nullptr, h->flips.as<int32_t>() + off, sizeof(int32_t) * (size_t)cnt));
            HN_TRY(d2h2(stream_out->margin_hits ? stream_out->margin_hits + off : nullptr, h->margin.as<int32_t>() + off, sizeof(int32_t) * (size_t)cnt));
            if (stream_out->distances && want_dist && h->P > 0 && !h->zc_dist)
                HN_TRY(d2h2(stream_out->distances + (size_t)off * h->P, h->dist.as<int32_t>() + (size_t)off * h->P, sizeof(int32_t) * (size_t)cnt * h->P));
            if (stream_out->history && hist) {
                const size_t per = (size_t)(h->cfg.max_relax_iterations + 1) * words;
                HN_TRY(d2h2(stream_out->history + (size_t)off * per, h->history.as<uint64_t>() + (size_t)off * per, sizeof(uint64_t) * (size_t)cnt * per));
            }
        }
    }
    HN_CUDA_TRY(cudaEventRecord(h->ev[2], h->st));

    if (want_energies) {
        // EnergyProfile of the final states (main.go:234): fresh fields by GEMM + epilogue, range by range
        HN_TRY(h->energies.reserve(sizeof(double) * (size_t)B * N));
        for (int c = 0; c < n_chunks; c++) {
            const int off = c * chunk, cnt = std::min(chunk, B - off);
            const uint64_t* fin = h->finals.as<uint64_t>() + (size_t)off * words;
            HN_TRY(launch_fields(h, fin, cnt, h->H.as<double>(), domain_values(h->cfg.domain)));
            energy_epilogue_kernel<<<cnt, RED_THREADS, 0, h->st>>>(h->H.as<double>(), h->ldw, fin, words,
                                                                  h->W, h->ldw, N, h->cfg.domain, h->exact_arith ? 0.0 : h->tau_base,
                                                                  h->cfg.max_relax_unstable_units, h->energies.as<double>() + (size_t)off * N,
                                                                  nullptr, nullptr, nullptr, nullptr);
            HN_CUDA_TRY(cudaGetLastError());
            launches += 2;
        }
    }
    h->result_unique = 0;
    if (dedup) {
        HN_TRY(run_dedup(h, h->finals.as<uint64_t>(), h->hash.as<uint64_t>(), h->zf.as<uint8_t>(), B, profile_keyed, nullptr, &launches));
    }
    HN_CUDA_TRY(cudaEventRecord(h->ev[3], h->st));
    HN_CUDA_TRY(cudaMemsetAsync(sc->totals, 0, sizeof(sc->totals), h->st));
    sum_counters_kernel<<<std::min((B + 255) / 256, h->sms * 4), 256, 0, h->st>>>(h->flips.as<int32_t>(), h->steps.as<int32_t>(),
                                                                                h->margin.as<int32_t>(), B, sc->totals);
    HN_CUDA_TRY(cudaGetLastError());
    Scratch host_sc;
    HN_CUDA_TRY(cudaMemcpyAsync(&host_sc, sc, sizeof(Scratch), cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st2));
    h->tot_flips = (int64_t)host_sc.totals[0]; h->tot_sweeps = (int64_t)host_sc.totals[1]; h->tot_margin = (int64_t)host_sc.totals[2];
    if (host_sc.status[0]) {
        set_error("permutation pool exhausted: %d rows supplied, at least %d needed", n_pool, host_sc.status[1]);
        return HN_E_POOL;
    }
    h->result_unique = dedup ? host_sc.unique_total : 0;
    h->result_profile_keyed = profile_keyed;
    double tf = 0, tr = 0, td = 0;
    for (int c = 0; c < n_chunks; c++) {
        cudaEvent_t* ce = &h->chunk_ev[6 * (size_t)c];
        float a = 0, b = 0, d = 0;
        cudaEventElapsedTime(&d, ce[0], ce[1]);
        cudaEventElapsedTime(&a, ce[1], ce[2]);
        cudaEventElapsedTime(&b, ce[2], ce[3]);
        td += d; tf += a; tr += b;
    }
    float q = 0;
    cudaEventElapsedTime(&q, h->ev[2], h->ev[3]);
    h->timing.fields_ms = tf; h->timing.relax_ms = tr; h->timing.post_ms = q; h->timing.launches = launches;
    h->timing.dense_ms = td; h->timing.dense_sweeps = dense_sweeps;
    h->timing.dense_kernel_ms = h->acc_dense_kernel_ms; h->timing.dense_gemm_ms = h->acc_dense_gemm_ms;
    h->timing.dense_scan_ms = h->acc_dense_scan_ms; h->timing.dense_flips = h->acc_dense_flips;
    // (distances written straight to the caller's memory are not on the device: a later download finds none)
    h->result_B = B; h->result_P = (want_dist && !h->zc_dist) ? h->P : 0; h->result_flags = flags | (want_energies ? 0x100u : 0u);
    h->zc_finals = nullptr; h->zc_dist = nullptr; h->zc_probes = nullptr;
    return HN_OK;
}

// streamed: the per-probe arrays (final states, steps, flags, counters, distances, history) were already copied range by
// range by relax_device; only what exists after the whole batch is left
int download_results(hn_handle h, hn_relax_out* out, bool streamed = false) {
    const int B = h->result_B, N = h->N, words = h->words;
    if (B <= 0) { set_error("no relaxation results on the device"); return HN_E_STATE; }
    auto d2h = [&](void* dst, const void* src, size_t bytes) -> int {
        if (dst && bytes) HN_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->st));
        return HN_OK;
    };
    if (!streamed) {
        HN_TRY(d2h(out->final_states, h->finals.p, sizeof(uint64_t) * (size_t)B * words));
        HN_TRY(d2h(out->num_steps, h->steps.p, sizeof(int32_t) * (size_t)B));
        HN_TRY(d2h(out->stable, h->stable.p, (size_t)B));
        HN_TRY(d2h(out->flips, h->flips.p, sizeof(int32_t) * (size_t)B));
        HN_TRY(d2h(out->margin_hits, h->margin.p, sizeof(int32_t) * (size_t)B));
    }
    if (out->distances && !streamed) {
        if (h->result_P != h->P) { set_error("distances were not computed"); return HN_E_STATE; }
        HN_TRY(d2h(out->distances, h->dist.p, sizeof(int32_t) * (size_t)B * h->P));
    }
    if (out->energies) {
        if (!(h->result_flags & 0x100u)) { set_error("energies were not computed"); return HN_E_STATE; }
        HN_TRY(d2h(out->energies, h->energies.p, sizeof(double) * (size_t)B * N));
    }
    out->n_unique = 0;
    if (h->result_flags & HN_RELAX_DEDUP) {
        out->n_unique = h->result_unique;
        HN_TRY(d2h(out->attractor_id, h->attractor.p, sizeof(int32_t) * (size_t)B));
        const int nu = std::min(h->result_unique, out->unique_cap);
        if (nu > 0) {
            HN_TRY(d2h(out->unique_first, h->ufirst.p, sizeof(int32_t) * (size_t)nu));
            HN_TRY(d2h(out->unique_hits, h->uhits.p, sizeof(int32_t) * (size_t)nu));
        }
    }
    if (out->state_hash) {
        if (!(h->result_flags & HN_RELAX_DEDUP)) { set_error("state hashes were not computed (HN_RELAX_DEDUP)"); return HN_E_STATE; }
        // the two words the groups were finally formed on: the exact-profile key where it exists, else the state hash
        const char* src = (const char*)h->hash.p + (h->result_profile_keyed ? 16 : 0);
        HN_CUDA_TRY(cudaMemcpy2DAsync(out->state_hash, 16, src, 32, 16, (size_t)B, cudaMemcpyDeviceToHost, h->st));
    }
    if (out->history) {
        if (!(h->result_flags & HN_RELAX_HISTORY)) { set_error("history was not captured"); return HN_E_STATE; }
        if (!streamed) HN_TRY(d2h(out->history, h->history.p, sizeof(uint64_t) * (size_t)B * (h->cfg.max_relax_iterations + 1) * words));
    }
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    out->total_flips = h->tot_flips; out->total_sweeps = h->tot_sweeps; out->total_margin_hits = h->tot_margin;   // summed on the device
    out->device_ms = h->timing.dense_ms + h->timing.fields_ms + h->timing.relax_ms + h->timing.post_ms;
    return HN_OK;
}

// dW = A^T-style contraction over the n states, written to h->dW
int contraction_hebbian(hn_handle h, const uint64_t* d_states, int n, ValueMap vm) {
    const int pwords = (n + 63) / 64;
    HN_TRY(h->bitsAT.reserve(sizeof(uint64_t) * (size_t)h->N * pwords));
    transpose_bits_kernel<<<grid_for((int64_t)h->N * pwords, 256, h->sms), 256, 0, h->st>>>(
        d_states, h->bitsAT.as<uint64_t>(), h->N, h->words, n, pwords);
    HN_CUDA_TRY(cudaGetLastError());
    OpBits a{h->bitsAT.as<uint64_t>(), pwords, h->N, n, vm.lo, vm.hi};
    EpiStore e{h->dW, h->ldw, h->N, h->N};
    return launch_gemm_f64(a, a, e, h->N, h->N, n, h->st);
}

struct OpBitsDiffG {  // scale * kscale[k] * (val_a(bitA) - val_b(bitB)); kscale == nullptr -> 1
    const uint64_t* a; const uint64_t* b; int words; int rows; int K; double alo, ahi, blo, bhi, scale; const double* kscale;
    __device__ __forceinline__ void load8(int row, int k0, double (&v)[8]) const {
        unsigned ba = 0, bb = 0;
        if (row < rows && k0 < K) {
            size_t o = (size_t)row * words + (k0 >> 6);
            ba = (unsigned)((a[o] >> (k0 & 63)) & 0xFFu);
            bb = (unsigned)((b[o] >> (k0 & 63)) & 0xFFu);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const double va = ((ba >> e) & 1u) ? ahi : alo, vb = ((bb >> e) & 1u) ? bhi : blo;
            v[e] = (row < rows && (k0 + e) < K) ? (kscale ? scale * kscale[k0 + e] : scale) * (va - vb) : 0.0;
        }
    }
};

// dW on the integer tensor path (tcgen05): the unit-major bit matrices are expanded to int8 operands A = t - r (or t),
// B = t, and dWi = A B^T is the exact int32 contraction; the reference's updatedMatrix is s * dWi (s = 0.5 for the Delta
// rule's 0.5 (t - r), 1 for Hebbian).  dWi overwrites the front of the float64 dW buffer.
int contraction_int(hn_handle h, const uint64_t* bitsT, const uint64_t* bitsR, int n, ValueMap tvm, ValueMap rvm) {
    const int N = h->N, pwords = (n + 63) / 64, pitch = round_up(n, 128);
    HN_TRY(h->opA8.reserve((size_t)N * pitch));
    HN_TRY(h->opB8.reserve((size_t)N * pitch));
    const int g = grid_for((int64_t)N * (pitch / 8), 256, h->sms);
    expand_units_i8_kernel<<<g, 256, 0, h->st>>>(bitsT, bitsR, N, pwords, n, pitch, (int)tvm.lo, (int)tvm.hi, (int)rvm.lo, (int)rvm.hi,
                                                 h->opA8.as<int8_t>());
    if (bitsR)
        expand_units_i8_kernel<<<g, 256, 0, h->st>>>(bitsT, nullptr, N, pwords, n, pitch, (int)tvm.lo, (int)tvm.hi, 0, 0, h->opB8.as<int8_t>());
    HN_CUDA_TRY(cudaGetLastError());
    const int8_t* B8 = bitsR ? h->opB8.as<int8_t>() : h->opA8.as<int8_t>();
    return launch_igemm_tc05_mats(h->opA8.as<int8_t>(), N, B8, N, n, pitch, reinterpret_cast<int32_t*>(h->dW), h->ldw, h->sms, h->st);
}
int all_reduce_dwi(hn_handle h) {   // exact integers: half the NVLink bytes of the float64 all-reduce
    if (!h->comm) return HN_OK;
    int32_t* d = reinterpret_cast<int32_t*>(h->dW);
    HN_NCCL_TRY(g_nccl.AllReduce(d, d, (size_t)h->N * h->ldw, NCCL_INT32, NCCL_SUM, h->comm, h->st));
    return HN_OK;
}
int apply_update_and_constraints_int(hn_handle h, double s) {
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    Scratch* sc = h->scratch();
    HN_CUDA_TRY(cudaMemsetAsync(&sc->maxabs_bits, 0, 8, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(&sc->not_quarters, 0, 4, h->st));
    HN_CUDA_TRY(cudaMemsetAsync(&sc->frac_flags, 0, 4, h->st));
    dim3 g((h->N + 31) / 32, (h->N + 31) / 32), b(32, 8);
    update_constraints_i32_kernel<<<g, b, 0, h->st>>>(h->W, reinterpret_cast<const int32_t*>(h->dW), h->N, h->ldw, h->cfg.learning_rate, s,
                                                     h->cfg.force_zero_diagonal, h->cfg.force_symmetric, &sc->maxabs_bits, &sc->not_quarters,
                                                     &sc->frac_flags);
    HN_CUDA_TRY(cudaGetLastError());
    h->w_symmetric = h->cfg.force_symmetric != 0;
    weights_changed(h);
    h->fused_flags_valid = true;
    return HN_OK;
}

int all_reduce_dw(hn_handle h) {
    if (!h->comm) return HN_OK;
    HN_NCCL_TRY(g_nccl.AllReduce(h->dW, h->dW, (size_t)h->N * h->ldw, NCCL_FLOAT64, NCCL_SUM, h->comm, h->st));
    return HN_OK;
}

// AllStatesAreStable across ranks (LearningMethod.go:56-58 with the targets sharded): minimum of the per-rank flags
int all_reduce_min_flag(hn_handle h, bool* flag) {
    if (!h->comm) return HN_OK;
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    int32_t* d = &h->scratch()->stable_flag;
    int32_t v = *flag ? 1 : 0;
    HN_CUDA_TRY(cudaMemcpyAsync(d, &v, 4, cudaMemcpyHostToDevice, h->st));
    HN_NCCL_TRY(g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_MIN, h->comm, h->st));
    HN_CUDA_TRY(cudaMemcpyAsync(&v, d, 4, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    *flag = v != 0;
    return HN_OK;
}

int apply_update_and_constraints(hn_handle h);
// a rank whose shard of this epoch's targets is empty still takes part in the all-reduce (with a zero dW) and applies
// the same update, so replicas stay identical (iterative-batch prefixes can leave late ranks without targets)
int empty_shard_epoch(hn_handle h, bool integer_path, double s) {
    HN_CUDA_TRY(cudaMemsetAsync(h->dW, 0, sizeof(double) * (size_t)h->N * h->ldw, h->st));
    if (integer_path) {
        HN_TRY(all_reduce_dwi(h));
        HN_TRY(apply_update_and_constraints_int(h, s));
    } else {
        HN_TRY(all_reduce_dw(h));
        HN_TRY(apply_update_and_constraints(h));
    }
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int apply_update_and_constraints(hn_handle h) {
    apply_update_kernel<<<grid_for((int64_t)h->N * h->ldw, 256, h->sms), 256, 0, h->st>>>(h->W, h->dW, h->N, h->ldw,
                                                                                       h->cfg.learning_rate);
    HN_CUDA_TRY(cudaGetLastError());
    if (h->cfg.force_zero_diagonal || h->cfg.force_symmetric) {
        dim3 g((h->N + SYM_T - 1) / SYM_T, (h->N + SYM_T - 1) / SYM_T), b(SYM_T, 8);
        constraints_kernel<<<g, b, 0, h->st>>>(h->W, h->N, h->ldw, h->cfg.force_zero_diagonal, h->cfg.force_symmetric);
        HN_CUDA_TRY(cudaGetLastError());
    }
    h->w_symmetric = h->cfg.force_symmetric != 0;
    weights_changed(h);
    return HN_OK;
}

int upload_states(hn_handle h, DevBuf& buf, const uint64_t* host, int n) {
    HN_TRY(buf.reserve(sizeof(uint64_t) * (size_t)std::max(n, 1) * h->words));
    HN_CUDA_TRY(cudaMemcpyAsync(buf.p, host, sizeof(uint64_t) * (size_t)n * h->words, cudaMemcpyHostToDevice, h->st));
    return HN_OK;
}

int hebbian_epoch_impl(hn_handle h, const uint64_t* states, int n, ValueMap vm) {
    HN_TRY(upload_states(h, h->bitsA, states, n));
    if (h->use_tc05) {   // exact integer contraction on the tcgen05 path, int32 all-reduce, fused update
        const int pwords = (n + 63) / 64;
        HN_TRY(h->bitsAT.reserve(sizeof(uint64_t) * (size_t)h->N * pwords));
        transpose_bits_kernel<<<grid_for((int64_t)h->N * pwords, 256, h->sms), 256, 0, h->st>>>(
            h->bitsA.as<uint64_t>(), h->bitsAT.as<uint64_t>(), h->N, h->words, n, pwords);
        HN_CUDA_TRY(cudaGetLastError());
        HN_TRY(contraction_int(h, h->bitsAT.as<uint64_t>(), nullptr, n, vm, vm));
        HN_TRY(all_reduce_dwi(h));
        return apply_update_and_constraints_int(h, 1.0);
    }
    HN_TRY(contraction_hebbian(h, h->bitsA.as<uint64_t>(), n, vm));
    HN_TRY(all_reduce_dw(h));
    return apply_update_and_constraints(h);
}

// tvm: value map of the target states as the rule sees them (mapped rules re-activate);
// the relaxed copies always carry the network domain's values.
// UpdateState (HopfieldNetwork.go:280-291) for n states: bitsB -> bitsC, one sweep each in its own fresh order
// (row p of `perms`), no stability test.
int sweep_once(hn_handle h, int n, const uint16_t* perms) {
    const int N = h->N, words = h->words;
    HN_TRY(ensure_margin(h));
    HN_TRY(h->bitsC.reserve(sizeof(uint64_t) * (size_t)n * words));
    if (h->maxabs == 0.0) {
        // all-zero matrix (the first Delta epoch of a fresh network): every field is an exact zero whatever the state,
        // h <= 0 selects the low value for every unit (BipolarDomainManager.go:10-22).  Without this every decision
        // would take the exact tie path: seconds per epoch at N = 16384.
        HN_CUDA_TRY(cudaMemsetAsync(h->bitsC.p, 0, sizeof(uint64_t) * (size_t)n * words, h->st));
        return HN_OK;
    }
    HN_TRY(refresh_wt(h));
    uint16_t *d_perm = nullptr, *d_inv = nullptr;
    HN_TRY(upload_pool(h, perms, n, h->perms, &d_perm, &d_inv));
    HN_TRY(h->offsets.reserve(sizeof(int32_t) * (size_t)n));   // offsets[p] = p : state p uses its own row
    std::vector<int32_t> off((size_t)n);
    for (int i = 0; i < n; i++) off[i] = i;
    HN_CUDA_TRY(cudaMemcpyAsync(h->offsets.p, off.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));  // `off` must outlive the copy
    HN_TRY(h->H.reserve(sizeof(double) * (size_t)n * h->ldw));
    HN_TRY(ensure_learning_image(h));
    const bool int_fields = integer_stream(h) && h->w_symmetric;
    HN_CUDA_TRY(cudaEventRecord(h->lev[0], h->st));
    if (int_fields) HN_TRY(integer_fields(h, h->bitsB.as<uint64_t>(), n, reinterpret_cast<int32_t*>(h->H.as<double>())));
    else HN_TRY(launch_fields(h, h->bitsB.as<uint64_t>(), n, h->H.as<double>(), domain_values(h->cfg.domain), integer_stream(h) ? h->Wraw : nullptr));
    HN_CUDA_TRY(cudaEventRecord(h->lev[1], h->st));
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    Scratch* sc = h->scratch();
    HN_CUDA_TRY(cudaMemsetAsync(sc, 0, offsetof(Scratch, asym_flag), h->st));
    RelaxParams p{};
    p.N = N; p.ldw = h->ldw; p.words = words;
    p.Wrow = h->W;
    p.H0 = h->H.as<double>(); p.ldh = h->ldw;
    p.H0i = int_fields ? reinterpret_cast<const int32_t*>(h->H.as<double>()) : nullptr;
    p.probes = h->bitsB.as<uint64_t>(); p.perm_pool = d_perm; p.inv_pool = d_inv; p.ldp = h->ldp; p.n_pool = n;
    p.offsets = h->offsets.as<int32_t>(); p.B = n; p.max_iter = 1; p.max_unstable = 0; p.domain = h->cfg.domain;
    p.check_stability = 0; p.serial_stream = 0;
    p.final_states = h->bitsC.as<uint64_t>();
    HN_TRY(h->flips.reserve(sizeof(int32_t) * (size_t)n));
    p.flips = h->flips.as<int32_t>();   // counter only (roofline of the sweep: one column per flip)
    p.work_counter = &sc->work_counter; p.status = sc->status;
    HN_TRY(launch_relax(h, p));
    HN_CUDA_TRY(cudaEventRecord(h->lev[2], h->st));
    HN_CUDA_TRY(cudaMemsetAsync(sc->totals, 0, sizeof(sc->totals), h->st));
    sum_counters_kernel<<<std::min((n + 255) / 256, h->sms * 4), 256, 0, h->st>>>(h->flips.as<int32_t>(), nullptr, nullptr, n, sc->totals);
    HN_CUDA_TRY(cudaGetLastError());
    h->sweep_timed = true;
    return HN_OK;
}

int delta_epoch_impl(hn_handle h, const uint64_t* states, const uint64_t* noised, const uint16_t* perms, int n,
                     uint64_t* relaxed_out, ValueMap tvm, bool thermal = false) {
    const int N = h->N, words = h->words;
    HN_TRY(upload_states(h, h->bitsA, states, n));
    HN_TRY(upload_states(h, h->bitsB, noised, n));
    h->sweep_timed = false;
    HN_TRY(sweep_once(h, n, perms));   // one sweep per noised copy in its own fresh order (HopfieldNetwork.go:281-282)
    if (relaxed_out)
        HN_CUDA_TRY(cudaMemcpyAsync(relaxed_out, h->bitsC.p, sizeof(uint64_t) * (size_t)n * words, cudaMemcpyDeviceToHost, h->st));
    // dW[i][j] = sum_mu 0.5*(t_mu,i - r_mu,i) * t_mu,j
    const int pwords = (n + 63) / 64;
    HN_TRY(h->bitsAT.reserve(sizeof(uint64_t) * (size_t)N * pwords));
    HN_TRY(h->bitsCT.reserve(sizeof(uint64_t) * (size_t)N * pwords));
    const int tg = grid_for((int64_t)N * pwords, 256, h->sms);
    transpose_bits_kernel<<<tg, 256, 0, h->st>>>(h->bitsA.as<uint64_t>(), h->bitsAT.as<uint64_t>(), N, words, n, pwords);
    transpose_bits_kernel<<<tg, 256, 0, h->st>>>(h->bitsC.as<uint64_t>(), h->bitsCT.as<uint64_t>(), N, words, n, pwords);
    HN_CUDA_TRY(cudaGetLastError());
    const ValueMap rvm = domain_values(h->cfg.domain);
    const double* kscale = nullptr;
    if (thermal) {
        // f_mu = exp(-||W t_mu||_2 / (1 + ||W||_F)) with the CURRENT W (LearningRule.go:134,149-150)
        const int np = 1024;
        HN_TRY(h->partial.reserve(sizeof(double) * (np + 2 + (size_t)n)));
        double* part = h->partial.as<double>();
        sumsq_partial_kernel<<<np, RED_THREADS, 0, h->st>>>(h->W, N, h->ldw, part);
        sumsq_final_kernel<<<1, RED_THREADS, 0, h->st>>>(part, np, part + np);
        HN_TRY(h->H.reserve(sizeof(double) * (size_t)n * h->ldw));
        OpBits ta{h->bitsA.as<uint64_t>(), words, n, N, tvm.lo, tvm.hi};
        OpF64 tb{h->W, h->ldw, N};
        EpiStore te{h->H.as<double>(), h->ldw, n, N};
        HN_TRY(launch_gemm_f64(ta, tb, te, n, N, N, h->st));
        thermal_factor_kernel<<<n, RED_THREADS, 0, h->st>>>(h->H.as<double>(), h->ldw, N, part + np, part + np + 2);
        HN_CUDA_TRY(cudaGetLastError());
        kscale = part + np + 2;
    }
    HN_CUDA_TRY(cudaEventRecord(h->lev[3], h->st));
    if (!thermal && h->use_tc05) {   // dW_ij = 0.5 * sum_mu (t - r)_i t_j with the integer sum on the tcgen05 path
        HN_TRY(contraction_int(h, h->bitsAT.as<uint64_t>(), h->bitsCT.as<uint64_t>(), n, tvm, rvm));
        HN_CUDA_TRY(cudaEventRecord(h->lev[4], h->st));
        HN_TRY(all_reduce_dwi(h));
        HN_CUDA_TRY(cudaEventRecord(h->lev[5], h->st));
        HN_TRY(apply_update_and_constraints_int(h, 0.5));
    } else {
        OpBitsDiffG a{h->bitsAT.as<uint64_t>(), h->bitsCT.as<uint64_t>(), pwords, N, n, tvm.lo, tvm.hi, rvm.lo, rvm.hi,
                      thermal ? 1.0 : 0.5, kscale};
        OpBits b{h->bitsAT.as<uint64_t>(), pwords, N, n, tvm.lo, tvm.hi};
        EpiStore e{h->dW, h->ldw, N, N};
        HN_TRY(launch_gemm_f64(a, b, e, N, N, n, h->st));
        HN_CUDA_TRY(cudaEventRecord(h->lev[4], h->st));
        HN_TRY(all_reduce_dw(h));
        HN_CUDA_TRY(cudaEventRecord(h->lev[5], h->st));
        HN_TRY(apply_update_and_constraints(h));
    }
    HN_CUDA_TRY(cudaEventRecord(h->lev[6], h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    {
        float t = 0;
        hn_learn_timing& lt = h->ltiming;
        lt.sweep_fields_ms = lt.sweep_ms = 0;
        if (h->sweep_timed) {
            cudaEventElapsedTime(&t, h->lev[0], h->lev[1]); lt.sweep_fields_ms = t;
            cudaEventElapsedTime(&t, h->lev[1], h->lev[2]); lt.sweep_ms = t;
        }
        cudaEventElapsedTime(&t, h->lev[3], h->lev[4]); lt.dw_gemm_ms = t;
        cudaEventElapsedTime(&t, h->lev[4], h->lev[5]); lt.allreduce_ms = t;
        cudaEventElapsedTime(&t, h->lev[5], h->lev[6]); lt.update_ms = t;
        lt.targets = n;
        lt.sweep_flips = 0;
        if (h->sweep_timed) {
            unsigned long long f = 0;
            HN_CUDA_TRY(cudaMemcpy(&f, &h->scratch()->totals[0], 8, cudaMemcpyDeviceToHost));
            lt.sweep_flips = (int64_t)f;
        }
    }
    return HN_OK;
}

int energy_profiles_impl(hn_handle h, const uint64_t* d_states, int n, double* energies_out, int32_t* unstable_out,
                         uint8_t* stable_out, double* state_energy_out) {
    const int N = h->N;
    HN_TRY(ensure_margin(h));
    HN_TRY(h->H.reserve(sizeof(double) * (size_t)n * h->ldw));
    HN_TRY(ensure_learning_image(h));
    // exact arithmetic (quarter-integer weights) with the integer image at hand: fields from the integer tensor path, the
    // float64 energies follow exactly (h = g / scale, every product and sum exact)
    const bool int_path = h->exact_arith && integer_stream(h) && h->w_symmetric;
    HN_CUDA_TRY(cudaEventRecord(h->lev[7], h->st));
    if (int_path) HN_TRY(integer_fields(h, d_states, n, reinterpret_cast<int32_t*>(h->H.as<double>())));
    else HN_TRY(launch_fields(h, d_states, n, h->H.as<double>(), domain_values(h->cfg.domain)));
    HN_CUDA_TRY(cudaEventRecord(h->lev[0], h->st));
    if (energies_out) HN_TRY(h->energies.reserve(sizeof(double) * (size_t)n * N));
    HN_TRY(h->eunst.reserve(sizeof(int32_t) * (size_t)n));
    HN_TRY(h->estab.reserve((size_t)n));
    HN_TRY(h->esum.reserve(sizeof(double) * (size_t)n));
    if (int_path)
        energy_epilogue_i32_kernel<<<n, RED_THREADS, 0, h->st>>>(reinterpret_cast<const int32_t*>(h->H.as<double>()), h->ldw, d_states, h->words, N,
                                                                h->cfg.domain, 1.0 / h->int_scale, h->cfg.max_relax_unstable_units,
                                                                energies_out ? h->energies.as<double>() : nullptr,
                                                                h->eunst.as<int32_t>(), h->estab.as<uint8_t>(), h->esum.as<double>());
    else
    energy_epilogue_kernel<<<n, RED_THREADS, 0, h->st>>>(h->H.as<double>(), h->ldw, d_states, h->words, h->W, h->ldw, N,
                                                        h->cfg.domain, h->exact_arith ? 0.0 : h->tau_base, h->cfg.max_relax_unstable_units,
                                                        energies_out ? h->energies.as<double>() : nullptr,
                                                        h->eunst.as<int32_t>(), h->estab.as<uint8_t>(), h->esum.as<double>(),
                                                        nullptr);
    HN_CUDA_TRY(cudaGetLastError());
    HN_CUDA_TRY(cudaEventRecord(h->lev[1], h->st));
    if (energies_out) HN_CUDA_TRY(cudaMemcpyAsync(energies_out, h->energies.p, sizeof(double) * (size_t)n * N, cudaMemcpyDeviceToHost, h->st));
    if (unstable_out) HN_CUDA_TRY(cudaMemcpyAsync(unstable_out, h->eunst.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
    if (stable_out) HN_CUDA_TRY(cudaMemcpyAsync(stable_out, h->estab.p, (size_t)n, cudaMemcpyDeviceToHost, h->st));
    if (state_energy_out) HN_CUDA_TRY(cudaMemcpyAsync(state_energy_out, h->esum.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    {
        float t = 0;
        cudaEventElapsedTime(&t, h->lev[7], h->lev[0]); h->ltiming.energy_gemm_ms = t;
        cudaEventElapsedTime(&t, h->lev[0], h->lev[1]); h->ltiming.energy_epilogue_ms = t;
    }
    return HN_OK;
}

int append_targets(hn_handle h, const uint64_t* states, int n) {
    const size_t each = sizeof(uint64_t) * (size_t)h->words;
    DevBuf nb;
    HN_TRY(nb.reserve(each * (size_t)(h->P + n)));
    if (h->P > 0) HN_CUDA_TRY(cudaMemcpyAsync(nb.p, h->targets.p, each * (size_t)h->P, cudaMemcpyDeviceToDevice, h->st));
    HN_CUDA_TRY(cudaMemcpyAsync((char*)nb.p + each * (size_t)h->P, states, each * (size_t)n, cudaMemcpyHostToDevice, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    h->targets.release();
    h->targets = nb;
    h->P += n;
    return HN_OK;
}

bool check_handle(hn_handle h, const char* fn) {
    if (!h) { set_error("%s: handle is NULL", fn); return false; }
    return true;
}

}  // namespace

extern "C" {

int hn_abi_version(void) { return HN_ABI_VERSION; }
const char* hn_last_error(void) { return hn::g_err; }

void hn_default_config(hn_config* cfg) {
    if (!cfg) return;
    cfg->dimension = 0;
    cfg->domain = HN_DOMAIN_BIPOLAR;
    cfg->force_symmetric = 1;
    cfg->force_zero_diagonal = 1;
    cfg->max_relax_unstable_units = 0;
    cfg->max_relax_iterations = 100;
    cfg->learning_rate = 1.0;
    cfg->device = 0;
    cfg->column_stream = HN_COLUMNS_AUTO;
}

int hn_create(const hn_config* cfg, hn_handle* out) {
    if (!cfg || !out) { set_error("hn_create: NULL argument"); return HN_E_INVALID; }
    *out = nullptr;
    if (cfg->dimension <= 0) {
        // HopfieldNetworkBuilder.go:215-217
        set_error("HopfieldNetworkBuilder encountered an error during build! Dimension must be explicitly set to a positive integer!");
        return HN_E_INVALID;
    }
    if (cfg->dimension > 65535) { set_error("dimension %d exceeds 65535 (uint16 unit indices)", cfg->dimension); return HN_E_INVALID; }
    if (cfg->domain != HN_DOMAIN_BIPOLAR && cfg->domain != HN_DOMAIN_BINARY) { set_error("unknown domain %d", cfg->domain); return HN_E_INVALID; }
    if (cfg->max_relax_iterations <= 0) { set_error("max_relax_iterations must be positive"); return HN_E_INVALID; }
    if (cfg->column_stream < HN_COLUMNS_F64 || cfg->column_stream > HN_COLUMNS_AUTO) { set_error("unknown column_stream %d", cfg->column_stream); return HN_E_INVALID; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libhopfield_b200 has no CPU fallback", ce != cudaSuccess ? cudaGetErrorString(ce) : "0 devices");
        return HN_E_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device ordinal %d out of range (0..%d)", cfg->device, ndev - 1); return HN_E_NO_DEVICE; }
    cudaDeviceProp prop;
    HN_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only and has no fallback", cfg->device, prop.major, prop.minor);
        return HN_E_NO_DEVICE;
    }
    hn_handle h = new (std::nothrow) hn_handle_s();
    if (!h) return HN_E_OOM;
    h->cfg = *cfg;
    h->N = cfg->dimension;
    h->ldw = round_up(h->N, 16);
    h->ldp = round_up(h->N, 16);
    h->words = hn_words(h->N);
    h->sms = prop.multiProcessorCount;
    int rc = HN_OK;
    do {
        if (cudaSetDevice(cfg->device) != cudaSuccess) { rc = HN_E_CUDA; break; }
        if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { rc = HN_E_CUDA; break; }
        if (cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking) != cudaSuccess) { rc = HN_E_CUDA; break; }
        const size_t wb = sizeof(double) * (size_t)h->N * h->ldw;
        if (cudaMalloc(&h->W, wb) != cudaSuccess || cudaMalloc(&h->dW, wb) != cudaSuccess) { rc = HN_E_OOM; break; }
        if (cudaMemsetAsync(h->W, 0, wb, h->st) != cudaSuccess) { rc = HN_E_CUDA; break; }  // matrix.Zero(), :248-250
        if (cudaMemsetAsync(h->dW, 0, wb, h->st) != cudaSuccess) { rc = HN_E_CUDA; break; }
        for (auto& e : h->ev) if (cudaEventCreate(&e) != cudaSuccess) { rc = HN_E_CUDA; break; }
        for (auto& e : h->lev) if (cudaEventCreate(&e) != cudaSuccess) { rc = HN_E_CUDA; break; }
        if (rc != HN_OK) break;
        if (cudaStreamSynchronize(h->st) != cudaSuccess) { rc = HN_E_CUDA; break; }
    } while (0);
    if (rc != HN_OK) {
        set_error("hn_create: CUDA setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        hn_destroy(h);
        return rc;
    }
    h->w_symmetric = true;  // the zero matrix
    if (const char* e = getenv("HN_IGEMM")) h->use_tc05 = std::strcmp(e, "mma_sync") != 0;
    if (const char* e = getenv("HN_DENSE")) h->use_dense = std::atoi(e) != 0;
    if (const char* e = getenv("HN_DENSE_ROWS")) h->dense_max_rows = std::max(1, std::atoi(e));
    if (const char* e = getenv("HN_DENSE_MIN_N")) h->dense_min_n = std::max(32, std::atoi(e));
    if (const char* e = getenv("HN_DENSE_MIN_FLIPS")) h->dense_min_flips = std::atof(e);
    if (const char* e = getenv("HN_ZERO_COPY")) h->zero_copy = std::atoi(e) != 0;
    if (const char* e = getenv("HN_EXACT_NORM")) h->exact_norm = std::atoi(e) != 0;
    if (const char* e = getenv("HN_DENSE_MAP")) h->dense_want_map3d = std::strcmp(e, "2d") != 0;
    if (const char* e = getenv("HN_TMA_BOX_DENSE")) { const int x = std::atoi(e); if (x == 8 || x == 16 || x == 32) h->dense_box_rows = x; }
    *out = h;
    return HN_OK;
}

int hn_destroy(hn_handle h) {
    if (!h) return HN_OK;
    cudaSetDevice(h->cfg.device);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->st) cudaStreamSynchronize(h->st);
    DevBuf* bufs[] = {&h->targets, &h->probes, &h->pool, &h->offsets, &h->H, &h->finals, &h->steps, &h->stable,
                      &h->flips, &h->margin, &h->dist, &h->hash, &h->energies, &h->history, &h->rep, &h->first, &h->hits,
                      &h->flag, &h->rank, &h->table, &h->rep1, &h->rep2, &h->pairs, &h->opA8, &h->opB8, &h->d_mperm, &h->d_mbb, &h->d_mx, &h->d_info,
                      &h->d_state, &h->d_sweeps, &h->d_dflips, &h->d_dmargin, &h->d_finished, &h->d_scan, &h->attractor, &h->ufirst, &h->uhits, &h->small, &h->partial,
                      &h->bitsA, &h->bitsB, &h->bitsC, &h->bitsAT, &h->bitsCT, &h->perms, &h->eunst, &h->estab, &h->esum,
                      &h->emargin};
    for (DevBuf* b : bufs) b->release();
    if (h->W) cudaFree(h->W);
    if (h->WT) cudaFree(h->WT);
    if (h->W32) cudaFree(h->W32);
    if (h->Wraw) cudaFree(h->Wraw);
    if (h->K16) cudaFree(h->K16);
    if (h->K8) cudaFree(h->K8);
    if (h->dW) cudaFree(h->dW);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->lev) if (e) cudaEventDestroy(e);
    for (auto& e : h->chunk_ev) if (e) cudaEventDestroy(e);
    for (auto& e : h->dense_ev) if (e) cudaEventDestroy(e);
    if (h->st2) { cudaStreamSynchronize(h->st2); cudaStreamDestroy(h->st2); }
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return HN_OK;
}

int hn_set_weights(hn_handle h, const double* w_host) {
    if (!check_handle(h, "hn_set_weights") || !w_host) { if (h) set_error("hn_set_weights: NULL matrix"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_CUDA_TRY(cudaMemsetAsync(h->W, 0, sizeof(double) * (size_t)h->N * h->ldw, h->st));
    HN_CUDA_TRY(cudaMemcpy2DAsync(h->W, sizeof(double) * h->ldw, w_host, sizeof(double) * h->N, sizeof(double) * h->N, h->N,
                                  cudaMemcpyHostToDevice, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    HN_TRY(detect_symmetry(h));
    weights_changed(h);
    return detect_integer_structure(h);
}

int hn_get_weights(hn_handle h, double* w_host) {
    if (!check_handle(h, "hn_get_weights") || !w_host) { if (h) set_error("hn_get_weights: NULL matrix"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_CUDA_TRY(cudaMemcpy2DAsync(w_host, sizeof(double) * h->N, h->W, sizeof(double) * h->ldw, sizeof(double) * h->N, h->N,
                                  cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int hn_set_integer_structure(hn_handle h, const double* w_raw) {
    if (!check_handle(h, "hn_set_integer_structure") || !w_raw) { if (h) set_error("hn_set_integer_structure: NULL matrix"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    const int N = h->N;
    // proportionality W = c * w_raw, c > 0, checked on the host against a copy of the resident weights
    std::vector<double> w((size_t)N * N);
    HN_TRY(hn_get_weights(h, w.data()));
    size_t imax = 0;
    for (size_t i = 1; i < w.size(); i++) if (std::fabs(w_raw[i]) > std::fabs(w_raw[imax])) imax = i;
    const double c = w_raw[imax] != 0.0 ? w[imax] / w_raw[imax] : 0.0;   // W = fl(c * Wraw): one rounding per element
    if (!(c > 0.0)) { set_error("hn_set_integer_structure: weights are not a positive multiple of the given matrix"); return HN_E_INVALID; }
    for (size_t i = 0; i < w.size(); i++)
        if (std::fabs(w[i] - c * w_raw[i]) > 1e-13 * std::fabs(c * w_raw[i])) {
            set_error("hn_set_integer_structure: weights are not proportional to the given matrix (element %zu)", i);
            return HN_E_INVALID;
        }
    const size_t wb = sizeof(double) * (size_t)N * h->ldw;
    if (!h->Wraw) HN_CUDA_TRY(cudaMalloc(&h->Wraw, wb));
    HN_CUDA_TRY(cudaMemsetAsync(h->Wraw, 0, wb, h->st));
    HN_CUDA_TRY(cudaMemcpy2DAsync(h->Wraw, sizeof(double) * h->ldw, w_raw, sizeof(double) * N, sizeof(double) * N, N,
                                  cudaMemcpyHostToDevice, h->st));
    HN_TRY(build_integer_structure(h, h->w_symmetric, c));
    if (!h->int_ok) {
        HN_TRY(detect_integer_structure(h));   // a refused declaration leaves whatever the matrix itself reveals in force
        set_error("hn_set_integer_structure: no multiple 0.5, 1, 2 or 4 of the matrix is integer valued within int16");
        return HN_E_INVALID;
    }
    return HN_OK;
}
int hn_integer_stream_active(hn_handle h, int32_t* active) {
    if (!check_handle(h, "hn_integer_stream_active") || !active) return HN_E_INVALID;
    *active = integer_stream(h) ? 1 : 0;
    return HN_OK;
}

int hn_set_targets(hn_handle h, const uint64_t* targets_packed, int32_t n_targets) {
    if (!check_handle(h, "hn_set_targets")) return HN_E_INVALID;
    if (n_targets < 0 || (n_targets > 0 && !targets_packed)) { set_error("hn_set_targets: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    h->targets.release();
    h->P = 0;
    if (n_targets > 0) HN_TRY(append_targets(h, targets_packed, n_targets));
    return HN_OK;
}
int hn_num_targets(hn_handle h, int32_t* n_targets) {
    if (!check_handle(h, "hn_num_targets") || !n_targets) return HN_E_INVALID;
    *n_targets = h->P;
    return HN_OK;
}
int hn_get_targets(hn_handle h, uint64_t* targets_packed) {
    if (!check_handle(h, "hn_get_targets") || !targets_packed) return HN_E_INVALID;
    HN_TRY(set_device(h));
    if (h->P > 0) {
        HN_CUDA_TRY(cudaMemcpyAsync(targets_packed, h->targets.p, sizeof(uint64_t) * (size_t)h->P * h->words, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    }
    return HN_OK;
}

int hn_hebbian_epoch(hn_handle h, const uint64_t* states_packed, int32_t n) {
    if (!check_handle(h, "hn_hebbian_epoch")) return HN_E_INVALID;
    if (n <= 0 || !states_packed) { set_error("hn_hebbian_epoch: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(hebbian_epoch_impl(h, states_packed, n, domain_values(h->cfg.domain)));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int hn_delta_epoch(hn_handle h, const uint64_t* states_packed, const uint64_t* noised_packed, const uint16_t* perms,
                   int32_t n, uint64_t* relaxed_out) {
    if (!check_handle(h, "hn_delta_epoch")) return HN_E_INVALID;
    if (n <= 0 || !states_packed || !noised_packed || !perms) { set_error("hn_delta_epoch: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(validate_pool_host(h, perms, n));
    return delta_epoch_impl(h, states_packed, noised_packed, perms, n, relaxed_out, domain_values(h->cfg.domain));
}

int hn_thermal_delta_epoch(hn_handle h, const uint64_t* states_packed, const uint64_t* noised_packed,
                           const uint16_t* perms, int32_t n, uint64_t* relaxed_out) {
    if (!check_handle(h, "hn_thermal_delta_epoch")) return HN_E_INVALID;
    if (n <= 0 || !states_packed || !noised_packed || !perms) { set_error("hn_thermal_delta_epoch: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(validate_pool_host(h, perms, n));
    return delta_epoch_impl(h, states_packed, noised_packed, perms, n, relaxed_out, domain_values(h->cfg.domain), true);
}

int hn_rule_epoch(hn_handle h, int32_t rule, const uint64_t* states_packed, const uint64_t* noised_packed,
                  const uint16_t* perms, int32_t n, uint64_t* relaxed_out) {
    if (!check_handle(h, "hn_rule_epoch")) return HN_E_INVALID;
    if (rule < 0 || rule > 5 || n <= 0 || !states_packed) { set_error("hn_rule_epoch: bad argument (rule %d, n %d)", rule, n); return HN_E_INVALID; }
    const bool is_thermal = rule == HN_RULE_THERMAL_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_THERMAL_DELTA;
    const bool is_delta = rule == HN_RULE_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_DELTA || is_thermal;
    if (is_delta && (!noised_packed || !perms)) { set_error("hn_rule_epoch: the delta rules need the noised copies and the update orders"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    // value map the rule sees (see learn_states_impl): mapped Hebbian re-activates with the BINARY manager (sic,
    // LearningRule.go:80), the mapped delta rules with the Bipolar manager (:119, :163)
    ValueMap tvm = domain_values(h->cfg.domain);
    if (rule == HN_RULE_BIPOLAR_MAPPED_HEBBIAN) tvm = ValueMap{0.0, 1.0};
    if (rule == HN_RULE_BIPOLAR_MAPPED_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_THERMAL_DELTA) tvm = ValueMap{-1.0, 1.0};
    if (!is_delta) {
        HN_TRY(hebbian_epoch_impl(h, states_packed, n, tvm));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        return HN_OK;
    }
    HN_TRY(validate_pool_host(h, perms, n));
    return delta_epoch_impl(h, states_packed, noised_packed, perms, n, relaxed_out, tvm, is_thermal);
}

int hn_enforce_constraints(hn_handle h) {
    if (!check_handle(h, "hn_enforce_constraints")) return HN_E_INVALID;
    HN_TRY(set_device(h));
    if (h->cfg.force_zero_diagonal || h->cfg.force_symmetric) {
        dim3 g((h->N + SYM_T - 1) / SYM_T, (h->N + SYM_T - 1) / SYM_T), b(SYM_T, 8);
        constraints_kernel<<<g, b, 0, h->st>>>(h->W, h->N, h->ldw, h->cfg.force_zero_diagonal, h->cfg.force_symmetric);
        HN_CUDA_TRY(cudaGetLastError());
        if (h->cfg.force_symmetric) h->w_symmetric = true;
        weights_changed(h);
    }
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

// ||W||_F exactly as the reference computes it — mat.Norm(W, 2) = Dlange('F'): gonum's Dlassq recurrence (scaled sum of
// squares) over the rows in order, then scale*sqrt(ssq) (HopfieldNetwork.go:260; oracle: orc_frobenius) — for a matrix
// of quarter-integers given as k = 4w.  The recurrence is sequential by nature (every addition rounds), so it runs on
// the host, once per LearnStates; the term a unit adds depends only on |k| and the current scale, hence the table.
// The resulting W = fl((1/norm) * Wraw) is then bit-identical to the reference's, which the exact re-evaluation of tied
// fields and the +0/-0 marks of the unique table depend on.
static double exact_frobenius_quarters(const int16_t* k, size_t count) {
    double scale = 0.0, ssq = 1.0;
    std::vector<double> term(32768);
    std::vector<uint8_t> have(32768, 0);
    for (size_t i = 0; i < count; i++) {
        const int a = k[i] < 0 ? -(int)k[i] : (int)k[i];
        if (a == 0) continue;
        const double absxi = 0.25 * (double)a;
        if (scale < absxi) {
            const double ratio = scale / absxi;
            double t = ssq * ratio;
            t = t * ratio;
            ssq = 1.0 + t;
            scale = absxi;
            std::fill(have.begin(), have.end(), 0);
        } else {
            if (!have[a]) { const double ratio = absxi / scale; term[a] = ratio * ratio; have[a] = 1; }
            ssq += term[a];
        }
    }
    return scale * std::sqrt(ssq);
}

int hn_set_norm_algorithm(hn_handle h, int32_t algorithm) {
    if (!check_handle(h, "hn_set_norm_algorithm")) return HN_E_INVALID;
    if (algorithm != HN_NORM_DLASSQ_CLASSIC && algorithm != HN_NORM_DLASSQ_LAPACK310) {
        set_error("hn_set_norm_algorithm: unknown algorithm %d", algorithm);
        return HN_E_INVALID;
    }
    h->exact_norm = algorithm == HN_NORM_DLASSQ_CLASSIC;
    return HN_OK;
}

int hn_normalize_frobenius(hn_handle h, double* norm_out) {
    if (!check_handle(h, "hn_normalize_frobenius")) return HN_E_INVALID;
    HN_TRY(set_device(h));
    const int np = 1024;
    HN_TRY(h->partial.reserve(sizeof(double) * (np + 2)));
    double* part = h->partial.as<double>();
    const bool capture = true;   // every stream: the integer structure also gives the exact unique-table key (relax.cuh)
    if (capture) {  // keep the un-normalised matrix: W = (1/norm) * Wraw is the integer structure of the new weights
        const size_t wb = sizeof(double) * (size_t)h->N * h->ldw;
        if (!h->Wraw) HN_CUDA_TRY(cudaMalloc(&h->Wraw, wb));
        HN_CUDA_TRY(cudaMemcpyAsync(h->Wraw, h->W, wb, cudaMemcpyDeviceToDevice, h->st));
    }
    sumsq_partial_kernel<<<np, RED_THREADS, 0, h->st>>>(h->W, h->N, h->ldw, part);
    sumsq_final_kernel<<<1, RED_THREADS, 0, h->st>>>(part, np, part + np);
    if (h->exact_norm) {   // quarter-integer weights (every Hebbian / Delta matrix learned from zero): the reference's own recurrence
        const size_t count = (size_t)h->N * h->N;
        HN_TRY(h->opA8.reserve(sizeof(int16_t) * count));
        HN_TRY(h->small.reserve(sizeof(Scratch)));
        int32_t* bad = &h->scratch()->int_flag;
        HN_CUDA_TRY(cudaMemsetAsync(bad, 0, 4, h->st));
        quarters_i16_kernel<<<grid_for((int64_t)count, 256, h->sms), 256, 0, h->st>>>(h->W, h->N, h->ldw, h->opA8.as<int16_t>(), bad);
        HN_CUDA_TRY(cudaGetLastError());
        int32_t b = 1;
        HN_CUDA_TRY(cudaMemcpyAsync(&b, bad, 4, cudaMemcpyDeviceToHost, h->st));
        HN_CUDA_TRY(cudaStreamSynchronize(h->st));
        if (!b) {
            std::vector<int16_t> k(count);
            HN_CUDA_TRY(cudaMemcpyAsync(k.data(), h->opA8.p, sizeof(int16_t) * count, cudaMemcpyDeviceToHost, h->st));
            HN_CUDA_TRY(cudaStreamSynchronize(h->st));
            const double nrm = exact_frobenius_quarters(k.data(), count);
            if (nrm > 0.0 && std::isfinite(nrm)) {
                const double both[2] = {nrm, 1.0 / nrm};
                HN_CUDA_TRY(cudaMemcpyAsync(part + np, both, sizeof(both), cudaMemcpyHostToDevice, h->st));
                HN_CUDA_TRY(cudaStreamSynchronize(h->st));   // `both` is a stack array
            }
        }
    }
    scale_kernel<<<grid_for((int64_t)h->N * h->ldw, 256, h->sms), 256, 0, h->st>>>(h->W, h->N, h->ldw, part + np);
    HN_CUDA_TRY(cudaGetLastError());
    double nrm[2] = {0, 0};
    HN_CUDA_TRY(cudaMemcpyAsync(nrm, part + np, sizeof(nrm), cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    if (norm_out) *norm_out = nrm[0];
    weights_changed(h);
    if (capture && nrm[0] > 0.0 && std::isfinite(nrm[0])) HN_TRY(build_integer_structure(h, h->w_symmetric, nrm[1]));
    return HN_OK;
}

int hn_energy_profiles(hn_handle h, const uint64_t* states_packed, int32_t n, double* energies_out,
                       int32_t* unstable_out, uint8_t* stable_out, double* state_energy_out) {
    if (!check_handle(h, "hn_energy_profiles")) return HN_E_INVALID;
    if (n <= 0 || !states_packed) { set_error("hn_energy_profiles: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(upload_states(h, h->bitsA, states_packed, n));
    return energy_profiles_impl(h, h->bitsA.as<uint64_t>(), n, energies_out, unstable_out, stable_out, state_energy_out);
}

int hn_unit_energy(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t unit_index, double* energy_out) {
    if (!check_handle(h, "hn_unit_energy")) return HN_E_INVALID;
    if (n <= 0 || !states_packed || !energy_out || unit_index < 0 || unit_index >= h->N) {
        set_error("hn_unit_energy: bad argument (unit index %d, dimension %d)", unit_index, h->N);
        return HN_E_INVALID;
    }
    HN_TRY(set_device(h));
    HN_TRY(upload_states(h, h->bitsA, states_packed, n));
    HN_TRY(h->esum.reserve(sizeof(double) * (size_t)n));
    unit_energy_kernel<<<(n + 63) / 64, 64, 0, h->st>>>(h->W, h->ldw, h->bitsA.as<uint64_t>(), h->words, h->N, n,
                                                       h->cfg.domain, unit_index, h->esum.as<double>());
    HN_CUDA_TRY(cudaGetLastError());
    HN_CUDA_TRY(cudaMemcpyAsync(energy_out, h->esum.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

// LearnStates over targets [shard_begin, shard_end) of the n given ones (the whole list when no communicator is attached).
// Every rank passes the SAME n states and an identically seeded rng: noise and update orders are drawn for all n targets
// in the reference's order (LearningRule.go:98-104) so the stream equals the single-process one, and each rank keeps its
// shard's.  dW is all-reduced inside the epoch (all_reduce_dw), the all-stable early exit (LearningMethod.go:56-58) is the
// minimum over ranks, so every replica runs the same number of epochs and W stays bitwise identical.
static int learn_states_impl(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t shard_begin, int32_t shard_end,
                             int32_t rule, int32_t method, int32_t epochs, int32_t noise_method, double noise_scale,
                             hn_rng rng, int32_t learn_cap, int32_t* learn_epoch_out, int32_t* learn_index_out,
                             uint8_t* learn_stable_out, double* learn_energy_out, int32_t* n_learn_rows) {
    const bool is_thermal = rule == HN_RULE_THERMAL_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_THERMAL_DELTA;
    const bool is_delta = rule == HN_RULE_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_DELTA || is_thermal;
    const int N = h->N, words = h->words;
    HN_TRY(append_targets(h, states_packed, n));  // HopfieldNetwork.go:258 (every replica holds the whole list)
    // value map the rule sees: bipolarMappedHebbian re-activates with the BINARY manager (sic,
    // LearningRule.go:80); bipolarMappedDelta with the Bipolar manager (:119).
    ValueMap tvm = domain_values(h->cfg.domain);
    if (rule == HN_RULE_BIPOLAR_MAPPED_HEBBIAN) tvm = ValueMap{0.0, 1.0};
    if (rule == HN_RULE_BIPOLAR_MAPPED_DELTA || rule == HN_RULE_BIPOLAR_MAPPED_THERMAL_DELTA) tvm = ValueMap{-1.0, 1.0};
    int produced = 0, last_epoch = 0;
    std::vector<uint64_t> noised;
    std::vector<uint16_t> perms;
    std::vector<double> tmp((size_t)N);
    std::vector<uint8_t> stab;
    std::vector<double> en;

    auto full_set = [&](int cnt, int epoch_offset) -> int {  // LearningMethod.go:38-61 over states[:cnt]
        const int lo = std::min(shard_begin, cnt), hi = std::min(shard_end, cnt), mine = hi - lo;
        const uint64_t* my_states = states_packed + (size_t)lo * words;
        for (int epoch = 0; epoch < epochs; epoch++) {
            if (!is_delta) {
                if (mine > 0) HN_TRY(hebbian_epoch_impl(h, my_states, mine, tvm));
                else HN_TRY(empty_shard_epoch(h, h->use_tc05, 1.0));
            } else {
                noised.assign((size_t)std::max(mine, 1) * words, 0);
                perms.resize((size_t)std::max(mine, 1) * N);
                std::vector<uint16_t> scratch_perm((size_t)N);
                for (int s = 0; s < cnt; s++) {  // LearningRule.go:98-104, draw order: noise then shuffle, ALL targets
                    const bool keep = s >= lo && s < hi;
                    const uint64_t* src = states_packed + (size_t)s * words;
                    for (int i = 0; i < N; i++) tmp[i] = ((src[i >> 6] >> (i & 63)) & 1ULL) ? tvm.hi : tvm.lo;
                    HN_TRY(hn_rng_apply_noise(rng, noise_method, noise_scale, tmp.data(), N));
                    if (keep) {
                        uint64_t* dst = noised.data() + (size_t)(s - lo) * words;
                        for (int i = 0; i < N; i++) if (tmp[i] > 0.0) dst[i >> 6] |= 1ULL << (i & 63);  // network activation
                    }
                    uint16_t* pr = keep ? perms.data() + (size_t)(s - lo) * N : scratch_perm.data();
                    for (int i = 0; i < N; i++) pr[i] = (uint16_t)i;
                    HN_TRY(hn_rng_shuffle_u16(rng, pr, N));
                }
                if (mine > 0) HN_TRY(delta_epoch_impl(h, my_states, noised.data(), perms.data(), mine, nullptr, tvm, is_thermal));
                else HN_TRY(empty_shard_epoch(h, h->use_tc05 && !is_thermal, 0.5));
            }
            // per-target diagnostics (LearningMethod.go:45-53) on the ORIGINAL states in the network domain
            bool all_stable = true;
            if (mine > 0) {
                stab.resize((size_t)mine);
                const bool want_e = learn_energy_out != nullptr && produced < learn_cap;
                if (want_e) en.resize((size_t)mine * N);
                HN_TRY(upload_states(h, h->bitsA, my_states, mine));
                HN_TRY(energy_profiles_impl(h, h->bitsA.as<uint64_t>(), mine, want_e ? en.data() : nullptr, nullptr, stab.data(), nullptr));
                for (int s = 0; s < mine; s++) {
                    if (!stab[s]) all_stable = false;
                    if (produced < learn_cap) {
                        if (learn_epoch_out) learn_epoch_out[produced] = epoch + epoch_offset;
                        if (learn_index_out) learn_index_out[produced] = lo + s;
                        if (learn_stable_out) learn_stable_out[produced] = stab[s];
                        if (want_e) std::memcpy(learn_energy_out + (size_t)produced * N, en.data() + (size_t)s * N, sizeof(double) * N);
                    }
                    produced++;
                }
            }
            last_epoch = epoch + epoch_offset;
            HN_TRY(all_reduce_min_flag(h, &all_stable));   // AllStatesAreStable over EVERY rank's targets
            if (all_stable) break;  // :56-58
        }
        return HN_OK;
    };
    if (method == HN_METHOD_FULL_SET) {
        HN_TRY(full_set(n, 0));
    } else {  // iterativeBatchLearningMethod, LearningMethod.go:70-89 (BATCHSIZE 5)
        const int batches = n / 5;
        int total = 0;
        for (int it = 1; it <= batches; it++) {
            HN_TRY(full_set(5 * it, total));
            total = last_epoch;
        }
    }
    HN_TRY(hn_normalize_frobenius(h, nullptr));  // HopfieldNetwork.go:260 (replicated: identical on every rank)
    if (n_learn_rows) *n_learn_rows = produced;
    return HN_OK;
}

static int learn_states_check(hn_handle h, const char* fn, const uint64_t* states_packed, int32_t n, int32_t rule,
                              int32_t method, int32_t epochs, double noise_scale, hn_rng rng) {
    if (!check_handle(h, fn)) return HN_E_INVALID;
    if (n <= 0 || !states_packed) { set_error("%s: bad argument", fn); return HN_E_INVALID; }
    if (epochs <= 0) {  // HopfieldNetworkBuilder.go:219-221
        set_error("HopfieldNetworkBuilder encountered an error during build! Epochs must be a positive integer!");
        return HN_E_INVALID;
    }
    if (noise_scale < 0.0 || noise_scale > 1.0) {  // :223-225
        set_error("HopfieldNetworkBuilder encountered an error during build! learningNoiseRatio must be in range [0.0, 1.0]!");
        return HN_E_INVALID;
    }
    if (rule < 0 || rule > 5 || (method != HN_METHOD_FULL_SET && method != HN_METHOD_ITERATIVE_BATCH)) {
        set_error("unknown learning rule %d / method %d", rule, method);
        return HN_E_INVALID;
    }
    const bool is_delta = rule >= HN_RULE_DELTA;
    if (is_delta && !rng) { set_error("%s: the delta rule needs an rng", fn); return HN_E_INVALID; }
    return HN_OK;
}

int hn_learn_states(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t rule, int32_t method,
                    int32_t epochs, int32_t noise_method, double noise_scale, hn_rng rng, int32_t learn_cap,
                    int32_t* learn_epoch_out, int32_t* learn_index_out, uint8_t* learn_stable_out,
                    double* learn_energy_out, int32_t* n_learn_rows) {
    HN_TRY(learn_states_check(h, "hn_learn_states", states_packed, n, rule, method, epochs, noise_scale, rng));
    if (h->comm) {   // a sharded replica would all-reduce a dW that already holds every target, once per rank
        set_error("hn_learn_states: a communicator is attached; use hn_learn_states_sharded (or hn_comm_detach first)");
        return HN_E_STATE;
    }
    HN_TRY(set_device(h));
    return learn_states_impl(h, states_packed, n, 0, n, rule, method, epochs, noise_method, noise_scale, rng, learn_cap,
                             learn_epoch_out, learn_index_out, learn_stable_out, learn_energy_out, n_learn_rows);
}

int hn_learn_states_sharded(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t shard_begin, int32_t shard_end,
                            int32_t rule, int32_t method, int32_t epochs, int32_t noise_method, double noise_scale,
                            hn_rng rng, int32_t learn_cap, int32_t* learn_epoch_out, int32_t* learn_index_out,
                            uint8_t* learn_stable_out, double* learn_energy_out, int32_t* n_learn_rows) {
    HN_TRY(learn_states_check(h, "hn_learn_states_sharded", states_packed, n, rule, method, epochs, noise_scale, rng));
    if (shard_begin < 0 || shard_end < shard_begin || shard_end > n) {
        set_error("hn_learn_states_sharded: shard [%d, %d) is not inside [0, %d)", shard_begin, shard_end, n);
        return HN_E_INVALID;
    }
    if (!h->comm && (shard_begin != 0 || shard_end != n)) {
        set_error("hn_learn_states_sharded: a proper shard needs a communicator (hn_comm_attach)");
        return HN_E_STATE;
    }
    HN_TRY(set_device(h));
    return learn_states_impl(h, states_packed, n, shard_begin, shard_end, rule, method, epochs, noise_method, noise_scale,
                             rng, learn_cap, learn_epoch_out, learn_index_out, learn_stable_out, learn_energy_out, n_learn_rows);
}

int hn_relax_batch(hn_handle h, const uint64_t* probes_packed, int32_t n_probes, const uint16_t* perm_pool,
                   int32_t n_pool, const int32_t* offsets, uint32_t flags, hn_relax_out* out) {
    if (!check_handle(h, "hn_relax_batch")) return HN_E_INVALID;
    if (n_probes < 0 || n_pool <= 0 || !perm_pool || !out || (n_probes > 0 && !probes_packed)) {
        set_error("hn_relax_batch: bad argument");
        return HN_E_INVALID;
    }
    if (n_probes == 0) {
        out->n_unique = 0; out->total_flips = out->total_sweeps = out->total_margin_hits = 0; out->device_ms = 0;
        return HN_OK;
    }
    HN_TRY(set_device(h));
    HN_TRY(h->probes.reserve(sizeof(uint64_t) * (size_t)n_probes * h->words));   // uploaded range by range inside relax_device
    HN_TRY(stage_pool(h, perm_pool, n_pool));
    h->resident_B = n_probes; h->resident_pool = n_pool;
    const int32_t* d_off = nullptr;
    if (offsets && !(flags & HN_RELAX_SERIAL_STREAM)) {
        for (int i = 0; i < n_probes; i++)
            if (offsets[i] < 0) { set_error("hn_relax_batch: negative offset"); return HN_E_INVALID; }
        HN_TRY(h->offsets.reserve(sizeof(int32_t) * (size_t)n_probes));
        HN_CUDA_TRY(cudaMemcpyAsync(h->offsets.p, offsets, sizeof(int32_t) * (size_t)n_probes, cudaMemcpyHostToDevice, h->st));
        d_off = h->offsets.as<int32_t>();
    }
    HN_TRY(relax_device(h, h->probes.as<uint64_t>(), n_probes, h->d_pool, h->d_inv, n_pool, d_off,
                        flags, out->energies != nullptr, out->distances != nullptr, probes_packed, out));
    return download_results(h, out, true);
}

int hn_update_states(hn_handle h, uint64_t* states_packed, int32_t n, const uint16_t* perms) {
    if (!check_handle(h, "hn_update_states")) return HN_E_INVALID;
    if (n <= 0 || !states_packed || !perms) { set_error("hn_update_states: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(validate_pool_host(h, perms, n));
    const int words = h->words;
    HN_TRY(upload_states(h, h->bitsB, states_packed, n));
    HN_TRY(sweep_once(h, n, perms));
    HN_CUDA_TRY(cudaMemcpyAsync(states_packed, h->bitsC.p, sizeof(uint64_t) * (size_t)n * words, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int hn_upload_perm_pool(hn_handle h, const uint16_t* perm_pool, int32_t n_pool) {
    if (!check_handle(h, "hn_upload_perm_pool")) return HN_E_INVALID;
    if (n_pool <= 0 || !perm_pool) { set_error("hn_upload_perm_pool: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(stage_pool(h, perm_pool, n_pool));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    h->resident_pool = n_pool;
    return HN_OK;
}
int hn_upload_probes(hn_handle h, const uint64_t* probes_packed, int32_t n_probes) {
    if (!check_handle(h, "hn_upload_probes")) return HN_E_INVALID;
    if (n_probes <= 0 || !probes_packed) { set_error("hn_upload_probes: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(upload_states(h, h->probes, probes_packed, n_probes));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    h->resident_B = n_probes;
    return HN_OK;
}
int hn_generate_probes(hn_handle h, hn_rng r, int32_t count, double rand_min, double rand_max) {
    if (!check_handle(h, "hn_generate_probes")) return HN_E_INVALID;
    if (!r || count <= 0 || !(rand_min < rand_max)) { set_error("hn_generate_probes: bad argument"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    uint64_t table[64][4], state[2];
    rng_jump_table(table);
    rng_get_state(r, state);
    HN_TRY(h->jump.reserve(sizeof(table)));
    HN_CUDA_TRY(cudaMemcpyAsync(h->jump.p, table, sizeof(table), cudaMemcpyHostToDevice, h->st));
    HN_TRY(h->probes.reserve(sizeof(uint64_t) * (size_t)count * h->words));
    const long long total = (long long)count * h->words;
    generate_states_kernel<<<grid_for(total, 256, h->sms), 256, 0, h->st>>>(state[0], state[1], h->jump.as<uint64_t>(), h->N,
                                                                        h->words, count, rand_min, rand_max, h->probes.as<uint64_t>());
    HN_CUDA_TRY(cudaGetLastError());
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));   // `table` is a stack buffer
    rng_advance(r, (uint64_t)count * (uint64_t)h->N);   // the host stream continues after the draws taken on the device
    h->resident_B = count;
    return HN_OK;
}

int hn_download_probes(hn_handle h, uint64_t* probes_packed_out, int32_t n_probes) {
    if (!check_handle(h, "hn_download_probes")) return HN_E_INVALID;
    if (!probes_packed_out || n_probes <= 0 || n_probes > h->resident_B) {
        set_error("hn_download_probes: %d probes asked, %d resident", n_probes, h->resident_B);
        return HN_E_INVALID;
    }
    HN_TRY(set_device(h));
    HN_CUDA_TRY(cudaMemcpyAsync(probes_packed_out, h->probes.p, sizeof(uint64_t) * (size_t)n_probes * h->words,
                                cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int hn_relax_resident(hn_handle h, uint32_t flags, double* device_ms) {
    if (!check_handle(h, "hn_relax_resident")) return HN_E_INVALID;
    if (h->resident_B <= 0 || h->resident_pool <= 0) { set_error("hn_relax_resident: upload probes and a pool first"); return HN_E_STATE; }
    if (flags & (HN_RELAX_SERIAL_STREAM | HN_RELAX_HISTORY)) { set_error("hn_relax_resident: unsupported flag"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    HN_TRY(relax_device(h, h->probes.as<uint64_t>(), h->resident_B, h->d_pool, h->d_inv,
                        h->resident_pool, nullptr, flags, false, true));
    if (device_ms) *device_ms = h->timing.dense_ms + h->timing.fields_ms + h->timing.relax_ms + h->timing.post_ms;
    return HN_OK;
}
int hn_download_results(hn_handle h, hn_relax_out* out) {
    if (!check_handle(h, "hn_download_results") || !out) return HN_E_INVALID;
    HN_TRY(set_device(h));
    return download_results(h, out);
}
int hn_last_relax_timing(hn_handle h, hn_relax_timing* t) {
    if (!check_handle(h, "hn_last_relax_timing") || !t) return HN_E_INVALID;
    *t = h->timing;
    return HN_OK;
}
int hn_export_unique_device(hn_handle h, int64_t index_offset, int64_t* d_first, int32_t* d_hits, uint64_t* d_keys,
                            uint64_t* d_states, uint8_t* d_zero_field, int32_t* n_unique, int32_t* profile_keyed) {
    if (!check_handle(h, "hn_export_unique_device")) return HN_E_INVALID;
    if (!(h->result_flags & HN_RELAX_DEDUP) || h->result_B <= 0) { set_error("hn_export_unique_device: no unique table on the device (HN_RELAX_DEDUP)"); return HN_E_STATE; }
    if (n_unique) *n_unique = h->result_unique;
    if (profile_keyed) *profile_keyed = h->result_profile_keyed ? 1 : 0;
    if (!d_first) return HN_OK;   // size query
    if (!d_hits || !d_keys || !d_states || !d_zero_field) { set_error("hn_export_unique_device: NULL output"); return HN_E_INVALID; }
    HN_TRY(set_device(h));
    const int nu = h->result_unique;
    if (nu > 0) {
        export_unique_kernel<<<grid_for((int64_t)nu * (h->words + 4), 256, h->sms), 256, 0, h->st>>>(
            h->ufirst.as<int32_t>(), h->uhits.as<int32_t>(), nu, h->finals.as<uint64_t>(), h->words, h->hash.as<uint64_t>(),
            h->zf.as<uint8_t>(), (long long)index_offset, reinterpret_cast<long long*>(d_first), d_hits, d_keys, d_states, d_zero_field);
        HN_CUDA_TRY(cudaGetLastError());
    }
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    return HN_OK;
}

int hn_merge_unique_device(hn_handle h, int64_t n, const int64_t* d_first, const int32_t* d_hits, uint64_t* d_keys,
                           const uint64_t* d_states, const uint8_t* d_zero_field, int32_t profile_keyed,
                           int64_t* first_out, int64_t* hits_out, int64_t* rep_out, int64_t* n_groups_out) {
    if (!check_handle(h, "hn_merge_unique_device")) return HN_E_INVALID;
    if (n < 0 || n > 0x7FFFFFFF || !n_groups_out || (n > 0 && (!d_first || !d_hits || !d_keys || !d_states || !d_zero_field || !first_out || !hits_out))) {
        set_error("hn_merge_unique_device: bad argument");
        return HN_E_INVALID;
    }
    *n_groups_out = 0;
    if (n == 0) return HN_OK;
    HN_TRY(set_device(h));
    HN_TRY(h->small.reserve(sizeof(Scratch)));
    HN_CUDA_TRY(cudaMemsetAsync(&h->scratch()->unique_total, 0, 4, h->st));
    // entries arrive in ascending global first index (rank by rank, each rank's table in first-seen order): position
    // order IS first-seen order, exactly what run_dedup's flag + scan produce
    HN_TRY(run_dedup(h, d_states, d_keys, d_zero_field, (int)n, profile_keyed != 0, d_hits, nullptr));
    int32_t nu = 0;
    HN_CUDA_TRY(cudaMemcpyAsync(&nu, &h->scratch()->unique_total, 4, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    HN_TRY(h->partial.reserve(sizeof(long long) * 3 * (size_t)std::max(nu, 1)));
    long long* out = h->partial.as<long long>();
    merge_rows_kernel<<<(nu + 255) / 256, 256, 0, h->st>>>(h->ufirst.as<int32_t>(), h->uhits.as<int32_t>(), nu,
                                                          reinterpret_cast<const long long*>(d_first), out, out + nu, out + 2 * (size_t)nu);
    HN_CUDA_TRY(cudaGetLastError());
    HN_CUDA_TRY(cudaMemcpyAsync(first_out, out, sizeof(long long) * (size_t)nu, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaMemcpyAsync(hits_out, out + nu, sizeof(long long) * (size_t)nu, cudaMemcpyDeviceToHost, h->st));
    if (rep_out) HN_CUDA_TRY(cudaMemcpyAsync(rep_out, out + 2 * (size_t)nu, sizeof(long long) * (size_t)nu, cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    *n_groups_out = nu;
    h->result_flags &= ~HN_RELAX_DEDUP;   // the handle's own table buffers were reused by the merge
    return HN_OK;
}

int hn_debug_integer_fields(hn_handle h, const uint64_t* states_packed, int32_t n, int32_t* fields_out, int32_t* kernel_used) {
    if (!check_handle(h, "hn_debug_integer_fields")) return HN_E_INVALID;
    if (n <= 0 || !states_packed || !fields_out) { set_error("hn_debug_integer_fields: bad argument"); return HN_E_INVALID; }
    if (!integer_stream(h) || !h->w_symmetric) { set_error("hn_debug_integer_fields: the integer stream is not active"); return HN_E_STATE; }
    HN_TRY(set_device(h));
    HN_TRY(upload_states(h, h->bitsA, states_packed, n));
    HN_TRY(h->H.reserve(sizeof(double) * (size_t)n * h->ldw));
    HN_TRY(integer_fields(h, h->bitsA.as<uint64_t>(), n, reinterpret_cast<int32_t*>(h->H.as<double>())));
    HN_CUDA_TRY(cudaMemcpy2DAsync(fields_out, sizeof(int32_t) * h->N, h->H.p, sizeof(int32_t) * h->ldw, sizeof(int32_t) * h->N, n,
                                  cudaMemcpyDeviceToHost, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    if (kernel_used) *kernel_used = (h->use_tc05 && h->k8_valid) ? 1 : 0;
    return HN_OK;
}
int hn_unique_table_stats(hn_handle h, int64_t* verified_merges, int64_t* refuted_merges) {
    if (!check_handle(h, "hn_unique_table_stats")) return HN_E_INVALID;
    if (verified_merges) *verified_merges = h->verified_pairs;
    if (refuted_merges) *refuted_merges = h->refuted_pairs;
    return HN_OK;
}
int hn_last_learn_timing(hn_handle h, hn_learn_timing* t) {
    if (!check_handle(h, "hn_last_learn_timing") || !t) return HN_E_INVALID;
    *t = h->ltiming;
    return HN_OK;
}
int hn_get_margin(hn_handle h, double* tau_base, double* tau_per_flip) {
    if (!check_handle(h, "hn_get_margin")) return HN_E_INVALID;
    HN_TRY(set_device(h));
    HN_TRY(ensure_margin(h));
    if (tau_base) *tau_base = h->tau_base;
    if (tau_per_flip) *tau_per_flip = h->tau_flip;
    return HN_OK;
}

int hn_host_alloc(void** ptr_out, size_t bytes) {
    if (!ptr_out || bytes == 0) { set_error("hn_host_alloc: bad argument"); return HN_E_INVALID; }
    *ptr_out = nullptr;
    const cudaError_t e = cudaHostAlloc(ptr_out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error("hn_host_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? HN_E_OOM : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? HN_E_NO_DEVICE : HN_E_CUDA);
    }
    return HN_OK;
}
int hn_host_free(void* ptr) {
    if (ptr) HN_CUDA_TRY(cudaFreeHost(ptr));
    return HN_OK;
}

int hn_pack_states_f64(int32_t dimension, int32_t n, const double* states_f64, uint64_t* packed) {
    if (dimension <= 0 || n < 0 || !states_f64 || !packed) { set_error("hn_pack_states_f64: bad argument"); return HN_E_INVALID; }
    const int words = hn_words(dimension);
    std::memset(packed, 0, sizeof(uint64_t) * (size_t)n * words);
    for (int s = 0; s < n; s++)
        for (int i = 0; i < dimension; i++)
            if (states_f64[(size_t)s * dimension + i] > 0.0) packed[(size_t)s * words + (i >> 6)] |= 1ULL << (i & 63);
    return HN_OK;
}
int hn_unpack_states_f64(int32_t dimension, int32_t domain, int32_t n, const uint64_t* packed, double* states_f64) {
    if (dimension <= 0 || n < 0 || !states_f64 || !packed) { set_error("hn_unpack_states_f64: bad argument"); return HN_E_INVALID; }
    const int words = hn_words(dimension);
    const double lo = domain == HN_DOMAIN_BINARY ? 0.0 : -1.0;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < dimension; i++)
            states_f64[(size_t)s * dimension + i] = ((packed[(size_t)s * words + (i >> 6)] >> (i & 63)) & 1ULL) ? 1.0 : lo;
    return HN_OK;
}

int hn_comm_unique_id(uint8_t id[HN_NCCL_UNIQUE_ID_BYTES]) {
    if (!id) return HN_E_INVALID;
    HN_TRY(load_nccl());
    HN_NCCL_TRY(g_nccl.GetUniqueId(id));
    return HN_OK;
}
int hn_comm_attach(hn_handle h, const uint8_t id[HN_NCCL_UNIQUE_ID_BYTES], int32_t rank, int32_t n_ranks) {
    if (!check_handle(h, "hn_comm_attach") || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) { set_error("hn_comm_attach: bad argument"); return HN_E_INVALID; }
    HN_TRY(load_nccl());
    HN_TRY(set_device(h));
    if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    Id128 uid;
    std::memcpy(uid.bytes, id, 128);
    HN_NCCL_TRY(g_nccl.CommInitRank(&h->comm, n_ranks, uid, rank));
    h->rank_id = rank; h->n_ranks = n_ranks;
    return HN_OK;
}
int hn_comm_detach(hn_handle h) {
    if (!check_handle(h, "hn_comm_detach")) return HN_E_INVALID;
    if (h->comm) { HN_TRY(set_device(h)); cudaStreamSynchronize(h->st); g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    h->rank_id = 0; h->n_ranks = 1;
    return HN_OK;
}
int hn_broadcast_weights(hn_handle h, int32_t root) {
    if (!check_handle(h, "hn_broadcast_weights")) return HN_E_INVALID;
    if (!h->comm) { set_error("hn_broadcast_weights: no communicator attached"); return HN_E_STATE; }
    HN_TRY(set_device(h));
    HN_NCCL_TRY(g_nccl.Broadcast(h->W, h->W, (size_t)h->N * h->ldw, NCCL_FLOAT64, root, h->comm, h->st));
    HN_CUDA_TRY(cudaStreamSynchronize(h->st));
    weights_changed(h);
    HN_TRY(detect_symmetry(h));
    return detect_integer_structure(h);
}

}  // extern "C"
