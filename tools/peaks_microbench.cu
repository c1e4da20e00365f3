// tools/peaks_microbench.cu — measured denominators for the rooflines bench.py reports ("report the roofline
// against the resource that binds").  Own kernels only; prints ONE JSON object (-> profiles/r02_peaks.json).
//   l2_read      : L2 -> SM read bandwidth with the relaxation kernel's access shape (one coalesced 128-bit load per
//                  thread and row, rows of 2N / 8N bytes picked pseudo-randomly from a matrix of 32 MiB / 128 MiB), every
//                  CTA on its own row sequence ("independent") or all CTAs on the same sequence ("lockstep")
//   tma_stream   : the same matrix streamed by TMA (box 32 rows x 128 B, 128-byte swizzle) into a shared-memory ring
//   tcgen05_i8   : tcgen05.mma.kind::i8 issue rate on resident operand tiles, M=128 with N = 256 / 64 / 32 / 16
//   imma_sync    : mma.sync.m16n8k32.s8 rate (the legacy integer tensor path)
//   dmma_sync    : mma.sync.m8n8k4.f64 rate (the only FP64 tensor path on sm_100)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/peaks tools/peaks_microbench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../hopfield_network_go_b200/csrc/tc05.cuh"

using namespace hn::tc05;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

// ---- L2 -> SM by plain loads -------------------------------------------------------------------------------------------
// Every CTA (128 threads) reads `rows_per_cta` rows of `row_bytes` (one uint4 per thread and 2 KiB of row); the row index
// follows an LCG; lockstep = every CTA uses the same seed.
__global__ void __launch_bounds__(128) l2_rows_kernel(const uint4* __restrict__ mat, int n_rows, int row_vec, int rows_per_cta,
                                                     int lockstep, uint4* __restrict__ sink) {
    uint32_t s = lockstep ? 12345u : 12345u + 7919u * blockIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < rows_per_cta; r++) {
        s = s * 1664525u + 1013904223u;
        const uint4* row = mat + (size_t)((s >> 8) % (uint32_t)n_rows) * row_vec;
        for (int v = threadIdx.x; v < row_vec; v += 128) {
            const uint4 x = __ldg(row + v);
            acc.x ^= x.x; acc.y += x.y; acc.z ^= x.z; acc.w += x.w;
        }
    }
    if (acc.x == 0x12345678u && acc.y == 42u) sink[blockIdx.x * 128 + threadIdx.x] = acc;   // never true: keeps the loads
}

// ---- L2 -> SM by TMA into a shared-memory ring ----------------------------------------------------------------------------
constexpr int TS_STAGES = 8, TS_BOX_ROWS = 32, TS_BYTES = TS_BOX_ROWS * 128;
__global__ void __launch_bounds__(64) tma_stream_kernel(const __grid_constant__ CUtensorMap map, int n_row_blocks, int k_chunks,
                                                       int passes, int lockstep, uint32_t* __restrict__ sink) {
    extern __shared__ unsigned char ts_raw[];
    const uint32_t raw = smem_u32(ts_raw);
    unsigned char* smem = ts_raw + (((raw + 1023u) & ~1023u) - raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TS_STAGES * TS_BYTES);
    uint64_t* empty = full + TS_STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    const int total = n_row_blocks * k_chunks * passes;
    const int start = lockstep ? 0 : (int)((blockIdx.x * 2654435761u) % (uint32_t)(n_row_blocks * k_chunks));
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < total; i++) {
            const int t = (start + i) % (n_row_blocks * k_chunks);
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], TS_BYTES);
            tma_load_2d(smem + stage * TS_BYTES, &map, (t % k_chunks) * 128, (t / k_chunks) * TS_BOX_ROWS, &full[stage]);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0; uint32_t x = 0;
        for (int i = 0; i < total; i++) {
            mbar_wait(&full[stage], phase);
            x ^= *reinterpret_cast<volatile uint32_t*>(smem + stage * TS_BYTES);
            mbar_arrive(&empty[stage]);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        }
        if (x == 0x12345678u) sink[blockIdx.x] = x;
    }
}

// ---- tcgen05.mma.kind::i8 on resident tiles ---------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) tc_i8_kernel(int iters, uint32_t* __restrict__ sink) {
    extern __shared__ unsigned char tc_raw[];
    const uint32_t raw = smem_u32(tc_raw);
    unsigned char* smem = tc_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* sA = smem;                   // 128 rows x 128 B
    unsigned char* sB = smem + 128 * 128;       // N rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 128 + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01FF01FFu;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_i8(128, N, true, true);
        const uint64_t da = smem_desc_k128(smem_u32(sA)), db = smem_desc_k128(smem_u32(sB));
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) mma_i8(tm + (uint32_t)((i & 1) * 256), da + 2 * j, db + 2 * j, idesc, 1u);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    uint32_t r[32];
    tmem_ld_32x32(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), r);
    tmem_ld_wait();
    if (r[0] == 0x7FFFFFFFu) sink[blockIdx.x] = r[1];
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// ---- legacy tensor paths ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) imma_kernel(int iters, int* __restrict__ sink) {
    int c[4][4] = {};
    uint32_t a[4] = {threadIdx.x, 0x01010101u, 0xFF01FF01u, threadIdx.x * 3u}, b[2] = {0x01FF01FFu, threadIdx.x};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[k][0]), "+r"(c[k][1]), "+r"(c[k][2]), "+r"(c[k][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
    for (int k = 0; k < 4; k++) for (int e = 0; e < 4; e++) s += c[k][e];
    if (s == 0x7FFFFFFF) sink[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dmma_kernel(int iters, double* __restrict__ sink) {
    double c[4][2] = {};
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 4; k++) s += c[k][0] + c[k][1];
    if (s == 1.2345) sink[blockIdx.x * 256 + threadIdx.x] = s;
}

template <class F> static double time_ms(F&& launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();   // warm-up
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    void* sink = nullptr; CK(cudaMalloc(&sink, 1 << 24));
    std::string out = "{";
    char buf[512];
    snprintf(buf, sizeof(buf), "\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d", prop.name, sms, prop.l2CacheSize); out += buf;

    // L2 -> SM, plain loads
    out += ", \"l2_read\": [";
    bool first = true;
    for (size_t mib : {32, 128}) {
        const int row_bytes = mib == 32 ? 8192 : 32768;     // int16 / float64 column of a 4096-unit network
        const int n_rows = (int)((mib << 20) / row_bytes);
        uint4* mat = nullptr; CK(cudaMalloc(&mat, mib << 20)); CK(cudaMemset(mat, 1, mib << 20));
        for (int lock = 0; lock < 2; lock++)
            for (int occ : {4, 9, 16}) {
                const int rows_per_cta = 4096, grid = sms * occ;
                const double ms = time_ms([&] { l2_rows_kernel<<<grid, 128>>>(mat, n_rows, row_bytes / 16, rows_per_cta, lock, (uint4*)sink); });
                const double gbs = (double)grid * rows_per_cta * row_bytes / (ms * 1e-3) / 1e9;
                snprintf(buf, sizeof(buf), "%s{\"matrix_mib\": %zu, \"row_bytes\": %d, \"ctas_per_sm\": %d, \"lockstep\": %d, \"gbs\": %.1f}",
                         first ? "" : ", ", mib, row_bytes, occ, lock, gbs);
                out += buf; first = false;
            }
        CK(cudaFree(mat));
    }
    out += "]";

    // L2 -> SM, TMA stream of a 16 MiB byte matrix [4096][4096]
    {
        const int n = 4096;
        unsigned char* m8 = nullptr; CK(cudaMalloc(&m8, (size_t)n * n)); CK(cudaMemset(m8, 1, (size_t)n * n));
        CUtensorMap map;
        if (make_map_bytes_2d(&map, m8, n, n, n, TS_BOX_ROWS) != 0) { fprintf(stderr, "tensor map failed\n"); return 3; }
        const size_t smem = TS_STAGES * TS_BYTES + 256 + 1024;
        CK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        out += ", \"tma_stream\": [";
        first = true;
        for (int lock = 0; lock < 2; lock++)
            for (int occ : {1, 2, 4}) {
                const int grid = sms * occ, passes = 2;
                const double ms = time_ms([&] { tma_stream_kernel<<<grid, 64, smem>>>(map, n / TS_BOX_ROWS, n / 128, passes, lock, (uint32_t*)sink); });
                const double gbs = (double)grid * passes * n * n / (ms * 1e-3) / 1e9;
                snprintf(buf, sizeof(buf), "%s{\"matrix_mib\": 16, \"box\": \"32x128B\", \"ctas_per_sm\": %d, \"lockstep\": %d, \"gbs\": %.1f}",
                         first ? "" : ", ", occ, lock, gbs);
                out += buf; first = false;
            }
        out += "]";
        CK(cudaFree(m8));
    }

    // tcgen05 int8
    {
        const size_t smem = 128 * 128 + 256 * 128 + 64 + 1024;
        const int iters = 20000;
        out += ", \"tcgen05_i8\": [";
        auto run = [&](auto kern, int n, bool last) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const double ms = time_ms([&] { kern<<<sms, 128, smem>>>(iters, (uint32_t*)sink); });
            const double ops = 2.0 * 128 * n * 32 * 4.0 * iters * sms;
            snprintf(buf, sizeof(buf), "{\"m\": 128, \"n\": %d, \"k\": 32, \"tops\": %.1f, \"clk_per_mma_at_1965mhz\": %.2f}%s", n,
                     ops / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (4.0 * iters), last ? "" : ", ");
            out += buf;
        };
        run(tc_i8_kernel<256>, 256, false);
        run(tc_i8_kernel<128>, 128, false);
        run(tc_i8_kernel<64>, 64, false);
        run(tc_i8_kernel<32>, 32, false);
        run(tc_i8_kernel<16>, 16, true);
        out += "]";
    }
    {
        const int iters = 20000;
        double ms = time_ms([&] { imma_kernel<<<sms * 8, 256>>>(iters, (int*)sink); });
        snprintf(buf, sizeof(buf), ", \"imma_sync_tops\": %.1f", 2.0 * 16 * 8 * 32 * 4.0 * iters * 8 * sms * 8 / (ms * 1e-3) / 1e12); out += buf;
        ms = time_ms([&] { dmma_kernel<<<sms * 8, 256>>>(iters, (double*)sink); });
        snprintf(buf, sizeof(buf), ", \"dmma_sync_tflops\": %.2f", 2.0 * 8 * 8 * 4 * 4.0 * iters * 8 * sms * 8 / (ms * 1e-3) / 1e12); out += buf;
    }
    out += "}";
    printf("%s\n", out.c_str());
    return 0;
}
