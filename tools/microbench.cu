// microbench.cu — pins the sm_100a facts the kernel design in DESIGN.md rests on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu && tools/microbench
// Prints: per-SM per-clock lane throughput of the fp64 pipe (DFMA/DADD/DMUL/min/compare), whether
// fp32 / integer / MUFU / conversion / shuffle instructions issue beside it, and the accuracy of
// the MUFU 64-bit reciprocal / rsqrt seeds.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

enum Op { DFMA_ONLY, DADD_ONLY, DMUL_ONLY, DMIN_ONLY, DSETP_SEL, DFMA_FFMA, DFMA_IMAD, DFMA_MUFU, DFMA_F2F, DFMA_SHFL, DFMA_LDS,
          FFMA_ONLY, MUFU64_ONLY, F2F_ONLY, DFMA_DEP1, DFMA_DEP2, DFMA_DEP4 };

__device__ __forceinline__ double mufu_rsqrt64(double x) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double mufu_rcp64(double x) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }

template <int OP>
__global__ void __launch_bounds__(256) k_tput(double* out, double seed, int iters) {
    double a[CH];
    float f[CH];
    int n[CH];
    __shared__ double sm[256 * 2];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed * 2;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c) { a[c] = seed + c + threadIdx.x * 1e-3; f[c] = (float)a[c]; n[c] = c + threadIdx.x; }
    const double m = 1.0000001, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == DFMA_ONLY || OP >= DFMA_FFMA && OP <= DFMA_LDS) a[c] = fma(a[c], m, b);
            if (OP == DADD_ONLY) a[c] = a[c] + b;
            if (OP == DMUL_ONLY) a[c] = a[c] * m;
            if (OP == DMIN_ONLY) a[c] = fmin(a[c], sm[(threadIdx.x + c + i) & 511]);
            if (OP == DSETP_SEL) a[c] = (a[c] < seed * i) ? a[c] + b : a[c];
            if (OP == DFMA_FFMA || OP == FFMA_ONLY) f[c] = fmaf(f[c], 1.0000001f, 1e-9f);
            if (OP == DFMA_IMAD) n[c] = n[c] * 3 + (n[c] >> 3);
            if (OP == DFMA_MUFU && (c & 3) == 0) f[c] = rsqrtf(f[c] + 1.0f);
            if (OP == DFMA_F2F && (c & 3) == 0) f[c] += (float)a[c];
            if (OP == DFMA_SHFL && (c & 3) == 0) n[c] += __shfl_xor_sync(0xffffffffu, n[c], 1);
            if (OP == DFMA_LDS && (c & 3) == 0) a[c] += sm[(threadIdx.x * 3 + i + c) & 511];
            if (OP == MUFU64_ONLY) a[c] = mufu_rsqrt64(a[c]) + 1.0;
            if (OP == F2F_ONLY) { f[c] = (float)a[c] + 1.0f; a[c] = (double)f[c]; }
        }
        if (OP == DFMA_DEP1) { a[0] = fma(a[0], m, b); }
        if (OP == DFMA_DEP2) { a[0] = fma(a[0], m, b); a[1] = fma(a[1], m, b); }
        if (OP == DFMA_DEP4) { a[0] = fma(a[0], m, b); a[1] = fma(a[1], m, b); a[2] = fma(a[2], m, b); a[3] = fma(a[3], m, b); }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += a[c] + f[c] + n[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double ops_per_iter_per_thread, int warps_per_sm, int nsm, double* d_out, double clk_hz) {
    int threads = 256;
    int blocks_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
    if (warps_per_sm * 32 < threads) { threads = warps_per_sm * 32; blocks_per_sm = 1; }
    int blocks = nsm * blocks_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_tput<OP><<<blocks, threads>>>(d_out, 1.0, ITERS);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        k_tput<OP><<<blocks, threads>>>(d_out, 1.0, ITERS);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double total = ops_per_iter_per_thread * ITERS * (double)blocks * threads;
    double per_s = total / (best * 1e-3);
    printf("%-34s warps/SM=%2d  %8.3f ms  %8.2f Gop/s  %6.2f lane-ops/clk/SM (at %.0f MHz)\n", name, warps_per_sm, best, per_s / 1e9,
           per_s / nsm / clk_hz, clk_hz / 1e6);
}

// fp64 tensor-core MMA (DMMA m8n8k4: 256 FMA per warp instruction = 8 per lane): own pipe or the vector fp64 pipe?
// MIX = 0: DMMA only; MIX = 1: one DFMA per DMMA beside it (8 + 1 lane-FMAs); MIX = 2: four DFMA per DMMA (8 + 4).
template <int MIX>
__global__ void __launch_bounds__(256) k_dmma(double* out, double seed, int iters) {
    double c0[4], c1[4], f[4];
    const double a = seed * 1.0000001, b = 1e-3 + threadIdx.x * 1e-9;
#pragma unroll
    for (int j = 0; j < 4; ++j) { c0[j] = seed + j; c1[j] = seed - j; f[j] = seed + 0.5 * j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(a), "d"(b));
            if (MIX == 1) f[j] = fma(f[j], 1.0000001, 1e-9);
            if (MIX == 2) {
                f[0] = fma(f[0], 1.0000001, 1e-9); f[1] = fma(f[1], 1.0000001, 1e-9);
                f[2] = fma(f[2], 1.0000001, 1e-9); f[3] = fma(f[3], 1.0000001, 1e-9);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c0[j] + c1[j] + f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
void run_dmma(const char* name, int warps_per_sm, int nsm, double* d_out, double clk_hz) {
    const int threads = 256, blocks = nsm * (warps_per_sm * 32 / threads);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_dmma<MIX><<<blocks, threads>>>(d_out, 1.0, ITERS);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        k_dmma<MIX><<<blocks, threads>>>(d_out, 1.0, ITERS);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double lanes = (double)blocks * threads * ITERS * 4;              // DMMA issued per lane
    const double mma_fma = 8.0 * lanes / (best * 1e-3), vec_fma = (MIX == 1 ? 1.0 : MIX == 2 ? 4.0 : 0.0) * lanes / (best * 1e-3);
    printf("%-34s warps/SM=%2d  %8.3f ms  tensor %7.2f + vector %6.2f FMA lane-ops/clk/SM\n", name, warps_per_sm, best,
           mma_fma / nsm / clk_hz, vec_fma / nsm / clk_hz);
}

__global__ void k_seed_accuracy(const double* x, double* rs, double* rc, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { rs[i] = mufu_rsqrt64(x[i]); rc[i] = mufu_rcp64(x[i]); }
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double clk = clk_khz * 1e3;
    printf("device %s  sm_%d%d  SMs=%d  clock(attr)=%.0f MHz  smem/SM=%zu KB  regs/SM=%d  L2=%d MB\n", p.name, p.major, p.minor,
           p.multiProcessorCount, clk / 1e6, p.sharedMemPerMultiprocessor / 1024, p.regsPerMultiprocessor, p.l2CacheSize >> 20);
    int nsm = p.multiProcessorCount;
    double* d_out;
    CK(cudaMalloc(&d_out, sizeof(double) * 256 * 8 * nsm * 4));
    run<DFMA_ONLY>("DFMA x8 chains", CH, 32, nsm, d_out, clk);
    run<DFMA_ONLY>("DFMA x8 chains", CH, 16, nsm, d_out, clk);
    run<DFMA_ONLY>("DFMA x8 chains", CH, 8, nsm, d_out, clk);
    run<DFMA_ONLY>("DFMA x8 chains", CH, 4, nsm, d_out, clk);
    run<DFMA_DEP1>("DFMA 1 dependent chain", 1, 4, nsm, d_out, clk);
    run<DFMA_DEP2>("DFMA 2 chains", 2, 4, nsm, d_out, clk);
    run<DFMA_DEP4>("DFMA 4 chains", 4, 4, nsm, d_out, clk);
    run<DFMA_DEP4>("DFMA 4 chains", 4, 8, nsm, d_out, clk);
    run<DFMA_DEP2>("DFMA 2 chains", 2, 8, nsm, d_out, clk);
    run<DADD_ONLY>("DADD", CH, 16, nsm, d_out, clk);
    run<DMUL_ONLY>("DMUL", CH, 16, nsm, d_out, clk);
    run<DMIN_ONLY>("fmin(double) + LDS", CH, 16, nsm, d_out, clk);
    run<DSETP_SEL>("DSETP + predicated DADD", CH, 16, nsm, d_out, clk);
    run<FFMA_ONLY>("FFMA", CH, 16, nsm, d_out, clk);
    run<DFMA_FFMA>("DFMA + FFMA 1:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<DFMA_IMAD>("DFMA + 2 int ops 1:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<DFMA_MUFU>("DFMA + MUFU.RSQ 4:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<DFMA_F2F>("DFMA + F2F.F32.F64 4:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<DFMA_SHFL>("DFMA + SHFL 4:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<DFMA_LDS>("DFMA + LDS.64 4:1 (count DFMA)", CH, 16, nsm, d_out, clk);
    run<MUFU64_ONLY>("MUFU.RSQ64H + DADD", CH, 16, nsm, d_out, clk);
    run<F2F_ONLY>("F2F f64->f32->f64 pair + FADD", CH, 16, nsm, d_out, clk);

    run_dmma<0>("DMMA m8n8k4 only", 16, nsm, d_out, clk);
    run_dmma<0>("DMMA m8n8k4 only", 8, nsm, d_out, clk);
    run_dmma<1>("DMMA + 1 DFMA", 16, nsm, d_out, clk);
    run_dmma<2>("DMMA + 4 DFMA", 16, nsm, d_out, clk);

    // seed accuracy
    const int N = 1 << 20;
    double* hx = (double*)malloc(sizeof(double) * N);
    double *dx, *drs, *drc;
    srand(1);
    for (int i = 0; i < N; ++i) {
        double u = (rand() + 1.0) / ((double)RAND_MAX + 2.0);
        double e = (rand() % 80) - 40;
        hx[i] = (1.0 + u) * pow(2.0, e);
    }
    CK(cudaMalloc(&dx, sizeof(double) * N)); CK(cudaMalloc(&drs, sizeof(double) * N)); CK(cudaMalloc(&drc, sizeof(double) * N));
    CK(cudaMemcpy(dx, hx, sizeof(double) * N, cudaMemcpyHostToDevice));
    k_seed_accuracy<<<N / 256, 256>>>(dx, drs, drc, N);
    double* hrs = (double*)malloc(sizeof(double) * N);
    double* hrc = (double*)malloc(sizeof(double) * N);
    CK(cudaMemcpy(hrs, drs, sizeof(double) * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hrc, drc, sizeof(double) * N, cudaMemcpyDeviceToHost));
    double ers = 0, erc = 0;
    for (int i = 0; i < N; ++i) {
        double t1 = 1.0 / sqrt(hx[i]), t2 = 1.0 / hx[i];
        ers = fmax(ers, fabs(hrs[i] - t1) / t1);
        erc = fmax(erc, fabs(hrc[i] - t2) / t2);
    }
    printf("MUFU.RSQ64H seed max rel err = %.3e (%.1f bits)   MUFU.RCP64H seed max rel err = %.3e (%.1f bits)\n", ers, -log2(ers), erc, -log2(erc));
    return 0;
}
