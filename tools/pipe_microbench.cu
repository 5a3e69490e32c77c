// Pipe-rate microbenchmark for sm_100a: how fast are the instruction classes the GRU
// recurrence kernels are built from?  (legacy mma.sync f16 / tf32, FFMA, MUFU ex2 / rcp)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu
// Prints ops per clock per SM for 4, 8 and 16 resident warps per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096

__global__ void k_hmma(float* out, unsigned a0, unsigned b0) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_tf32(float* out, unsigned a0, unsigned b0) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, float x, float y) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], x, y);
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mufu(float* out, float x) {
    float c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(c[i]));
            asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(c[i]) : "f"(e));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
    printf("%s, %d SMs, nominal %d MHz (rates below assume the clock measured by FFMA = 128/clk/SM)\n", p.name, sms, khz / 1000);
    for (int warps : {4, 8, 16, 32}) {
        int thr = warps * 32;
        float ms_f = time_ms([&] { k_ffma<<<sms, thr>>>(out, 1.0001f, 0.5f); });
        float ms_h = time_ms([&] { k_hmma<<<sms, thr>>>(out, 0x3c003c00u, 0x3c003c00u); });
        float ms_t = time_ms([&] { k_tf32<<<sms, thr>>>(out, 0x3f800000u, 0x3f800000u); });
        float ms_m = time_ms([&] { k_mufu<<<sms, thr>>>(out, 0.3f); });
        double ffma = (double)ITERS * 16 * thr / (ms_f * 1e-3);           // FMA lanes / s / SM
        double clk = ffma / 128.0;                                        // implied clock if FFMA saturates
        double hmma = (double)ITERS * 8 * warps * (16 * 8 * 16) / (ms_h * 1e-3);   // MAC / s / SM
        double tf32 = (double)ITERS * 8 * warps * (16 * 8 * 8) / (ms_t * 1e-3);
        double mufu = (double)ITERS * 8 * 2 * thr / (ms_m * 1e-3);
        printf("warps/SM=%2d  FFMA %.1f GFMA/s/SM (implied clk %.0f MHz) | HMMA.f16 %.0f MAC/clk/SM | TF32 %.0f MAC/clk/SM | MUFU %.1f ops/clk/SM\n",
               warps, ffma / 1e9, clk / 1e6, hmma / clk, tf32 / clk, mufu / clk);
    }
    return 0;
}
