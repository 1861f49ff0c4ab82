// Issue rate of the FP32 / ALU instructions the epilogues are made of, on one SM: cycles per warp-instruction per
// SM sub-partition for N warps per sub-partition, 8 independent chains per thread.   nvcc -arch=sm_100a -O3 --fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); }                                  // FFMA (3 registers)
            if (OP == 1) { a[i] = __ffma2_rn(a[i], m, c); }                                          // FFMA2
            if (OP == 2) { a[i] = __fadd2_rn(a[i], c); }                                             // FADD2
            if (OP == 3) { a[i] = __fmul2_rn(a[i], m); }                                             // FMUL2
            if (OP == 4) { a[i].x = fminf(fmaxf(a[i].x, c.x), 1e30f); }                              // 2 x FMNMX
            if (OP == 5) { a[i].x = __int2float_rn(__float_as_int(a[i].x) - 7); }                    // IADD + I2FP
            if (OP == 6) { a[i].x = __uint_as_float(__byte_perm(__float_as_uint(a[i].x), 0x4B000000u, 0x7650)); }  // PRMT
            if (OP == 7) { a[i].x = __fadd_rn(a[i].x, 12582912.0f); }                                // FADD imm
            if (OP == 8) { a[i].x = __fmul_rn(a[i].x, m.x); }                                        // FMUL
            if (OP == 9) { a[i] = __fadd2_rn(a[i], make_float2(12582912.0f, 12582912.0f)); }         // FADD2 imm
            if (OP == 10) { a[i].x = __fmaf_rn(a[i].x, 1.0001f, 0.5f); }                             // FFMA imm
            if (OP == 11) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i].x)); a[i].x = r; }   // MUFU.EX2
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    printf("%-14s", name);
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int threads = warps_per_smsp * 4 * 32;
        k<OP><<<1, threads>>>(out, cyc, iters);
        k<OP><<<1, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // warp-instructions issued per sub-partition = iters * 8 * per_iter * warps_per_smsp
        printf("  %dw: %5.2f", warps_per_smsp, (double)c / ((double)iters * 8 * per_iter * warps_per_smsp));
    }
    printf("   cycles per warp-instruction per sub-partition\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA", 1); run<10>("FFMA imm", 1); run<1>("FFMA2", 1); run<2>("FADD2", 1); run<9>("FADD2 imm", 1); run<3>("FMUL2", 1);
    run<8>("FMUL", 1); run<7>("FADD imm", 1); run<4>("FMNMX x2", 2); run<5>("IADD+I2FP", 2); run<6>("PRMT", 1); run<11>("MUFU.EX2", 1);
    return 0;
}
