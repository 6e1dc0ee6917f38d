// Measurement helper (not part of the library): issue rate of the integer instructions the int16 kernels lean on,
// in thread-instructions per clock per SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void rate_kernel(uint32_t* out, uint32_t a0, uint32_t b0, long long* cycles)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = a0 + threadIdx.x * 7 + i;
    const uint32_t b = b0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(b), "r"(a0));
            if (OP == 1) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(a0));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (OP == 4) asm volatile("shr.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b & 3));
            if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(v[i]) : "r"(b));
            if (OP == 6) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(b), "r"(a0));
            if (OP == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(a0));
            if (OP == 8) asm volatile("vadd2.s32.s32.s32.sat %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(a0));
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int threads)
{
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    rate_kernel<OP><<<148, threads>>>(out, 3, 0x01ff, cyc);
    rate_kernel<OP><<<148, threads>>>(out, 3, 0x01ff, cyc);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s threads/SM %4d  %.1f thread-instr/clk/SM\n", name, threads, (double)threads * ITERS * CHAINS / (double)h);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    for (int threads : {512, 1024}) {
        run<0>("dp2a", threads);
        run<6>("dp4a", threads);
        run<1>("mul.hi", threads);
        run<2>("mad.lo", threads);
        run<3>("add", threads);
        run<4>("shr", threads);
        run<5>("prmt", threads);
        run<7>("lop3", threads);
        run<8>("vadd2", threads);
    }
    return cudaGetLastError() != cudaSuccess;
}
