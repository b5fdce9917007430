// Microbenchmark behind the scatter design choice (DESIGN.md "scatter"): throughput of
// shared-memory atomicMax vs plain store + read-back on sm_100a, conflict-free and 4-way conflicted.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_scatter smem_scatter.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 2048;

template <int MODE, int STRIDE>
__global__ void __launch_bounds__(256) k(unsigned *out, int iters) {
    __shared__ unsigned keys[W * 4];
    for (int i = threadIdx.x; i < W * 4; i += 256) keys[i] = 0;
    __syncthreads();
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int x = ((threadIdx.x >> 5) * 8 + j) * 32 + (threadIdx.x & 31);       // interleaved pixels
            int idx = ((x + (it & 7)) & (W - 1)) * STRIDE;
            unsigned key = (it + j) & 63;
            if (MODE == 0) atomicMax(&keys[idx], key);                          // ATOMS.MAX
            else if (MODE == 1) keys[idx] = key;                                // STS
            else { keys[idx] = key; acc += keys[(idx + 32 * STRIDE) & (W * 4 - 1)]; }   // STS + LDS
        }
    }
    __syncthreads();
    out[blockIdx.x * 256 + threadIdx.x] = acc + keys[threadIdx.x];
}

template <int MODE, int STRIDE>
void run(const char *name) {
    unsigned *out;
    int sms = 148, occ = 4, iters = 2000;
    cudaMalloc(&out, sms * occ * 256 * 4);
    k<MODE, STRIDE><<<sms * occ, 256>>>(out, 10);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE, STRIDE><<<sms * occ, 256>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double warp_instr_per_sm = (double)occ * 8 /*warps*/ * 8 * iters * (MODE == 2 ? 2 : 1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %6.2f ns per warp-instr per SM  (~%.2f cyc @%d MHz nominal)\n", name, ms,
           ms * 1e6 / warp_instr_per_sm, ms * 1e-3 * clk * 1e3 / warp_instr_per_sm, clk / 1000);
    cudaFree(out);
}

int main() {
    run<0, 1>("atomicMax conflict-free");
    run<0, 4>("atomicMax 4-way conflict");
    run<1, 1>("plain STS conflict-free");
    run<1, 4>("plain STS 4-way conflict");
    run<2, 1>("STS + LDS conflict-free");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
