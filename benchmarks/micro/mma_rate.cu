// mma_rate.cu -- microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, K=8, operands in
// shared memory, 128B swizzle) as a function of N and of how many independent TMEM
// accumulators consecutive MMAs rotate over.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -I pvnet_b200/csrc -o gpurun_out/mma_rate benchmarks/micro/mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

// STYLE 0: one thread (lane 0 inside a divergent branch) issues; STYLE 1: the whole warp runs the
// loop converged and an elected lane issues (operands can live in uniform registers).
template <int N, int GROUPS, int STYLE>
__global__ void __launch_bounds__(128, 1) k_rate(int reps, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        ptx::tmem_alloc(&slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::make_idesc_tf32(128, N);
        const uint64_t a = ptx::make_kmajor_desc(ptx::smem_u32(smem), 128);
        const uint64_t b = ptx::make_kmajor_desc(ptx::smem_u32(smem + 16384), 128);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (STYLE == 0) {
            if (threadIdx.x == 0) {
                t0 = clock64();
                for (int r = 0; r < reps; r += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::mma_tf32_ss(tmem + (uint32_t)((k % GROUPS) * N), a + 2 * k, b + 2 * k, idesc, 1);
                }
                t1 = clock64();
                ptx::mma_commit(&bar);
                ptx::mbar_wait(&bar, 0);
                t2 = clock64();
            }
        } else {
            t0 = clock64();
            for (int r = 0; r < reps; r += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (ptx::elect_one())
                        ptx::mma_tf32_ss(tmem + (uint32_t)((k % GROUPS) * N), a + 2 * k, b + 2 * k, idesc, 1);
                __syncwarp();
            }
            t1 = clock64();
            if (ptx::elect_one()) ptx::mma_commit(&bar);
            __syncwarp();
            ptx::mbar_wait(&bar, 0);
            t2 = clock64();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

template <int N, int GROUPS, int STYLE>
void run(int grid, long long *d)
{
    const int reps = 4096;
    long long h[2];
    cudaFuncSetAttribute(k_rate<N, GROUPS, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_rate<N, GROUPS, STYLE><<<grid, 128, 64 * 1024>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%4d %3d %6d %5d | %10.1f | %10.1f | %6.1f\n", grid, N, GROUPS, STYLE, (double)h[0] / reps,
           (double)h[1] / reps, 128.0 * N / 256.0);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 16);
    printf("grid  N  groups style | issue cyc/mma | complete cyc/mma | ideal(128*N/256)\n");
    for (int grid : {1, 148}) {
        run<32, 1, 0>(grid, d); run<32, 1, 1>(grid, d); run<32, 4, 0>(grid, d); run<32, 4, 1>(grid, d);
        run<64, 1, 0>(grid, d); run<64, 1, 1>(grid, d); run<64, 4, 1>(grid, d);
        run<128, 1, 0>(grid, d); run<128, 1, 1>(grid, d); run<128, 2, 1>(grid, d);
        run<256, 1, 0>(grid, d); run<256, 1, 1>(grid, d); run<256, 2, 1>(grid, d);
    }
    return 0;
}
