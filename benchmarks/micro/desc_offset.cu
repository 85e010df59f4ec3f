// desc_offset.cu -- microtest: may a tcgen05 K-major SWIZZLE_128B operand descriptor start at a row
// that is NOT a multiple of the 1024-byte swizzle repeat, and may the 8-row groups be spaced by an
// SBO that is not a multiple of 1024 bytes?  (That is what a convolution needs to take all KW x KH
// taps from ONE haloed (TH+2) x (TW+2) activation box instead of re-loading the box per kw shift.)
//
// Shared memory is filled the way TMA SWIZZLE_128B fills a 1024-byte-aligned box of 128-byte rows:
// chunk c (16 B) of row r sits at r*128 + ((c ^ (r & 7)) << 4).  Then D = A_shift * B^T is computed
// with the A descriptor pointing at row s (and, in test 2, a 10-row pitch between 8-row groups), with
// the descriptor's base-offset field (bits 49..51) either 0 or (start >> 7) & 7, and compared with
// the host result.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I pvnet_b200/csrc
// -o gpurun_out/desc_offset benchmarks/micro/desc_offset.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "ptx.cuh"

constexpr int R_FULL = 296;      // rows of the haloed A box
constexpr int NB = 32;           // N

__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 7 + k * 3) % 11 - 5); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k) % 7 - 3); }

__device__ __host__ inline int swz(int o, int rowb) { return o ^ (((o >> 7) & (rowb / 16 - 1)) << 4); }

// rowb = bytes per K-major row = swizzle span: 128, 64 or 32
__global__ void __launch_bounds__(128, 1) k_test(int rowb, int shift_rows, int sbo_bytes, int base_off, float *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + 40 * 1024;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int kk = rowb / 4;
    for (int i = threadIdx.x; i < R_FULL * kk; i += blockDim.x) {
        const int r = i / kk, k = i % kk;
        reinterpret_cast<float *>(sA + swz(r * rowb + (k >> 2) * 16, rowb))[k & 3] = a_val(r, k);
    }
    for (int i = threadIdx.x; i < NB * kk; i += blockDim.x) {
        const int r = i / kk, k = i % kk;
        reinterpret_cast<float *>(sB + swz(r * rowb + (k >> 2) * 16, rowb))[k & 3] = b_val(r, k);
    }
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        ptx::tmem_alloc(&slot, 32);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::make_idesc_tf32(128, NB);
        const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)(shift_rows * rowb);
        uint64_t a = ptx::make_kmajor_desc(a_addr, rowb);
        a = (a & ~(0x3fffull << 32)) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)(base_off & 7) << 49);
        const uint64_t b = ptx::make_kmajor_desc(ptx::smem_u32(sB), rowb);
        for (int k = 0; k < rowb / 32; ++k) ptx::mma_tf32_ss(tmem, a + 2 * k, b + 2 * k, idesc, k != 0 ? 1u : 0u);
        ptx::mma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    uint32_t r[32];
    const int warp = threadIdx.x >> 5;
    ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(r[j]);
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 32);
}

static double run(int rowb, int shift, int pitch_rows, int base_off, float *d_out)
{
    k_test<<<1, 128, 64 * 1024>>>(rowb, shift, pitch_rows * rowb, base_off, d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
        exit(1);
    }
    std::vector<float> h(128 * 32);
    cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < 128; ++m) {
        const int row = (m / 8) * pitch_rows + (m % 8) + shift;
        for (int n = 0; n < NB; ++n) {
            double ref = 0;
            for (int k = 0; k < rowb / 4; ++k) ref += (double)a_val(row, k) * b_val(n, k);
            err = fmax(err, fabs(ref - h[m * 32 + n]));
        }
    }
    return err;
}

int main()
{
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    float *d_out;
    cudaMalloc(&d_out, 128 * 32 * 4);
    printf("# max |D - ref| ; 0 means the descriptor addressed the rows we meant\n");
    printf("# test            shift  base_offset=0   base_offset=(start>>7)&7\n");
    for (int rowb : {128, 64, 32})
        for (int pitch : {8, 10, 11}) {
            for (int s = 0; s < 8; ++s) {
                const double e0 = run(rowb, s, pitch, 0, d_out);
                const int bo = ((s * rowb) >> 7) & 7;
                const double e1 = bo ? run(rowb, s, pitch, bo, d_out) : e0;
                printf("rowb=%3d pitch=%2d rows   %d      %-14g  %g\n", rowb, pitch, s, e0, e1);
            }
        }
    return 0;
}
