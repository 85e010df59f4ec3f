// vote_mix.cu -- microbenchmark of the vote kernel's inner loop in isolation (every form it went through): which instruction mix per inlier
// test does the SM sustain?  One CTA per SM-slot, no global traffic inside the timed loop: every lane keeps
// HPL hypotheses in registers and sweeps a 64-pixel record tile from shared memory over and over.
// Variants (template V):
//   0  4 FFMA + FADD + FSETP(count) + IADD + FSETP(band)            (round 2, first form: ALU-heavy)
//   1  4 FFMA + FADD + FFMA.SAT + FADD + FSETP(band)                 (scalar FMA, count on the FMA pipe)
//   2  2 FFMA2 + FADD + FFMA.SAT + FADD + FSETP(band)                (shipped form)
//   3  2 FFMA2 + FADD + FFMA.SAT + FADD                              (no band check: what the check costs)
//   4  2 FFMA2 + FADD + FSETP(count) + IADD + FSETP(band)            (packed FMAs, ALU count)
//   5  2 FFMA2 only                                                  (raw packed-FMA rate)
//   6  4 FFMA only                                                   (raw scalar-FMA rate)
// Output: cycles per test-warp (32 tests) per SM sub-partition, for 2/3/4/6/8 warps per sub-partition.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/vote_mix benchmarks/micro/vote_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min_nan(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float min3_nan_abs(float a, float b, float c) { float r; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c))); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void count_if_gt(int &cnt, float a, float b)
{
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "f"(a), "f"(b));
}

constexpr int HPL = 8, PIX = 64;

template <int V>
__global__ void __launch_bounds__(256) k_mix(int reps, float seed, float *out, long long *cycles)
{
    __shared__ float4 rec[3 * PIX];
    for (int i = threadIdx.x; i < 3 * PIX; i += blockDim.x)
        rec[i] = make_float4(0.01f * i + seed, 0.02f * i - seed, 0.5f - 0.003f * i, 0.25f + 0.001f * i);
    __syncthreads();
    float hx[HPL], hy[HPL], bd[HPL], nb2[HPL], cntf[HPL];
    int cnt[HPL];
    f32x2 hx2[HPL / 2], hy2[HPL / 2];
    for (int j = 0; j < HPL; ++j) {
        hx[j] = seed + 0.1f * j + 0.01f * threadIdx.x;
        hy[j] = seed - 0.2f * j + 0.02f * threadIdx.x;
        bd[j] = 1e-5f * (1 + j);
        nb2[j] = -bd[j] * 18446744073709551616.f;
        cntf[j] = 0.f;
        cnt[j] = 0;
    }
    for (int j = 0; j < HPL / 2; ++j) {
        hx2[j] = pk2(hx[2 * j], hx[2 * j + 1]);
        hy2[j] = pk2(hy[2 * j], hy[2 * j + 1]);
    }
    bool unc = false, unc1 = false, unc2 = false, unc3 = false;
    f32x2 acc2[HPL / 2];
    float accs[HPL];
    for (int j = 0; j < HPL / 2; ++j) acc2[j] = pk2(0.f, 0.f);
    for (int j = 0; j < HPL; ++j) accs[j] = 0.f;
    float ev[2 * HPL];
    for (int j = 0; j < 2 * HPL; ++j) ev[j] = 1e30f;
    float mab[HPL], mabB[HPL];
    for (int j = 0; j < HPL; ++j) mab[j] = mabB[j] = 1e30f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (V == 18 || V == 19) {
#pragma unroll 2
            for (int p = 0; p < PIX; p += 2) {
                const float4 a0 = rec[3 * p], b0 = rec[3 * p + 1], c0 = rec[3 * p + 2];
                const float4 a1 = rec[3 * p + 3], b1 = rec[3 * p + 4], c1 = rec[3 * p + 5];
                const f32x2 AP0 = pk2(a0.x, a0.y), BP0 = pk2(a0.z, a0.w), CP0 = pk2(b0.x, b0.y), AM0 = pk2(b0.z, b0.w),
                            BM0 = pk2(c0.x, c0.y), CM0 = pk2(c0.z, c0.w);
                const f32x2 AP1 = pk2(a1.x, a1.y), BP1 = pk2(a1.z, a1.w), CP1 = pk2(b1.x, b1.y), AM1 = pk2(b1.z, b1.w),
                            BM1 = pk2(c1.x, c1.y), CM1 = pk2(c1.z, c1.w);
#pragma unroll
                for (int j = 0; j < HPL / 2; ++j) {
                    if (V == 19) asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");     // does ptxas schedule across it?
                    // ALU pipe: the previous pair's e values of this hypothesis pair
                    cnt[2 * j] += __float_as_uint(ev[4 * j]) >> 31;
                    cnt[2 * j + 1] += __float_as_uint(ev[4 * j + 1]) >> 31;
                    cnt[2 * j] += __float_as_uint(ev[4 * j + 2]) >> 31;
                    cnt[2 * j + 1] += __float_as_uint(ev[4 * j + 3]) >> 31;
                    mab[2 * j] = min3_nan_abs(mab[2 * j], ev[4 * j], ev[4 * j + 2]);
                    mab[2 * j + 1] = min3_nan_abs(mab[2 * j + 1], ev[4 * j + 1], ev[4 * j + 3]);
                    // FMA pipe: this pair
                    const f32x2 p0 = fma2(hx2[j], AP0, fma2(hy2[j], BP0, CP0));
                    const f32x2 q0 = fma2(hx2[j], AM0, fma2(hy2[j], BM0, CM0));
                    const f32x2 p1 = fma2(hx2[j], AP1, fma2(hy2[j], BP1, CP1));
                    const f32x2 q1 = fma2(hx2[j], AM1, fma2(hy2[j], BM1, CM1));
                    float p0a, p0b, q0a, q0b, p1a, p1b, q1a, q1b;
                    upk2(p0, p0a, p0b);
                    upk2(q0, q0a, q0b);
                    upk2(p1, p1a, p1b);
                    upk2(q1, q1a, q1b);
                    ev[4 * j] = fabsf(q0a) - p0a;
                    ev[4 * j + 1] = fabsf(q0b) - p0b;
                    ev[4 * j + 2] = fabsf(q1a) - p1a;
                    ev[4 * j + 3] = fabsf(q1b) - p1b;
                }
            }
            continue;
        }
        if (V >= 7 && V != 13) {
#pragma unroll 2
            for (int p = 0; p < PIX; p += 2) {
                const float4 a0 = rec[3 * p], b0 = rec[3 * p + 1], c0 = rec[3 * p + 2];
                const float4 a1 = rec[3 * p + 3], b1 = rec[3 * p + 4], c1 = rec[3 * p + 5];
                const f32x2 AP0 = pk2(a0.x, a0.y), BP0 = pk2(a0.z, a0.w), CP0 = pk2(b0.x, b0.y), AM0 = pk2(b0.z, b0.w),
                            BM0 = pk2(c0.x, c0.y), CM0 = pk2(c0.z, c0.w);
                const f32x2 AP1 = pk2(a1.x, a1.y), BP1 = pk2(a1.z, a1.w), CP1 = pk2(b1.x, b1.y), AM1 = pk2(b1.z, b1.w),
                            BM1 = pk2(c1.x, c1.y), CM1 = pk2(c1.z, c1.w);
#pragma unroll
                for (int j = 0; j < HPL / 2; ++j) {
                    const f32x2 p0 = fma2(hx2[j], AP0, fma2(hy2[j], BP0, CP0));
                    const f32x2 q0 = fma2(hx2[j], AM0, fma2(hy2[j], BM0, CM0));
                    const f32x2 p1 = fma2(hx2[j], AP1, fma2(hy2[j], BP1, CP1));
                    const f32x2 q1 = fma2(hx2[j], AM1, fma2(hy2[j], BM1, CM1));
                    float p0a, p0b, q0a, q0b, p1a, p1b, q1a, q1b;
                    upk2(p0, p0a, p0b);
                    upk2(q0, q0a, q0b);
                    upk2(p1, p1a, p1b);
                    upk2(q1, q1a, q1b);
                    float m0a, m0b, m1a, m1b;
                    if (V == 16 || V == 17) {               // -m: negative exactly when num > |perp|
                        m0a = fabsf(q0a) - p0a, m0b = fabsf(q0b) - p0b, m1a = fabsf(q1a) - p1a, m1b = fabsf(q1b) - p1b;
                        cnt[2 * j] += __float_as_uint(m0a) >> 31;
                        cnt[2 * j + 1] += __float_as_uint(m0b) >> 31;
                        cnt[2 * j] += __float_as_uint(m1a) >> 31;
                        cnt[2 * j + 1] += __float_as_uint(m1b) >> 31;
                        if (V == 16) {
                            mab[2 * j] = min3_nan_abs(mab[2 * j], m0a, m1a);
                            mab[2 * j + 1] = min3_nan_abs(mab[2 * j + 1], m0b, m1b);
                        }
                        continue;
                    }
                    if (V == 11 || V == 15) {
                        m0a = p0a - fabsf(q0a), m0b = p0b - fabsf(q0b), m1a = p1a - fabsf(q1a), m1b = p1b - fabsf(q1b);
                    } else {
                        m0a = min_nan(p0a, q0a), m0b = min_nan(p0b, q0b), m1a = min_nan(p1a, q1a), m1b = min_nan(p1b, q1b);
                    }
                    const float s0a = fma_sat(m0a, 18446744073709551616.f, nb2[2 * j]);
                    const float s0b = fma_sat(m0b, 18446744073709551616.f, nb2[2 * j + 1]);
                    const float s1a = fma_sat(m1a, 18446744073709551616.f, nb2[2 * j]);
                    const float s1b = fma_sat(m1b, 18446744073709551616.f, nb2[2 * j + 1]);
                    if (V == 8) {
                        cnt[2 * j] += __float_as_int(s0a) + __float_as_int(s1a);
                        cnt[2 * j + 1] += __float_as_int(s0b) + __float_as_int(s1b);
                    } else if (V == 10) {
                        acc2[j] = add2(acc2[j], pk2(s0a, s0b));
                        acc2[j] = add2(acc2[j], pk2(s1a, s1b));
                    } else {
                        cntf[2 * j] += s0a;
                        cntf[2 * j + 1] += s0b;
                        cntf[2 * j] += s1a;
                        cntf[2 * j + 1] += s1b;
                    }
                    if (V == 11 || V == 12) {
                        // two accumulators per hypothesis, or ptxas fuses the nested mins back into FMNMX3
                        mab[2 * j] = min_nan(mab[2 * j], fabsf(m0a));
                        mabB[2 * j] = min_nan(mabB[2 * j], fabsf(m1a));
                        mab[2 * j + 1] = min_nan(mab[2 * j + 1], fabsf(m0b));
                        mabB[2 * j + 1] = min_nan(mabB[2 * j + 1], fabsf(m1b));
                    } else if (V == 14) {
                        accs[2 * j] += fma_sat(m0a, 18446744073709551616.f, bd[2 * j]);
                        accs[2 * j + 1] += fma_sat(m0b, 18446744073709551616.f, bd[2 * j + 1]);
                        accs[2 * j] += fma_sat(m1a, 18446744073709551616.f, bd[2 * j]);
                        accs[2 * j + 1] += fma_sat(m1b, 18446744073709551616.f, bd[2 * j + 1]);
                    } else if (V != 9) {
                        mab[2 * j] = min3_nan_abs(mab[2 * j], m0a, m1a);
                        mab[2 * j + 1] = min3_nan_abs(mab[2 * j + 1], m0b, m1b);
                    }
                }
            }
            continue;
        }
#pragma unroll 4
        for (int p = 0; p < PIX; ++p) {
            const float4 a = rec[3 * p], b = rec[3 * p + 1], c = rec[3 * p + 2];
            if (V == 0 || V == 1 || V == 6) {
#pragma unroll
                for (int j = 0; j < HPL; ++j) {
                    const float num = fmaf(hx[j], a.x, fmaf(hy[j], a.z, b.x));
                    const float per = fmaf(hx[j], b.z, fmaf(hy[j], c.x, c.z));
                    if (V == 6) {
                        accs[j] += 0.f;
                        hx[j] = num;
                        hy[j] = per;
                        continue;
                    }
                    const float m = num - fabsf(per);
                    if (V == 0) count_if_gt(cnt[j], m, bd[j]);
                    else cntf[j] += fma_sat(m, 18446744073709551616.f, nb2[j]);
                    unc |= !(fabsf(m) > bd[j]);
                }
            } else {
                const f32x2 SX = pk2(a.x, a.y), SY = pk2(a.z, a.w), NS = pk2(b.x, b.y), CX = pk2(b.z, b.w), CY = pk2(c.x, c.y),
                            NC = pk2(c.z, c.w);
#pragma unroll
                for (int j = 0; j < HPL / 2; ++j) {
                    const f32x2 num2 = fma2(hx2[j], SX, fma2(hy2[j], SY, NS));
                    const f32x2 per2 = fma2(hx2[j], CX, fma2(hy2[j], CY, NC));
                    if (V == 5) {
                        hx2[j] = num2;
                        hy2[j] = per2;
                        continue;
                    }
                    float n0, n1, q0, q1;
                    upk2(num2, n0, n1);
                    upk2(per2, q0, q1);
                    const float m0 = n0 - fabsf(q0), m1 = n1 - fabsf(q1);
                    if (V == 4) {
                        count_if_gt(cnt[2 * j], m0, bd[2 * j]);
                        count_if_gt(cnt[2 * j + 1], m1, bd[2 * j + 1]);
                    } else {
                        cntf[2 * j] += fma_sat(m0, 18446744073709551616.f, nb2[2 * j]);
                        cntf[2 * j + 1] += fma_sat(m1, 18446744073709551616.f, nb2[2 * j + 1]);
                    }
                    if (V == 13) {
                        if (j & 1) {
                            unc2 |= !(fabsf(m0) > bd[2 * j]);
                            unc3 |= !(fabsf(m1) > bd[2 * j + 1]);
                        } else {
                            unc |= !(fabsf(m0) > bd[2 * j]);
                            unc1 |= !(fabsf(m1) > bd[2 * j + 1]);
                        }
                    } else if (V != 3) {
                        unc |= !(fabsf(m0) > bd[2 * j]);
                        unc |= !(fabsf(m1) > bd[2 * j + 1]);
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = (unc ? 1.f : 0.f) + (unc1 ? 2.f : 0.f) + (unc2 ? 4.f : 0.f) + (unc3 ? 8.f : 0.f);
    for (int j = 0; j < HPL; ++j) s += cntf[j] + (float)cnt[j] + hx[j] + hy[j] + accs[j] + mab[j] + mabB[j] + ev[j] + ev[HPL + j];
    for (int j = 0; j < HPL / 2; ++j) {
        float x, y;
        upk2(acc2[j], x, y);
        s += x + y;
    }
    for (int j = 0; j < HPL / 2; ++j) {
        float x, y;
        upk2(hx2[j], x, y);
        s += x + y;
        upk2(hy2[j], x, y);
        s += x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name, int sms)
{
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 256 * 2);
    cudaMalloc(&cyc, sizeof(long long) * sms * 8);
    const int reps = 400;
    printf("%-58s", name);
    for (int ctas = 1; ctas <= 4; ++ctas) {                      // 256-thread CTAs per SM: 2, 4, 6, 8 warps per sub-partition
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mix<V>, 256, 0);
        if (ctas > occ) { printf("  %6s", "-"); continue; }
        k_mix<V><<<sms * ctas, 256>>>(reps, 0.37f, out, cyc);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_mix<V><<<sms * ctas, 256>>>(reps, 0.37f, out, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        long long h[8 * 200];
        cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms * ctas; ++i) mean += (double)h[i];
        mean /= sms * ctas;
        // test-warps per sub-partition in that time: ctas * 8 warps / 4 sub-partitions * reps * PIX * HPL
        const double tw = (double)ctas * 2 * reps * PIX * HPL;
        printf("  %6.2f", mean / tw);
    }
    printf("\n");
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("cycles per test-warp (32 inlier tests) per SM sub-partition; columns: 2, 4, 6, 8 resident warps per sub-partition\n");
    run<0>("0: 4 FFMA + FADD + FSETP + IADD + FSETP(band)", sms);
    run<1>("1: 4 FFMA + FADD + FFMA.SAT + FADD + FSETP(band)", sms);
    run<2>("2: 2 FFMA2 + FADD + FFMA.SAT + FADD + FSETP(band)  [round 2, first kernel]", sms);
    run<3>("3: 2 FFMA2 + FADD + FFMA.SAT + FADD  (no band check)", sms);
    run<4>("4: 2 FFMA2 + FADD + FSETP + IADD + FSETP(band)", sms);
    run<5>("5: 2 FFMA2 only", sms);
    run<6>("6: 4 FFMA only", sms);
    run<7>("7: 2 FFMA2 + FMNMX + FFMA.SAT + FADD + 1/2 FMNMX3  (cone edges)", sms);
    run<8>("8: 2 FFMA2 + FMNMX + FFMA.SAT + 1/2 IADD3 + 1/2 FMNMX3", sms);
    run<9>("9: 2 FFMA2 + FMNMX + FFMA.SAT + FADD  (no band tracking)", sms);
    run<10>("10: 2 FFMA2 + FMNMX + FFMA.SAT + 1/2 FADD2 + 1/2 FMNMX3", sms);
    run<11>("11: 2 FFMA2 + FADD + FFMA.SAT + FADD + FMNMX(band)", sms);
    run<12>("12: 2 FFMA2 + FMNMX + FFMA.SAT + FADD + FMNMX(band)", sms);
    run<13>("13: = 2 with four band predicates", sms);
    run<14>("14: 2 FFMA2 + FMNMX + 2 FFMA.SAT + 2 FADD", sms);
    run<15>("15: 2 FFMA2 + FADD + FFMA.SAT + FADD + 1/2 FMNMX3", sms);
    run<16>("16: 2 FFMA2 + FADD + LEA.HI + 1/2 FMNMX3  [k_vote3, shipped]", sms);
    run<17>("17: 2 FFMA2 + FADD + LEA.HI  (no band tracking)", sms);
    run<18>("18: = 16 software-pipelined (ALU ops of pair i-1 beside FMA ops of pair i)", sms);
    run<19>("19: = 18 with a warp barrier between hypothesis pairs (short scheduling regions)", sms);
    return 0;
}
