// microbench.cu -- per-instruction issue throughput on the SM (warp-instructions / clk / SM) for the
// integer ops the DP kernels are made of.  Not part of the product; results inform DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define ITER 4096
#define NACC 8

enum Op { IADD, IMAD, LOP3, MAX2, MAX3, ADDMAX, PRMT, SHF, ADD16X2, MAX3_16X2, ADDMAX_16X2, MAX2_16X2,
          FADD, FMNMX, SHFL, LDS, MIX_ADDMAX_IMAD, MIX_MAX3_LOP3, MIX_ADDMAX_FADD, MIX_MAX3_IMAD_LOP3, ISETP_SEL,
          MIX_MAX3_16_IMAD, NOPS };
static const char *names[] = {"IADD(add.s32)", "IMAD", "LOP3", "VIMNMX(max2)", "VIMNMX3", "VIADDMNMX", "PRMT", "SHF",
                              "VIADD.16x2", "VIMNMX3.S16x2", "VIADDMNMX.S16x2", "VIMNMX.S16x2", "FADD", "FMNMX",
                              "SHFL.UP", "LDS", "mix VIADDMNMX+IMAD", "mix VIMNMX3+LOP3", "mix VIADDMNMX+FADD",
                              "mix VIMNMX3+IMAD+LOP3", "ISETP+SEL", "mix VIMNMX3.S16x2+IMAD"};
static const int ops_per_iter[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 3, 2, 2};

template <int OP> __device__ __forceinline__ void step(int &x, int a, int b, float &f, int *sm, int lane, int y, float g)
{
    if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == MAX2) asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == MAX3) asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == ADDMAX) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 6;" : "+r"(x) : "r"(a));
    if (OP == ADD16X2) asm volatile("add.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == MAX3_16X2) asm volatile("{.reg .b32 t; max.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == ADDMAX_16X2) asm volatile("{.reg .b32 t; add.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
    if (OP == MAX2_16X2) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(f) : "f"(__int_as_float(a)));
    if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(g));
    if (OP == SHFL) x = __shfl_up_sync(0xffffffffu, x, 1);
    if (OP == LDS) x = sm[(x & 1023)];
    if (OP == MIX_ADDMAX_IMAD) {
        asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == MIX_MAX3_LOP3) {
        asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == MIX_ADDMAX_FADD) {
        asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(f) : "f"(__int_as_float(a)));
    }
    if (OP == MIX_MAX3_IMAD_LOP3) {
        asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == ISETP_SEL) {
        asm volatile("{.reg .pred p; setp.ge.s32 p, %0, %1; selp.s32 %0, %0, %2, p;}" : "+r"(x) : "r"(a), "r"(b));
    }
    if (OP == MIX_MAX3_16_IMAD) {
        asm volatile("{.reg .b32 t; max.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(x) : "r"(a), "r"(b));
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    }
}

template <int OP> __global__ void bench(int *out, long long *cycles, int a, int b)
{
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    int x[NACC];
    float f[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) { x[k] = threadIdx.x + k * 17 + a; f[k] = (float)(threadIdx.x + k); }
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) step<OP>(x[k], a, b, f[k], sm, lane, x[(k + 1) % NACC], f[(k + 1) % NACC]);
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc += x[k] + (int)f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(int sms, int threads, int blocks_per_sm, int *d_out, long long *d_cyc)
{
    const int grid = sms * blocks_per_sm;
    bench<OP><<<grid, threads>>>(d_out, d_cyc, 3, 5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<grid, threads>>>(d_out, d_cyc, 3, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(grid);
    cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : cyc) avg += (double)c;
    avg /= grid;
    const double winstr_per_sm = (double)blocks_per_sm * (threads / 32) * ITER * NACC * ops_per_iter[OP];
    printf("%-26s threads=%4d bps=%d  ms=%8.3f  cycles/block=%10.0f  warp-instr/clk/SM=%6.3f  (by events @1.9GHz: %6.3f)\n",
           names[OP], threads, blocks_per_sm, ms, avg, winstr_per_sm / avg,
           winstr_per_sm / (ms * 1e-3 * 1.9e9));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

template <int OP> void sweep(int sms, int *d_out, long long *d_cyc)
{
    run<OP>(sms, 256, 1, d_out, d_cyc);  //  8 warps/SM  (2 per SMSP)
    run<OP>(sms, 512, 2, d_out, d_cyc);  // 32 warps/SM  (8 per SMSP)
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(int) * p.multiProcessorCount * 4 * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * p.multiProcessorCount * 4);
    const int sms = p.multiProcessorCount;
    sweep<IADD>(sms, d_out, d_cyc);
    sweep<IMAD>(sms, d_out, d_cyc);
    sweep<LOP3>(sms, d_out, d_cyc);
    sweep<MAX2>(sms, d_out, d_cyc);
    sweep<MAX3>(sms, d_out, d_cyc);
    sweep<ADDMAX>(sms, d_out, d_cyc);
    sweep<PRMT>(sms, d_out, d_cyc);
    sweep<SHF>(sms, d_out, d_cyc);
    sweep<ADD16X2>(sms, d_out, d_cyc);
    sweep<MAX3_16X2>(sms, d_out, d_cyc);
    sweep<ADDMAX_16X2>(sms, d_out, d_cyc);
    sweep<MAX2_16X2>(sms, d_out, d_cyc);
    sweep<FADD>(sms, d_out, d_cyc);
    sweep<FMNMX>(sms, d_out, d_cyc);
    sweep<SHFL>(sms, d_out, d_cyc);
    sweep<LDS>(sms, d_out, d_cyc);
    sweep<ISETP_SEL>(sms, d_out, d_cyc);
    sweep<MIX_ADDMAX_IMAD>(sms, d_out, d_cyc);
    sweep<MIX_MAX3_LOP3>(sms, d_out, d_cyc);
    sweep<MIX_ADDMAX_FADD>(sms, d_out, d_cyc);
    sweep<MIX_MAX3_IMAD_LOP3>(sms, d_out, d_cyc);
    sweep<MIX_MAX3_16_IMAD>(sms, d_out, d_cyc);
    return 0;
}
