// microbench2.cu -- the packed 16-bit DP cell in isolation: SMSP cycles per packed cell for the two formulations of
// gnx_fill16.cuh (three-op chain / one-op chain) and for the bare DPX ops, against ILP (independent rows per thread)
// and warps per SM.  Not part of the product; results under profiles/.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench2 tools/microbench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#define ITER 2048

enum Op { MAX3U, ADDMAXU, MAX3S, ADDMAXS, CELL3, CELL1, CELL3_NOIMAD, NOPS };
static const char *names[] = {"VIMNMX3.U16x2", "VIADDMNMX.U16x2", "VIMNMX3.S32", "VIADDMNMX.S32",
                              "cell: max3 + IMAD + 2 addmax + IADD", "cell: 4 addmax + IADD (one-op chain)",
                              "cell: max3 + IADD + 2 addmax + IADD"};
static const int alu_per_cell[] = {1, 1, 1, 1, 3, 4, 3};

template <int OP, int ILP> __global__ void bench(unsigned *out, long long *cycles, unsigned e, unsigned oe, int one, unsigned s)
{
    unsigned It[ILP], Dt[ILP], MH[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        It[k] = threadIdx.x * 3 + k;
        Dt[k] = threadIdx.x * 5 + k * 7;
        MH[k] = threadIdx.x + k * 11;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == MAX3U) It[k] = __vimax3_u16x2(It[k], e, oe + u);
                if (OP == ADDMAXU) It[k] = __viaddmax_u16x2(It[k], e, oe + u);
                if (OP == MAX3S) It[k] = (unsigned)__vimax3_s32((int)It[k], (int)e, (int)oe + u);
                if (OP == ADDMAXS) It[k] = (unsigned)__viaddmax_s32((int)It[k], (int)e, (int)oe + u);
                if (OP == CELL3) {
                    const unsigned H = __vimax3_u16x2(MH[k], It[k], Dt[k]);
                    const unsigned Ho = (unsigned)((int)H * one + (int)oe);
                    It[k] = __viaddmax_u16x2(It[k], e, Ho);
                    Dt[k] = __viaddmax_u16x2(Dt[k], e, Ho);
                    MH[k] = H + s;
                }
                if (OP == CELL3_NOIMAD) {
                    const unsigned H = __vimax3_u16x2(MH[k], It[k], Dt[k]);
                    const unsigned Ho = H + oe;
                    It[k] = __viaddmax_u16x2(It[k], e, Ho);
                    Dt[k] = __viaddmax_u16x2(Dt[k], e, Ho);
                    MH[k] = H + s;
                }
                if (OP == CELL1) {
                    const unsigned Y = __viaddmax_u16x2(Dt[k], oe, MH[k]);
                    const unsigned Hs = __viaddmax_u16x2(It[k], oe, Y);
                    It[k] = __viaddmax_u16x2(It[k], e, Y);
                    Dt[k] = __viaddmax_u16x2(Dt[k], e, Hs);
                    MH[k] = (unsigned)((int)Hs * one + (int)s);
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc += It[k] + Dt[k] + MH[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP, int ILP> void run(int sms, int warps_per_sm, unsigned *d_out, long long *d_cyc)
{
    const int grid = sms * warps_per_sm; // one warp per block, like the DP kernels
    bench<OP, ILP><<<grid, 32>>>(d_out, d_cyc, 0xff6aff6au, 0xfd12fd12u, 1, 91u * 65537u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP, ILP><<<grid, 32>>>(d_out, d_cyc, 0xff6aff6au, 0xfd12fd12u, 1, 91u * 65537u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cells_per_smsp = (double)warps_per_sm / 4 * ITER * 4 * ILP;
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-40s ILP=%d warps/SM=%2d  ms=%7.3f  SMSP-cycles per cell=%6.2f  per ALU op=%5.2f\n", names[OP], ILP, warps_per_sm, ms,
           cyc / cells_per_smsp, cyc / cells_per_smsp / alu_per_cell[OP]);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

template <int OP> void sweep(int sms, unsigned *d_out, long long *d_cyc)
{
    for (int w : {8, 12, 16, 32}) {
        run<OP, 1>(sms, w, d_out, d_cyc);
        run<OP, 2>(sms, w, d_out, d_cyc);
        run<OP, 4>(sms, w, d_out, d_cyc);
    }
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s  SMs=%d\n", p.name, p.multiProcessorCount);
    unsigned *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(unsigned) * p.multiProcessorCount * 32 * 32);
    cudaMalloc(&d_cyc, sizeof(long long) * p.multiProcessorCount * 32);
    const int sms = p.multiProcessorCount;
    sweep<MAX3U>(sms, d_out, d_cyc);
    sweep<ADDMAXU>(sms, d_out, d_cyc);
    sweep<MAX3S>(sms, d_out, d_cyc);
    sweep<ADDMAXS>(sms, d_out, d_cyc);
    sweep<CELL3>(sms, d_out, d_cyc);
    sweep<CELL3_NOIMAD>(sms, d_out, d_cyc);
    sweep<CELL1>(sms, d_out, d_cyc);
    return 0;
}
