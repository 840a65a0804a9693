// Single-warp latency / throughput probes for the fp64 serial chains of the solve kernels (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/microbench_fp64 tools/microbench_fp64.cu
// Each probe runs in one warp of one CTA (or many CTAs for the contention rows) and reports cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>

#define REP 512

template <int LANES, int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    long long t0 = 0, t1 = 0;
    if ((threadIdx.x & 31) < LANES) {
        t0 = clock64();
#pragma unroll 16
        for (int r = 0; r < REP; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
        t1 = clock64();
    }
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
__global__ void op_chain(double* out, long long* cyc, double a) {
    double x = 1.5 + threadIdx.x * 1e-3, y = 0.0;
    long long t0 = clock64();
#pragma unroll 4
    for (int r = 0; r < REP; ++r) {
        if (OP == 0) x = 1.0 / (x + a);
        if (OP == 1) x = rsqrt(x + a);
        if (OP == 2) x = sqrt(x + a);
        if (OP == 3) { double s, c; sincos(x, &s, &c); x = s + c + a; }
        if (OP == 4) x = x + a;
        if (OP == 5) x = x * a;
        if (OP == 6) { x = fmax(x, a) + (isfinite(x) ? 1.0 : 0.0); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + y;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_chain(double* out, long long* cyc) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int r = 0; r < REP; ++r) p = nxt[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// a stage-like mix: 8 smem loads feeding 2 dependent length-4 dot products
__global__ void lds_fma_mix(double* out, long long* cyc) {
    __shared__ double rec[64 * 8];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) rec[i] = 1.0 / (1 + i);
    __syncthreads();
    double x0 = 1, x1 = 2, x2 = 3, x3 = 4;
    const int t = threadIdx.x & 7;
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r) {
        const double* q = rec + ((r & 3) * 16) * 8 + t;
        double y0 = q[0], y1 = q[8], y2 = q[16], y3 = q[24];
        y0 = fma(q[32], x0, y0); y0 = fma(q[40], x1, y0); y0 = fma(q[48], x2, y0); y0 = fma(q[56], x3, y0);
        y1 = fma(q[64], x0, y1); y1 = fma(q[72], x1, y1); y1 = fma(q[80], x2, y1); y1 = fma(q[88], x3, y1);
        y2 = fma(q[96], x0, y2); y2 = fma(q[104], x1, y2); y2 = fma(q[112], x2, y2); y2 = fma(q[120], x3, y2);
        y3 = fma(q[32], x3, y3); y3 = fma(q[40], x2, y3); y3 = fma(q[48], x1, y3); y3 = fma(q[56], x0, y3);
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 22);
    cudaMalloc(&cyc, 8);
#define RUN(label, per, ...)                                   \
    do {                                                       \
        __VA_ARGS__;                                           \
        __VA_ARGS__;                                           \
        cudaDeviceSynchronize();                               \
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);        \
        printf("%-52s %8.2f cyc/op\n", label, (double)h / (per)); \
    } while (0)
    RUN("DFMA dependent chain, 32 lanes, 1 warp", REP, dfma_chain<32, 1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA dependent chain, 8 lanes, 1 warp", REP, dfma_chain<8, 1><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 2 chains, 32 lanes (per DFMA)", REP * 2, dfma_chain<32, 2><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 4 chains, 32 lanes (per DFMA)", REP * 4, dfma_chain<32, 4><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 8 chains, 32 lanes (per DFMA)", REP * 8, dfma_chain<32, 8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 8 chains, 8 lanes (per DFMA)", REP * 8, dfma_chain<8, 8><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 16 chains, 32 lanes (per DFMA)", REP * 16, dfma_chain<32, 16><<<1, 32>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 8 chains, 4 warps/CTA (per DFMA per warp)", REP * 8, dfma_chain<32, 8><<<1, 128>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 8 chains, 8 warps/CTA (per DFMA per warp)", REP * 8, dfma_chain<32, 8><<<1, 256>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 8 chains, 16 warps/CTA (per DFMA per warp)", REP * 8, dfma_chain<32, 8><<<1, 512>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 1 chain, 8 warps/CTA", REP, dfma_chain<32, 1><<<1, 256>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DFMA 1 chain, 8 lanes, 8 warps/CTA", REP, dfma_chain<8, 1><<<1, 256>>>(out, cyc, 1.0000001, 1e-9));
    RUN("DADD dependent chain", REP, op_chain<4><<<1, 32>>>(out, cyc, 1e-9));
    RUN("DMUL dependent chain", REP, op_chain<5><<<1, 32>>>(out, cyc, 1.0000001));
    RUN("1.0/x dependent chain", REP, op_chain<0><<<1, 32>>>(out, cyc, 1e-9));
    RUN("rsqrt(x) dependent chain", REP, op_chain<1><<<1, 32>>>(out, cyc, 1e-9));
    RUN("sqrt(x) dependent chain", REP, op_chain<2><<<1, 32>>>(out, cyc, 1e-9));
    RUN("sincos(x) dependent chain", REP, op_chain<3><<<1, 32>>>(out, cyc, 1e-9));
    RUN("fmax+isfinite chain", REP, op_chain<6><<<1, 32>>>(out, cyc, 1e-9));
    RUN("LDS dependent chain (pointer chase)", REP, lds_chain<<<1, 32>>>(out, cyc));
    RUN("stage mix: 16 LDS + 16 DFMA (4 chains of 4) per iter", REP, lds_fma_mix<<<1, 32>>>(out, cyc));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
