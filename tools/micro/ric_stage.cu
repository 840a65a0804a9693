// Developer microbenchmark: cycles per Riccati stage of ONE warp (lane = instance), records in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mahi_mpc_b200/csrc/generated -o /tmp/ric_stage tools/micro/ric_stage.cu && /tmp/ric_stage
#include <cstdio>
#include "mmpc_model_double_pendulum.cuh"
#include "../kernels/mmpc_kernels.cuh"
using M = mmpc_model_double_pendulum;
using C = mmpc::T3<M, 32, 7, 5>;

template <int MODE>
__global__ void __launch_bounds__(256) k_ric(int nstage, double* kg, double* pg, long long* cycles, double* sink) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    // 7 records: Fx = I + small, Fu small, H = diag-dominant, c, q small
    for (int e = tid; e < 7 * C::RB * 32; e += blockDim.x) {
        const int f = (e / 32) % C::RB;
        double v = 1e-3 * ((e * 2654435761u) % 1000) / 1000.0;
        if (f >= C::RB_H && f < C::RB_C) { int h = f - C::RB_H; int i = 0, rem = h; while (rem >= 6 - i) { rem -= 6 - i; ++i; } if (rem == 0) v += 2.0; }
        smem[e] = v;
    }
    __syncthreads();
    if (tid >= 32) return;
    mmpc::T3Riccati<M> V;
    for (int a = 0; a < 16; ++a) V.Pxx[a] = (a % 5 == 0) ? 1.0 : 0.0;
    for (int a = 0; a < 8; ++a) V.Pxv[a] = 0.0;
    for (int a = 0; a < 4; ++a) { V.Pvv[a] = 0.0; V.px[a] = 0.1; }
    V.pv[0] = V.pv[1] = 0.0;
    double R[2] = {5.0, 5.0};
    bool ok = true;
    const long long t0 = clock64();
    for (int s = 0; s < nstage; ++s) {
        double* kgp = MODE == 1 ? kg + tid : kg + (size_t)(s % 50) * C::KG * 32 + tid;
        double* pgp = MODE == 1 ? pg + tid : pg + (size_t)(s % 50) * C::PG * 32 + tid;
        ok = mmpc::t3_riccati_stage<M, C>(V, smem, s % 7, tid, R, kgp, pgp) && ok;
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    double acc = ok ? 1.0 : 0.0;
    for (int a = 0; a < 16; ++a) acc += V.Pxx[a];
    sink[tid] = acc + V.px[0] + V.pv[0] + V.Pxv[0] + V.Pvv[0];
}

int main() {
    double *kg, *pg, *sink; long long* cyc;
    cudaMalloc(&kg, 50 * 64 * 32 * 8); cudaMalloc(&pg, 50 * 64 * 32 * 8); cudaMalloc(&sink, 256 * 8); cudaMalloc(&cyc, 8);
    const size_t sm = 7 * C::RB * 32 * 8;
    cudaFuncSetAttribute(k_ric<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(k_ric<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
            long long h;
            k_ric<0><<<1, threads, sm>>>(2000, kg, pg, cyc, sink); cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads %d  stores to 50 rotating node records: %.1f cycles / stage   (%s)\n", threads, h / 2000.0, cudaGetErrorString(cudaGetLastError()));
            k_ric<1><<<1, threads, sm>>>(2000, kg, pg, cyc, sink); cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("threads %d  stores to one record:                   %.1f cycles / stage\n", threads, h / 2000.0);
        }
    }
    return 0;
}
