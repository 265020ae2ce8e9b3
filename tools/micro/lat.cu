// Latency microbenchmarks that guide the warp kernel's design: dependent-issue latencies on sm_100a (one warp, one CTA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lat tools/micro/lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* t, double a, double b, int mode) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a + lane, y = b, c0 = 0, c1 = 0;
    int idx = lane;
    long long t0 = clock64();
    if (mode == 0) { for (int i = 0; i < N; ++i) x = fma(x, y, a); }                                       // dependent DFMA
    if (mode == 1) { for (int i = 0; i < N; ++i) { x = fma(x, y, a); c0 = fma(c0, y, b); } }               // 2 chains
    if (mode == 2) { for (int i = 0; i < N; ++i) { x = fma(x, y, a); c0 = fma(c0, y, b); c1 = fma(c1, y, a); y = fma(y, b, b);} }  // 4 chains
    if (mode == 3) { for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31); }       // dependent SHFL.64 (2 SHFL)
    if (mode == 4) { for (int i = 0; i < N; ++i) { idx = (int)sm[idx & 63] & 31; } x = idx; }             // dependent LDS.64 (+cvt)
    if (mode == 5) { for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + a; __syncwarp(); } }  // STS -> sync -> LDS -> DADD -> sync
    if (mode == 6) { for (int i = 0; i < N; ++i) dmma(c0, c1, a, b); x = c0 + c1; }                         // dependent DMMA
    if (mode == 7) { double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
                     for (int i = 0; i < N; ++i) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); }
                     x = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1; }                                           // 4 independent DMMA chains
    if (mode == 8) { for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + a; } }   // rcp.approx + DADD
    if (mode == 9) { for (int i = 0; i < N; ++i) { __syncwarp(); x = x + a; } }                              // syncwarp + DADD
    if (mode == 10) { for (int i = 0; i < N; ++i) x = x + a; }                                              // dependent DADD
    if (mode == 11) { for (int i = 0; i < N; ++i) { double2 v = *reinterpret_cast<double2*>(sm + ((idx * 2) & 62)); idx = (int)v.x & 31; } x = idx; }  // dependent LDS.128
    long long t1 = clock64();
    out[threadIdx.x] = x + c0 + c1 + y;
    if (threadIdx.x == 0) t[0] = t1 - t0;
}
int main() {
    double* out; long long* t;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&t, 8);
    const char* names[] = {"dependent DFMA", "2 DFMA chains (per pair)", "4 DFMA chains (per quad)", "dependent SHFL of a double", "dependent LDS.64 + cvt",
                           "STS->syncwarp->LDS->DADD->syncwarp", "dependent DMMA m8n8k4", "4 independent DMMA chains (per 4)", "rcp.approx.f64 + DADD", "syncwarp + DADD", "dependent DADD", "dependent LDS.128 + cvt"};
    for (int warps = 1; warps <= 16; warps *= 4)
    for (int m = 0; m < 12; ++m) {
        k<<<1, 32 * warps>>>(out, t, 1.0000001, 0.9999999, m);
        cudaDeviceSynchronize();
        k<<<1, 32 * warps>>>(out, t, 1.0000001, 0.9999999, m);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
        printf("warps/CTA %2d  %-40s %6.1f cycles per iteration\n", warps, names[m], (double)h / N);
    }
    return 0;
}
