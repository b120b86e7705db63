// Developer micro-benchmark: dependent-issue latency of the instructions the serial sections of k_riccati_bdf are made of.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
    __shared__ double s[64];
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999, c = 1e-7;
    for (int i = threadIdx.x; i < 64; i += 32) s[i] = 1.0 + i * 1e-3;
    __syncwarp();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < n; ++i) d = d + c;
    long long t2 = clock64();
    double e = d;
    for (int i = 0; i < n; ++i) e = e * b;
    long long t3 = clock64();
    double f = e;
    for (int i = 0; i < n; ++i) f = __shfl_xor_sync(0xffffffffu, f, 1) + c;
    long long t4 = clock64();
    double g = f; int idx = threadIdx.x;
    for (int i = 0; i < n; ++i) { g = s[(idx + (int)g) & 63]; }
    long long t5 = clock64();
    double h = g + 2.0;
    for (int i = 0; i < n; ++i) h = 1.0 / h + 1.5;
    long long t6 = clock64();
    double r = h + 2.0;
    for (int i = 0; i < n; ++i) r = rsqrt(r) + 1.5;
    long long t7 = clock64();
    double q = r + 2.0;
    for (int i = 0; i < n; ++i) q = sqrt(q) + 1.5;
    long long t8 = clock64();
    double w = q;
    for (int i = 0; i < n; ++i) { s[threadIdx.x] = w; __syncwarp(); w = s[threadIdx.x ^ 1] + c; __syncwarp(); }
    long long t9 = clock64();
    if (threadIdx.x == 0) { out[0] = w; cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7; cyc[8] = t9 - t8; }
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 8); cudaMalloc(&c, 80);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(o, c, n);
    long long h[9]; cudaMemcpy(h, c, 72, cudaMemcpyDeviceToHost);
    const char* nm[9] = {"dfma", "dadd", "dmul", "shfl64+dadd", "lds(dep, incl f2i)", "ddiv+dadd", "drsqrt+dadd", "dsqrt+dadd", "sts+sync+lds+dadd+sync"};
    for (int i = 0; i < 9; ++i) printf("%-24s %.1f cycles\n", nm[i], (double)h[i] / n);
    return 0;
}
