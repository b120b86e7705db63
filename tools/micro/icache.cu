// Developer micro-benchmark: cycles per instruction of straight-line code of N instructions (16 B each) inside a loop, for
// 1, 2, 4, 8, 16 warps per SM: where do the instruction-cache tiers end on this part and what does a miss cost a lone warp?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu ; run: ./icache
#include <cstdio>
#include <cuda_runtime.h>
#define I1(k) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(b), "r"(c));
#define D1(k) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(db), "d"(dc));
#define I8 I1(0) I1(1) I1(2) I1(3) I1(4) I1(5) I1(6) I1(7)
#define D8 D1(0) D1(1) D1(2) D1(3) D1(4) D1(5) D1(6) D1(7)
#define R64(X) X X X X X X X X
template <int N, bool DBL>
__global__ void k(unsigned* out, long long* cyc, int iters, unsigned b, unsigned c, int phase) {
    unsigned a[8];
    double d[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; d[i] = threadIdx.x + i; }
    const double db = 1.0 + 1e-9 * b, dc = 1e-9 * c;
    // optional de-phasing of the warps: warp w spins w * phase cycles first
    if (phase) { const long long t = clock64() + (long long)(threadIdx.x >> 5) * phase; while (clock64() < t) {} }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N / 64; ++i) {
            if (DBL) { R64(D8) } else { R64(I8) }
        }
    }
    const long long t1 = clock64();
    unsigned s = 0; double sd = 0;
    for (int i = 0; i < 8; ++i) { s += a[i]; sd += d[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (unsigned)sd;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}
template <int N, bool DBL>
void run(unsigned* out, long long* cyc, int phase) {
    const long long total = 1 << 22;                // instructions per warp
    const int iters = (int)(total / N);
    for (int w : {1, 2, 4, 8, 16}) {
        k<N, DBL><<<148, 32 * w>>>(out, cyc, 2, 3, 1, 0);      // warm
        k<N, DBL><<<148, 32 * w>>>(out, cyc, iters, 3, 1, phase);
        cudaDeviceSynchronize();
        long long h[148 * 16]; cudaMemcpy(h, cyc, sizeof(long long) * 148 * w, cudaMemcpyDeviceToHost);
        double m = 0; for (int i = 0; i < 148 * w; ++i) m += h[i];
        m /= 148.0 * w;
        printf("{\"op\": \"%s\", \"instrs\": %d, \"kb\": %d, \"warps_per_sm\": %d, \"phase\": %d, \"cycles_per_instr_per_warp\": %.3f, \"sm_ipc\": %.3f}\n",
               DBL ? "dfma" : "imad", N, N * 16 / 1024, w, phase, m / ((double)iters * N), w * ((double)iters * N) / m);
    }
}
int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
    run<ICN, false>(out, cyc, 0); run<ICN, false>(out, cyc, 997); run<ICN, true>(out, cyc, 0);
    return 0;
}
