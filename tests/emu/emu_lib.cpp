// Host emulation of the CUDA kernels for the CPU test-suite (no GPU in the build container).
// Every CUDA thread of a CTA is a std::thread, __syncthreads() is a std::barrier, CTAs run one after another.
// Compile with -fsanitize=thread to have missing barriers reported as data races.
// TEST INFRASTRUCTURE ONLY: the product (lfsd_b200) never loads this library.
#include <barrier>
#include <thread>
#include <vector>
#include <cstdio>

#include CPDP_MODEL_HEADER_PORT
#include CPDP_MODEL_HEADER
#include "cpdp_kernels.cuh"
#include "cpdp_aux.cuh"
#include "cpdp_bdf.cuh"
#include "cpdp_optim.cuh"

thread_local cpdp_emu_dim3 threadIdx;
cpdp_emu_dim3 blockIdx, blockDim, gridDim;
double* cpdp_emu_dyn_smem = nullptr;   // (hidden visibility: private to this library)
static std::barrier<>* g_bar = nullptr;
void cpdp_emu_syncthreads() { g_bar->arrive_and_wait(); }

template <class F, class... Args>
static void emu_launch(F kernel, int grid, int block, size_t smem, Args... args) {
    std::vector<double> dyn(smem / sizeof(double) + 16);
    cpdp_emu_dyn_smem = dyn.data();
    gridDim = {grid, 1, 1};
    blockDim = {block, 1, 1};
    for (int b = 0; b < grid; ++b) {
        blockIdx = {b, 0, 0};
        std::barrier<> bar(block);
        g_bar = &bar;
        std::vector<std::thread> th;
        th.reserve(block);
        for (int t = 0; t < block; ++t)
            th.emplace_back([=] {
                threadIdx = {t, 0, 0};
                kernel(args...);
                g_bar->arrive_and_drop();
            });
        for (auto& x : th) x.join();
    }
}

#define CPDP_LAUNCH(kernel, grid, block, smem, stream, ...) emu_launch(CPDP_NS::kernel, grid, block, smem, __VA_ARGS__)
#define CPDP_READ_INT(dst, src, stream) (dst) = *(src)
#define CPDP_NUM_SMS() 1
#define CPDP_PREPARE_SMEM(kernel, bytes) (void)0
#define CPDP_LAST_ERROR() 0
#include "cpdp_api.inl"

// ---- unit-test hook (emulation only): Schur form, "LU" event and one Newton solve of the BDF kernel on a given L, C
#ifdef CPDP_WITH_BDF
namespace CPDP_NS {
void k_emu_bdf_linear(const double* L, const double* Cm, double c, const double* rhs, double* out_TZ, double* out_sol, int* ok) {
    const int tid = threadIdx.x, nt = blockDim.x;
    BDF_LAYOUT();
    {
        int* s_ti = (int*)s.ti; int* s_tj = (int*)s.tj;
        for (int q = tid; q < NT; q += nt) {
            int i = 0, rem = q;
            while (rem >= NX - i) { rem -= NX - i; ++i; }
            s_ti[q] = i; s_tj[q] = i + rem;
        }
    }
    for (int i = tid; i < NX * NX; i += nt) bs.Lm[i] = L[i];
    for (int i = tid; i < NX * NP; i += nt) bs.Cm[i] = Cm[i];
    for (int i = tid; i < NYR; i += nt) bs.dy[i] = rhs[i];
    __syncthreads();
    bool good = bdf_schur();
    if (good) good = bdf_factor(c);
    if (good) bdf_solve(c);
    __syncthreads();
    for (int i = tid; i < NX * NX; i += nt) {
        out_TZ[i] = bs.Tr[i]; out_TZ[NX * NX + i] = bs.Ti[i]; out_TZ[2 * NX * NX + i] = bs.Zr[i]; out_TZ[3 * NX * NX + i] = (i < 4 * NX) ? bs.ga[i] : 0.0;   // ga | gbr | gbi | pi (contiguous)
        out_TZ[4 * NX * NX + i] = bs.Winv[i];
    }
    for (int i = tid; i < NYR; i += nt) out_sol[i] = bs.dy[i];
    if (tid == 0) *ok = good ? 1 : 0;
}
}  // namespace CPDP_NS

extern "C" CPDP_API int cpdp_emu_bdf_linear(const double* L, const double* Cm, double c, const double* rhs, double* out_TZ, double* out_sol) {
    int ok = 0;
    emu_launch(CPDP_NS::k_emu_bdf_linear, 1, CPDP_NS::BDF_THREADS, CPDP_NS::BDF_SMEM_BYTES,
               L, Cm, c, rhs, out_TZ, out_sol, &ok);
    return ok;
}
#endif
