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
#include "cpdp_fwd.cuh"
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
