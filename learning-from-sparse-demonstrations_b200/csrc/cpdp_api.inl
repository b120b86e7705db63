// Host-side orchestration behind the C ABI declared in include/cpdp.h.  Included by cpdp_lib.cu (nvcc, the shipped
// library) and by tests/emu/emu_lib.cpp (g++, thread-per-CUDA-thread emulation used only by the CPU test-suite).
// The includer provides:
//   CPDP_LAUNCH(kernel, grid, block, smem_bytes, stream, ...)   launch
//   CPDP_READ_INT(dst_host_int, src_dev_ptr, stream)            blocking read of one device int
//   CPDP_NUM_SMS()                                              multiprocessor count
//   CPDP_PREPARE_SMEM(kernel, bytes)                            opt in to > 48 KB dynamic shared memory
//   CPDP_LAST_ERROR()                                           0 if no launch error
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstdint>

namespace CPDP_NS {

struct WsLayout {
    SolveArgs sa;
    double* PW;
    double* Dbdf;
    size_t bytes;
};

static inline size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

// Carves the caller-provided workspace.  With base == nullptr only the size is computed.
static WsLayout ws_carve(char* base, int B, int N, int S) {
    WsLayout w;
    size_t off = 0;
    auto takeD = [&](size_t n) { double* p = base ? (double*)(base + off) : nullptr; off = al(off + n * sizeof(double)); return p; };
    auto takeI = [&](size_t n) { int* p = base ? (int*)(base + off) : nullptr; off = al(off + n * sizeof(int)); return p; };
    const size_t BN = (size_t)B * N, BN1 = (size_t)B * (N + 1);
    SolveArgs& a = w.sa;
    const size_t BNT = (BN + STAGE_TILE - 1) / STAGE_TILE * STAGE_TILE;     // stage arrays are tiled by STAGE_TILE intervals
    a.xs = takeD(BNT * 4 * S * NX);
    a.mu = takeD(BNT * 4 * S * NX);
    a.AB = takeD(BN * NX * NZ);
    a.H = takeD(BN * NZ * NZ);
    a.gL = takeD(BN * NZ);
    a.dfc = takeD(BN1 * NX);
    a.cost = takeD(BN);
    a.Vs = takeD(BN1 * NX * NX);
    a.vs = takeD(BN1 * NX);
    a.Kf = takeD(BN * NU * NX);
    a.kf = takeD(BN * NU);
    a.gq = takeD(BN * NZ);
    a.dX = takeD(BN1 * NX);
    a.dU = takeD(BN * NU);
    a.lamn = takeD(BN1 * NX);
    a.th_init = takeD(B);
    a.filt = takeD((size_t)B * FILTER_CAP * 2);
    a.nfilt = takeI(B);
    a.dlast = takeD(B);
    a.J = takeD(B);
    a.kkt = takeD(B);
    a.act = takeI(B);
    a.nact = takeI(1);
    w.PW = takeD(BN1 * NYR);
    w.Dbdf = takeD((size_t)B * BDF_WS_DOUBLES);      // k_riccati_bdf: differences array + scale, psi, d rows per problem
    w.bytes = off;
    return w;
}

}  // namespace CPDP_NS

static thread_local int g_last_rounds = 0;

extern "C" {

// Newton rounds launched by the most recent cpdp_solve call of the CALLING THREAD (each round = 4 kernel launches).
CPDP_API int cpdp_last_rounds(void) { return g_last_rounds; }

CPDP_API int cpdp_model_dims(int* n, int* m, int* r, int* q) {
    if (q) *q = CPDP_NS::NQ;
    if (n) *n = CPDP_NS::NX;
    if (m) *m = CPDP_NS::NU;
    if (r) *r = CPDP_NS::NP;
    return 0;
}

CPDP_API int cpdp_riccati_state_dim(void) { return CPDP_NS::NYR; }

// Present when the library contains the as-shipped BDF backward sweep (mode 1 of cpdp_aux).
CPDP_API int cpdp_has_bdf(void) { return 1; }

CPDP_API int cpdp_num_counters(void) { return CPDP_NS::NCOUNTERS; }

CPDP_API size_t cpdp_workspace_bytes(int B, int N, int S) {
    if (B <= 0 || N <= 0 || S <= 0) return 0;
    return CPDP_NS::ws_carve(nullptr, B, N, S).bytes;
}

// Forward optimal-control solve for B problems (COCSys.cocSolver, CPDP.py:92-198).
CPDP_API int cpdp_solve(void* ws, size_t ws_bytes, int B, int N, int S, double T,
               const double* x0, const double* theta, int theta_stride, const double* pdata,
               double tol, int max_iter, int rounds,
               double* X, double* U, double* Lam, int* status, int* iters,
               double* kkt_out, double* cost_out, void* stream) {
    using namespace CPDP_NS;
    if (!ws || B <= 0 || N <= 0 || S <= 0 || !x0 || !theta || !X || !U || !Lam || !status || !iters) return -1;
    if (theta_stride != 0 && theta_stride != NP) return -2;
    WsLayout w = ws_carve((char*)ws, B, N, S);
    if (w.bytes > ws_bytes) return -3;
    SolveArgs a = w.sa;
    if (max_iter < 0 || max_iter >= FILTER_CAP) return -10;      // the per-problem filter holds FILTER_CAP corners (one per iteration at most)
    a.B = B; a.N = N; a.S = S; a.T = T; a.tol = tol; a.max_iter = max_iter;
    if (NQ > 0 && !pdata) return -8;
    a.x0 = x0; a.theta = theta; a.theta_stride = theta_stride; a.pdata = pdata;
    a.X = X; a.U = U; a.Lam = Lam; a.status = status; a.iters = iters;
    if (kkt_out) a.kkt = kkt_out;
    if (cost_out) a.J = cost_out;
    const int sms = CPDP_NUM_SMS();
    cudaStream_t st = (cudaStream_t)stream;
    CPDP_LAUNCH(k_solve_init, sms * 4, 256, 0, st, a);
    CPDP_LAUNCH(k_compact, 1, COMPACT_THREADS, 0, st, a);
    const int total_rounds = (rounds > 0) ? rounds : max_iter + 1;
    g_last_rounds = 0;
    for (int it = 0; it < total_rounds; ++it) {
        ++g_last_rounds;
        CPDP_LAUNCH(k_stage_adjoint, sms * 8, 128, 0, st, a);
        CPDP_LAUNCH(k_stage_hessian, sms * CPDP_HESS_GRID_MULT, HESS_THREADS, 0, st, a);
        CPDP_LAUNCH(k_newton_step, sms * 16, NEWTON_THREADS, 0, st, a);      // persistent: strides over the active list
        CPDP_LAUNCH(k_compact, 1, COMPACT_THREADS, 0, st, a);
        if (rounds <= 0 && it >= 3) {     // adaptive mode: poll the active count (host sync)
            int nact = 0;
            CPDP_READ_INT(nact, a.nact, st);
            if (nact == 0) break;
        }
    }
    return CPDP_LAST_ERROR();
}

// Auxiliary system + loss (COCSys.auxSysSolver, CPDP.py:301-381; loss closures QuadAlgorithm.py:616-639).
// mode 0: backward Riccati sweep with RK45 (rtol_b, atol_b); mode 1: BDF scheme of the as-shipped reference.
// phases: bit 0 = backward Riccati sweep (node table into the workspace), bit 1 = forward sweep + loss.
static int cpdp_aux_impl(void* ws, size_t ws_bytes, int B, int N, int S, double T,
             const double* theta, int theta_stride, const double* pdata,
             const double* X, const double* U, const double* Lam, const int* solve_status,
             int mode, double rtol_b, double atol_b, double rtol_f, double atol_f,
             int W, int D, const int* sel_host, const double* taus, int taus_stride, const double* wp,
             double* Xa, double* Ua, double* loss, double* dtheta, int* aux_status, int* counters, void* stream, int phases) {
    using namespace CPDP_NS;
    if (!ws || B <= 0 || N <= 0 || !theta || !X || !U || !Lam || !Xa || !Ua || !loss || !dtheta || !aux_status || !counters) return -1;
    if (theta_stride != 0 && theta_stride != NP) return -2;
    if (W < 0 || D < 0 || D > MAX_SEL || (W > 0 && (!taus || !wp || !sel_host))) return -4;
    if (W > 0 && taus_stride != 0 && taus_stride != W) return -5;
    if (mode != 0 && mode != 1) return -7;
    WsLayout w = ws_carve((char*)ws, B, N, S);
    if (w.bytes > ws_bytes) return -3;
    AuxArgs a;
    if (NQ > 0 && !pdata) return -8;
    a.B = B; a.N = N; a.T = T; a.theta = theta; a.theta_stride = theta_stride; a.pdata = pdata;
    a.X = X; a.U = U; a.Lam = Lam;
    a.rtol_b = rtol_b; a.atol_b = atol_b; a.rtol_f = rtol_f; a.atol_f = atol_f;
    a.PW = w.PW; a.Dws = w.Dbdf; a.Xa = Xa; a.Ua = Ua;
    a.W = W; a.D = D;
    for (int i = 0; i < MAX_SEL; ++i) a.sel[i] = (i < D) ? sel_host[i] : 0;
    for (int i = 0; i < D; ++i) if (a.sel[i] < 0 || a.sel[i] >= NX) return -6;
    a.taus = taus; a.taus_stride = taus_stride; a.wp = wp;
    a.loss = loss; a.dtheta = dtheta; a.solve_status = solve_status; a.aux_status = aux_status; a.counters = counters;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fwd_bytes = FW_SMEM_BYTES;
    if (phases & 1) {
        if (mode == 0) {
            CPDP_PREPARE_SMEM(k_riccati_rk45, RKW_SMEM_BYTES);
            CPDP_LAUNCH(k_riccati_rk45, B, BDF_THREADS, RKW_SMEM_BYTES, st, a);
        } else {
            size_t bdf_bytes = BDF_SMEM_BYTES;
#ifdef CPDP_BDF_OCCUPANCY_KNOB
            // developer knob (tools/prof_bdf_occupancy.py): pad the dynamic shared memory to limit the resident problems per SM
            if (const char* pad = getenv("CPDP_BDF_SMEM_PAD")) bdf_bytes += (size_t)atol(pad);
#endif
            CPDP_PREPARE_SMEM(k_riccati_bdf, bdf_bytes);
            CPDP_LAUNCH(k_riccati_bdf, B, BDF_THREADS, bdf_bytes, st, a);
        }
    }
    if (phases & 2) {
        CPDP_PREPARE_SMEM(k_aux_forward, fwd_bytes);
        CPDP_LAUNCH(k_aux_forward, B, FW_THREADS, fwd_bytes, st, a);
    }
    return CPDP_LAST_ERROR();
}

CPDP_API int cpdp_aux(void* ws, size_t ws_bytes, int B, int N, int S, double T,
             const double* theta, int theta_stride, const double* pdata,
             const double* X, const double* U, const double* Lam, const int* solve_status,
             int mode, double rtol_b, double atol_b, double rtol_f, double atol_f,
             int W, int D, const int* sel_host, const double* taus, int taus_stride, const double* wp,
             double* Xa, double* Ua, double* loss, double* dtheta, int* aux_status, int* counters, void* stream) {
    return cpdp_aux_impl(ws, ws_bytes, B, N, S, T, theta, theta_stride, pdata, X, U, Lam, solve_status, mode, rtol_b, atol_b,
                         rtol_f, atol_f, W, D, sel_host, taus, taus_stride, wp, Xa, Ua, loss, dtheta, aux_status, counters, stream, 3);
}

// Same arguments as cpdp_aux plus `phases` (1 = backward sweep only, 2 = forward sweep + loss only, 3 = both), so
// that a caller can time or overlap the two sweeps separately.  Phase 2 needs the node table phase 1 left in `ws`.
CPDP_API int cpdp_aux_phases(void* ws, size_t ws_bytes, int B, int N, int S, double T,
             const double* theta, int theta_stride, const double* pdata,
             const double* X, const double* U, const double* Lam, const int* solve_status,
             int mode, double rtol_b, double atol_b, double rtol_f, double atol_f,
             int W, int D, const int* sel_host, const double* taus, int taus_stride, const double* wp,
             double* Xa, double* Ua, double* loss, double* dtheta, int* aux_status, int* counters, void* stream, int phases) {
    if (phases < 1 || phases > 3) return -9;
    return cpdp_aux_impl(ws, ws_bytes, B, N, S, T, theta, theta_stride, pdata, X, U, Lam, solve_status, mode, rtol_b, atol_b,
                         rtol_f, atol_f, W, D, sel_host, taus, taus_stride, wp, Xa, Ua, loss, dtheta, aux_status, counters, stream, phases);
}

// FP64 FMA throughput probe (roofline denominator of bench.py: MEASURED_PEAKS.json carries no fp64 figure).
// Launches one kernel of blocks x 256 threads, each thread running 8 independent chains of `iters` DFMAs;
// flops = blocks * 256 * 8 * iters * 2.  sink: >= 1 double of device memory.
CPDP_API int cpdp_dfma_probe(double* sink, int blocks, int iters, void* stream) {
    if (!sink || blocks <= 0 || iters <= 0) return -1;
    CPDP_LAUNCH(k_dfma_probe, blocks, 256, 0, (cudaStream_t)stream, sink, iters);
    return CPDP_LAST_ERROR();
}

// Fixed-shape binary-tree sum over B rows of [loss | dL/dtheta] -> out[1+NP].  scratch: nextpow2(B)*(1+NP) doubles.
CPDP_API int cpdp_reduce(const double* loss, const double* dtheta, int B, double* scratch, double* out, void* stream) {
    using namespace CPDP_NS;
    if (!loss || !dtheta || !out || !scratch || B <= 0) return -1;
    CPDP_LAUNCH(k_reduce_tree, 1, 256, 0, (cudaStream_t)stream, loss, dtheta, B, scratch, out);
    return CPDP_LAST_ERROR();
}

// Rows [loss | dL/dtheta | bad] (r + 2 doubles per problem) for the cross-GPU all-gather; bad = 1 when the forward solve of
// the problem did not converge (solve_status != converged; may be null) or an auxiliary sweep failed (aux_status != 0).
CPDP_API int cpdp_pack_rows(const double* loss, const double* dtheta, const int* solve_status, const int* aux_status, int B,
                            double* rows, void* stream) {
    using namespace CPDP_NS;
    if (!loss || !dtheta || !rows || B <= 0) return -1;
    const int grid = (int)(((size_t)B * (NP + 2) + 255) / 256);
    CPDP_LAUNCH(k_pack_rows, grid < 1024 ? grid : 1024, 256, 0, (cudaStream_t)stream, loss, dtheta, solve_status, aux_status, B, rows);
    return CPDP_LAST_ERROR();
}

// Fixed binary-tree sum (over the row index) of B rows of C doubles -> out[C]; scratch: nextpow2(B)*C doubles.  With the rows
// of cpdp_pack_rows: out = [sum loss | sum dL/dtheta | number of failed problems].
CPDP_API int cpdp_reduce_rows(const double* rows, int B, int C, double* scratch, double* out, void* stream) {
    using namespace CPDP_NS;
    if (!rows || !out || !scratch || B <= 0 || C <= 0) return -1;
    CPDP_LAUNCH(k_reduce_rows, 1, 256, 0, (cudaStream_t)stream, rows, B, C, scratch, out);
    return CPDP_LAST_ERROR();
}

// Learner step on the device (lib/QuadAlgorithm.py:239-257, 454-578).  phase 0: theta_eval <- evaluation point of the next
// gradient iteration (theta + mu*velocity for Nesterov, theta otherwise).  phase 1: update from red = [loss | dL/dtheta | ..],
// projection theta[0] >= 1e-8, traces, stop rule.  phase 2: Nesterov's true_loss_print_flag second evaluation (records loss
// and applies the stop rule with the gradient at the new point; no update).  With phase 1 followed by phase 2 pass
// defer_close = 1 to phase 1.  method: 0 Vanilla, 1 Nesterov, 2 Adam, 3 Nadam, 4 AMSGrad.
// state[3][r] zero-initialised by the caller; it[2] = {iterations done, stop flag} zero-initialised; param_trace row 0 = theta0.
CPDP_API int cpdp_optim_step(int phase, int method, double lr, double mu, double beta1, double beta2, double eps,
                             double loss_stop, double grad_stop, double* theta, double* theta_eval, double* state,
                             const double* red, int* it, double* loss_trace, double* param_trace, int cap, int defer_close,
                             void* stream) {
    using namespace CPDP_NS;
    if (!theta || !theta_eval || !state || !it || phase < 0 || phase > 2 || method < 0 || method > 4) return -1;
    if (phase > 0 && (!red || !loss_trace || !param_trace || cap <= 0)) return -1;
    OptimArgs a;
    a.method = method; a.lr = lr; a.mu = mu; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.loss_stop = loss_stop; a.grad_stop = grad_stop;
    a.theta = theta; a.theta_eval = theta_eval; a.state = state; a.red = red; a.it = it;
    a.loss_trace = loss_trace; a.param_trace = param_trace; a.cap = cap;
    if (phase == 0) { a.record_only = defer_close ? 1 : 0; CPDP_LAUNCH(k_optim_pre, 1, 32, 0, (cudaStream_t)stream, a); }
    else { a.record_only = (phase == 2) ? 1 : (defer_close ? -1 : 0); CPDP_LAUNCH(k_optim_post, 1, 32, 0, (cudaStream_t)stream, a); }
    return CPDP_LAST_ERROR();
}

}  // extern "C"
