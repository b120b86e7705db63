// Shared pieces of the auxiliary-system kernels: argument block, interp1d rules, Dormand-Prince tableau, the PMP evaluation of one
// stage time, the fixed-tree reduction kernel.
// Reference: COCSys.raccatiODE / auxSysODE / auxSysSolver, /root/reference/CPDP/CPDP.py:253-381, and the loss
// closures (e.g. /root/reference/lib/QuadAlgorithm.py:616-639).  The sweeps themselves: cpdp_bdf.cuh (backward: k_riccati_bdf as
// shipped, k_riccati_rk45 for mode 0) and cpdp_fwd.cuh (forward + loss).  The integrators re-implement the control logic of
// scipy.integrate.solve_ivp (scipy/integrate/_ivp/rk.py:14-16,111-170; bdf.py; common.py:63-134) so that step sequences, and
// therefore results, follow the reference run: per-interval restart, select_initial_step, RMS error norm, SAFETY .9,
// MIN/MAX_FACTOR .2/10.
//
// Node table handed from the backward to the forward sweep: P is symmetric, only its upper triangle is stored,
//     PW[k] = [ P_ij (i<=j, row-major) | W (NX x NP row-major) ].
#pragma once
#include "cpdp_kernels.cuh"

namespace CPDP_NS {

constexpr int NT = NX * (NX + 1) / 2;          // packed upper triangle of P
constexpr int NYR = NT + NX * NP;              // Riccati state
constexpr int NYF = NX * NP;                   // forward state
constexpr int NFULL_R = NX * NX + NX * NP;
constexpr int MSZ = Model::PMP_SIZE + NU * NU; // dense PMP matrices + inverse of Huu
constexpr int NSLOT = 5;                       // distinct stage times of one Dormand-Prince step
#ifndef CPDP_AUX_THREADS
#define CPDP_AUX_THREADS 64
#endif
constexpr int AUX_THREADS = CPDP_AUX_THREADS;
constexpr int MAX_SEL = 16;
constexpr int NCOUNTERS = 6;                    // per-problem counters: back rhs, back steps, fwd rhs, fwd steps, back LU, back Jacobians

struct AuxArgs {
    int B, N;
    double T;
    const double* theta; int theta_stride;
    const double* pdata;   // [B][NQ]
    const double* X; const double* U; const double* Lam;   // [B][N+1][.]
    double rtol_b, atol_b, rtol_f, atol_f;
    double* PW;            // [B][N+1][NYR]   packed Riccati nodes
    double* Dws;           // [B][11][NYR]    BDF differences arrays + scale, psi, d rows (L2-resident workspace; mode 1 only)
    double* Xa;            // [B][N+1][NX*NP] aux state nodes  (dx/dtheta)
    double* Ua;            // [B][N+1][NU*NP] aux control nodes
    int W, D;              // waypoints per problem, observed dims
    int sel[MAX_SEL];      // observed state indices
    const double* taus; int taus_stride;    // [B or 1][W]
    const double* wp;      // [B][W][D]
    double* loss;          // [B]
    double* dtheta;        // [B][NP]
    const int* solve_status;   // [B] (problems that did not converge are skipped; may be null)
    int* aux_status;       // [B]  0 ok, 1 step too small, 2 non-finite
    int* counters;         // [B][NCOUNTERS]
};

CPDP_HD int tri(int i, int j) { return i * NX - (i * (i - 1)) / 2 + (j - i); }   // i <= j

// scipy.interpolate.interp1d(kind='linear') index rule: lo = clip(searchsorted(grid, t, 'left'), 1, N) - 1
CPDP_HD int interp_lo(double t, double dt, int N) {
    int lo = (int)(t / dt);
    if (lo < 0) lo = 0;
    if (lo > N - 1) lo = N - 1;
    while (lo < N - 1 && dt * (lo + 1) < t) ++lo;
    while (lo > 0 && !(dt * lo < t)) --lo;
    return lo;
}
CPDP_HD double interp_val(double ylo, double yhi, double xlo, double xhi, double t) {
    const double slope = (yhi - ylo) / (xhi - xlo);
    return slope * (t - xlo) + ylo;
}

// Gauss-Jordan inverse with partial pivoting of a small matrix.  Every loop has a compile-time trip count and every
// array index is static (the pivot row is swapped in through predicated exchanges), so the 2n^2 work array lives in
// registers instead of local memory: it sits on the single-thread critical path of every ODE step.
template <int n>
CPDP_HD bool inv_small(const double* A, double* Ai) {
    double M[n][2 * n];
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) { M[i][j] = A[i * n + j]; M[i][n + j] = (i == j) ? 1.0 : 0.0; }
    bool ok = true;
#pragma unroll
    for (int c = 0; c < n; ++c) {
        int p = c; double best = fabs(M[c][c]);
#pragma unroll
        for (int i = c + 1; i < n; ++i) if (fabs(M[i][c]) > best) { best = fabs(M[i][c]); p = i; }
        if (!(best > 0.0)) ok = false;
#pragma unroll
        for (int i = c + 1; i < n; ++i) {
            const bool sw = (p == i);
#pragma unroll
            for (int j = 0; j < 2 * n; ++j) { const double a = M[c][j], b = M[i][j]; M[c][j] = sw ? b : a; M[i][j] = sw ? a : b; }
        }
        const double d = 1.0 / M[c][c];
#pragma unroll
        for (int j = 0; j < 2 * n; ++j) M[c][j] *= d;
#pragma unroll
        for (int i = 0; i < n; ++i) if (i != c) {
            const double f = M[i][c];
#pragma unroll
            for (int j = 0; j < 2 * n; ++j) M[i][j] = (f != 0.0) ? M[i][j] - f * M[c][j] : M[i][j];
        }
    }
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) Ai[i * n + j] = M[i][n + j];
    return ok;
}

// Dormand-Prince 5(4) tableau as in scipy/integrate/_ivp/rk.py (class RK45)
CPDP_HD double dp_C(int s) { const double c[6] = {0.0, 0.2, 0.3, 0.8, 8.0 / 9.0, 1.0}; return c[s]; }
CPDP_HD double dp_A(int s, int j) {
    const double a[6][5] = {
        {0, 0, 0, 0, 0},
        {1.0 / 5, 0, 0, 0, 0},
        {3.0 / 40, 9.0 / 40, 0, 0, 0},
        {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
        {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
        {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
    return a[s][j];
}
CPDP_HD double dp_B(int j) { const double b[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}; return b[j]; }
CPDP_HD double dp_E(int j) {
    const double e[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
    return e[j];
}
// which PMP slot a stage uses: stages 1..4 -> slots 0..3, stage 5 and f_new (t+h) -> slot 4
CPDP_HD int dp_slot(int s) { return s <= 4 ? s - 1 : 4; }

struct AuxProblem {
    const double* X; const double* U; const double* Lam; const double* th; const double* pd;
    const double* PW;          // node table (forward sweep)
    double dt; int N;
};

// (x, u, lam) at time t by scipy's linear interp1d rule; element e of [x | u | lam]
CPDP_D double xul_at(const AuxProblem& p, double t, int e) {
    const int lo = interp_lo(t, p.dt, p.N);
    const double xlo = p.dt * lo, xhi = p.dt * (lo + 1);
    if (e < NX) return interp_val(p.X[(size_t)lo * NX + e], p.X[(size_t)(lo + 1) * NX + e], xlo, xhi, t);
    if (e < NX + NU) return interp_val(p.U[(size_t)lo * NU + e - NX], p.U[(size_t)(lo + 1) * NU + e - NX], xlo, xhi, t);
    return interp_val(p.Lam[(size_t)lo * NX + e - NX - NU], p.Lam[(size_t)(lo + 1) * NX + e - NX - NU], xlo, xhi, t);
}
// PMP matrices + inv(Huu) of one slot from its interpolated (x, u, lam); executed by ONE thread
CPDP_D bool pmp_eval(const AuxProblem& p, const double* xul, double* M, const double t) {
    PdBuf pdb;
    Model::pmp(xul, xul + NX, xul + NX + NU, p.th, pd_at(p.pd, t, pdb), M);
    if (Model::HUU_DIAG) {           // every JinEnv model: Huu = diag (control-effort weights); the general inverse stays for user models
        bool ok = true;
        double* Hi = M + Model::PMP_SIZE;
        for (int i = 0; i < NU; ++i) {
            const double d = M[Model::PMP_HUU + i * NU + i];
            if (!(fabs(d) > 0.0)) ok = false;
            for (int j = 0; j < NU; ++j) Hi[i * NU + j] = (i == j) ? 1.0 / d : 0.0;
        }
        return ok;
    }
    return inv_small<NU>(M + Model::PMP_HUU, M + Model::PMP_SIZE);
}

// dynamic shared memory of a kernel (the host emulation hands every CTA its own block)
#ifdef __CUDACC__
#define CPDP_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#else
#define CPDP_DYN_SMEM(name) double* name = ::cpdp_emu_dyn_smem
#endif

// ------------------------------------------------------------------------------------------------
// k_reduce_tree: canonical pairwise (binary-tree over the problem index) sum of rows [loss | dL/dtheta].
// The tree shape depends only on the TOTAL number of rows, never on how they were sharded over GPUs, so the
// result is bit-identical for 1/2/4/8 ranks (rows are all-gathered before this kernel).
// ------------------------------------------------------------------------------------------------
CPDP_GLOBAL void __launch_bounds__(256) k_reduce_tree(const double* loss, const double* dtheta, int B, double* scratch, double* out) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int C = NP + 1;
    int P2 = 1; while (P2 < B) P2 <<= 1;
    for (int q = tid; q < P2 * C; q += nt) {
        const int r_ = q / C, c = q % C;
        scratch[q] = (r_ < B) ? (c == 0 ? loss[r_] : dtheta[(size_t)r_ * NP + c - 1]) : 0.0;
    }
    __syncthreads();
    for (int stride = 1; stride < P2; stride <<= 1) {
        const int pairs = P2 / (2 * stride);
        for (int q = tid; q < pairs * C; q += nt) {
            const int pr = q / C, c = q % C;
            scratch[(size_t)(2 * stride * pr) * C + c] += scratch[(size_t)(2 * stride * pr + stride) * C + c];
        }
        __syncthreads();
    }
    for (int c = tid; c < C; c += nt) out[c] = scratch[c];
}

}  // namespace CPDP_NS
