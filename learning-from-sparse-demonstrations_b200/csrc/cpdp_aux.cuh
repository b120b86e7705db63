// Auxiliary-system kernels: backward matrix-Riccati ODE, forward sensitivity ODE, fused loss / dL/dtheta.
// Reference: COCSys.raccatiODE / auxSysODE / auxSysSolver, /root/reference/CPDP/CPDP.py:253-381, and the loss
// closures (e.g. /root/reference/lib/QuadAlgorithm.py:616-639).  The integrators re-implement the control logic of
// scipy.integrate.solve_ivp's RK45 (scipy/integrate/_ivp/rk.py:14-16,111-170; common.py:63-134) so that step
// sequences, and therefore results, follow the reference run: per-interval restart, select_initial_step, RMS
// error norm, SAFETY .9, MIN/MAX_FACTOR .2/10.  One CTA per problem.
//
// State layout.  P is symmetric along the exact solution and every operation applied to it here preserves that
// bit-for-bit, so only its upper triangle is stored:  y = [ P_ij (i<=j, row-major) | W (NX x NP row-major) ].
// Norms weight the off-diagonal entries twice and divide by NX*NX + NX*NP, i.e. they equal the reference's RMS
// norm over the full vec(P), vec(W) state.
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

// ------------------------------------------------------------------------------------------------
// Shared state of one problem's auxiliary-system integration
// ------------------------------------------------------------------------------------------------
struct AuxShared {
    double* M;        // [NSLOT][MSZ]
    double* xul;      // [NSLOT][2NX+NU]
    double* red;      // [AUX_THREADS+1]
    double* P;        // [NX*NX]
    double* Y;        // [NU*NX]
    double* Yp;       // [NU*NX]
    double* Z;        // [NU*NP]
    int* ti; int* tj; // [NT]
    // sparsity tables of fx, fu, fe copied to shared memory (CSR: rowptr/colidx, CSC: colptr/rowidx)
    const int* fx_rowptr; const int* fx_colidx; const int* fx_colptr; const int* fx_rowidx;
    const int* fu_rowptr; const int* fu_colidx; const int* fu_colptr; const int* fu_rowidx;
    const int* fe_colptr; const int* fe_rowidx;
    // forward only
    double* PWt;      // [NSLOT][NYR]
    double* HY;       // [NSLOT][NU*NX]
    double* HZ;       // [NSLOT][NU*NP]
    double* Uc;       // [NU*NP]
};

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

// weighted RMS norm pieces ------------------------------------------------------------------------
CPDP_D double ric_wgt(const AuxShared& s, int i) { return (i < NT && s.ti[i] != s.tj[i]) ? 2.0 : 1.0; }

// ------------------------------------------------------------------------------------------------
// Riccati right-hand side (CPDP.py:262-274), written without forming A, R, Q:
//   Y = fu'P + Hux,  Z = fu'W + Hue,  Y' = Huu^{-1} Y
//   Pdot = -(Hxx + fx'P + P fx - Y' Huu^{-1} Y)          Wdot = -fx'W - P fe - Hxe + Y'^T Z
// fx, fu, fe enter through their static sparsity tables.
// ------------------------------------------------------------------------------------------------
CPDP_D void riccati_rhs(const AuxShared& s, const double* M, const double* yin, double* ydot) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* fx = M + Model::PMP_FX; const double* fu = M + Model::PMP_FU; const double* fe = M + Model::PMP_FE;
    const double* Hxx = M + Model::PMP_HXX; const double* Hxu = M + Model::PMP_HXU; const double* Hxe = M + Model::PMP_HXE;
    const double* Hue = M + Model::PMP_HUE; const double* Hinv = M + Model::PMP_SIZE;
    const double* Wm = yin + NT;
    CPDP_LOOP for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        s.P[i] = yin[r_ <= c ? tri(r_, c) : tri(c, r_)];
    }
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < NU * NX + NU * NP; i += nt) {
        if (i < NU * NX) {
            const int a = i / NX, j = i % NX;
            double acc = Hxu[j * NU + a];
            CPDP_LOOP for (int p = s.fu_colptr[a]; p < s.fu_colptr[a + 1]; ++p) {
                const int r_ = s.fu_rowidx[p];
                acc += fu[r_ * NU + a] * s.P[r_ * NX + j];
            }
            s.Y[i] = acc;
        } else {
            const int q = i - NU * NX, a = q / NP, k = q % NP;
            double acc = Hue[a * NP + k];
            CPDP_LOOP for (int p = s.fu_colptr[a]; p < s.fu_colptr[a + 1]; ++p) {
                const int r_ = s.fu_rowidx[p];
                acc += fu[r_ * NU + a] * Wm[r_ * NP + k];
            }
            s.Z[q] = acc;
        }
    }
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < NU * NX; i += nt) {
        const int a = i / NX, j = i % NX;
        double acc = 0.0;
        CPDP_LOOP for (int b2 = 0; b2 < NU; ++b2) acc += Hinv[a * NU + b2] * s.Y[b2 * NX + j];
        s.Yp[i] = acc;
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NYR; q += nt) {
        if (q < NT) {
            const int i = s.ti[q], j = s.tj[q];
            double acc = Hxx[i * NX + j];
            CPDP_LOOP for (int p = s.fx_colptr[i]; p < s.fx_colptr[i + 1]; ++p) {
                const int a = s.fx_rowidx[p];
                acc += fx[a * NX + i] * s.P[a * NX + j];
            }
            CPDP_LOOP for (int p = s.fx_colptr[j]; p < s.fx_colptr[j + 1]; ++p) {
                const int a = s.fx_rowidx[p];
                acc += s.P[i * NX + a] * fx[a * NX + j];
            }
            // symmetric evaluation of Y' Hinv Y: average of (i,j) and (j,i) orderings is not needed because
            // sum_a Y[a][i]*Yp[a][j] and sum_a Yp[a][i]*Y[a][j] agree to rounding; use the mean to be exact-symmetric
            double yy = 0.0;
            CPDP_LOOP for (int a = 0; a < NU; ++a) yy += 0.5 * (s.Y[a * NX + i] * s.Yp[a * NX + j] + s.Yp[a * NX + i] * s.Y[a * NX + j]);
            ydot[q] = -(acc - yy);
        } else {
            const int e = q - NT, i = e / NP, k = e % NP;
            double acc = -Hxe[i * NP + k];
            CPDP_LOOP for (int p = s.fx_colptr[i]; p < s.fx_colptr[i + 1]; ++p) {
                const int a = s.fx_rowidx[p];
                acc -= fx[a * NX + i] * Wm[a * NP + k];
            }
            CPDP_LOOP for (int p = s.fe_colptr[k]; p < s.fe_colptr[k + 1]; ++p) {
                const int a = s.fe_rowidx[p];
                acc -= s.P[i * NX + a] * fe[a * NP + k];
            }
            CPDP_LOOP for (int a = 0; a < NU; ++a) acc += s.Yp[a * NX + i] * s.Z[a * NP + k];
            ydot[q] = acc;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Forward auxiliary right-hand side (CPDP.py:295-297) with the time-only factors hoisted into prepare():
//   HY = -Huu^{-1}(fu'P + Hux),  HZ = -Huu^{-1}(fu'W + Hue);   Ua = HY X + HZ;   Xdot = fx X + fu Ua + fe
// ------------------------------------------------------------------------------------------------
CPDP_D void forward_rhs(const AuxShared& s, int slot, const double* Xin, double* Xdot) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* M = s.M + (size_t)slot * MSZ;
    const double* fx = M + Model::PMP_FX; const double* fu = M + Model::PMP_FU; const double* fe = M + Model::PMP_FE;
    const double* HY = s.HY + (size_t)slot * NU * NX; const double* HZ = s.HZ + (size_t)slot * NU * NP;
    CPDP_LOOP for (int i = tid; i < NU * NP; i += nt) {
        const int a = i / NP, k = i % NP;
        double acc = HZ[i];
        CPDP_LOOP for (int c = 0; c < NX; ++c) acc += HY[a * NX + c] * Xin[c * NP + k];
        s.Uc[i] = acc;
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NYF; q += nt) {
        const int i = q / NP, k = q % NP;
        double acc = fe[q];
        CPDP_LOOP for (int p = s.fx_rowptr[i]; p < s.fx_rowptr[i + 1]; ++p) {
            const int a = s.fx_colidx[p];
            acc += fx[i * NX + a] * Xin[a * NP + k];
        }
        CPDP_LOOP for (int p = s.fu_rowptr[i]; p < s.fu_rowptr[i + 1]; ++p) {
            const int a = s.fu_colidx[p];
            acc += fu[i * NU + a] * s.Uc[a * NP + k];
        }
        Xdot[q] = acc;
    }
    __syncthreads();
}

// PMP evaluation for `cnt` times (thread i < cnt handles time i), then (forward only) P,W interpolation and HY/HZ.
template <bool FWD>
CPDP_D bool aux_prepare(const AuxShared& s, const AuxProblem& p, const double* times, int cnt) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    double bad = 0.0;
    // the interpolations (one division each) are spread over the CTA; only the generated model code and the m x m
    // inverse stay on a single thread per slot
    CPDP_LOOP for (int q = tid; q < cnt * (2 * NX + NU); q += nt) {
        const int sl = q / (2 * NX + NU), e = q % (2 * NX + NU);
        s.xul[q] = xul_at(p, times[sl], e);
    }
    __syncthreads();
    if (tid < cnt) {
        if (!pmp_eval(p, s.xul + (size_t)tid * (2 * NX + NU), s.M + (size_t)tid * MSZ, times[tid])) bad = 1.0;
    }
    if (FWD) {
        CPDP_LOOP for (int q = tid; q < cnt * NYR; q += nt) {
            const int sl = q / NYR, e = q % NYR;
            const double t = times[sl];
            const int lo = interp_lo(t, p.dt, p.N);
            s.PWt[q] = interp_val(p.PW[(size_t)lo * NYR + e], p.PW[(size_t)(lo + 1) * NYR + e], p.dt * lo, p.dt * (lo + 1), t);
        }
    }
    bad = block_reduce(bad, s.red, true);
    if (bad != 0.0) return false;
    if (FWD) {
        // Y = fu'P + Hux ; Z = fu'W + Hue   (stored temporarily in HY/HZ), then multiplied by -Hinv
        CPDP_LOOP for (int q = tid; q < cnt * (NU * NX + NU * NP); q += nt) {
            const int sl = q / (NU * NX + NU * NP), i = q % (NU * NX + NU * NP);
            const double* M = s.M + (size_t)sl * MSZ;
            const double* fu = M + Model::PMP_FU;
            const double* PWt = s.PWt + (size_t)sl * NYR;
            if (i < NU * NX) {
                const int a = i / NX, j = i % NX;
                double acc = M[Model::PMP_HXU + j * NU + a];
                CPDP_LOOP for (int pp = s.fu_colptr[a]; pp < s.fu_colptr[a + 1]; ++pp) {
                    const int r_ = s.fu_rowidx[pp];
                    acc += fu[r_ * NU + a] * PWt[r_ <= j ? tri(r_, j) : tri(j, r_)];
                }
                s.HY[(size_t)sl * NU * NX + i] = acc;
            } else {
                const int e = i - NU * NX, a = e / NP, k = e % NP;
                double acc = M[Model::PMP_HUE + a * NP + k];
                CPDP_LOOP for (int pp = s.fu_colptr[a]; pp < s.fu_colptr[a + 1]; ++pp) {
                    const int r_ = s.fu_rowidx[pp];
                    acc += fu[r_ * NU + a] * PWt[NT + r_ * NP + k];
                }
                s.HZ[(size_t)sl * NU * NP + e] = acc;
            }
        }
        __syncthreads();
        // in-place multiply by -Hinv, one thread per (slot, column)
        CPDP_LOOP for (int q = tid; q < cnt * (NX + NP); q += nt) {
            const int sl = q / (NX + NP), c = q % (NX + NP);
            const double* Hinv = s.M + (size_t)sl * MSZ + Model::PMP_SIZE;
            double col[NU], out[NU];
            if (c < NX) { for (int a = 0; a < NU; ++a) col[a] = s.HY[(size_t)sl * NU * NX + a * NX + c]; }
            else { for (int a = 0; a < NU; ++a) col[a] = s.HZ[(size_t)sl * NU * NP + a * NP + (c - NX)]; }
            CPDP_LOOP for (int a = 0; a < NU; ++a) {
                double acc = 0.0;
                CPDP_LOOP for (int b2 = 0; b2 < NU; ++b2) acc += Hinv[a * NU + b2] * col[b2];
                out[a] = -acc;
            }
            if (c < NX) { for (int a = 0; a < NU; ++a) s.HY[(size_t)sl * NU * NX + a * NX + c] = out[a]; }
            else { for (int a = 0; a < NU; ++a) s.HZ[(size_t)sl * NU * NP + a * NP + (c - NX)] = out[a]; }
        }
        __syncthreads();
    }
    return true;
}

// dynamic shared memory carve-up ---------------------------------------------------------------------
constexpr int RK_COMMON_DOUBLES = NSLOT * MSZ + NSLOT * (2 * NX + NU) + (AUX_THREADS + 1) + NX * NX + 2 * NU * NX + NU * NP;
constexpr int RIC_SMEM_DOUBLES = RK_COMMON_DOUBLES + 3 * NYR + 7 * NYR + 8;
constexpr int FWD_SMEM_DOUBLES = RK_COMMON_DOUBLES + NSLOT * NYR + NSLOT * NU * NX + NSLOT * NU * NP + NU * NP + 3 * NYF + 7 * NYF + 8;

#ifdef __CUDACC__
#define CPDP_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#else
#define CPDP_DYN_SMEM(name) double* name = ::cpdp_emu_dyn_smem
#endif

CPDP_D double* carve(double*& ptr, int n) { double* r_ = ptr; ptr += n; return r_; }

// shared-memory copy of the model's static sparsity tables (divergent lookups are cheap there)
constexpr int SPTAB_INTS = 2 * (NX + 1) + 2 * Model::FX_nnz + (NX + 1) + (NU + 1) + 2 * Model::FU_nnz + (NP + 1) + Model::FE_nnz + 8;
// pointers into the table block (pure address arithmetic: folds to constants when `tab` is a constant address)
CPDP_D void aux_table_ptrs(AuxShared& s, int* tab) {
    int* p = tab;
    s.fx_rowptr = p; p += NX + 1; s.fx_colidx = p; p += Model::FX_nnz; s.fx_colptr = p; p += NX + 1; s.fx_rowidx = p; p += Model::FX_nnz;
    s.fu_rowptr = p; p += NX + 1; s.fu_colidx = p; p += Model::FU_nnz; s.fu_colptr = p; p += NU + 1; s.fu_rowidx = p; p += Model::FU_nnz;
    s.fe_colptr = p; p += NP + 1; s.fe_rowidx = p; p += Model::FE_nnz;
}
CPDP_D void aux_tables(AuxShared& s, int* tab) {
    const int tid = threadIdx.x, nt = blockDim.x;
    aux_table_ptrs(s, tab);
    int* fx_rowptr = (int*)s.fx_rowptr; int* fx_colidx = (int*)s.fx_colidx; int* fx_colptr = (int*)s.fx_colptr; int* fx_rowidx = (int*)s.fx_rowidx;
    int* fu_rowptr = (int*)s.fu_rowptr; int* fu_colidx = (int*)s.fu_colidx; int* fu_colptr = (int*)s.fu_colptr; int* fu_rowidx = (int*)s.fu_rowidx;
    int* fe_colptr = (int*)s.fe_colptr; int* fe_rowidx = (int*)s.fe_rowidx;
    for (int i = tid; i <= NX; i += nt) { fx_rowptr[i] = Model::FX_rowptr(i); fx_colptr[i] = Model::FX_colptr(i); fu_rowptr[i] = Model::FU_rowptr(i); }
    for (int i = tid; i <= NU; i += nt) fu_colptr[i] = Model::FU_colptr(i);
    for (int i = tid; i <= NP; i += nt) fe_colptr[i] = Model::FE_colptr(i);
    for (int i = tid; i < Model::FX_nnz; i += nt) { fx_colidx[i] = Model::FX_colidx(i); fx_rowidx[i] = Model::FX_rowidx(i); }
    for (int i = tid; i < Model::FU_nnz; i += nt) { fu_colidx[i] = Model::FU_colidx(i); fu_rowidx[i] = Model::FU_rowidx(i); }
    for (int i = tid; i < Model::FE_nnz; i += nt) fe_rowidx[i] = Model::FE_rowidx(i);
}

// Shared-memory layout of the RK45 kernels: every array at a compile-time offset of the dynamic block (doubles first,
// then the int tables), so that the out-of-line pieces rebuild their views from constants.
struct RkWork { double* y; double* yn; double* ys; double* K; double* tms; };
constexpr int AUX_SMEM_INTS = 2 * NT + SPTAB_INTS;
template <bool FWD> constexpr size_t rk_smem_bytes() {
    return (size_t)(FWD ? FWD_SMEM_DOUBLES : RIC_SMEM_DOUBLES) * sizeof(double) + (size_t)((AUX_SMEM_INTS + 3) & ~3) * sizeof(int);
}
template <bool FWD>
CPDP_D void rk_layout(double* smem, AuxShared& s, RkWork& w) {
    constexpr int NY = FWD ? NYF : NYR;
    double* ptr = smem;
    s.M = carve(ptr, NSLOT * MSZ);
    s.xul = carve(ptr, NSLOT * (2 * NX + NU));
    s.red = carve(ptr, AUX_THREADS + 1);
    s.P = carve(ptr, NX * NX);
    s.Y = carve(ptr, NU * NX);
    s.Yp = carve(ptr, NU * NX);
    s.Z = carve(ptr, NU * NP);
    if (FWD) {
        s.PWt = carve(ptr, NSLOT * NYR); s.HY = carve(ptr, NSLOT * NU * NX); s.HZ = carve(ptr, NSLOT * NU * NP);
        s.Uc = carve(ptr, NU * NP);
    } else {
        s.PWt = nullptr; s.HY = nullptr; s.HZ = nullptr; s.Uc = nullptr;
    }
    w.y = carve(ptr, NY); w.yn = carve(ptr, NY); w.ys = carve(ptr, NY);
    w.K = carve(ptr, 7 * NY); w.tms = carve(ptr, 8);
    int* ip = (int*)(smem + (FWD ? FWD_SMEM_DOUBLES : RIC_SMEM_DOUBLES));
    s.ti = ip; s.tj = ip + NT;
    aux_table_ptrs(s, ip + 2 * NT);
}
#define RK_LAYOUT(FWD) CPDP_DYN_SMEM(smem); AuxShared s; RkWork w; rk_layout<FWD>(smem, s, w)

// fills ti/tj and the sparsity tables, clears the structural zeros of the PMP slots (all threads; caller syncs)
CPDP_D void aux_shared_fill(const AuxShared& s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int* ti = (int*)s.ti; int* tj = (int*)s.tj;
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {
        int i = 0, rem = q;
        while (rem >= NX - i) { rem -= NX - i; ++i; }
        ti[q] = i; tj[q] = i + rem;
    }
    CPDP_LOOP for (int q = tid; q < NSLOT * MSZ; q += nt) s.M[q] = 0.0;     // structural zeros of the PMP matrices
}

// ------------------------------------------------------------------------------------------------
// RK45 over one grid interval [t0, t1] (either direction), scipy semantics.  y (in/out), K[7][NY] and the work
// vectors live in shared memory.  Returns 0 ok, 1 step too small, 2 non-finite.
// ------------------------------------------------------------------------------------------------
// out-of-line pieces (one copy each per kernel; see the instruction-cache note in cpdp_bdf.cuh)
template <bool FWD>
CPDP_D_NOINLINE bool rk_prepare(const AuxProblem p, int cnt) {
    RK_LAYOUT(FWD);
    return aux_prepare<FWD>(s, p, w.tms, cnt);
}
template <bool FWD>
CPDP_D_NOINLINE void rk_rhs(int slot, const double* yin, double* yout) {
    RK_LAYOUT(FWD);
    if (FWD) forward_rhs(s, slot, yin, yout);
    else riccati_rhs(s, s.M + (size_t)slot * MSZ, yin, yout);
}

template <bool FWD, int NY>
CPDP_D int rk45_interval(const AuxShared& s, const AuxProblem& p, double t0, double t1, double rtol, double atol,
                         double* y, double* yn, double* ys, double* K, double* tms, int& nrhs, int& nsteps) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double NF = FWD ? (double)NYF : (double)NFULL_R;
    auto rhs = [&](int slot, const double* yin, double* yout) {
        rk_rhs<FWD>(slot, yin, yout);
        ++nrhs;
    };
    auto wgt = [&](int i) { return FWD ? 1.0 : ric_wgt(s, i); };
    double* f = K;                       // K[0] holds f(t, y)
    // f0
    if (tid == 0) tms[0] = t0;
    if (!rk_prepare<FWD>(p, 1)) return 2;
    rhs(0, y, f);
    // select_initial_step (common.py:68-134), order = 4
    double h_abs;
    {
        const double interval_length = fabs(t1 - t0);
        double a0 = 0.0, a1 = 0.0;
        CPDP_LOOP for (int i = tid; i < NY; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            a0 += wgt(i) * (y[i] / sc) * (y[i] / sc);
            a1 += wgt(i) * (f[i] / sc) * (f[i] / sc);
        }
        const double d0 = sqrt(block_reduce(a0, s.red, false) / NF);
        const double d1 = sqrt(block_reduce(a1, s.red, false) / NF);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval_length);
        CPDP_LOOP for (int i = tid; i < NY; i += nt) ys[i] = y[i] + h0 * dir * f[i];
        if (tid == 0) tms[0] = t0 + h0 * dir;
        if (!rk_prepare<FWD>(p, 1)) return 2;
        rhs(0, ys, yn);                   // f1 in yn
        double a2 = 0.0;
        CPDP_LOOP for (int i = tid; i < NY; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            const double v = (yn[i] - f[i]) / sc;
            a2 += wgt(i) * v * v;
        }
        const double d2 = sqrt(block_reduce(a2, s.red, false) / NF) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), interval_length);
    }
    double t = t0;
    while (dir * (t - t1) < 0) {
        const double min_step = 10 * fabs(nextafter(t, dir * INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false;
        double t_new = t;
        while (true) {
            if (h_abs < min_step) return 1;
            double h = h_abs * dir;
            t_new = t + h;
            if (dir * (t_new - t1) > 0) t_new = t1;
            h = t_new - t;
            h_abs = fabs(h);
            __syncthreads();
            if (tid < NSLOT) tms[tid] = t + dp_C(tid + 1) * h;
            if (!rk_prepare<FWD>(p, NSLOT)) return 2;
            CPDP_LOOP for (int st = 1; st < 6; ++st) {
                CPDP_LOOP for (int i = tid; i < NY; i += nt) {
                    double acc = 0.0;
                    CPDP_LOOP for (int j = 0; j < st; ++j) acc += K[(size_t)j * NY + i] * dp_A(st, j);
                    ys[i] = y[i] + acc * h;
                }
                __syncthreads();
                rhs(dp_slot(st), ys, K + (size_t)st * NY);
            }
            CPDP_LOOP for (int i = tid; i < NY; i += nt) {
                double acc = 0.0;
                CPDP_LOOP for (int j = 0; j < 6; ++j) acc += K[(size_t)j * NY + i] * dp_B(j);
                yn[i] = y[i] + h * acc;
            }
            __syncthreads();
            rhs(4, yn, K + (size_t)6 * NY);
            double ae = 0.0, fin = 0.0;
            CPDP_LOOP for (int i = tid; i < NY; i += nt) {
                double acc = 0.0;
                CPDP_LOOP for (int j = 0; j < 7; ++j) acc += K[(size_t)j * NY + i] * dp_E(j);
                const double sc = atol + fmax(fabs(y[i]), fabs(yn[i])) * rtol;
                const double v = acc * h / sc;
                ae += wgt(i) * v * v;
                if (!(fabs(yn[i]) < 1e300)) fin = 1.0;
            }
            const double error_norm = sqrt(block_reduce(ae, s.red, false) / NF);
            fin = block_reduce(fin, s.red, true);
            if (fin != 0.0 || !(error_norm == error_norm)) return 2;
            ++nsteps;
            if (error_norm < 1) {
                double factor = (error_norm == 0) ? 10.0 : fmin(10.0, 0.9 * pow(error_norm, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(error_norm, -0.2));
            rejected = true;
        }
        t = t_new;
        CPDP_LOOP for (int i = tid; i < NY; i += nt) { y[i] = yn[i]; K[i] = K[(size_t)6 * NY + i]; }
        __syncthreads();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// k_riccati_rk45: backward sweep, P(T)=hxx, W(T)=hxe, RK45 per interval (CPDP.py:327-336 with method RK45).
// ------------------------------------------------------------------------------------------------
CPDP_GLOBAL void __launch_bounds__(AUX_THREADS) k_riccati_rk45(AuxArgs a) {
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    // The reference never looks at IPOPT's return status (CPDP.py:183); here trajectories that are not a solution
    // at all (NaN / still iterating) are skipped and flagged, max-iter / line-search exits are integrated as they are.
    if (a.solve_status && (a.solve_status[b] == ST_NUMERIC || a.solve_status[b] == ST_RUNNING)) {
        if (tid == 0) a.aux_status[b] = 3;
        return;
    }
    RK_LAYOUT(false);
    aux_shared_fill(s);
    aux_tables(s, (int*)s.ti + 2 * NT);
    double* y = w.y; double* yn = w.yn; double* ys = w.ys; double* K = w.K; double* tms = w.tms;
    double* s_hxx = K; double* s_hxe = K + NX * NX;       // terminal condition staged in the (still unused) stage array
    const int N = a.N;
    AuxProblem p;
    p.X = a.X + (size_t)b * (N + 1) * NX; p.U = a.U + (size_t)b * (N + 1) * NU; p.Lam = a.Lam + (size_t)b * (N + 1) * NX;
    p.th = a.theta + (size_t)b * a.theta_stride; p.pd = a.pdata + (size_t)b * NQ; p.PW = nullptr; p.dt = a.T / N; p.N = N;
    double* PW = a.PW + (size_t)b * (N + 1) * NYR;
    if (tid == 0) {
        // terminal condition at opt_sol(time_grid[-1]) (CPDP.py:327-331); interp1d at the last node returns
        // slope*(x_hi-x_lo)+y_lo of the last interval
        const double tN = p.dt * N;
        double xT[NX];
        const int lo = interp_lo(tN, p.dt, N);
        for (int i = 0; i < NX; ++i) xT[i] = interp_val(p.X[(size_t)lo * NX + i], p.X[(size_t)(lo + 1) * NX + i], p.dt * lo, p.dt * (lo + 1), tN);
        PdBuf pdb;
        Model::term2(xT, p.th, pd_at(p.pd, tN, pdb), s_hxx, s_hxe);
    }
    __syncthreads();
    for (int q = tid; q < NYR; q += nt) {
        const double v = (q < NT) ? 0.5 * (s_hxx[s.ti[q] * NX + s.tj[q]] + s_hxx[s.tj[q] * NX + s.ti[q]]) : s_hxe[q - NT];
        y[q] = v;
        PW[(size_t)N * NYR + q] = v;
    }
    __syncthreads();
    int nrhs = 0, nsteps = 0, st = 0;
    for (int k = N; k >= 1 && st == 0; --k) {
        st = rk45_interval<false, NYR>(s, p, p.dt * k, p.dt * (k - 1), a.rtol_b, a.atol_b, y, yn, ys, K, tms, nrhs, nsteps);
        for (int q = tid; q < NYR; q += nt) PW[(size_t)(k - 1) * NYR + q] = y[q];
        __syncthreads();
    }
    if (tid == 0) { a.aux_status[b] = st; a.counters[b * NCOUNTERS + 0] = nrhs; a.counters[b * NCOUNTERS + 1] = nsteps; }
}

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
