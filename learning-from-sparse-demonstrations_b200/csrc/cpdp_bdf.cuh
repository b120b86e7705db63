// Backward Riccati sweep with the integrator of the as-shipped reference: COCSys.auxSysSolver calls
// scipy.integrate.solve_ivp(method='BDF') once per grid interval at scipy's default tolerances
// (/root/reference/CPDP/CPDP.py:333-336).  This file re-implements that solver's control logic per problem
// (one CTA per problem), following scipy/integrate/_ivp/bdf.py of the scipy the oracle runs (1.18.1; the control
// flow is unchanged since the 1.6.1 the reference pins):
//   variable-order (1..5) NDF in quasi-constant-step form on the differences array D   bdf.py:314-453
//   change_D / compute_R on every step-size change                                       bdf.py:18-33
//   simplified Newton, <= 4 iterations, rate test, tol = max(10 eps/rtol, min(.03, sqrt(rtol)))   bdf.py:36-75,217
//   select_initial_step with order 1                                                     common.py:68-134
//   RMS error norm over the full vec(P), vec(W) state                                    common.py:63-65
//   LU kept across an error-test rejection, dropped on Newton failure / order change     bdf.py:377-407
//
// One deliberate difference.  scipy approximates the Jacobian of the right-hand side by finite differences
// (num_jac) and factorises the dense (n^2+nr)^2 matrix I - cJ.  The Riccati right-hand side is quadratic, so its
// Jacobian is known in closed form and has Kronecker structure: with L = A' - P R and C = R W - r_,
//     d(Pdot)[dP]     = -(L dP + dP L')          d(Wdot)[dP, dW] = dP C - L dW .
// The Newton systems (I - cJ) dy = b are therefore solved exactly as
//     X + c (L X + X L') = B_P        a linear system in the n(n+1)/2 packed unknowns of the symmetric X
//     (I + c L) dW = B_W + c X C      an n x n system with r right-hand sides,
// i.e. with the same matrix the reference factorises up to its finite-difference error (~1e-8 relative).  Measured
// with scipy itself (jac=closed form vs jac=None on the stored quadrotor run): dL/dtheta moves by 2.2e-7 relative.
#pragma once
#include "cpdp_aux.cuh"

namespace CPDP_NS {

constexpr int BDF_THREADS = 128;
constexpr int BDF_MAX_ORDER = 5;
constexpr int BDF_NEWTON_MAXITER = 4;
constexpr int BDF_NROWS = BDF_MAX_ORDER + 3;
constexpr int BDF_GJ_NP = ((NT + 15) / 16) * 16 > ((NT + 7) / 8) * 8 ? ((NT + 15) / 16) * 16 : ((NT + 7) / 8) * 8;   // = GJ_NP below

struct BdfShared {
    double* G;      // [NT*NT]  LU of the packed operator X -> X + c (L X + X L')
    double* Wl;     // [NX*NX]  LU of I + c L
    double* Lm;     // [NX*NX]  L at the Jacobian point
    double* Cm;     // [NX*NP]  C at the Jacobian point
    double* GH;     // [NX*NU]  scratch: fu Huu^{-1}
    double* Am;     // [NX*NX]  scratch: A
    double* Rm;     // [NX*NX]  scratch: R
    double* D;      // [BDF_NROWS][NYR]
    double* ypred; double* scale; double* psi; double* d; double* y; double* f; double* dy;
    double* RU;     // [6*6]
    double* Winv;   // [NX*NX]  inverse of I + c L
    double* gjbuf;  // [2*GJ_NP] pivot column / pivot row exchange of the Gauss-Jordan steps
    double* tmp;    // [NYR]    scratch of bdf_solve
    int* piv;       // [3*GJ_NP] used flags, perm, kidx
    int* wpiv;      // [NX]
};

CPDP_HD double bdf_kappa(int k) { const double v[6] = {0.0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0.0}; return v[k]; }
CPDP_HD double bdf_gamma(int k) { double g = 0.0; for (int i = 1; i <= k; ++i) g += 1.0 / i; return g; }
CPDP_HD double bdf_alpha(int k) { return (1.0 - bdf_kappa(k)) * bdf_gamma(k); }
CPDP_HD double bdf_error_const(int k) { return bdf_kappa(k) * bdf_gamma(k) + 1.0 / (k + 1); }

// (value, index) arg-max over the CTA, ties to the smaller index; same butterfly on GPU and in the emulation.
CPDP_D int block_argmax(double v, int idx, double* red, double& vmax) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int* redi = (int*)(red + nt);
#ifdef __CUDACC__
    for (int o = 16; o > 0; o >>= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, v, o);
        const int xi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (x > v || (x == v && xi < idx)) { v = x; idx = xi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = v; redi[tid >> 5] = idx; }
    __syncthreads();
    double r = red[0]; int ri = redi[0];
    for (int i = 1; i < (nt >> 5); ++i) if (red[i] > r || (red[i] == r && redi[i] < ri)) { r = red[i]; ri = redi[i]; }
#else
    for (int o = 16; o > 0; o >>= 1) {
        __syncthreads();
        red[tid] = v; redi[tid] = idx;
        __syncthreads();
        const double x = red[tid ^ o]; const int xi = redi[tid ^ o];
        if (x > v || (x == v && xi < idx)) { v = x; idx = xi; }
    }
    __syncthreads();
    red[tid] = v; redi[tid] = idx;
    __syncthreads();
    double r = red[0]; int ri = redi[0];
    for (int i = 32; i < nt; i += 32) if (red[i] > r || (red[i] == r && redi[i] < ri)) { r = red[i]; ri = redi[i]; }
#endif
    vmax = r;
    return ri;
}

// In-place LU with partial pivoting of the row-major n x n matrix A in shared memory by the whole CTA
// (right-looking, one column per step; rows over warps, columns over lanes).  Returns false if singular.
CPDP_D bool block_lu(double* A, const int n, int* piv, double* red) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, wrp = tid >> 5, nw = nt >> 5;
    for (int k = 0; k < n; ++k) {
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < n; i += nt) { const double v = fabs(A[i * n + k]); if (v > best) { best = v; bi = i; } }
        double vmax;
        const int p = block_argmax(best, bi, red, vmax);
        if (!(vmax > 0.0)) return false;
        if (tid == 0) piv[k] = p;
        if (p != k) for (int j = tid; j < n; j += nt) { const double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
        __syncthreads();
        const double inv = 1.0 / A[k * n + k];
        for (int i = k + 1 + tid; i < n; i += nt) A[i * n + k] *= inv;
        __syncthreads();
        for (int i = k + 1 + wrp; i < n; i += nw) {
            const double l = A[i * n + k];
            for (int j = k + 1 + lane; j < n; j += 32) A[i * n + j] -= l * A[k * n + j];
        }
        __syncthreads();
    }
    return true;
}

// Closed-form Jacobian data at (PMP matrices M, packed state yJ):  L = A' - P R,  C = R W - r_
// with A = fx - fu Huu^{-1} Hxu', R = fu Huu^{-1} fu', r_ = fe - fu Huu^{-1} Hue  (CPDP.py:262-270).
CPDP_D void bdf_jacobian(const AuxShared& s, const BdfShared& bs, const double* M, const double* yJ) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* fx = M + Model::PMP_FX; const double* fu = M + Model::PMP_FU; const double* fe = M + Model::PMP_FE;
    const double* Hxu = M + Model::PMP_HXU; const double* Hue = M + Model::PMP_HUE; const double* Hinv = M + Model::PMP_SIZE;
    const double* Wm = yJ + NT;
    __syncthreads();
    for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        s.P[i] = yJ[r_ <= c ? tri(r_, c) : tri(c, r_)];
    }
    for (int i = tid; i < NX * NU; i += nt) {
        const int r_ = i / NU, a = i % NU;
        double acc = 0.0;
        for (int b2 = 0; b2 < NU; ++b2) acc += fu[r_ * NU + b2] * Hinv[b2 * NU + a];
        bs.GH[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        double a1 = fx[i], a2 = 0.0;
        for (int a = 0; a < NU; ++a) { a1 -= bs.GH[r_ * NU + a] * Hxu[c * NU + a]; a2 += bs.GH[r_ * NU + a] * fu[c * NU + a]; }
        bs.Am[i] = a1; bs.Rm[i] = a2;
    }
    __syncthreads();
    for (int i = tid; i < NX * NX + NX * NP; i += nt) {
        if (i < NX * NX) {
            const int r_ = i / NX, a = i % NX;
            double acc = bs.Am[a * NX + r_];
            for (int b2 = 0; b2 < NX; ++b2) acc -= s.P[r_ * NX + b2] * bs.Rm[b2 * NX + a];
            bs.Lm[i] = acc;
        } else {
            const int e = i - NX * NX, r_ = e / NP, k = e % NP;
            double acc = -fe[e];
            for (int a = 0; a < NU; ++a) acc += bs.GH[r_ * NU + a] * Hue[a * NP + k];
            for (int b2 = 0; b2 < NX; ++b2) acc += bs.Rm[r_ * NX + b2] * Wm[b2 * NP + k];
            bs.Cm[e] = acc;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Inverse of the packed Newton operator  X -> X + c (L X + X L')  (NT x NT, NT = n(n+1)/2) held in REGISTERS.
// The CTA is an 8 x 16 grid of threads (thread t: row group tr = t % 8, column group tc = t / 8); element (i, j)
// lives in thread (i % 8, j % 16), local slot (i / 8, j / 16).  In-place Gauss-Jordan with implicit partial pivoting:
// rows are never swapped, the pivot row of step k is only marked as used.  Per step the pivot column and the pivot row
// travel through shared memory (2 barriers), the rank-1 update runs on register tiles (FP64-pipe bound instead of
// shared-memory bound).  The result is written to shared memory with both permutations folded in, so that a Newton
// solve is one plain matrix-vector product.
// ------------------------------------------------------------------------------------------------
constexpr int GJ_TR = 8, GJ_TC = 16;
constexpr int GJ_RT = (NT + GJ_TR - 1) / GJ_TR;        // rows per thread
constexpr int GJ_CT = (NT + GJ_TC - 1) / GJ_TC;        // columns per thread
constexpr int GJ_NP = GJ_RT * GJ_TR > GJ_CT * GJ_TC ? GJ_RT * GJ_TR : GJ_CT * GJ_TC;   // padded extent
static_assert(BDF_THREADS == GJ_TR * GJ_TC, "thread grid of the Gauss-Jordan tiles");
static_assert(GJ_NP == BDF_GJ_NP, "padded extent");

// coefficient of X_uv (u <= v) in row (i <= j) of  X + c (L X + X L')  for symmetric X
CPDP_D double gj_entry(const double* Lm, const double c, int i, int j, int u, int v) {
    double acc = 0.0;
    if (j == v) acc += Lm[i * NX + u];
    if (j == u && u != v) acc += Lm[i * NX + v];
    if (i == u) acc += Lm[j * NX + v];
    if (i == v && u != v) acc += Lm[j * NX + u];
    return ((i == u && j == v) ? 1.0 : 0.0) + c * acc;
}

// arg-max of |col[i]| over rows that are not used yet; every thread returns the same (p, value); ties -> smallest i
CPDP_D int gj_pivot(const double* col, const int* used, double& pval) {
#ifdef __CUDACC__
    const int lane = threadIdx.x & 31;
    double best = -1.0; int bi = 0x7fffffff;
    for (int i = lane; i < NT; i += 32) {
        const double v = used[i] ? -1.0 : fabs(col[i]);
        if (v > best) { best = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, best, o);
        const int xi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (x > best || (x == best && xi < bi)) { best = x; bi = xi; }
    }
#else
    double best = -1.0; int bi = 0x7fffffff;
    for (int i = 0; i < NT; ++i) {
        const double v = used[i] ? -1.0 : fabs(col[i]);
        if (v > best) { best = v; bi = i; }
    }
#endif
    pval = best;
    return bi;
}

// Assemble and invert I - cJ in its structured form: bs.G <- inverse of the packed operator (row-major NT x NT),
// bs.Wl <- inverse of I + c L.
CPDP_D bool bdf_factor(const AuxShared& s, const BdfShared& bs, const double c) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tr = tid % GJ_TR, tc = tid / GJ_TR;
    double* colbuf = bs.gjbuf;                  // [GJ_NP] pivot column of the current step
    double* rowbuf = bs.gjbuf + GJ_NP;          // [GJ_NP] pivot row of the current step
    int* used = bs.piv;                         // [GJ_NP] row already used as a pivot
    int* perm = bs.piv + GJ_NP;                 // [NT]    perm[k] = pivot row of step k
    int* kidx = bs.piv + 2 * GJ_NP;             // [GJ_NP] kidx[i] = step at which row i was the pivot
    double A[GJ_RT][GJ_CT];
    __syncthreads();
    // ---- assemble the register tiles; publish column 0
#pragma unroll
    for (int a = 0; a < GJ_RT; ++a) {
        const int q = tr + GJ_TR * a;
#pragma unroll
        for (int cc = 0; cc < GJ_CT; ++cc) {
            const int col = tc + GJ_TC * cc;
            double v = 0.0;
            if (q < NT && col < NT) v = gj_entry(bs.Lm, c, s.ti[q], s.tj[q], s.ti[col], s.tj[col]);
            A[a][cc] = v;
            if (col == 0) colbuf[q] = v;
        }
    }
    for (int i = tid; i < GJ_NP; i += nt) { used[i] = (i >= NT) ? 1 : 0; rowbuf[i] = 0.0; }
    for (int i = tid; i < NX * NX; i += nt) bs.Wl[i] = ((i / NX == i % NX) ? 1.0 : 0.0) + c * bs.Lm[i];
    __syncthreads();
    bool singular = false;
    for (int k = 0; k < NT; ++k) {
        double pval;
        const int p = gj_pivot(colbuf, used, pval);           // every thread computes the same pivot
        if (!(pval > 0.0)) { singular = true; break; }
        const int ap = p / GJ_TR;
        const bool rowowner = (tr == p % GJ_TR);
        if (rowowner) {
#pragma unroll
            for (int a = 0; a < GJ_RT; ++a) if (a == ap) {
#pragma unroll
                for (int cc = 0; cc < GJ_CT; ++cc) rowbuf[tc + GJ_TC * cc] = A[a][cc];
            }
        }
        double f[GJ_RT];
#pragma unroll
        for (int a = 0; a < GJ_RT; ++a) f[a] = colbuf[tr + GJ_TR * a];
        const double inv = 1.0 / colbuf[p];
        __syncthreads();                                       // rowbuf complete; everybody has read colbuf
        if (tid == 0) { used[p] = 1; perm[k] = p; kidx[p] = k; }
        const int kc = k / GJ_TC, kn = k + 1, knc = kn / GJ_TC;
        const bool colowner = (tc == k % GJ_TC), nextowner = (tc == kn % GJ_TC) && (kn < NT);
        double rinv[GJ_CT];
#pragma unroll
        for (int cc = 0; cc < GJ_CT; ++cc) rinv[cc] = rowbuf[tc + GJ_TC * cc] * inv;
#pragma unroll
        for (int a = 0; a < GJ_RT; ++a) {
            const bool prow = rowowner && (a == ap);
#pragma unroll
            for (int cc = 0; cc < GJ_CT; ++cc) {
                double v = prow ? rinv[cc] : (A[a][cc] - f[a] * rinv[cc]);
                if (colowner && cc == kc) v = prow ? inv : -f[a] * inv;
                A[a][cc] = v;
                if (nextowner && cc == knc) colbuf[tr + GJ_TR * a] = v;      // look-ahead: publish column k+1
            }
        }
        __syncthreads();                                       // colbuf of step k+1 complete; rowbuf free again
    }
    if (singular) return false;
    // ---- write the inverse with both permutations folded in:  Ginv[kidx[i]][perm[j]] = Z[i][j]
#pragma unroll
    for (int a = 0; a < GJ_RT; ++a) {
        const int i = tr + GJ_TR * a;
#pragma unroll
        for (int cc = 0; cc < GJ_CT; ++cc) {
            const int j = tc + GJ_TC * cc;
            if (i < NT && j < NT) bs.G[(size_t)kidx[i] * NT + perm[j]] = A[a][cc];
        }
    }
    // ---- n x n block: LU with partial pivoting, then the explicit inverse (one column per thread)
    if (!block_lu(bs.Wl, NX, bs.wpiv, s.red)) return false;
    if (tid < NX) {
        double col[NX];
        for (int q = 0; q < NX; ++q) col[q] = (q == tid) ? 1.0 : 0.0;
        for (int q = 0; q < NX; ++q) { const int pp = bs.wpiv[q]; if (pp != q) { const double t = col[q]; col[q] = col[pp]; col[pp] = t; } }
        for (int q = 0; q < NX; ++q) {
            double acc = col[q];
            for (int e = 0; e < q; ++e) acc -= bs.Wl[q * NX + e] * col[e];
            col[q] = acc;
        }
        for (int q = NX - 1; q >= 0; --q) {
            double acc = col[q];
            for (int e = q + 1; e < NX; ++e) acc -= bs.Wl[q * NX + e] * col[e];
            col[q] = acc / bs.Wl[q * NX + q];
        }
        for (int q = 0; q < NX; ++q) bs.Winv[q * NX + tid] = col[q];
    }
    __syncthreads();
    return true;
}

// dy <- (I - cJ)^{-1} dy   (dy holds the right-hand side on entry; tmp: NYR doubles of scratch)
CPDP_D void bdf_solve(const AuxShared& s, const BdfShared& bs, const double c, double* dy, double* tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    for (int k = tid; k < NT; k += nt) {                       // X = Ginv * B_P
        const double* g = bs.G + (size_t)k * NT;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int m = 0;
        for (; m + 3 < NT; m += 4) { a0 += g[m] * dy[m]; a1 += g[m + 1] * dy[m + 1]; a2 += g[m + 2] * dy[m + 2]; a3 += g[m + 3] * dy[m + 3]; }
        for (; m < NT; ++m) a0 += g[m] * dy[m];
        tmp[k] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    double* dW = dy + NT;
    for (int e = tid; e < NX * NP; e += nt) {                  // B_W + c X C
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        for (int a = 0; a < NX; ++a) acc += tmp[i <= a ? tri(i, a) : tri(a, i)] * bs.Cm[a * NP + k];
        tmp[NT + e] = dW[e] + c * acc;
    }
    for (int k = tid; k < NT; k += nt) dy[k] = tmp[k];
    __syncthreads();
    for (int e = tid; e < NX * NP; e += nt) {                  // dW = (I + cL)^{-1} (...)
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        for (int a = 0; a < NX; ++a) acc += bs.Winv[i * NX + a] * tmp[NT + a * NP + k];
        dW[e] = acc;
    }
    __syncthreads();
}

// change_D (bdf.py:18-33): D[:order+1] <- (R U)' D[:order+1]
CPDP_D void bdf_change_D(const BdfShared& bs, const int order, const double factor) {
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    if (tid == 0) {
        double R[6][6], U[6][6];
        for (int j = 0; j <= order; ++j) { R[0][j] = 1.0; U[0][j] = 1.0; }
        for (int i = 1; i <= order; ++i) {
            R[i][0] = 0.0; U[i][0] = 0.0;
            for (int j = 1; j <= order; ++j) {
                R[i][j] = R[i - 1][j] * (((double)(i - 1) - factor * j) / i);
                U[i][j] = U[i - 1][j] * (((double)(i - 1) - (double)j) / i);
            }
        }
        for (int i = 0; i <= order; ++i)
            for (int j = 0; j <= order; ++j) {
                double acc = 0.0;
                for (int k = 0; k <= order; ++k) acc += R[i][k] * U[k][j];
                bs.RU[i * 6 + j] = acc;
            }
    }
    __syncthreads();
    for (int q = tid; q < NYR; q += nt) {
        double v[6], o[6];
        for (int i = 0; i <= order; ++i) v[i] = bs.D[(size_t)i * NYR + q];
        for (int i = 0; i <= order; ++i) {
            double acc = 0.0;
            for (int j = 0; j <= order; ++j) acc += bs.RU[j * 6 + i] * v[j];
            o[i] = acc;
        }
        for (int i = 0; i <= order; ++i) bs.D[(size_t)i * NYR + q] = o[i];
    }
    __syncthreads();
}

// RMS norm of v/scale over the FULL (n^2 + n r) state (off-diagonal entries of the packed P count twice)
CPDP_D double bdf_norm(const AuxShared& s, const double* v, const double* scale, const double mul) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double a = 0.0;
    for (int i = tid; i < NYR; i += nt) { const double x = mul * v[i] / scale[i]; a += ric_wgt(s, i) * x * x; }
    return sqrt(block_reduce(a, s.red, false) / (double)NFULL_R);
}

// One grid interval [t0, t1] with scipy's BDF.  y in/out (shared memory).  Returns 0 ok, 1 step too small,
// 2 non-finite, 4 singular Newton matrix.  cnt: [rhs evaluations, steps (accepted), LU factorisations, Jacobians]
CPDP_D int bdf_interval(const AuxShared& s, const BdfShared& bs, const AuxProblem& p, const double t0, const double t1,
                        const double rtol, const double atol, double* y, double* tms, int* cnt) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double EPS = 2.220446049250313e-16;
    double* M0 = s.M;                  // PMP matrices of the current evaluation time (single slot)
    // ---- __init__ (bdf.py:200-257)
    if (tid == 0) tms[0] = t0;
    if (!aux_prepare<false>(s, p, tms, 1)) return 2;
    riccati_rhs(s, M0, y, bs.f); ++cnt[0];
    bdf_jacobian(s, bs, M0, y); ++cnt[3];
    double h_abs;
    {
        const double interval_length = fabs(t1 - t0);
        double a0 = 0.0, a1 = 0.0;
        for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            a0 += ric_wgt(s, i) * (y[i] / sc) * (y[i] / sc);
            a1 += ric_wgt(s, i) * (bs.f[i] / sc) * (bs.f[i] / sc);
        }
        const double d0 = sqrt(block_reduce(a0, s.red, false) / (double)NFULL_R);
        const double d1 = sqrt(block_reduce(a1, s.red, false) / (double)NFULL_R);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval_length);
        for (int i = tid; i < NYR; i += nt) bs.ypred[i] = y[i] + h0 * dir * bs.f[i];
        if (tid == 0) tms[0] = t0 + h0 * dir;
        if (!aux_prepare<false>(s, p, tms, 1)) return 2;
        riccati_rhs(s, M0, bs.ypred, bs.dy); ++cnt[0];
        double a2 = 0.0;
        for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            const double v = (bs.dy[i] - bs.f[i]) / sc;
            a2 += ric_wgt(s, i) * v * v;
        }
        const double d2 = sqrt(block_reduce(a2, s.red, false) / (double)NFULL_R) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 2.0);       // order 1
        h_abs = fmin(fmin(100 * h0, h1), interval_length);
    }
    const double newton_tol = fmax(10 * EPS / rtol, fmin(0.03, sqrt(rtol)));
    for (int i = tid; i < NYR; i += nt) { bs.D[i] = y[i]; bs.D[NYR + i] = bs.f[i] * h_abs * dir; }
    __syncthreads();
    int order = 1, n_equal_steps = 0;
    bool lu_valid = false;
    double c_lu = 0.0;                 // the c the current LU was built with (scipy keeps a stale LU after an error rejection)
    double t = t0;

    while (dir * (t - t1) < 0) {
        // ---- _step_impl (bdf.py:314-453)
        const double min_step = 10 * fabs(nextafter(t, dir * INFINITY) - t);
        if (h_abs < min_step) {
            bdf_change_D(bs, order, min_step / h_abs);
            h_abs = min_step;
            n_equal_steps = 0;
        }
        bool current_jac = false;
        double t_new = t, error_norm = 0.0, safety = 0.0;
        int n_iter = 0;
        while (true) {
            if (h_abs < min_step) return 1;
            double h = h_abs * dir;
            t_new = t + h;
            if (dir * (t_new - t1) > 0) {
                t_new = t1;
                bdf_change_D(bs, order, fabs(t_new - t) / h_abs);
                n_equal_steps = 0;
                lu_valid = false;
            }
            h = t_new - t;
            h_abs = fabs(h);
            double gam[6];
            for (int k = 1; k <= order; ++k) gam[k] = bdf_gamma(k);
            const double al = bdf_alpha(order);
            for (int i = tid; i < NYR; i += nt) {
                double yp = 0.0, ps = 0.0;
                for (int k = 0; k <= order; ++k) yp += bs.D[(size_t)k * NYR + i];
                for (int k = 1; k <= order; ++k) ps += bs.D[(size_t)k * NYR + i] * gam[k];
                bs.ypred[i] = yp;
                bs.scale[i] = atol + rtol * fabs(yp);
                bs.psi[i] = ps / al;
            }
            if (tid == 0) tms[0] = t_new;
            if (!aux_prepare<false>(s, p, tms, 1)) return 2;      // PMP matrices at t_new (every Newton iterate shares them)
            const double c = h / al;
            bool converged = false;
            while (!converged) {
                if (!lu_valid) {
                    if (!bdf_factor(s, bs, c)) return 4;
                    lu_valid = true; c_lu = c; ++cnt[2];
                }
                // ---- solve_bdf_system (bdf.py:36-75)
                for (int i = tid; i < NYR; i += nt) { bs.d[i] = 0.0; bs.y[i] = bs.ypred[i]; }
                __syncthreads();
                double dy_norm_old = -1.0;
                int k = 0;
                for (k = 0; k < BDF_NEWTON_MAXITER; ++k) {
                    riccati_rhs(s, M0, bs.y, bs.f); ++cnt[0];
                    double fin = 0.0;
                    for (int i = tid; i < NYR; i += nt) {
                        if (!(fabs(bs.f[i]) < 1e300)) fin = 1.0;
                        bs.dy[i] = c * bs.f[i] - bs.psi[i] - bs.d[i];
                    }
                    fin = block_reduce(fin, s.red, true);
                    if (fin != 0.0) break;
                    bdf_solve(s, bs, c_lu, bs.dy, bs.tmp);
                    const double dy_norm = bdf_norm(s, bs.dy, bs.scale, 1.0);
                    const bool have_rate = dy_norm_old >= 0.0;
                    const double rate = have_rate ? dy_norm / dy_norm_old : 0.0;
                    if (have_rate && (rate >= 1 || pow(rate, (double)(BDF_NEWTON_MAXITER - k)) / (1 - rate) * dy_norm > newton_tol)) break;
                    for (int i = tid; i < NYR; i += nt) { bs.y[i] += bs.dy[i]; bs.d[i] += bs.dy[i]; }
                    __syncthreads();
                    if (dy_norm == 0 || (have_rate && rate / (1 - rate) * dy_norm < newton_tol)) { converged = true; break; }
                    dy_norm_old = dy_norm;
                }
                n_iter = (k < BDF_NEWTON_MAXITER) ? k + 1 : BDF_NEWTON_MAXITER;
                if (!converged) {
                    if (current_jac) break;
                    bdf_jacobian(s, bs, M0, bs.ypred); ++cnt[3];
                    lu_valid = false;
                    current_jac = true;
                }
            }
            if (!converged) {
                h_abs *= 0.5;
                bdf_change_D(bs, order, 0.5);
                n_equal_steps = 0;
                lu_valid = false;
                continue;
            }
            safety = 0.9 * (2 * BDF_NEWTON_MAXITER + 1) / (double)(2 * BDF_NEWTON_MAXITER + n_iter);
            for (int i = tid; i < NYR; i += nt) bs.scale[i] = atol + rtol * fabs(bs.y[i]);
            __syncthreads();
            error_norm = bdf_norm(s, bs.d, bs.scale, bdf_error_const(order));
            if (!(error_norm == error_norm)) return 2;
            if (error_norm > 1) {
                const double factor = fmax(0.2, safety * pow(error_norm, -1.0 / (order + 1)));
                h_abs *= factor;
                bdf_change_D(bs, order, factor);
                n_equal_steps = 0;
                // LU deliberately kept (bdf.py:404-405)
            } else {
                break;
            }
        }
        ++n_equal_steps;
        ++cnt[1];
        t = t_new;
        // ---- update the differences (bdf.py:417-421)
        for (int i = tid; i < NYR; i += nt) {
            const double dv = bs.d[i];
            bs.D[(size_t)(order + 2) * NYR + i] = dv - bs.D[(size_t)(order + 1) * NYR + i];
            bs.D[(size_t)(order + 1) * NYR + i] = dv;
            for (int k = order; k >= 0; --k) bs.D[(size_t)k * NYR + i] += bs.D[(size_t)(k + 1) * NYR + i];
        }
        __syncthreads();
        if (n_equal_steps < order + 1) continue;
        double error_m_norm = INFINITY, error_p_norm = INFINITY;
        if (order > 1) error_m_norm = bdf_norm(s, bs.D + (size_t)order * NYR, bs.scale, bdf_error_const(order - 1));
        if (order < BDF_MAX_ORDER) error_p_norm = bdf_norm(s, bs.D + (size_t)(order + 2) * NYR, bs.scale, bdf_error_const(order + 1));
        const double fm = pow(error_m_norm, -1.0 / order);
        const double f0 = pow(error_norm, -1.0 / (order + 1));
        const double fp = pow(error_p_norm, -1.0 / (order + 2));
        int delta_order = -1; double fmaxv = fm;          // np.argmax: first maximum
        if (f0 > fmaxv) { fmaxv = f0; delta_order = 0; }
        if (fp > fmaxv) { fmaxv = fp; delta_order = 1; }
        order += delta_order;
        const double factor = fmin(10.0, safety * fmaxv);
        h_abs *= factor;
        bdf_change_D(bs, order, factor);
        n_equal_steps = 0;
        lu_valid = false;
    }
    // solve_ivp(t_eval=[t1]) returns the dense output at the step end = D[0] (bdf.py:462-484)
    for (int i = tid; i < NYR; i += nt) y[i] = bs.D[i];
    __syncthreads();
    return 0;
}

constexpr int BDF_SMEM_DOUBLES = MSZ + (2 * NX + NU) + (2 * BDF_THREADS + 2) + NX * NX + 2 * NU * NX + NU * NP
                                 + NT * NT + 4 * NX * NX + NX * NP + NX * NU + BDF_NROWS * NYR + 7 * NYR + 36 + 8 + NYR
                                 + NX * NX + 2 * BDF_GJ_NP + NYR + (3 * BDF_GJ_NP + NX + 4) / 2 + 2;

// k_riccati_bdf: backward sweep of COCSys.auxSysSolver as shipped (CPDP.py:327-338).
CPDP_GLOBAL void __launch_bounds__(BDF_THREADS) k_riccati_bdf(AuxArgs a) {
    CPDP_DYN_SMEM(smem);
    CPDP_SHARED int s_ti[NT], s_tj[NT], s_tab[SPTAB_INTS];
    CPDP_SHARED double s_hxx[NX * NX], s_hxe[NX * NP];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (a.solve_status && (a.solve_status[b] == ST_NUMERIC || a.solve_status[b] == ST_RUNNING)) {
        if (tid == 0) a.aux_status[b] = 3;
        return;
    }
    double* ptr = smem;
    AuxShared s;
    s.M = carve(ptr, MSZ);                       // one PMP slot: every Newton iterate of a step shares t_new
    s.xul = carve(ptr, 2 * NX + NU);
    s.red = carve(ptr, 2 * BDF_THREADS + 2);
    s.P = carve(ptr, NX * NX);
    s.Y = carve(ptr, NU * NX);
    s.Yp = carve(ptr, NU * NX);
    s.Z = carve(ptr, NU * NP);
    s.ti = s_ti; s.tj = s_tj;
    for (int q = tid; q < NT; q += nt) {
        int i = 0, rem = q;
        while (rem >= NX - i) { rem -= NX - i; ++i; }
        s_ti[q] = i; s_tj[q] = i + rem;
    }
    for (int q = tid; q < MSZ; q += nt) s.M[q] = 0.0;
    aux_tables(s, s_tab);
    BdfShared bs;
    bs.G = carve(ptr, NT * NT); bs.Wl = carve(ptr, NX * NX); bs.Lm = carve(ptr, NX * NX); bs.Am = carve(ptr, NX * NX);
    bs.Rm = carve(ptr, NX * NX); bs.Cm = carve(ptr, NX * NP); bs.GH = carve(ptr, NX * NU);
    bs.D = carve(ptr, BDF_NROWS * NYR);
    bs.ypred = carve(ptr, NYR); bs.scale = carve(ptr, NYR); bs.psi = carve(ptr, NYR); bs.d = carve(ptr, NYR);
    bs.y = carve(ptr, NYR); bs.f = carve(ptr, NYR); bs.dy = carve(ptr, NYR);
    bs.RU = carve(ptr, 36);
    double* tms = carve(ptr, 8);
    double* y = carve(ptr, NYR);
    bs.Winv = carve(ptr, NX * NX); bs.gjbuf = carve(ptr, 2 * BDF_GJ_NP); bs.tmp = carve(ptr, NYR);
    bs.piv = (int*)carve(ptr, (3 * BDF_GJ_NP + NX + 4) / 2 + 2);
    bs.wpiv = bs.piv + 3 * BDF_GJ_NP + 1;
    const int N = a.N;
    AuxProblem p;
    p.X = a.X + (size_t)b * (N + 1) * NX; p.U = a.U + (size_t)b * (N + 1) * NU; p.Lam = a.Lam + (size_t)b * (N + 1) * NX;
    p.th = a.theta + (size_t)b * a.theta_stride; p.pd = a.pdata + (size_t)b * NQ; p.PW = nullptr; p.dt = a.T / N; p.N = N;
    double* PW = a.PW + (size_t)b * (N + 1) * NYR;
    if (tid == 0) {
        const double tN = p.dt * N;
        double xT[NX];
        const int lo = interp_lo(tN, p.dt, N);
        for (int i = 0; i < NX; ++i) xT[i] = interp_val(p.X[(size_t)lo * NX + i], p.X[(size_t)(lo + 1) * NX + i], p.dt * lo, p.dt * (lo + 1), tN);
        Model::term2(xT, p.th, p.pd, s_hxx, s_hxe);
    }
    __syncthreads();
    for (int q = tid; q < NYR; q += nt) {
        const double v = (q < NT) ? 0.5 * (s_hxx[s_ti[q] * NX + s_tj[q]] + s_hxx[s_tj[q] * NX + s_ti[q]]) : s_hxe[q - NT];
        y[q] = v;
        PW[(size_t)N * NYR + q] = v;
    }
    __syncthreads();
    int cnt[4] = {0, 0, 0, 0};
    int st = 0;
    for (int k = N; k >= 1 && st == 0; --k) {
        st = bdf_interval(s, bs, p, p.dt * k, p.dt * (k - 1), a.rtol_b, a.atol_b, y, tms, cnt);
        for (int q = tid; q < NYR; q += nt) PW[(size_t)(k - 1) * NYR + q] = y[q];
        __syncthreads();
    }
    if (tid == 0) {
        a.aux_status[b] = st;
        a.counters[b * NCOUNTERS + 0] = cnt[0]; a.counters[b * NCOUNTERS + 1] = cnt[1];
        a.counters[b * NCOUNTERS + 4] = cnt[2]; a.counters[b * NCOUNTERS + 5] = cnt[3];
    }
}

}  // namespace CPDP_NS
