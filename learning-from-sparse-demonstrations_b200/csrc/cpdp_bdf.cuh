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

// Large, rarely interleaved pieces are kept out of line (CPDP_D_NOINLINE): the fully inlined kernel was 33 k SASS
// instructions (535 KB) and its warps spent ~30 % of their stall samples waiting for instruction fetch (ncu,
// profiles/r01_*schur*); the B200 instruction caches hold 6 KB (L0) / 32 KB (L1.5).

namespace CPDP_NS {

constexpr int BDF_THREADS = 64;
constexpr int BDF_MAX_ORDER = 5;
constexpr int BDF_NEWTON_MAXITER = 4;
constexpr int BDF_NROWS = BDF_MAX_ORDER + 3;

struct BdfShared {
    double* Tr; double* Ti;   // [NX*NX]  Schur form L = Z T Z^H (T upper triangular)
    double* Zr; double* Zi;   // [NX*NX]
    double* Fr; double* Fi;   // [NX*NX]  scratch of the factor / solve steps
    double* Gr; double* Gi;   // [NX*NX]
    double* Dr; double* Di;   // [NT]     1 / ((1/2 + c t_ii) + conj(1/2 + c t_jj))
    double* Lm;     // [NX*NX]  L at the Jacobian point            (aliases Fr: dead once bdf_schur has copied it)
    double* Cm;     // [NX*NP]  C at the Jacobian point
    double* GH;     // [NX*NU]  scratch: fu Huu^{-1}               (aliases Gi)
    double* Am;     // [NX*NX]  scratch: A                         (aliases Fi)
    double* Rm;     // [NX*NX]  scratch: R                         (aliases Gr)
    double* D;      // [BDF_NROWS][NYR]  differences array, GLOBAL memory (L2-resident workspace): element (k, i) is
                    //                   only ever touched by the thread that owns column i (i % blockDim.x == tid)
    double* ypred;  // [NYR]  predictor; between intervals it carries the interval's start / end state
    double* scale; double* psi; double* d; double* y; double* dy;
    double* RU;     // [6*6]
    double* Winv;   // [NX*NX]  inverse of I + c L
    double* tmp;    // [NYR]    scratch of bdf_solve; f(t0, y0) during the start-up of an interval
    int* flag;      // [2]
};

CPDP_HD double bdf_kappa(int k) { const double v[6] = {0.0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0.0}; return v[k]; }
CPDP_HD double bdf_gamma(int k) { double g = 0.0; for (int i = 1; i <= k; ++i) g += 1.0 / i; return g; }
CPDP_HD double bdf_alpha(int k) { return (1.0 - bdf_kappa(k)) * bdf_gamma(k); }
CPDP_HD double bdf_error_const(int k) { return bdf_kappa(k) * bdf_gamma(k) + 1.0 / (k + 1); }

// Shared-memory layout.  Every array sits at a COMPILE-TIME offset of the dynamic shared-memory block, so the
// out-of-line pieces below rebuild their views from constants (no pointer structs in local memory, no registers).
constexpr int BDF_SMEM_DOUBLES = MSZ + (2 * NX + NU) + (2 * BDF_THREADS + 2) + NX * NX + 2 * NU * NX + NU * NP   // AuxShared
                                 + 8 * NX * NX + 2 * NT + NX * NP + NX * NX                                     // Schur data, Cm, Winv
                                 + 6 * NYR + 36 + 8 + NYR + 2;                                                  // work vectors, RU, tms, tmp, flag
constexpr int BDF_SMEM_INTS = 2 * NT + SPTAB_INTS;
constexpr size_t BDF_SMEM_BYTES = (size_t)BDF_SMEM_DOUBLES * sizeof(double) + (size_t)((BDF_SMEM_INTS + 3) & ~3) * sizeof(int);
static_assert(NX * NU <= NX * NX, "GH aliases an NX x NX scratch matrix");

CPDP_D void bdf_layout(double* smem, AuxShared& s, BdfShared& bs, double*& tms) {
    double* ptr = smem;
    s.M = carve(ptr, MSZ);                       // one PMP slot: every Newton iterate of a step shares t_new
    s.xul = carve(ptr, 2 * NX + NU);
    s.red = carve(ptr, 2 * BDF_THREADS + 2);
    s.P = carve(ptr, NX * NX);
    s.Y = carve(ptr, NU * NX);
    s.Yp = carve(ptr, NU * NX);
    s.Z = carve(ptr, NU * NP);
    bs.Tr = carve(ptr, NX * NX); bs.Ti = carve(ptr, NX * NX); bs.Zr = carve(ptr, NX * NX); bs.Zi = carve(ptr, NX * NX);
    bs.Fr = carve(ptr, NX * NX); bs.Fi = carve(ptr, NX * NX); bs.Gr = carve(ptr, NX * NX); bs.Gi = carve(ptr, NX * NX);
    bs.Dr = carve(ptr, NT); bs.Di = carve(ptr, NT);
    bs.Lm = bs.Fr; bs.Am = bs.Fi; bs.Rm = bs.Gr; bs.GH = bs.Gi;      // Jacobian scratch, dead once bdf_schur has copied Lm
    bs.Cm = carve(ptr, NX * NP); bs.Winv = carve(ptr, NX * NX);
    bs.ypred = carve(ptr, NYR); bs.scale = carve(ptr, NYR); bs.psi = carve(ptr, NYR); bs.d = carve(ptr, NYR);
    bs.y = carve(ptr, NYR); bs.dy = carve(ptr, NYR);
    bs.RU = carve(ptr, 36);
    tms = carve(ptr, 8);
    bs.tmp = carve(ptr, NYR);
    bs.flag = (int*)carve(ptr, 2);
    int* ip = (int*)(smem + BDF_SMEM_DOUBLES);
    s.ti = ip; s.tj = ip + NT;
    aux_table_ptrs(s, ip + 2 * NT);
    bs.D = nullptr;
}
#define BDF_LAYOUT() CPDP_DYN_SMEM(smem); AuxShared s; BdfShared bs; double* tms; bdf_layout(smem, s, bs, tms); (void)tms

// out-of-line instances of the shared right-hand side / PMP evaluation (one copy each instead of three)
CPDP_D_NOINLINE void bdf_rhs(const double* yin, double* ydot) { BDF_LAYOUT(); riccati_rhs(s, s.M, yin, ydot); }
CPDP_D_NOINLINE bool bdf_prepare(const AuxProblem p) { BDF_LAYOUT(); return aux_prepare<false>(s, p, tms, 1); }

// Closed-form Jacobian data at (PMP matrices M, packed state yJ):  L = A' - P R,  C = R W - r_
// with A = fx - fu Huu^{-1} Hxu', R = fu Huu^{-1} fu', r_ = fe - fu Huu^{-1} Hue  (CPDP.py:262-270).
CPDP_D_NOINLINE void bdf_jacobian(const double* yJ) {
    BDF_LAYOUT();
    const double* M = s.M;
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* fx = M + Model::PMP_FX; const double* fu = M + Model::PMP_FU; const double* fe = M + Model::PMP_FE;
    const double* Hxu = M + Model::PMP_HXU; const double* Hue = M + Model::PMP_HUE; const double* Hinv = M + Model::PMP_SIZE;
    const double* Wm = yJ + NT;
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        s.P[i] = yJ[r_ <= c ? tri(r_, c) : tri(c, r_)];
    }
    CPDP_LOOP for (int i = tid; i < NX * NU; i += nt) {
        const int r_ = i / NU, a = i % NU;
        double acc = 0.0;
        CPDP_LOOP for (int b2 = 0; b2 < NU; ++b2) acc += fu[r_ * NU + b2] * Hinv[b2 * NU + a];
        bs.GH[i] = acc;
    }
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        double a1 = fx[i], a2 = 0.0;
        CPDP_LOOP for (int a = 0; a < NU; ++a) { a1 -= bs.GH[r_ * NU + a] * Hxu[c * NU + a]; a2 += bs.GH[r_ * NU + a] * fu[c * NU + a]; }
        bs.Am[i] = a1; bs.Rm[i] = a2;
    }
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < NX * NX + NX * NP; i += nt) {
        if (i < NX * NX) {
            const int r_ = i / NX, a = i % NX;
            double acc = bs.Am[a * NX + r_];
            CPDP_LOOP for (int b2 = 0; b2 < NX; ++b2) acc -= s.P[r_ * NX + b2] * bs.Rm[b2 * NX + a];
            bs.Lm[i] = acc;
        } else {
            const int e = i - NX * NX, r_ = e / NP, k = e % NP;
            double acc = -fe[e];
            CPDP_LOOP for (int a = 0; a < NU; ++a) acc += bs.GH[r_ * NU + a] * Hue[a * NP + k];
            CPDP_LOOP for (int b2 = 0; b2 < NX; ++b2) acc += bs.Rm[r_ * NX + b2] * Wm[b2 * NP + k];
            bs.Cm[e] = acc;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Newton systems through ONE Schur form per Jacobian (Bartels-Stewart).  With L = Z T Z^H (Z unitary, T upper
// triangular, complex) the packed operator  X -> X + c (L X + X L')  becomes, for Y = Z^H X Z, C = Z^H B Z,
//     (1/2 + c T) Y + Y (1/2 + c T)^H = C          solved entry by entry along anti-diagonals,
// and I + c L = Z (I + c T) Z^H.  Nothing has to be re-factorised when the step size (c) changes: scipy's "LU"
// events only recompute the reciprocals 1 / (1 + c (t_ii + conj t_jj)) and the n x n inverse of I + c L.
// The Schur form is computed by warp 0: Givens reduction to Hessenberg form, Francis double-shift QR in real
// arithmetic (2 x 2 blocks left as they come), then one complex Givens rotation per 2 x 2 block.
// ------------------------------------------------------------------------------------------------
static_assert(NX <= 16, "warp-0 sections map one lane per row/column and 16 + lane per row of Z");
#ifdef __CUDACC__
#define CPDP_W0_SYNC() __syncwarp()
#else
#define CPDP_W0_SYNC() __syncthreads()
#endif

// Real Schur form by warp 0 (every lane runs the same control flow on the same shared-memory values).
// H: in L, out quasi-upper-triangular T (exact zeros below the sub-diagonal); Z: out orthogonal, L = Z T Z'.
CPDP_D bool schur_real_w0(double* H, double* Z) {
    constexpr int n = NX;
    const int lane = threadIdx.x;
    const double EPS = 2.220446049250313e-16;
#define h_(i, j) H[(i) * n + (j)]
    if (lane < 32) for (int i = lane; i < n * n; i += 32) Z[i] = (i / n == i % n) ? 1.0 : 0.0;
    CPDP_W0_SYNC();
    // ---- Hessenberg form by Givens rotations in the planes (i-1, i)
    CPDP_LOOP for (int j = 0; j < n - 2; ++j)
        CPDP_LOOP for (int i = n - 1; i >= j + 2; --i) {
            const double a = h_(i - 1, j), b = h_(i, j);
            if (b == 0.0) continue;
            const double sc = fabs(a) + fabs(b), isc = 1.0 / sc;      // reciprocals: one division per dependent stage
            const double rr = sc * sqrt((a * isc) * (a * isc) + (b * isc) * (b * isc)), irr = 1.0 / rr;
            const double c = a * irr, s = b * irr;
            CPDP_W0_SYNC();
            if (lane >= j && lane < n) {
                const double t1 = h_(i - 1, lane), t2 = h_(i, lane);
                h_(i - 1, lane) = c * t1 + s * t2;
                h_(i, lane) = (lane == j) ? 0.0 : c * t2 - s * t1;
            }
            CPDP_W0_SYNC();
            if (lane < n) {
                const double t1 = h_(lane, i - 1), t2 = h_(lane, i);
                h_(lane, i - 1) = c * t1 + s * t2;
                h_(lane, i) = c * t2 - s * t1;
            } else if (lane >= 16 && lane < 16 + n) {
                const int k = lane - 16;
                const double t1 = Z[k * n + i - 1], t2 = Z[k * n + i];
                Z[k * n + i - 1] = c * t1 + s * t2;
                Z[k * n + i] = c * t2 - s * t1;
            }
            CPDP_W0_SYNC();
        }
    double norm = 0.0;
    CPDP_LOOP for (int i = 0; i < n; ++i) for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) norm += fabs(h_(i, j));
    // ---- Francis double-shift QR sweeps, full Schur form (rows/columns updated over the whole matrix)
    int en = n - 1;
    while (en >= 0) {
        int its = 0;
        while (true) {
            int l;
            CPDP_LOOP for (l = en; l >= 1; --l) {
                double s = fabs(h_(l - 1, l - 1)) + fabs(h_(l, l));
                if (s == 0.0) s = norm;
                if (fabs(h_(l, l - 1)) <= EPS * s) break;
            }
            if (l >= 1) {
                CPDP_W0_SYNC();
                if (lane == 0) h_(l, l - 1) = 0.0;
                CPDP_W0_SYNC();
            }
            if (l == en) { en -= 1; break; }
            if (l == en - 1) { en -= 2; break; }
            if (its >= 600) return false;
            double x = h_(en, en), y = h_(en - 1, en - 1), w = h_(en, en - 1) * h_(en - 1, en);
            // exceptional shift every 14th sweep: blocks holding two nearly identical complex pairs (the x/y symmetry
            // of the quadrotor) converge only linearly under the standard shifts and can need > 100 sweeps
            if (its > 0 && its % 14 == 0) {
                const double s = fabs(h_(en, en - 1)) + fabs(h_(en - 1, en - 2));
                x = y = 0.75 * s + h_(en, en);
                w = -0.4375 * s * s;
            }
            ++its;
            int m;
            double p = 0.0, q = 0.0, r = 0.0;
            CPDP_LOOP for (m = en - 2; m >= l; --m) {
                const double z = h_(m, m), r0 = x - z, s0 = y - z;
                p = (r0 * s0 - w) / h_(m + 1, m) + h_(m, m + 1);
                q = h_(m + 1, m + 1) - z - r0 - s0;
                r = h_(m + 2, m + 1);
                const double s = fabs(p) + fabs(q) + fabs(r);
                if (s != 0.0) { const double is = 1.0 / s; p *= is; q *= is; r *= is; }
                if (m == l) break;
                const double u = fabs(h_(m, m - 1)) * (fabs(q) + fabs(r));
                const double v = fabs(p) * (fabs(h_(m - 1, m - 1)) + fabs(z) + fabs(h_(m + 1, m + 1)));
                if (u <= EPS * v) break;
            }
            CPDP_LOOP for (int k = m; k <= en - 1; ++k) {
                const bool notlast = (k != en - 1);
                double x2 = 0.0;
                if (k != m) {
                    p = h_(k, k - 1); q = h_(k + 1, k - 1); r = notlast ? h_(k + 2, k - 1) : 0.0;
                    x2 = fabs(p) + fabs(q) + fabs(r);
                    if (x2 == 0.0) continue;
                    const double ix2 = 1.0 / x2;
                    p *= ix2; q *= ix2; r *= ix2;
                }
                double s = sqrt(p * p + q * q + r * r);
                if (s == 0.0) continue;
                if (p < 0.0) s = -s;
                CPDP_W0_SYNC();
                if (lane == 0) {
                    if (k != m) { h_(k, k - 1) = -s * x2; h_(k + 1, k - 1) = 0.0; if (notlast) h_(k + 2, k - 1) = 0.0; }
                    else if (l != m) h_(k, k - 1) = -h_(k, k - 1);
                }
                p += s;
                const double is = 1.0 / s, ip = 1.0 / p;
                const double xx = p * is, yy = q * is, zz = r * is;
                q *= ip; r *= ip;
                if (lane >= k && lane < n) {
                    double pp = h_(k, lane) + q * h_(k + 1, lane);
                    if (notlast) { pp += r * h_(k + 2, lane); h_(k + 2, lane) -= pp * zz; }
                    h_(k, lane) -= pp * xx;
                    h_(k + 1, lane) -= pp * yy;
                }
                CPDP_W0_SYNC();
                const int imax = (en < k + 3) ? en : k + 3;
                if (lane <= imax) {
                    double pp = xx * h_(lane, k) + yy * h_(lane, k + 1);
                    if (notlast) { pp += zz * h_(lane, k + 2); h_(lane, k + 2) -= pp * r; }
                    h_(lane, k) -= pp;
                    h_(lane, k + 1) -= pp * q;
                } else if (lane >= 16 && lane < 16 + n) {
                    const int i = lane - 16;
                    double pp = xx * Z[i * n + k] + yy * Z[i * n + k + 1];
                    if (notlast) { pp += zz * Z[i * n + k + 2]; Z[i * n + k + 2] -= pp * r; }
                    Z[i * n + k] -= pp;
                    Z[i * n + k + 1] -= pp * q;
                }
                CPDP_W0_SYNC();
            }
        }
    }
#undef h_
    return true;
}

// Complex Schur form L = Z T Z^H from bs.Lm, by warp 0; result in (Tr,Ti), (Zr,Zi).  Returns false (uniformly over
// the CTA) if the QR iteration did not converge.
CPDP_D_NOINLINE bool bdf_schur() {
    BDF_LAYOUT();
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    CPDP_LOOP for (int i = tid; i < n * n; i += nt) { bs.Tr[i] = bs.Lm[i]; bs.Ti[i] = 0.0; bs.Zi[i] = 0.0; }
    if (tid == 0) bs.flag[0] = 1;
    __syncthreads();
#ifdef __CUDACC__
    if (tid < 32)
#endif
    {
        const int lane = tid;
        const bool ok = schur_real_w0(bs.Tr, bs.Zr);
        CPDP_W0_SYNC();
        if (!ok && lane == 0) bs.flag[0] = 0;
        // ---- one unitary rotation per 2 x 2 block:  G = [[c, s], [-conj(s), c]],  T <- G T G^H,  Z <- Z G^H
        CPDP_LOOP for (int j = 0; ok && j < n - 1; ++j) {
            const double cc = bs.Tr[(j + 1) * n + j];
            if (cc == 0.0) continue;
            const double a = bs.Tr[j * n + j], b = bs.Tr[j * n + j + 1], d = bs.Tr[(j + 1) * n + j + 1];
            const double hd = 0.5 * (a - d), disc = hd * hd + b * cc;
            double v1r, v1i;                                           // eigenvector [lambda - d, cc]
            if (disc >= 0.0) { v1r = hd + (hd >= 0.0 ? sqrt(disc) : -sqrt(disc)); v1i = 0.0; }
            else { v1r = hd; v1i = sqrt(-disc); }
            const double av1 = sqrt(v1r * v1r + v1i * v1i);
            const double rho = sqrt(av1 * av1 + cc * cc);
            double cr, sr, si;
            if (av1 == 0.0) { cr = 0.0; sr = (cc >= 0.0) ? 1.0 : -1.0; si = 0.0; }
            else { cr = av1 / rho; sr = v1r * cc / (av1 * rho); si = v1i * cc / (av1 * rho); }
            CPDP_W0_SYNC();
            if (lane >= j && lane < n) {                               // rows j, j+1
                const int k = lane;
                const double xr = bs.Tr[j * n + k], xi = bs.Ti[j * n + k], yr = bs.Tr[(j + 1) * n + k], yi = bs.Ti[(j + 1) * n + k];
                bs.Tr[j * n + k] = cr * xr + (sr * yr - si * yi);
                bs.Ti[j * n + k] = cr * xi + (sr * yi + si * yr);
                bs.Tr[(j + 1) * n + k] = cr * yr - (sr * xr + si * xi);
                bs.Ti[(j + 1) * n + k] = cr * yi - (sr * xi - si * xr);
            }
            CPDP_W0_SYNC();
            if (lane <= j + 1) {                                       // columns j, j+1 of T
                const int k = lane;
                const double xr = bs.Tr[k * n + j], xi = bs.Ti[k * n + j], yr = bs.Tr[k * n + j + 1], yi = bs.Ti[k * n + j + 1];
                bs.Tr[k * n + j] = cr * xr + (sr * yr + si * yi);
                bs.Ti[k * n + j] = cr * xi + (sr * yi - si * yr);
                bs.Tr[k * n + j + 1] = cr * yr - (sr * xr - si * xi);
                bs.Ti[k * n + j + 1] = cr * yi - (sr * xi + si * xr);
            } else if (lane >= 16 && lane < 16 + n) {                  // columns j, j+1 of Z
                const int k = lane - 16;
                const double xr = bs.Zr[k * n + j], xi = bs.Zi[k * n + j], yr = bs.Zr[k * n + j + 1], yi = bs.Zi[k * n + j + 1];
                bs.Zr[k * n + j] = cr * xr + (sr * yr + si * yi);
                bs.Zi[k * n + j] = cr * xi + (sr * yi - si * yr);
                bs.Zr[k * n + j + 1] = cr * yr - (sr * xr - si * xi);
                bs.Zi[k * n + j + 1] = cr * yi - (sr * xi + si * xr);
            }
            CPDP_W0_SYNC();
            if (lane == 0) { bs.Tr[(j + 1) * n + j] = 0.0; bs.Ti[(j + 1) * n + j] = 0.0; }
            CPDP_W0_SYNC();
        }
    }
    __syncthreads();
    return bs.flag[0] != 0;
}

// scipy's "LU" event for a new c: reciprocals of the Lyapunov pivots and Winv = (I + c L)^{-1} = Re(Z (I + c T)^{-1} Z^H).
CPDP_D_NOINLINE bool bdf_factor(const double c) {
    BDF_LAYOUT();
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    double bad = 0.0;
    if (tid < n) {                                   // column tid of S = (I + c T)^{-1}  (upper triangular) -> (Fr, Fi)
        const int j = tid;
        CPDP_LOOP for (int i = j; i >= 0; --i) {
            double nr = (i == j) ? 1.0 : 0.0, ni = 0.0;
            CPDP_LOOP for (int k = i + 1; k <= j; ++k) {
                const double tr = c * bs.Tr[i * n + k], ti = c * bs.Ti[i * n + k];
                const double sr = bs.Fr[k * n + j], si = bs.Fi[k * n + j];
                nr -= tr * sr - ti * si;
                ni -= tr * si + ti * sr;
            }
            const double dr = 1.0 + c * bs.Tr[i * n + i], di = c * bs.Ti[i * n + i];
            const double dd = dr * dr + di * di;
            if (!(dd > 0.0)) bad = 1.0;
            bs.Fr[i * n + j] = (nr * dr + ni * di) / dd;
            bs.Fi[i * n + j] = (ni * dr - nr * di) / dd;
        }
        CPDP_LOOP for (int i = j + 1; i < n; ++i) { bs.Fr[i * n + j] = 0.0; bs.Fi[i * n + j] = 0.0; }
    }
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // 1 / ((1/2 + c t_ii) + conj(1/2 + c t_jj))
        const int i = s.ti[q], j = s.tj[q];
        const double dr = 1.0 + c * (bs.Tr[i * n + i] + bs.Tr[j * n + j]), di = c * (bs.Ti[i * n + i] - bs.Ti[j * n + j]);
        const double dd = dr * dr + di * di;
        if (!(dd > 0.0)) bad = 1.0;
        bs.Dr[q] = dr / dd; bs.Di[q] = -di / dd;
    }
    bad = block_reduce(bad, s.red, true);
    if (bad != 0.0) return false;
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // G = Z S
        const int i = e / n, k = e % n;
        double gr = 0.0, gi = 0.0;
        CPDP_LOOP for (int j = 0; j <= k; ++j) {
            const double zr = bs.Zr[i * n + j], zi = bs.Zi[i * n + j], sr = bs.Fr[j * n + k], si = bs.Fi[j * n + k];
            gr += zr * sr - zi * si;
            gi += zr * si + zi * sr;
        }
        bs.Gr[e] = gr; bs.Gi[e] = gi;
    }
    __syncthreads();
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // Winv = Re(G Z^H)
        const int i = e / n, l = e % n;
        double acc = 0.0;
        CPDP_LOOP for (int k = 0; k < n; ++k) acc += bs.Gr[i * n + k] * bs.Zr[l * n + k] + bs.Gi[i * n + k] * bs.Zi[l * n + k];
        bs.Winv[e] = acc;
    }
    __syncthreads();
    return true;
}

// dy <- (I - cJ)^{-1} dy   (dy holds the right-hand side on entry; tmp: NYR doubles of scratch)
CPDP_D_NOINLINE void bdf_solve(const double c) {
    BDF_LAYOUT();
    double* dy = bs.dy; double* tmp = bs.tmp;
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // F = B Z, B = sym(dy[0:NT])
        const int i = e / n, k = e % n;
        double fr = 0.0, fi = 0.0;
        CPDP_LOOP for (int j = 0; j < n; ++j) {
            const double b = dy[i <= j ? tri(i, j) : tri(j, i)];
            fr += b * bs.Zr[j * n + k]; fi += b * bs.Zi[j * n + k];
        }
        bs.Fr[e] = fr; bs.Fi[e] = fi;
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // C = Z^H F (upper triangle) -> (Gr, Gi)
        const int i = s.ti[q], j = s.tj[q];
        double cr = 0.0, ci = 0.0;
        CPDP_LOOP for (int k = 0; k < n; ++k) {
            const double zr = bs.Zr[k * n + i], zi = bs.Zi[k * n + i], fr = bs.Fr[k * n + j], fi = bs.Fi[k * n + j];
            cr += zr * fr + zi * fi;
            ci += zr * fi - zi * fr;
        }
        bs.Gr[i * n + j] = cr; bs.Gi[i * n + j] = ci;
    }
    __syncthreads();
    // ---- (1/2 + cT) Y + Y (1/2 + cT)^H = C along anti-diagonals i + j = d (warp 0; 4 lanes per entry); Y overwrites C,
    //      both triangles are kept (Y is Hermitian)
#ifdef __CUDACC__
    if (tid < 32)
#endif
    {
        const int e = tid >> 2, sub = tid & 3;
        CPDP_LOOP for (int d = 2 * (n - 1); d >= 0; --d) {
            const int ilo = (d > n - 1) ? d - (n - 1) : 0;
            const int i = ilo + e, j = d - i;
            const bool valid = (tid < 32) && (i <= j);
            double ar = 0.0, ai = 0.0;
            if (valid) {
                CPDP_LOOP for (int k = i + 1 + sub; k < n; k += 4) {           // T_ik Y_kj
                    const double tr = bs.Tr[i * n + k], ti = bs.Ti[i * n + k], yr = bs.Gr[k * n + j], yi = bs.Gi[k * n + j];
                    ar += tr * yr - ti * yi;
                    ai += tr * yi + ti * yr;
                }
                CPDP_LOOP for (int k = j + 1 + sub; k < n; k += 4) {           // Y_ik conj(T_jk)
                    const double tr = bs.Tr[j * n + k], ti = bs.Ti[j * n + k], yr = bs.Gr[i * n + k], yi = bs.Gi[i * n + k];
                    ar += yr * tr + yi * ti;
                    ai += yi * tr - yr * ti;
                }
            }
#ifdef __CUDACC__
            ar += __shfl_xor_sync(0xffffffffu, ar, 1); ai += __shfl_xor_sync(0xffffffffu, ai, 1);
            ar += __shfl_xor_sync(0xffffffffu, ar, 2); ai += __shfl_xor_sync(0xffffffffu, ai, 2);
#else
            __syncthreads();
            if (tid < 32) { s.red[tid] = ar; s.red[32 + tid] = ai; }
            __syncthreads();
            if (tid < 32 && sub == 0) {
                ar = (s.red[tid] + s.red[tid + 1]) + (s.red[tid + 2] + s.red[tid + 3]);
                ai = (s.red[32 + tid] + s.red[32 + tid + 1]) + (s.red[32 + tid + 2] + s.red[32 + tid + 3]);
            }
#endif
            if (valid && sub == 0) {
                const double rr = bs.Gr[i * n + j] - c * ar, ri = bs.Gi[i * n + j] - c * ai;
                const int q = tri(i, j);
                const double yr = rr * bs.Dr[q] - ri * bs.Di[q], yi = rr * bs.Di[q] + ri * bs.Dr[q];
                bs.Gr[i * n + j] = yr; bs.Gi[i * n + j] = (i == j) ? 0.0 : yi;
                if (i != j) { bs.Gr[j * n + i] = yr; bs.Gi[j * n + i] = -yi; }
            }
            CPDP_W0_SYNC();
        }
    }
    __syncthreads();
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // F = Z Y
        const int i = e / n, k = e % n;
        double fr = 0.0, fi = 0.0;
        CPDP_LOOP for (int j = 0; j < n; ++j) {
            const double zr = bs.Zr[i * n + j], zi = bs.Zi[i * n + j], yr = bs.Gr[j * n + k], yi = bs.Gi[j * n + k];
            fr += zr * yr - zi * yi;
            fi += zr * yi + zi * yr;
        }
        bs.Fr[e] = fr; bs.Fi[e] = fi;
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // X = Re(F Z^H), upper triangle
        const int i = s.ti[q], l = s.tj[q];
        double acc = 0.0;
        CPDP_LOOP for (int k = 0; k < n; ++k) acc += bs.Fr[i * n + k] * bs.Zr[l * n + k] + bs.Fi[i * n + k] * bs.Zi[l * n + k];
        tmp[q] = acc;
    }
    __syncthreads();
    double* dW = dy + NT;
    CPDP_LOOP for (int e = tid; e < NX * NP; e += nt) {                  // B_W + c X C
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        CPDP_LOOP for (int a = 0; a < NX; ++a) acc += tmp[i <= a ? tri(i, a) : tri(a, i)] * bs.Cm[a * NP + k];
        tmp[NT + e] = dW[e] + c * acc;
    }
    CPDP_LOOP for (int k = tid; k < NT; k += nt) dy[k] = tmp[k];
    __syncthreads();
    CPDP_LOOP for (int e = tid; e < NX * NP; e += nt) {                  // dW = (I + cL)^{-1} (...)
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        CPDP_LOOP for (int a = 0; a < NX; ++a) acc += bs.Winv[i * NX + a] * tmp[NT + a * NP + k];
        dW[e] = acc;
    }
    __syncthreads();
}

// change_D (bdf.py:18-33): D[:order+1] <- (R U)' D[:order+1]
CPDP_D_NOINLINE void bdf_change_D(double* D, const int order, const double factor) {
    BDF_LAYOUT();
    bs.D = D;
    const int tid = threadIdx.x, nt = blockDim.x;
    __syncthreads();
    if (tid == 0) {
        double R[6][6], U[6][6];
        CPDP_LOOP for (int j = 0; j <= order; ++j) { R[0][j] = 1.0; U[0][j] = 1.0; }
        CPDP_LOOP for (int i = 1; i <= order; ++i) {
            R[i][0] = 0.0; U[i][0] = 0.0;
            CPDP_LOOP for (int j = 1; j <= order; ++j) {
                R[i][j] = R[i - 1][j] * (((double)(i - 1) - factor * j) / i);
                U[i][j] = U[i - 1][j] * (((double)(i - 1) - (double)j) / i);
            }
        }
        CPDP_LOOP for (int i = 0; i <= order; ++i)
            CPDP_LOOP for (int j = 0; j <= order; ++j) {
                double acc = 0.0;
                CPDP_LOOP for (int k = 0; k <= order; ++k) acc += R[i][k] * U[k][j];
                bs.RU[i * 6 + j] = acc;
            }
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NYR; q += nt) {
        double v[6], o[6];
        CPDP_LOOP for (int i = 0; i <= order; ++i) v[i] = bs.D[(size_t)i * NYR + q];
        CPDP_LOOP for (int i = 0; i <= order; ++i) {
            double acc = 0.0;
            CPDP_LOOP for (int j = 0; j <= order; ++j) acc += bs.RU[j * 6 + i] * v[j];
            o[i] = acc;
        }
        CPDP_LOOP for (int i = 0; i <= order; ++i) bs.D[(size_t)i * NYR + q] = o[i];
    }
    __syncthreads();
}

// RMS norm of v/scale over the FULL (n^2 + n r) state (off-diagonal entries of the packed P count twice)
CPDP_D_NOINLINE double bdf_norm(const double* v, const double* scale, const double mul) {
    BDF_LAYOUT();
    const int tid = threadIdx.x, nt = blockDim.x;
    double a = 0.0;
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { const double x = mul * v[i] / scale[i]; a += ric_wgt(s, i) * x * x; }
    return sqrt(block_reduce(a, s.red, false) / (double)NFULL_R);
}

// One grid interval [t0, t1] with scipy's BDF.  y in/out (shared memory).  Returns 0 ok, 1 step too small,
// 2 non-finite, 4 singular Newton matrix.  cnt: [rhs evaluations, steps (accepted), LU factorisations, Jacobians]
CPDP_D int bdf_interval(const AuxShared& s, const BdfShared& bs, const AuxProblem& p, const double t0, const double t1,
                        const double rtol, const double atol, double* y, double* tms, int* cnt) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double EPS = 2.220446049250313e-16;
    // ---- __init__ (bdf.py:200-257)
    if (tid == 0) tms[0] = t0;
    if (!bdf_prepare(p)) return 2;
    double* f0 = bs.tmp;               // f(t0, y0): bdf_solve (the other user of tmp) is not called during start-up
    bdf_rhs(y, f0); ++cnt[0];
    bdf_jacobian(y); ++cnt[3];
    if (!bdf_schur()) return 4;
    double h_abs;
    {
        const double interval_length = fabs(t1 - t0);
        double a0 = 0.0, a1 = 0.0;
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            a0 += ric_wgt(s, i) * (y[i] / sc) * (y[i] / sc);
            a1 += ric_wgt(s, i) * (f0[i] / sc) * (f0[i] / sc);
        }
        const double d0 = sqrt(block_reduce(a0, s.red, false) / (double)NFULL_R);
        const double d1 = sqrt(block_reduce(a1, s.red, false) / (double)NFULL_R);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval_length);
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) bs.y[i] = y[i] + h0 * dir * f0[i];
        if (tid == 0) tms[0] = t0 + h0 * dir;
        if (!bdf_prepare(p)) return 2;
        bdf_rhs(bs.y, bs.dy); ++cnt[0];
        double a2 = 0.0;
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            const double v = (bs.dy[i] - f0[i]) / sc;
            a2 += ric_wgt(s, i) * v * v;
        }
        const double d2 = sqrt(block_reduce(a2, s.red, false) / (double)NFULL_R) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 2.0);       // order 1
        h_abs = fmin(fmin(100 * h0, h1), interval_length);
    }
    const double newton_tol = fmax(10 * EPS / rtol, fmin(0.03, sqrt(rtol)));
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { bs.D[i] = y[i]; bs.D[NYR + i] = f0[i] * h_abs * dir; }
    __syncthreads();
    int order = 1, n_equal_steps = 0;
    bool lu_valid = false;
    double c_lu = 0.0;                 // the c the current LU was built with (scipy keeps a stale LU after an error rejection)
    double t = t0;

    while (dir * (t - t1) < 0) {
        // ---- _step_impl (bdf.py:314-453)
        const double min_step = 10 * fabs(nextafter(t, dir * INFINITY) - t);
        if (h_abs < min_step) {
            bdf_change_D(bs.D, order, min_step / h_abs);
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
                bdf_change_D(bs.D, order, fabs(t_new - t) / h_abs);
                n_equal_steps = 0;
                lu_valid = false;
            }
            h = t_new - t;
            h_abs = fabs(h);
            double gam[6];
            CPDP_LOOP for (int k = 1; k <= order; ++k) gam[k] = bdf_gamma(k);
            const double al = bdf_alpha(order);
            CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
                double yp = 0.0, ps = 0.0;
                CPDP_LOOP for (int k = 0; k <= order; ++k) yp += bs.D[(size_t)k * NYR + i];
                CPDP_LOOP for (int k = 1; k <= order; ++k) ps += bs.D[(size_t)k * NYR + i] * gam[k];
                bs.ypred[i] = yp;
                bs.scale[i] = atol + rtol * fabs(yp);
                bs.psi[i] = ps / al;
            }
            if (tid == 0) tms[0] = t_new;
            if (!bdf_prepare(p)) return 2;      // PMP matrices at t_new (every Newton iterate shares them)
            const double c = h / al;
            bool converged = false;
            while (!converged) {
                if (!lu_valid) {
                    if (!bdf_factor(c)) return 4;
                    lu_valid = true; c_lu = c; ++cnt[2];
                }
                // ---- solve_bdf_system (bdf.py:36-75)
                CPDP_LOOP for (int i = tid; i < NYR; i += nt) { bs.d[i] = 0.0; bs.y[i] = bs.ypred[i]; }
                __syncthreads();
                double dy_norm_old = -1.0;
                int k = 0;
                CPDP_LOOP for (k = 0; k < BDF_NEWTON_MAXITER; ++k) {
                    bdf_rhs(bs.y, bs.dy); ++cnt[0];
                    double fin = 0.0;
                    CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
                        const double fv = bs.dy[i];
                        if (!(fabs(fv) < 1e300)) fin = 1.0;
                        bs.dy[i] = c * fv - bs.psi[i] - bs.d[i];
                    }
                    fin = block_reduce(fin, s.red, true);
                    if (fin != 0.0) break;
                    bdf_solve(c_lu);
                    const double dy_norm = bdf_norm(bs.dy, bs.scale, 1.0);
                    const bool have_rate = dy_norm_old >= 0.0;
                    const double rate = have_rate ? dy_norm / dy_norm_old : 0.0;
                    if (have_rate && (rate >= 1 || pow(rate, (double)(BDF_NEWTON_MAXITER - k)) / (1 - rate) * dy_norm > newton_tol)) break;
                    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { bs.y[i] += bs.dy[i]; bs.d[i] += bs.dy[i]; }
                    __syncthreads();
                    if (dy_norm == 0 || (have_rate && rate / (1 - rate) * dy_norm < newton_tol)) { converged = true; break; }
                    dy_norm_old = dy_norm;
                }
                n_iter = (k < BDF_NEWTON_MAXITER) ? k + 1 : BDF_NEWTON_MAXITER;
                if (!converged) {
                    if (current_jac) break;
                    bdf_jacobian(bs.ypred); ++cnt[3];
                    if (!bdf_schur()) return 4;
                    lu_valid = false;
                    current_jac = true;
                }
            }
            if (!converged) {
                h_abs *= 0.5;
                bdf_change_D(bs.D, order, 0.5);
                n_equal_steps = 0;
                lu_valid = false;
                continue;
            }
            safety = 0.9 * (2 * BDF_NEWTON_MAXITER + 1) / (double)(2 * BDF_NEWTON_MAXITER + n_iter);
            CPDP_LOOP for (int i = tid; i < NYR; i += nt) bs.scale[i] = atol + rtol * fabs(bs.y[i]);
            __syncthreads();
            error_norm = bdf_norm(bs.d, bs.scale, bdf_error_const(order));
            if (!(error_norm == error_norm)) return 2;
            if (error_norm > 1) {
                const double factor = fmax(0.2, safety * pow(error_norm, -1.0 / (order + 1)));
                h_abs *= factor;
                bdf_change_D(bs.D, order, factor);
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
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
            const double dv = bs.d[i];
            bs.D[(size_t)(order + 2) * NYR + i] = dv - bs.D[(size_t)(order + 1) * NYR + i];
            bs.D[(size_t)(order + 1) * NYR + i] = dv;
            CPDP_LOOP for (int k = order; k >= 0; --k) bs.D[(size_t)k * NYR + i] += bs.D[(size_t)(k + 1) * NYR + i];
        }
        __syncthreads();
        if (n_equal_steps < order + 1) continue;
        double error_m_norm = INFINITY, error_p_norm = INFINITY;
        if (order > 1) error_m_norm = bdf_norm(bs.D + (size_t)order * NYR, bs.scale, bdf_error_const(order - 1));
        if (order < BDF_MAX_ORDER) error_p_norm = bdf_norm(bs.D + (size_t)(order + 2) * NYR, bs.scale, bdf_error_const(order + 1));
        const double fm = pow(error_m_norm, -1.0 / order);
        const double f0 = pow(error_norm, -1.0 / (order + 1));
        const double fp = pow(error_p_norm, -1.0 / (order + 2));
        int delta_order = -1; double fmaxv = fm;          // np.argmax: first maximum
        if (f0 > fmaxv) { fmaxv = f0; delta_order = 0; }
        if (fp > fmaxv) { fmaxv = fp; delta_order = 1; }
        order += delta_order;
        const double factor = fmin(10.0, safety * fmaxv);
        h_abs *= factor;
        bdf_change_D(bs.D, order, factor);
        n_equal_steps = 0;
        lu_valid = false;
    }
    // solve_ivp(t_eval=[t1]) returns the dense output at the step end = D[0] (bdf.py:462-484)
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) y[i] = bs.D[i];
    __syncthreads();
    return 0;
}

// k_riccati_bdf: backward sweep of COCSys.auxSysSolver as shipped (CPDP.py:327-338).
CPDP_GLOBAL void __launch_bounds__(BDF_THREADS, 6) k_riccati_bdf(AuxArgs a) {
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (a.solve_status && (a.solve_status[b] == ST_NUMERIC || a.solve_status[b] == ST_RUNNING)) {
        if (tid == 0) a.aux_status[b] = 3;
        return;
    }
    BDF_LAYOUT();
    {
        int* s_ti = (int*)s.ti; int* s_tj = (int*)s.tj;
        CPDP_LOOP for (int q = tid; q < NT; q += nt) {
            int i = 0, rem = q;
            while (rem >= NX - i) { rem -= NX - i; ++i; }
            s_ti[q] = i; s_tj[q] = i + rem;
        }
    }
    CPDP_LOOP for (int q = tid; q < MSZ; q += nt) s.M[q] = 0.0;
    aux_tables(s, (int*)s.ti + 2 * NT);
    bs.D = a.Dws + (size_t)b * BDF_NROWS * NYR;
    double* y = bs.ypred;                        // state at the interval boundaries (the predictor is dead there)
    const int N = a.N;
    AuxProblem p;
    p.X = a.X + (size_t)b * (N + 1) * NX; p.U = a.U + (size_t)b * (N + 1) * NU; p.Lam = a.Lam + (size_t)b * (N + 1) * NX;
    p.th = a.theta + (size_t)b * a.theta_stride; p.pd = a.pdata + (size_t)b * NQ; p.PW = nullptr; p.dt = a.T / N; p.N = N;
    double* PW = a.PW + (size_t)b * (N + 1) * NYR;
    double* s_hxx = bs.Fr; double* s_hxe = bs.dy;     // terminal condition staged in scratch (NX*NX and NX*NP doubles)
    if (tid == 0) {
        const double tN = p.dt * N;
        double xT[NX];
        const int lo = interp_lo(tN, p.dt, N);
        CPDP_LOOP for (int i = 0; i < NX; ++i) xT[i] = interp_val(p.X[(size_t)lo * NX + i], p.X[(size_t)(lo + 1) * NX + i], p.dt * lo, p.dt * (lo + 1), tN);
        Model::term2(xT, p.th, p.pd, s_hxx, s_hxe);
    }
    __syncthreads();
    CPDP_LOOP for (int q = tid; q < NYR; q += nt) {
        const double v = (q < NT) ? 0.5 * (s_hxx[s.ti[q] * NX + s.tj[q]] + s_hxx[s.tj[q] * NX + s.ti[q]]) : s_hxe[q - NT];
        y[q] = v;
        PW[(size_t)N * NYR + q] = v;
    }
    __syncthreads();
    int cnt[4] = {0, 0, 0, 0};
    int st = 0;
    CPDP_LOOP for (int k = N; k >= 1 && st == 0; --k) {
        st = bdf_interval(s, bs, p, p.dt * k, p.dt * (k - 1), a.rtol_b, a.atol_b, y, tms, cnt);
        CPDP_LOOP for (int q = tid; q < NYR; q += nt) PW[(size_t)(k - 1) * NYR + q] = y[q];
        __syncthreads();
    }
#ifdef CPDP_DEBUG_DUMP
    if (st == 4) {
        double* dst = a.Xa + (size_t)b * (N + 1) * NYF;
        CPDP_LOOP for (int i = tid; i < NX * NX; i += nt) {
            dst[i] = bs.Lm[i]; dst[NX * NX + i] = bs.Tr[i]; dst[2 * NX * NX + i] = bs.Ti[i];
            dst[3 * NX * NX + i] = bs.Zr[i]; dst[4 * NX * NX + i] = bs.Zi[i];
        }
        if (tid == 0) dst[5 * NX * NX] = (double)bs.flag[0];
    }
#endif
    if (tid == 0) {
        a.aux_status[b] = st;
        a.counters[b * NCOUNTERS + 0] = cnt[0]; a.counters[b * NCOUNTERS + 1] = cnt[1];
        a.counters[b * NCOUNTERS + 4] = cnt[2]; a.counters[b * NCOUNTERS + 5] = cnt[3];
    }
}

}  // namespace CPDP_NS
