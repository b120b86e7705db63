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

#ifndef CPDP_BDF_THREADS
#define CPDP_BDF_THREADS 32
#endif
#ifndef CPDP_BDF_MINB
#define CPDP_BDF_MINB 8
#endif
constexpr int BDF_THREADS = CPDP_BDF_THREADS;
// With one warp per problem (the shipped shape) every CTA-wide barrier of this file is a warp barrier.
#if defined(__CUDACC__) && CPDP_BDF_THREADS == 32
#define BDF_SYNC() __syncwarp()
#else
#define BDF_SYNC() __syncthreads()
#endif
constexpr int BDF_MAX_ORDER = 5;
constexpr int BDF_NEWTON_MAXITER = 4;
constexpr int BDF_NROWS = BDF_MAX_ORDER + 3;
constexpr int BDF_WS_ROWS = BDF_NROWS + 3;      // workspace rows per problem: differences array + scale, psi, d
constexpr int BDF_WS_DOUBLES = BDF_WS_ROWS * NYR;

struct BdfShared {
    double* Tr; double* Ti;   // [NX*NX]  Schur form L = Z T Z^H (T upper triangular)
    double* Zr;               // [NX*NX]  Q: REAL Schur vectors
    double* ga; double* gbr; double* gbi; double* pi;   // [NX] each: block rotations G (Z = Q G^H), see bdf_schur
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
    double* scale; double* psi; double* d;   // [NYR] each, GLOBAL memory (rows 8..10 of the workspace block), thread-private columns like D
    double* y;      // [NYR]  Newton iterate (starts as the predictor); between intervals the interval's start / end state
    double* dy;     // [NYR]
    double* RU;     // [3][6*6]  change_D: RU, R, U
    double* Winv;   // [NX*NX]  inverse of I + c L
    double* tmp;    // [NYR]    scratch of bdf_solve (aliases Gr/Gi, dead by then) and of the start-up probe
    int* flag;      // [2]
};

CPDP_HD double bdf_kappa(int k) { const double v[6] = {0.0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0.0}; return v[k]; }
CPDP_HD constexpr double bdf_gamma(int k) { double g = 0.0; for (int i = 1; i <= k; ++i) g += 1.0 / i; return g; }
CPDP_HD double bdf_alpha(int k) { return (1.0 - bdf_kappa(k)) * bdf_gamma(k); }
CPDP_HD double bdf_error_const(int k) { return bdf_kappa(k) * bdf_gamma(k) + 1.0 / (k + 1); }

// Shared-memory layout.  Every array sits at a COMPILE-TIME offset of the dynamic shared-memory block, so the
// out-of-line pieces below rebuild their views from constants (no pointer structs in local memory, no registers).
constexpr int BDF_GSZ = (2 * NX * NX > NYR) ? 2 * NX * NX : NYR;      // (Gr, Gi) block; doubles as bdf_solve's NYR scratch
constexpr int BDF_RED = (BDF_THREADS > 64 ? BDF_THREADS : 64) + 2;   // block_reduce scratch; the emulated sweep uses 64 entries
constexpr int BDF_SMEM_DOUBLES = MSZ + (2 * NX + NU) + BDF_RED + NX * NX + 2 * NU * NX + NU * NP       // AuxShared
                                 + 5 * NX * NX + 4 * NX + BDF_GSZ + 2 * NT + NX * NP + NX * NX                           // Schur data, Cm, Winv
                                 + 2 * NYR + 108 + 8 + 2;                                                        // y, dy, RU, tms, flag
constexpr int BDF_SMEM_INTS = 2 * NT + SPTAB_INTS;
constexpr size_t BDF_SMEM_BYTES = (size_t)BDF_SMEM_DOUBLES * sizeof(double) + (size_t)((BDF_SMEM_INTS + 3) & ~3) * sizeof(int);
static_assert(NX * NU <= NX * NX, "GH aliases an NX x NX scratch matrix");

CPDP_D void bdf_layout(double* smem, AuxShared& s, BdfShared& bs, double*& tms) {
    double* ptr = smem;
    s.M = carve(ptr, MSZ);                       // one PMP slot: every Newton iterate of a step shares t_new (kept in shared
                                                 // memory: in the L2 workspace it bought 10 CTAs per SM and no throughput)
    s.xul = carve(ptr, 2 * NX + NU);
    s.red = carve(ptr, BDF_RED);
    s.P = carve(ptr, NX * NX);
    s.Y = carve(ptr, NU * NX);
    s.Yp = carve(ptr, NU * NX);
    s.Z = carve(ptr, NU * NP);
    bs.Tr = carve(ptr, NX * NX); bs.Ti = carve(ptr, NX * NX); bs.Zr = carve(ptr, NX * NX);
    bs.ga = carve(ptr, NX); bs.gbr = carve(ptr, NX); bs.gbi = carve(ptr, NX); bs.pi = carve(ptr, NX);
    bs.Fr = carve(ptr, NX * NX); bs.Fi = carve(ptr, NX * NX); bs.Gr = carve(ptr, BDF_GSZ); bs.Gi = bs.Gr + NX * NX;
    bs.Dr = carve(ptr, NT); bs.Di = carve(ptr, NT);
    bs.Lm = bs.Fr; bs.Am = bs.Fi; bs.Rm = bs.Gr; bs.GH = bs.Gi;      // Jacobian scratch, dead once bdf_schur has copied Lm
    bs.Cm = carve(ptr, NX * NP); bs.Winv = carve(ptr, NX * NX);
    bs.y = carve(ptr, NYR); bs.dy = carve(ptr, NYR);
    bs.RU = carve(ptr, 108);                     // RU | R | U, 6 x 6 each
    tms = carve(ptr, 8);
    bs.tmp = bs.Gr;
    bs.scale = nullptr; bs.psi = nullptr; bs.d = nullptr;
    bs.flag = (int*)carve(ptr, 2);
    int* ip = (int*)(smem + BDF_SMEM_DOUBLES);
    s.ti = ip; s.tj = ip + NT;
    aux_table_ptrs(s, ip + 2 * NT);
    bs.D = nullptr;
}
#define BDF_LAYOUT() CPDP_DYN_SMEM(smem); AuxShared s; BdfShared bs; double* tms; bdf_layout(smem, s, bs, tms); (void)tms

CPDP_D double bdf_reduce(double v, bool is_max) {
#if defined(__CUDACC__) && CPDP_BDF_THREADS == 32
    // one warp: block_reduce's xor butterfly alone (identical operation order, no shared memory, no barrier)
    CPDP_LOOP for (int o = 16; o > 0; o >>= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, x) : (v + x);
    }
    return v;
#else
    BDF_LAYOUT();
    return block_reduce(v, s.red, is_max);
#endif
}
CPDP_D double bdf_pow(double x, double y) { return pow(x, y); }

// out-of-line instances of the shared right-hand side / PMP evaluation (one copy each instead of three)
CPDP_D_NOINLINE void bdf_rhs(const double* M, const double* yin, double* ydot) { BDF_LAYOUT(); riccati_rhs(s, M, yin, ydot); }
CPDP_D_NOINLINE bool bdf_prepare(const AuxProblem p, double* M) { BDF_LAYOUT(); (void)M; return aux_prepare<false>(s, p, tms, 1); }

// Closed-form Jacobian data at (PMP matrices M, packed state yJ):  L = A' - P R,  C = R W - r_
// with A = fx - fu Huu^{-1} Hxu', R = fu Huu^{-1} fu', r_ = fe - fu Huu^{-1} Hue  (CPDP.py:262-270).
CPDP_D_NOINLINE void bdf_jacobian(const double* M, const double* yJ) {
    BDF_LAYOUT();
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* fx = M + Model::PMP_FX; const double* fu = M + Model::PMP_FU; const double* fe = M + Model::PMP_FE;
    const double* Hxu = M + Model::PMP_HXU; const double* Hue = M + Model::PMP_HUE; const double* Hinv = M + Model::PMP_SIZE;
    const double* Wm = yJ + NT;
    BDF_SYNC();
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
    BDF_SYNC();
    CPDP_LOOP for (int i = tid; i < NX * NX; i += nt) {
        const int r_ = i / NX, c = i % NX;
        double a1 = fx[i], a2 = 0.0;
        CPDP_LOOP for (int a = 0; a < NU; ++a) { a1 -= bs.GH[r_ * NU + a] * Hxu[c * NU + a]; a2 += bs.GH[r_ * NU + a] * fu[c * NU + a]; }
        bs.Am[i] = a1; bs.Rm[i] = a2;
    }
    BDF_SYNC();
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
    BDF_SYNC();
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
#ifdef CPDP_SCHUR_NONORM
                    x2 = 1.0;          // tuning knob: skip EISPACK's overflow guard (entries of L are O(1e3) at most here)
#else
                    const double ix2 = 1.0 / x2;
                    p *= ix2; q *= ix2; r *= ix2;
#endif
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
    BDF_SYNC();
    CPDP_LOOP for (int i = tid; i < n * n; i += nt) { bs.Tr[i] = bs.Lm[i]; bs.Ti[i] = 0.0; }
    if (tid == 0) bs.flag[0] = 1;
    BDF_SYNC();
#ifdef __CUDACC__
    if (tid < 32)
#endif
    {
        const int lane = tid;
        const bool ok = schur_real_w0(bs.Tr, bs.Zr);
        CPDP_W0_SYNC();
        if (!ok && lane == 0) bs.flag[0] = 0;
        // ---- one unitary rotation per 2 x 2 block:  G = [[c, s], [-conj(s), c]],  T <- G T G^H.  The Schur vectors stay
        //      REAL (Q = Zr); the block-diagonal unitary factor is kept as per-index coefficients (Z = Q G^H):
        //      G[i][i] = ga[i] (real), G[i][pi[i]] = gb[i] (complex), pi[i] = the other index of i's block (or i).
        if (lane < n) { bs.ga[lane] = 1.0; bs.gbr[lane] = 0.0; bs.gbi[lane] = 0.0; bs.pi[lane] = (double)lane; }
        CPDP_W0_SYNC();
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
            } else if (lane == 16) {
                bs.ga[j] = cr; bs.gbr[j] = sr; bs.gbi[j] = si; bs.pi[j] = (double)(j + 1);
                bs.ga[j + 1] = cr; bs.gbr[j + 1] = -sr; bs.gbi[j + 1] = si; bs.pi[j + 1] = (double)j;      // -conj(s)
            }
            CPDP_W0_SYNC();
            if (lane == 0) { bs.Tr[(j + 1) * n + j] = 0.0; bs.Ti[(j + 1) * n + j] = 0.0; }
            CPDP_W0_SYNC();
        }
    }
    BDF_SYNC();
    return bs.flag[0] != 0;
}

// inner 13-term loops of the small products: rolled by default, -DCPDP_MM_UNROLL=k unrolls them k times (tuning knob)
#ifdef CPDP_MM_UNROLL
#define CPDP_MM_PRAGMA(k) _Pragma(#k)
#define CPDP_MM_LOOP2(k) CPDP_MM_PRAGMA(unroll k)
#define CPDP_MM_LOOP CPDP_MM_LOOP2(CPDP_MM_UNROLL)
#else
#define CPDP_MM_LOOP CPDP_LOOP
#endif

// Small dense products with three (two) outputs per thread in flight: the rolled 13-term dot products are latency
// bound, independent accumulators overlap their shared-memory loads and DFMA chains without unrolling the loop.
//   mm_nn:      out[i][k] = sum_j A[i][j] B[j][k]                       all n*n outputs
//   mm_tri<TA>: v(i,j)    = sum_k (TA ? A[k][i] : A[i][k]) * (TA ? B[k][j] : B[j][k])   for the NT pairs i <= j
CPDP_D void mm_nn(const double* A, const double* B, double* out) {
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    CPDP_LOOP for (int base = 0; base < n * n; base += 3 * nt) {
        const int e0 = base + tid, e1 = e0 + nt, e2 = e1 + nt;
        const int f0 = e0 < n * n ? e0 : 0, f1 = e1 < n * n ? e1 : 0, f2 = e2 < n * n ? e2 : 0;
        const double* a0 = A + (f0 / n) * n; const double* a1 = A + (f1 / n) * n; const double* a2 = A + (f2 / n) * n;
        const double* b0 = B + f0 % n; const double* b1 = B + f1 % n; const double* b2 = B + f2 % n;
        double c0 = 0.0, c1 = 0.0, c2 = 0.0;
        CPDP_MM_LOOP for (int j = 0; j < n; ++j) { c0 += a0[j] * b0[j * n]; c1 += a1[j] * b1[j * n]; c2 += a2[j] * b2[j * n]; }
        if (e0 < n * n) out[e0] = c0;
        if (e1 < n * n) out[e1] = c1;
        if (e2 < n * n) out[e2] = c2;
    }
}
template <bool TA, class Store>
CPDP_D void mm_tri(const AuxShared& s, const double* A, const double* B, Store store) {
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    CPDP_LOOP for (int base = 0; base < NT; base += 2 * nt) {
        const int q0 = base + tid, q1 = q0 + nt;
        const int r0 = q0 < NT ? q0 : 0, r1 = q1 < NT ? q1 : 0;
        const int i0 = s.ti[r0], j0 = s.tj[r0], i1 = s.ti[r1], j1 = s.tj[r1];
        const double* a0 = TA ? A + i0 : A + i0 * n; const double* b0 = TA ? B + j0 : B + j0 * n;
        const double* a1 = TA ? A + i1 : A + i1 * n; const double* b1 = TA ? B + j1 : B + j1 * n;
        constexpr int st = TA ? n : 1;
        double c0 = 0.0, c1 = 0.0;
        CPDP_MM_LOOP for (int k = 0; k < n; ++k) { c0 += a0[k * st] * b0[k * st]; c1 += a1[k * st] * b1[k * st]; }
        if (q0 < NT) store(q0, i0, j0, c0);
        if (q1 < NT) store(q1, i1, j1, c1);
    }
}

// scipy's "LU" event for a new c: reciprocals of the Lyapunov pivots and Winv = (I + c L)^{-1} = Re(Z (I + c T)^{-1} Z^H).
CPDP_D_NOINLINE bool bdf_factor(const double c) {
    BDF_LAYOUT();
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    BDF_SYNC();
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
    bad = bdf_reduce(bad, true);
    BDF_SYNC();                                      // S, Dr, Di visible to every lane (the one-warp reduction has no barrier of its own)
    if (bad != 0.0) return false;
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // Sr = Re(G^H S G) -> Gr   (real: (I + cL)^{-1} and Q are)
        const int i = e / n, j = e % n;
        const int pi = (int)bs.pi[i], pj = (int)bs.pi[j];
        // conj(G[a][i]) for a in {i, pi}; G[b][j] for b in {j, pj}
        const double ai_r = bs.ga[i], api_r = bs.gbr[pi], api_i = -bs.gbi[pi];
        const double bj_r = bs.ga[j], bpj_r = bs.gbr[pj], bpj_i = bs.gbi[pj];
        // t_b = conj(G[i][i]) S[i][b] + conj(G[pi][i]) S[pi][b]   for b = j and b = pj
        double acc = 0.0;
        {
            const double s1r = bs.Fr[i * n + j], s1i = bs.Fi[i * n + j], s2r = bs.Fr[pi * n + j], s2i = bs.Fi[pi * n + j];
            const double tr = ai_r * s1r + ((pi != i) ? (api_r * s2r - api_i * s2i) : 0.0);
            const double ti = ai_r * s1i + ((pi != i) ? (api_r * s2i + api_i * s2r) : 0.0);
            acc += tr * bj_r;  (void)ti;
        }
        if (pj != j) {
            const double s1r = bs.Fr[i * n + pj], s1i = bs.Fi[i * n + pj], s2r = bs.Fr[pi * n + pj], s2i = bs.Fi[pi * n + pj];
            const double tr = ai_r * s1r + ((pi != i) ? (api_r * s2r - api_i * s2i) : 0.0);
            const double ti = ai_r * s1i + ((pi != i) ? (api_r * s2i + api_i * s2r) : 0.0);
            acc += tr * bpj_r - ti * bpj_i;
        }
        bs.Gr[e] = acc;
    }
    BDF_SYNC();
    mm_nn(bs.Zr, bs.Gr, bs.Fr);                                // F = Q Sr
    BDF_SYNC();
    CPDP_LOOP for (int e = tid; e < n * n; e += nt) {          // Winv = F Q^T
        const int i = e / n, l = e % n;
        double acc = 0.0;
        CPDP_LOOP for (int k = 0; k < n; ++k) acc += bs.Fr[i * n + k] * bs.Zr[l * n + k];
        bs.Winv[e] = acc;
    }
    BDF_SYNC();
    return true;
}

// dy <- (I - cJ)^{-1} dy   (dy holds the right-hand side on entry; tmp: NYR doubles of scratch)
CPDP_D_NOINLINE void bdf_solve(const double c) {
    BDF_LAYOUT();
    double* dy = bs.dy; double* tmp = bs.tmp;
    constexpr int n = NX;
    const int tid = threadIdx.x, nt = blockDim.x;
    BDF_SYNC();
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // B = sym(dy[0:NT]) expanded -> P
        const int i = s.ti[q], j = s.tj[q];
        const double v = dy[q];
        s.P[i * n + j] = v; s.P[j * n + i] = v;
    }
    BDF_SYNC();
    mm_nn(s.P, bs.Zr, bs.Fr);                                  // F = B Q   (real)
    BDF_SYNC();
    {                                                          // Cr = Q^T F  (real symmetric, both triangles) -> Fi
        double* Cr = bs.Fi;
        mm_tri<true>(s, bs.Zr, bs.Fr, [Cr](int, int i, int j, double v) { Cr[i * n + j] = v; Cr[j * n + i] = v; });
    }
    BDF_SYNC();
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // C = G Cr G^H (upper triangle) -> (Gr, Gi)
        const int i = s.ti[q], j = s.tj[q];
        const int pi = (int)bs.pi[i], pj = (int)bs.pi[j];
        const double gi_a = bs.ga[i], gi_br = bs.gbr[i], gi_bi = bs.gbi[i];       // G[i][i], G[i][pi]
        const double gj_a = bs.ga[j], gj_br = bs.gbr[j], gj_bi = -bs.gbi[j];      // conj(G[j][j]), conj(G[j][pj])
        // u_b = G[i][i] Cr[i][b] + G[i][pi] Cr[pi][b]    for b = j, pj
        const double c1 = bs.Fi[i * n + j], c2 = bs.Fi[pi * n + j], c3 = bs.Fi[i * n + pj], c4 = bs.Fi[pi * n + pj];
        const double ujr = gi_a * c1 + gi_br * c2, uji = gi_bi * c2;
        const double upr = gi_a * c3 + gi_br * c4, upi = gi_bi * c4;
        bs.Gr[i * n + j] = ujr * gj_a + (upr * gj_br - upi * gj_bi);
        bs.Gi[i * n + j] = uji * gj_a + (upr * gj_bi + upi * gj_br);
    }
    BDF_SYNC();
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
            {
                // every lane issues the same unconditional (index-clamped) loads, out-of-range terms are zeroed by a
                // select: no branches, all loads in flight together, one short DFMA chain per term
                constexpr int M = (n + 2) / 4;
                const int ic = valid ? i : 0, jc = valid ? j : 0;
                double pr[2 * M], pi_[2 * M];
#pragma unroll
                for (int m = 0; m < M; ++m) {                                      // T_ik Y_kj,  k = i + 1 + sub + 4 m
                    const int k = ic + 1 + sub + 4 * m;
                    const bool in = valid && (k < n);
                    const int kc = in ? k : 0;
                    const double tr = bs.Tr[ic * n + kc], ti = bs.Ti[ic * n + kc], yr = bs.Gr[kc * n + jc], yi = bs.Gi[kc * n + jc];
                    const double vr = tr * yr - ti * yi, vi = tr * yi + ti * yr;
                    pr[m] = in ? vr : 0.0; pi_[m] = in ? vi : 0.0;
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {                                      // Y_ik conj(T_jk),  k = j + 1 + sub + 4 m
                    const int k = jc + 1 + sub + 4 * m;
                    const bool in = valid && (k < n);
                    const int kc = in ? k : 0;
                    const double tr = bs.Tr[jc * n + kc], ti = bs.Ti[jc * n + kc], yr = bs.Gr[ic * n + kc], yi = bs.Gi[ic * n + kc];
                    const double vr = yr * tr + yi * ti, vi = yi * tr - yr * ti;
                    pr[M + m] = in ? vr : 0.0; pi_[M + m] = in ? vi : 0.0;
                }
#pragma unroll
                for (int m = 0; m < 2 * M; ++m) { ar += pr[m]; ai += pi_[m]; }
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
    BDF_SYNC();
    CPDP_LOOP for (int q = tid; q < NT; q += nt) {             // Yr = Re(G^H Y G)  (real symmetric, both triangles) -> Fr
        const int i = s.ti[q], j = s.tj[q];
        const int pi = (int)bs.pi[i], pj = (int)bs.pi[j];
        const double ai_r = bs.ga[i], ap_r = bs.gbr[pi], ap_i = -bs.gbi[pi];      // conj(G[i][i]), conj(G[pi][i])
        const double bj_r = bs.ga[j], bp_r = bs.gbr[pj], bp_i = bs.gbi[pj];       // G[j][j], G[pj][j]
        const double y1r = bs.Gr[i * n + j], y1i = bs.Gi[i * n + j], y2r = bs.Gr[pi * n + j], y2i = bs.Gi[pi * n + j];
        const double y3r = bs.Gr[i * n + pj], y3i = bs.Gi[i * n + pj], y4r = bs.Gr[pi * n + pj], y4i = bs.Gi[pi * n + pj];
        const double tjr = ai_r * y1r + (ap_r * y2r - ap_i * y2i);
        const double tpr = ai_r * y3r + (ap_r * y4r - ap_i * y4i), tpi = ai_r * y3i + (ap_r * y4i + ap_i * y4r);
        const double v = tjr * bj_r + (tpr * bp_r - tpi * bp_i);
        bs.Fr[i * n + j] = v; bs.Fr[j * n + i] = v;
    }
    BDF_SYNC();
    mm_nn(bs.Zr, bs.Fr, bs.Fi);                                // F2 = Q Yr -> Fi
    BDF_SYNC();
    {                                                          // X = F2 Q^T, upper triangle -> tmp and expanded -> P
        double* Pm = s.P;
        mm_tri<false>(s, bs.Fi, bs.Zr, [tmp, Pm](int q, int i, int l, double v) { tmp[q] = v; Pm[i * n + l] = v; Pm[l * n + i] = v; });
    }
    BDF_SYNC();
    double* dW = dy + NT;
    CPDP_LOOP for (int e = tid; e < NX * NP; e += nt) {                  // B_W + c X C
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        CPDP_MM_LOOP for (int a = 0; a < NX; ++a) acc += s.P[i * NX + a] * bs.Cm[a * NP + k];
        tmp[NT + e] = dW[e] + c * acc;
    }
    CPDP_LOOP for (int k = tid; k < NT; k += nt) dy[k] = tmp[k];
    BDF_SYNC();
    CPDP_LOOP for (int e = tid; e < NX * NP; e += nt) {                  // dW = (I + cL)^{-1} (...)
        const int i = e / NP, k = e % NP;
        double acc = 0.0;
        CPDP_MM_LOOP for (int a = 0; a < NX; ++a) acc += bs.Winv[i * NX + a] * tmp[NT + a * NP + k];
        dW[e] = acc;
    }
    BDF_SYNC();
}

// change_D (bdf.py:18-33): D[:order+1] <- (R U)' D[:order+1]
CPDP_D_NOINLINE void bdf_change_D(double* D, const int order, const double factor) {
    BDF_LAYOUT();
    bs.D = D;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* R = bs.RU + 36; double* U = bs.RU + 72;         // 6 x 6 scratch
    BDF_SYNC();
    if (tid <= order) {                                     // column tid of R and U: cumprod down the rows (compute_R)
        const int j = tid;
        R[j] = 1.0; U[j] = 1.0;
        CPDP_LOOP for (int i = 1; i <= order; ++i) {
            R[i * 6 + j] = (j == 0) ? 0.0 : R[(i - 1) * 6 + j] * (((double)(i - 1) - factor * j) / i);
            U[i * 6 + j] = (j == 0) ? 0.0 : U[(i - 1) * 6 + j] * (((double)(i - 1) - (double)j) / i);
        }
    }
    BDF_SYNC();
    CPDP_LOOP for (int e = tid; e < 36; e += nt) {
        const int i = e / 6, j = e % 6;
        if (i <= order && j <= order) {
            double acc = 0.0;
            CPDP_LOOP for (int k = 0; k <= order; ++k) acc += R[i * 6 + k] * U[k * 6 + j];
            bs.RU[i * 6 + j] = acc;
        }
    }
    BDF_SYNC();
    CPDP_LOOP for (int q = tid; q < NYR; q += nt) {
        double v[BDF_MAX_ORDER + 1];
#pragma unroll
        for (int i = 0; i <= BDF_MAX_ORDER; ++i) v[i] = (i <= order) ? bs.D[(size_t)i * NYR + q] : 0.0;
#pragma unroll
        for (int i = 0; i <= BDF_MAX_ORDER; ++i) {
            if (i <= order) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j <= BDF_MAX_ORDER; ++j) if (j <= order) acc += bs.RU[j * 6 + i] * v[j];
                bs.D[(size_t)i * NYR + q] = acc;
            }
        }
    }
    BDF_SYNC();
}

// RMS norm of v/scale over the FULL (n^2 + n r) state (off-diagonal entries of the packed P count twice)
CPDP_D_NOINLINE double bdf_norm(const double* v, const double* scale, const double mul) {
    BDF_LAYOUT();
    const int tid = threadIdx.x, nt = blockDim.x;
    double a = 0.0;
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { const double x = mul * v[i] / scale[i]; a += ric_wgt(s, i) * x * x; }
    return sqrt(bdf_reduce(a, false) / (double)NFULL_R);
}

// Developer instrumentation (-DCPDP_BDF_TIMING): per-phase clock64() totals of each problem, written over the Ua rows of
// the problem at kernel end (phases = 1 runs only; tools/prof_bdf_phases.py).  Off in the shipped library.
#ifdef CPDP_BDF_TIMING
#define BDF_T(ph, expr) do { const long long t0__ = clock64(); expr; tp[ph] += clock64() - t0__; } while (0)
#define BDF_TB(ph, var, expr) do { const long long t0__ = clock64(); var = (expr); tp[ph] += clock64() - t0__; } while (0)
#define BDF_TP_PARAM , long long* tp
#define BDF_TP_ARG , tp
#else
#define BDF_T(ph, expr) do { expr; } while (0)
#define BDF_TB(ph, var, expr) do { var = (expr); } while (0)
#define BDF_TP_PARAM
#define BDF_TP_ARG
#endif

// One grid interval [t0, t1] with scipy's BDF.  y in/out (shared memory).  Returns 0 ok, 1 step too small,
// 2 non-finite, 4 singular Newton matrix.  cnt: [rhs evaluations, steps (accepted), LU factorisations, Jacobians]
CPDP_D int bdf_interval(const AuxShared& s, const BdfShared& bs, const AuxProblem& p, const double t0, const double t1,
                        const double rtol, const double atol, double* y, double* tms, int* cnt BDF_TP_PARAM) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double dir = (t1 >= t0) ? 1.0 : -1.0;
    const double EPS = 2.220446049250313e-16;
    // ---- __init__ (bdf.py:200-257)
    if (tid == 0) tms[0] = t0;
    { bool okp__; BDF_TB(0, okp__, bdf_prepare(p, s.M)); if (!okp__) return 2; }
    double* f0 = bs.d;                 // f(t0, y0) parked in the (not yet used) d row
    BDF_T(1, bdf_rhs(s.M, y, f0)); ++cnt[0];
    BDF_T(2, bdf_jacobian(s.M, y)); ++cnt[3];
    { bool oks__; BDF_TB(3, oks__, bdf_schur()); if (!oks__) return 4; }
    double h_abs;
    {
        const double interval_length = fabs(t1 - t0);
        double a0 = 0.0, a1 = 0.0;
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            a0 += ric_wgt(s, i) * (y[i] / sc) * (y[i] / sc);
            a1 += ric_wgt(s, i) * (f0[i] / sc) * (f0[i] / sc);
        }
        const double d0 = sqrt(bdf_reduce(a0, false) / (double)NFULL_R);
        const double d1 = sqrt(bdf_reduce(a1, false) / (double)NFULL_R);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval_length);
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) bs.dy[i] = y[i] + h0 * dir * f0[i];
        if (tid == 0) tms[0] = t0 + h0 * dir;
        { bool okp__; BDF_TB(0, okp__, bdf_prepare(p, s.M)); if (!okp__) return 2; }
        BDF_T(1, bdf_rhs(s.M, bs.dy, bs.tmp)); ++cnt[0];
        double a2 = 0.0;
        CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
            const double sc = atol + fabs(y[i]) * rtol;
            const double v = (bs.tmp[i] - f0[i]) / sc;
            a2 += ric_wgt(s, i) * v * v;
        }
        const double d2 = sqrt(bdf_reduce(a2, false) / (double)NFULL_R) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = bdf_pow(0.01 / fmax(d1, d2), 1.0 / 2.0);       // order 1
        h_abs = fmin(fmin(100 * h0, h1), interval_length);
    }
    const double newton_tol = fmax(10 * EPS / rtol, fmin(0.03, sqrt(rtol)));
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { bs.D[i] = y[i]; bs.D[NYR + i] = f0[i] * h_abs * dir; }
    BDF_SYNC();
    int order = 1, n_equal_steps = 0;
    bool lu_valid = false;
    double c_lu = 0.0;                 // the c the current LU was built with (scipy keeps a stale LU after an error rejection)
    double t = t0;

    while (dir * (t - t1) < 0) {
        // ---- _step_impl (bdf.py:314-453)
        const double min_step = 10 * fabs(nextafter(t, dir * INFINITY) - t);
        if (h_abs < min_step) {
            BDF_T(7, bdf_change_D(bs.D, order, min_step / h_abs));
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
                BDF_T(7, bdf_change_D(bs.D, order, fabs(t_new - t) / h_abs));
                n_equal_steps = 0;
                lu_valid = false;
            }
            h = t_new - t;
            h_abs = fabs(h);
            const double al = bdf_alpha(order);
            CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
                double yp = 0.0, ps = 0.0;
#pragma unroll
                for (int k = 0; k <= BDF_MAX_ORDER; ++k) {
                    if (k <= order) {
                        const double v = bs.D[(size_t)k * NYR + i];
                        yp += v;
                        if (k >= 1) ps += v * bdf_gamma(k);
                    }
                }
                bs.y[i] = yp;                                     // y_predict: the Newton iteration starts from it
                bs.scale[i] = atol + rtol * fabs(yp);
                bs.psi[i] = ps / al;
            }
            if (tid == 0) tms[0] = t_new;
            { bool okp__; BDF_TB(0, okp__, bdf_prepare(p, s.M)); if (!okp__) return 2; }      // PMP matrices at t_new (every Newton iterate shares them)
            const double c = h / al;
            bool converged = false;
            while (!converged) {
                if (!lu_valid) {
                    { bool okf__; BDF_TB(4, okf__, bdf_factor(c)); if (!okf__) return 4; }
                    lu_valid = true; c_lu = c; ++cnt[2];
                }
                // ---- solve_bdf_system (bdf.py:36-75)
                CPDP_LOOP for (int i = tid; i < NYR; i += nt) bs.d[i] = 0.0;
                BDF_SYNC();
                double dy_norm_old = -1.0;
                int k = 0;
                CPDP_LOOP for (k = 0; k < BDF_NEWTON_MAXITER; ++k) {
                    BDF_T(1, bdf_rhs(s.M, bs.y, bs.dy)); ++cnt[0];
                    double fin = 0.0;
                    CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
                        const double fv = bs.dy[i];
                        if (!(fabs(fv) < 1e300)) fin = 1.0;
                        bs.dy[i] = c * fv - bs.psi[i] - bs.d[i];
                    }
                    fin = bdf_reduce(fin, true);
                    if (fin != 0.0) break;
                    BDF_T(5, bdf_solve(c_lu));
                    double dy_norm; BDF_TB(6, dy_norm, bdf_norm(bs.dy, bs.scale, 1.0));
                    const bool have_rate = dy_norm_old >= 0.0;
                    const double rate = have_rate ? dy_norm / dy_norm_old : 0.0;
                    if (have_rate && (rate >= 1 || bdf_pow(rate, (double)(BDF_NEWTON_MAXITER - k)) / (1 - rate) * dy_norm > newton_tol)) break;
                    CPDP_LOOP for (int i = tid; i < NYR; i += nt) { bs.y[i] += bs.dy[i]; bs.d[i] += bs.dy[i]; }
                    BDF_SYNC();
                    if (dy_norm == 0 || (have_rate && rate / (1 - rate) * dy_norm < newton_tol)) { converged = true; break; }
                    dy_norm_old = dy_norm;
                }
                n_iter = (k < BDF_NEWTON_MAXITER) ? k + 1 : BDF_NEWTON_MAXITER;
                if (!converged) {
                    if (current_jac) break;
                    // back to y_predict (same sum, same order => same bits as above), Jacobian there (bdf.py:372)
                    CPDP_LOOP for (int i = tid; i < NYR; i += nt) {
                        double yp = 0.0;
                        CPDP_LOOP for (int k = 0; k <= order; ++k) yp += bs.D[(size_t)k * NYR + i];
                        bs.y[i] = yp;
                    }
                    BDF_T(2, bdf_jacobian(s.M, bs.y)); ++cnt[3];
                    { bool oks__; BDF_TB(3, oks__, bdf_schur()); if (!oks__) return 4; }
                    lu_valid = false;
                    current_jac = true;
                }
            }
            if (!converged) {
                h_abs *= 0.5;
                BDF_T(7, bdf_change_D(bs.D, order, 0.5));
                n_equal_steps = 0;
                lu_valid = false;
                continue;
            }
            safety = 0.9 * (2 * BDF_NEWTON_MAXITER + 1) / (double)(2 * BDF_NEWTON_MAXITER + n_iter);
            CPDP_LOOP for (int i = tid; i < NYR; i += nt) bs.scale[i] = atol + rtol * fabs(bs.y[i]);
            BDF_SYNC();
            BDF_TB(6, error_norm, bdf_norm(bs.d, bs.scale, bdf_error_const(order)));
            if (!(error_norm == error_norm)) return 2;
            if (error_norm > 1) {
                const double factor = fmax(0.2, safety * bdf_pow(error_norm, -1.0 / (order + 1)));
                h_abs *= factor;
                BDF_T(7, bdf_change_D(bs.D, order, factor));
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
        BDF_SYNC();
        if (n_equal_steps < order + 1) continue;
        double error_m_norm = INFINITY, error_p_norm = INFINITY;
        if (order > 1) BDF_TB(6, error_m_norm, bdf_norm(bs.D + (size_t)order * NYR, bs.scale, bdf_error_const(order - 1)));
        if (order < BDF_MAX_ORDER) BDF_TB(6, error_p_norm, bdf_norm(bs.D + (size_t)(order + 2) * NYR, bs.scale, bdf_error_const(order + 1)));
        const double fm = bdf_pow(error_m_norm, -1.0 / order);
        const double f0 = bdf_pow(error_norm, -1.0 / (order + 1));
        const double fp = bdf_pow(error_p_norm, -1.0 / (order + 2));
        int delta_order = -1; double fmaxv = fm;          // np.argmax: first maximum
        if (f0 > fmaxv) { fmaxv = f0; delta_order = 0; }
        if (fp > fmaxv) { fmaxv = fp; delta_order = 1; }
        order += delta_order;
        const double factor = fmin(10.0, safety * fmaxv);
        h_abs *= factor;
        BDF_T(7, bdf_change_D(bs.D, order, factor));
        n_equal_steps = 0;
        lu_valid = false;
    }
    // solve_ivp(t_eval=[t1]) returns the dense output at the step end = D[0] (bdf.py:462-484)
    CPDP_LOOP for (int i = tid; i < NYR; i += nt) y[i] = bs.D[i];
    BDF_SYNC();
    return 0;
}

// k_riccati_bdf: backward sweep of COCSys.auxSysSolver as shipped (CPDP.py:327-338).
CPDP_GLOBAL void __launch_bounds__(BDF_THREADS, CPDP_BDF_MINB) k_riccati_bdf(AuxArgs a) {
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
    bs.D = a.Dws + (size_t)b * BDF_WS_DOUBLES;
    bs.scale = bs.D + (size_t)BDF_NROWS * NYR; bs.psi = bs.scale + NYR; bs.d = bs.psi + NYR;
    double* y = bs.y;                            // state at the interval boundaries
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
    BDF_SYNC();
    CPDP_LOOP for (int q = tid; q < NYR; q += nt) {
        const double v = (q < NT) ? 0.5 * (s_hxx[s.ti[q] * NX + s.tj[q]] + s_hxx[s.tj[q] * NX + s.ti[q]]) : s_hxe[q - NT];
        y[q] = v;
        PW[(size_t)N * NYR + q] = v;
    }
    BDF_SYNC();
    int cnt[4] = {0, 0, 0, 0};
    int st = 0;
#ifdef CPDP_BDF_TIMING
    long long tp[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long tstart__ = clock64();
#endif
    CPDP_LOOP for (int k = N; k >= 1 && st == 0; --k) {
        st = bdf_interval(s, bs, p, p.dt * k, p.dt * (k - 1), a.rtol_b, a.atol_b, y, tms, cnt BDF_TP_ARG);
        CPDP_LOOP for (int q = tid; q < NYR; q += nt) PW[(size_t)(k - 1) * NYR + q] = y[q];
        BDF_SYNC();
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
#ifdef CPDP_BDF_TIMING
    if (tid == 0) {
        tp[8] = clock64() - tstart__;
        double* dst = a.Ua + (size_t)b * (N + 1) * NU * NP;
        for (int i = 0; i < 10; ++i) dst[i] = (double)tp[i];
    }
#endif
    if (tid == 0) {
        a.aux_status[b] = st;
        a.counters[b * NCOUNTERS + 0] = cnt[0]; a.counters[b * NCOUNTERS + 1] = cnt[1];
        a.counters[b * NCOUNTERS + 4] = cnt[2]; a.counters[b * NCOUNTERS + 5] = cnt[3];
    }
}

}  // namespace CPDP_NS
