// Forward auxiliary sweep + fused loss / dL/dtheta, one warp per problem (round 2).
// Reference: COCSys.auxSysODE / auxSysSolver, /root/reference/CPDP/CPDP.py:281-298, 341-381 (solve_ivp's default RK45 restarted on
// every grid interval, X(0) = 0, linear interp1d of x, u, lambda, P, W) and the loss closures
// (/root/reference/lib/QuadAlgorithm.py:616-639, Examples/rocket_groundtruth.py:45-70).  The integrator control logic is scipy's
// (rk.py:14-16,111-170; common.py:63-134), exactly as in rk45_interval of cpdp_aux.cuh, which this kernel replaces for the
// forward direction; every sum is taken in the same order, so the results are bit-identical to the first version.
//
// Shape: lane k < NP owns COLUMN k of X = dx/dtheta (NX values in registers); the right-hand side
//     Ua = HY X + HZ,   Xdot = fx X + fu Ua + fe,     HY = -Huu^{-1}(fu'P + Hux),  HZ = -Huu^{-1}(fu'W + Hue)
// is "small matrix in shared memory (broadcast loads) times my register column", the sparse fx / fu walked through compile-time
// COO lists.  Per stage time only what that needs is kept (fx, fu values, fe, HY, HZ: ~230 doubles instead of the dense 684-double
// PMP set; the code generator emits Model::pmp_fwd for it), the P / W node rows are interpolated only where fu has non-zero rows,
// and the node's stage data serves both the aux control at the node and the first derivative of the next interval.
// 17 KB of shared memory per problem instead of 50.
#pragma once
#include "cpdp_bdf.cuh"

namespace CPDP_NS {

constexpr int FW_THREADS = 32;
static_assert(NP <= 32 && NX + NP <= 32, "k_aux_forward maps one lane per column of X and of [Y | Z]");
constexpr int NPS = NP | 1;                                     // odd row stride of the column buffers
constexpr int FW_SLOT = ((Model::FX_nnz + Model::FU_nnz + NX * NP + NU * NX + NU * NP) + 1) & ~1;   // fxc | fuc | fe | HY | HZ
constexpr int FW_SCR = NX * NU + NU * NP + 2 * NU * NU;         // per-slot scratch: Hxu | Hue | Huu | Huu^{-1}
namespace fo {
constexpr int XUL = 0;                                          // [NSLOT][2NX+NU]
constexpr int SL = XUL + ((NSLOT * (2 * NX + NU) + 1) & ~1);    // [NSLOT][FW_SLOT]
constexpr int SCR = SL + NSLOT * FW_SLOT;                       // [NSLOT][FW_SCR]
constexpr int KS = SCR + NSLOT * FW_SCR;                        // [7][NX][NPS]  Runge-Kutta stages
constexpr int XS = KS + 7 * NX * NPS;                           // [NX][NPS]     stage input / new state columns
constexpr int TMS = XS + NX * NPS;                              // [8]
constexpr int RED = TMS + 8;                                    // [66]
constexpr int FLAG = RED + 66;
constexpr int TAB = FLAG + 2;                                   // Dormand-Prince A[6][5] (a dynamically indexed local table would live in local memory)
constexpr int MBAR = (TAB + 30 + 1) & ~1;                                // mbarrier of the node-row ring (8 B) | ring_lo, ring_hi (2 ints)
constexpr int RING = MBAR + 2;                                  // [3][NYR] rows of the P / W node table, row r in slot r % 3
constexpr int END = RING + ((3 * NYR + 1) & ~1);
static_assert(RING % 2 == 0, "bulk copies need a 16-byte aligned destination");
}  // namespace fo
constexpr size_t FW_SMEM_BYTES = (size_t)fo::END * sizeof(double);
constexpr int FWS_FXC = 0, FWS_FUC = Model::FX_nnz, FWS_FE = FWS_FUC + Model::FU_nnz, FWS_HY = FWS_FE + NX * NP, FWS_HZ = FWS_HY + NU * NX;

CPDP_D_NOINLINE double fw_reduce(double v, bool is_max) {
#ifdef __CUDACC__
    BDF_UNROLL for (int o = 16; o > 0; o >>= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, x) : (v + x);
    }
    return v;
#else
    CPDP_DYN_SMEM(sm);
    return block_reduce(v, sm + fo::RED, is_max);
#endif
}

// ------------------------------------------------------------------------------------------------
// Node-row ring.  The forward right-hand side interpolates P, W between the two node rows that bracket the stage time; the rows of
// grid interval k (2.9 KB for the quadrotor) are the same for every stage of every step of that interval.  Row k+2 is fetched into
// shared memory by the TMA engine (cp.async.bulk, completion on an mbarrier) while interval k integrates; fw_prepare reads the ring
// and falls back to global memory for a row that is not in it.  Rows whose byte size is not a multiple of 16 (odd NYR) and the host
// emulation copy with plain loads.
// ------------------------------------------------------------------------------------------------
constexpr bool FW_RING_TMA = (NYR % 2 == 0);
CPDP_D unsigned fw_smem_addr(const void* ptr) {
#ifdef __CUDACC__
    return (unsigned)__cvta_generic_to_shared(ptr);
#else
    (void)ptr; return 0u;
#endif
}
// rows [row, row + nrows) -> their slots (contiguous in the ring: the caller never wraps); tma: this launch may use bulk copies
CPDP_D void fw_ring_issue(double* sm, const AuxProblem& p, const int row, const int nrows, const bool tma) {
    double* dst = sm + fo::RING + (row % 3) * NYR;
    const double* src = p.PW + (size_t)row * NYR;
#ifdef __CUDACC__
    if (tma) {
        if (threadIdx.x == 0) {
            const unsigned bar = fw_smem_addr(sm + fo::MBAR), bytes = (unsigned)(nrows * NYR * sizeof(double));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(fw_smem_addr(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
        }
        return;
    }
#else
    (void)tma;
#endif
    CPDP_LOOP for (int q = threadIdx.x; q < nrows * NYR; q += FW_THREADS) dst[q] = src[q];
}
// the copy issued `parity` completions ago ... has landed (tma) / is visible to the warp (plain copy)
CPDP_D void fw_ring_wait(double* sm, const unsigned parity, const bool tma) {
#ifdef __CUDACC__
    if (tma) {
        const unsigned bar = fw_smem_addr(sm + fo::MBAR);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        return;
    }
#else
    (void)parity; (void)tma;
#endif
    BDF_SYNC();
}
// row `lo` and `lo + 1` of the node table: ring slots when both are held, global memory otherwise
CPDP_D void fw_rows(const double* sm, const AuxProblem& p, const int lo, const double*& P0, const double*& P1) {
    const int* rg = (const int*)(sm + fo::MBAR) + 2;
    if (lo >= rg[0] && lo + 1 <= rg[1]) { P0 = sm + fo::RING + (lo % 3) * NYR; P1 = sm + fo::RING + ((lo + 1) % 3) * NYR; }
    else { P0 = p.PW + (size_t)lo * NYR; P1 = P0 + NYR; }
}

// Stage data for `cnt` times tms[0..cnt): (x, u, lambda) by interp1d, Model::pmp_fwd on lane s, then the columns of
// [HY | HZ] on lanes j < NX + NP (one (slot, column) pair per lane and round).
CPDP_D_NOINLINE bool fw_prepare(const AuxProblem p, const int cnt) {
    CPDP_DYN_SMEM(sm);
    const int lane = threadIdx.x;
    const double* tms = sm + fo::TMS;
    int* flag = (int*)(sm + fo::FLAG);
    BDF_SYNC();
    CPDP_LOOP for (int q = lane; q < cnt * (2 * NX + NU); q += FW_THREADS) {
        const int sl = q / (2 * NX + NU), e = q % (2 * NX + NU);
        sm[fo::XUL + q] = xul_at(p, tms[sl], e);
    }
    if (lane == 0) flag[0] = 1;
    BDF_SYNC();
    if (lane < cnt) {
        const double* xul = sm + fo::XUL + lane * (2 * NX + NU);
        double* sl = sm + fo::SL + lane * FW_SLOT;
        double* sc = sm + fo::SCR + lane * FW_SCR;
        PdBuf pdb;
        Model::pmp_fwd(xul, xul + NX, xul + NX + NU, p.th, pd_at(p.pd, tms[lane], pdb), sl + FWS_FXC, sl + FWS_FUC, sl + FWS_FE,
                       sc, sc + NX * NU, sc + NX * NU + NU * NP);
        double* huu = sc + NX * NU + NU * NP; double* hinv = huu + NU * NU;
        bool ok = true;
        if (Model::HUU_DIAG) {
            for (int i = 0; i < NU; ++i) {
                const double d = huu[i * NU + i];
                if (!(fabs(d) > 0.0)) ok = false;
                for (int jj = 0; jj < NU; ++jj) hinv[i * NU + jj] = (i == jj) ? 1.0 / d : 0.0;
            }
        } else {
            ok = inv_small<NU>(huu, hinv);
        }
        if (!ok) flag[0] = 0;
    }
    BDF_SYNC();
    if (flag[0] == 0) return false;
    CPDP_LOOP for (int q = lane; q < cnt * (NX + NP); q += FW_THREADS) {
        const int s = q / (NX + NP), j = q % (NX + NP);
        const double t = tms[s];
        const int lo = interp_lo(t, p.dt, p.N);
        const double xlo = p.dt * lo, xhi = p.dt * (lo + 1);
        const double* P0; const double* P1;
        fw_rows(sm, p, lo, P0, P1);
        double* sl = sm + fo::SL + s * FW_SLOT;
        const double* sc = sm + fo::SCR + s * FW_SCR;
        const double* fuc = sl + FWS_FUC;
        double yz[NU];
        BDF_UNROLL for (int a = 0; a < NU; ++a) yz[a] = (j < NX) ? sc[j * NU + a] : sc[NX * NU + a * NP + (j - NX)];
        double pv = 0.0;
        BDF_UNROLL for (int pp = 0; pp < Model::FU_nnz; ++pp) {
            if (pp == 0 || Model::FU_coor(pp) != Model::FU_coor(pp - 1)) {      // (compile-time: one interpolation per non-zero row of fu)
                const int a = Model::FU_coor(pp);
                const int e = (j < NX) ? (a <= j ? tri(a, j) : tri(j, a)) : NT + a * NP + (j - NX);
                pv = interp_val(P0[e], P1[e], xlo, xhi, t);
            }
            yz[Model::FU_cooc(pp)] += fuc[pp] * pv;
        }
        const double* hinv = sc + NX * NU + NU * NP + NU * NU;
        BDF_UNROLL for (int a = 0; a < NU; ++a) {
            double acc = 0.0;
            BDF_UNROLL for (int b2 = 0; b2 < NU; ++b2) acc += hinv[a * NU + b2] * yz[b2];
            if (j < NX) sl[FWS_HY + a * NX + j] = -acc; else sl[FWS_HZ + a * NP + (j - NX)] = -acc;
        }
    }
    BDF_SYNC();
    return true;
}

// kout[:, k] = fx xin[:, k] + fu (HY xin[:, k] + HZ[:, k]) + fe[:, k]     (columns through shared memory, stride NPS)
CPDP_D_NOINLINE void fw_rhs(const int slot, const double* __restrict__ xin, double* __restrict__ kout) {
    CPDP_DYN_SMEM(sm);
    const int k = threadIdx.x;
    if (k >= NP) return;
    const double* sl = sm + fo::SL + slot * FW_SLOT;
    const double* fxc = sl + FWS_FXC; const double* fuc = sl + FWS_FUC; const double* fe = sl + FWS_FE;
    const double* HY = sl + FWS_HY; const double* HZ = sl + FWS_HZ;
    double x[NX], uc[NU], xd[NX];
    BDF_UNROLL for (int i = 0; i < NX; ++i) x[i] = xin[i * NPS + k];
    BDF_UNROLL for (int a = 0; a < NU; ++a) {
        double acc = HZ[a * NP + k];
        BDF_UNROLL for (int c = 0; c < NX; ++c) acc += HY[a * NX + c] * x[c];
        uc[a] = acc;
    }
    BDF_UNROLL for (int i = 0; i < NX; ++i) xd[i] = fe[i * NP + k];
    BDF_UNROLL for (int pp = 0; pp < Model::FX_nnz; ++pp) xd[Model::FX_coor(pp)] += fxc[pp] * x[Model::FX_cooc(pp)];
    BDF_UNROLL for (int pp = 0; pp < Model::FU_nnz; ++pp) xd[Model::FU_coor(pp)] += fuc[pp] * uc[Model::FU_cooc(pp)];
    BDF_UNROLL for (int i = 0; i < NX; ++i) kout[i * NPS + k] = xd[i];
}

// One grid interval [t0, t1] forward with scipy's RK45; y: my column (registers, in/out).  On entry slot 0 holds the stage data
// of t0 (the caller prepared it for the node's aux control).  Returns 0 ok, 1 step too small, 2 non-finite.
CPDP_D int fw_interval(double* sm, const AuxProblem& p, const double t0, const double t1, const double rtol, const double atol,
                       double (&y)[NX], int& nrhs, int& nsteps) {
    const int k = threadIdx.x;
    const bool act = k < NP;
    const int kc = act ? k : 0;
    double* K = sm + fo::KS; double* XSb = sm + fo::XS; double* tms = sm + fo::TMS;
    const double NF = (double)NYF;
#define K_(j, i) K[((j) * NX + (i)) * NPS + kc]
    // f0: slot 0 holds the data of t0 (left by the previous interval / the caller)
    if (act) { BDF_UNROLL for (int i = 0; i < NX; ++i) XSb[i * NPS + kc] = y[i]; }
    BDF_SYNC();
    fw_rhs(0, XSb, K); ++nrhs;
    BDF_SYNC();
    // select_initial_step (common.py:68-134), order = 4
    double h_abs;
    {
        const double interval_length = fabs(t1 - t0);
        double a0 = 0.0, a1 = 0.0;
        if (act) {
            BDF_UNROLL for (int i = 0; i < NX; ++i) {
                const double sc = atol + fabs(y[i]) * rtol;
                const double f = K_(0, i);
                a0 += (y[i] / sc) * (y[i] / sc);
                a1 += (f / sc) * (f / sc);
            }
        }
        const double d0 = sqrt(fw_reduce(a0, false) / NF);
        const double d1 = sqrt(fw_reduce(a1, false) / NF);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval_length);
        if (act) { BDF_UNROLL for (int i = 0; i < NX; ++i) XSb[i * NPS + kc] = y[i] + h0 * K_(0, i); }
        BDF_SYNC();
        if (k == 0) tms[0] = t0 + h0;                           // the data of t0 is no longer needed: K[0] = f(t0, y) is in hand
        if (!fw_prepare(p, 1)) return 2;
        fw_rhs(0, XSb, K + 1 * NX * NPS); ++nrhs;              // f1 parked in K[1]
        BDF_SYNC();
        double a2 = 0.0;
        if (act) {
            BDF_UNROLL for (int i = 0; i < NX; ++i) {
                const double sc = atol + fabs(y[i]) * rtol;
                const double v = (K_(1, i) - K_(0, i)) / sc;
                a2 += v * v;
            }
        }
        const double d2 = sqrt(fw_reduce(a2, false) / NF) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), interval_length);
    }
    double t = t0;
    while (t - t1 < 0) {
        const double min_step = 10 * fabs(nextafter(t, (double)INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        bool rejected = false;
        double t_new = t;
        double yn[NX];
        while (true) {
            if (h_abs < min_step) return 1;
            double h = h_abs;
            t_new = t + h;
            if (t_new - t1 > 0) t_new = t1;
            h = t_new - t;
            h_abs = fabs(h);
            BDF_SYNC();
            if (k < NSLOT) tms[k] = t + dp_C(k + 1) * h;
            if (!fw_prepare(p, NSLOT)) return 2;
            CPDP_LOOP for (int st = 1; st < 6; ++st) {
                if (act) {
                    BDF_UNROLL for (int i = 0; i < NX; ++i) {
                        double acc = 0.0;
                        CPDP_LOOP for (int j = 0; j < st; ++j) acc += K_(j, i) * sm[fo::TAB + st * 5 + j];
                        XSb[i * NPS + kc] = y[i] + acc * h;
                    }
                }
                BDF_SYNC();
                fw_rhs(dp_slot(st), XSb, K + st * NX * NPS); ++nrhs;
                BDF_SYNC();
            }
            if (act) {
                BDF_UNROLL for (int i = 0; i < NX; ++i) {
                    double acc = 0.0;
                    BDF_UNROLL for (int j = 0; j < 6; ++j) acc += K_(j, i) * dp_B(j);
                    yn[i] = y[i] + h * acc;
                    XSb[i * NPS + kc] = yn[i];
                }
            }
            BDF_SYNC();
            fw_rhs(4, XSb, K + 6 * NX * NPS); ++nrhs;
            BDF_SYNC();
            double ae = 0.0, fin = 0.0;
            if (act) {
                BDF_UNROLL for (int i = 0; i < NX; ++i) {
                    double acc = 0.0;
                    BDF_UNROLL for (int j = 0; j < 7; ++j) acc += K_(j, i) * dp_E(j);
                    const double sc = atol + fmax(fabs(y[i]), fabs(yn[i])) * rtol;
                    const double v = acc * h / sc;
                    ae += v * v;
                    if (!(fabs(yn[i]) < 1e300)) fin = 1.0;
                }
            }
            const double error_norm = sqrt(fw_reduce(ae, false) / NF);
            fin = fw_reduce(fin, true);
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
        if (act) { BDF_UNROLL for (int i = 0; i < NX; ++i) { y[i] = yn[i]; K_(0, i) = K_(6, i); } }
        BDF_SYNC();
    }
#undef K_
    return 0;
}

CPDP_GLOBAL void __launch_bounds__(FW_THREADS) k_aux_forward(AuxArgs a) {
    const int b = blockIdx.x, k = threadIdx.x;
    if (a.aux_status[b] != 0) {
        if (k == 0) a.loss[b] = 0.0;
        for (int i = k; i < NP; i += FW_THREADS) a.dtheta[(size_t)b * NP + i] = 0.0;
        return;
    }
    CPDP_DYN_SMEM(sm);
    const bool act = k < NP;
    const int kc = act ? k : 0;
    const int N = a.N;
    AuxProblem p;
    p.X = a.X + (size_t)b * (N + 1) * NX; p.U = a.U + (size_t)b * (N + 1) * NU; p.Lam = a.Lam + (size_t)b * (N + 1) * NX;
    p.th = a.theta + (size_t)b * a.theta_stride; p.pd = a.pdata + (size_t)b * NQ; p.PW = a.PW + (size_t)b * (N + 1) * NYR; p.dt = a.T / N; p.N = N;
    double* Xa = a.Xa + (size_t)b * (N + 1) * NYF;
    double* Ua = a.Ua + (size_t)b * (N + 1) * NU * NP;
    CPDP_LOOP for (int q = k; q < NSLOT * FW_SLOT; q += FW_THREADS) sm[fo::SL + q] = 0.0;       // structural zeros of fe
    CPDP_LOOP for (int q = k; q < NSLOT * FW_SCR; q += FW_THREADS) sm[fo::SCR + q] = 0.0;      //   and of Hxu, Hue, Huu
    if (k < 30) sm[fo::TAB + k] = dp_A(k / 5, k % 5);
    // node-row ring: rows 0 and 1 now, row node + 2 while interval `node` integrates
    int* rg = (int*)(sm + fo::MBAR) + 2;                        // ring_lo, ring_hi
#ifdef __CUDACC__
    const bool tma = FW_RING_TMA && (((size_t)p.PW & 15) == 0);
    if (tma && k == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(fw_smem_addr(sm + fo::MBAR)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#else
    const bool tma = false;
#endif
    if (k == 0) { rg[0] = 0; rg[1] = -1; }
    BDF_SYNC();
    unsigned n_issued = 0, n_waited = 0;                        // copies requested / consumed (mbarrier phase = count & 1)
    fw_ring_issue(sm, p, 0, 2, tma); ++n_issued;                // (N >= 1: rows 0 and 1 exist)
    fw_ring_wait(sm, n_waited & 1u, tma); ++n_waited;
    if (k == 0) rg[1] = 1;
    double y[NX];
    BDF_UNROLL for (int i = 0; i < NX; ++i) y[i] = 0.0;
    if (act) { BDF_UNROLL for (int i = 0; i < NX; ++i) Xa[i * NP + kc] = 0.0; }
    BDF_SYNC();
    int nrhs = 0, nsteps = 0, st = 0;
    CPDP_LOOP for (int node = 0; node <= N && st == 0; ++node) {
        // stage data of the node time in slot 0: aux control at the node (CPDP.py:363-364,370-378) and, for node < N, the first
        // derivative of the next interval
        if (node >= 1 && node + 1 <= N) {                       // row node + 1, requested while the previous interval ran
            fw_ring_wait(sm, n_waited & 1u, tma); ++n_waited;
            if (k == 0) rg[1] = node + 1;
        }
        if (k == 0) sm[fo::TMS] = p.dt * node;
        if (!fw_prepare(p, 1)) { st = 2; break; }
        if (act) {
            const double* sl = sm + fo::SL;
            BDF_UNROLL for (int aa = 0; aa < NU; ++aa) {
                double acc = sl[FWS_HZ + aa * NP + kc];
                BDF_UNROLL for (int c = 0; c < NX; ++c) acc += sl[FWS_HY + aa * NX + c] * y[c];
                Ua[(size_t)node * NU * NP + aa * NP + kc] = acc;
            }
        }
        if (node == N) break;
        if (node + 2 <= N) {                                    // its slot holds row node - 1, last read by the prepare above
            BDF_SYNC();
            if (k == 0 && node >= 1) rg[0] = node;
            BDF_SYNC();
            fw_ring_issue(sm, p, node + 2, 1, tma); ++n_issued;
        }
        st = fw_interval(sm, p, p.dt * node, p.dt * (node + 1), a.rtol_f, a.atol_f, y, nrhs, nsteps);
        if (act) { BDF_UNROLL for (int i = 0; i < NX; ++i) Xa[(size_t)(node + 1) * NYF + i * NP + kc] = y[i]; }
        BDF_SYNC();
    }
    while (n_waited < n_issued) { fw_ring_wait(sm, n_waited & 1u, tma); ++n_waited; }       // (a failed interval leaves its prefetch in flight)
    if (k == 0) { a.aux_status[b] = st; a.counters[b * NCOUNTERS + 2] = nrhs; a.counters[b * NCOUNTERS + 3] = nsteps; }
#ifdef __CUDACC__
    __threadfence_block();
#endif
    BDF_SYNC();
    // ---- loss and gradient: lane 0 the loss, parameter i by lane i % 32.  A waypoint time outside [0, T] is an error (scipy's
    //      interp1d raises ValueError in the reference): status 5.
    const double* taus = a.taus + (size_t)b * a.taus_stride;
    const double* wp = a.wp + (size_t)b * a.W * a.D;
    bool tau_bad = false;
    for (int w = 0; w < a.W; ++w) if (!(taus[w] >= 0.0 && taus[w] <= p.dt * N)) tau_bad = true;
    if (tau_bad && st == 0) { st = 5; if (k == 0) a.aux_status[b] = 5; }
    if (k == 0) {
        double lo_ = 0.0;
        for (int w = 0; w < a.W && st == 0; ++w) {
            const double t = taus[w];
            const int lo = interp_lo(t, p.dt, N);
            const double xlo = p.dt * lo, xhi = p.dt * (lo + 1);
            for (int d = 0; d < a.D; ++d) {
                const int si = a.sel[d];
                const double yv = interp_val(p.X[(size_t)lo * NX + si], p.X[(size_t)(lo + 1) * NX + si], xlo, xhi, t);
                const double diff = yv - wp[(size_t)w * a.D + d];
                lo_ += diff * diff;
            }
        }
        a.loss[b] = lo_;
    }
    for (int i = k; i < NP; i += FW_THREADS) {
        double acc = 0.0;
        for (int w = 0; w < a.W && st == 0; ++w) {
            const double t = taus[w];
            const int lo = interp_lo(t, p.dt, N);
            const double xlo = p.dt * lo, xhi = p.dt * (lo + 1);
            for (int d = 0; d < a.D; ++d) {
                const int si = a.sel[d];
                const double yv = interp_val(p.X[(size_t)lo * NX + si], p.X[(size_t)(lo + 1) * NX + si], xlo, xhi, t);
                const double diff = yv - wp[(size_t)w * a.D + d];
                const double xa = interp_val(Xa[(size_t)lo * NYF + si * NP + i], Xa[(size_t)(lo + 1) * NYF + si * NP + i], xlo, xhi, t);
                acc += diff * xa;
            }
        }
        a.dtheta[(size_t)b * NP + i] = acc;
    }
}

}  // namespace CPDP_NS
