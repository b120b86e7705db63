// Learner kernels: the parameter update of the reference's quadrotor learner on the device, so that a whole learning run
// (evaluation point -> CPDP gradient iteration -> update -> projection -> stop rule -> traces) is a stream of kernel
// launches with no host round trip (CUDA-graph capturable).
// Reference: /root/reference/lib/QuadAlgorithm.py
//   update rules   Vanilla :454-466, Nesterov :469-494, Adam :497-520, Nadam :523-548, AMSGrad :551-578
//   loop           stop rule `loss > 0.9 and |dL| > 0.05`, projection theta[0] = max(theta[0], 1e-8)   :239-257
// Every expression is evaluated in the reference's association with separately rounded products (no FMA contraction), so a
// run reproduces the numpy arithmetic of the stored parameter_trace bit for bit given the same gradients.
#pragma once
#include "cpdp_aux.cuh"

namespace CPDP_NS {

enum OptimMethod { OPT_VANILLA = 0, OPT_NESTEROV = 1, OPT_ADAM = 2, OPT_NADAM = 3, OPT_AMSGRAD = 4 };

#ifdef __CUDACC__
CPDP_D double o_mul(double a, double b) { return __dmul_rn(a, b); }
CPDP_D double o_add(double a, double b) { return __dadd_rn(a, b); }
CPDP_D double o_sub(double a, double b) { return __dsub_rn(a, b); }
#else
CPDP_D double o_mul(double a, double b) { volatile double r = a * b; return r; }
CPDP_D double o_add(double a, double b) { volatile double r = a + b; return r; }
CPDP_D double o_sub(double a, double b) { volatile double r = a - b; return r; }
#endif

struct OptimArgs {
    int method;
    double lr, mu, beta1, beta2, eps;
    double loss_stop, grad_stop;       // the loop continues while loss > loss_stop and |dL| > grad_stop
    double* theta;                     // [NP]    current parameter (in/out)
    double* theta_eval;                // [NP]    point the next gradient iteration is evaluated at
    double* state;                     // [3][NP] Nesterov velocity | momentum vector ; velocity vector ; AMSGrad running max
    const double* red;                 // [1+NP(+..)]  [sum loss | sum dL/dtheta | ...] of the gradient iteration
    int* it;                           // [2]     iterations done ; stop flag
    double* loss_trace;                // [cap]
    double* param_trace;               // [cap+1][NP]  row 0 = initial parameter (written by the caller)
    int cap;
    int record_only;                   // 0: update + close the iteration; -1: update, leave the iteration open (a second evaluation
                                       // follows); 1: second evaluation (Nesterov true_loss_print_flag): record loss / gradient norm, close
};

// evaluation point of the next gradient iteration: theta + mu * velocity for Nesterov (QuadAlgorithm.py:480), theta otherwise
CPDP_GLOBAL void __launch_bounds__(32) k_optim_pre(OptimArgs a) {
    const int tid = threadIdx.x;
    for (int i = tid; i < NP; i += 32) {
        double v = a.theta[i];
        if (a.method == OPT_NESTEROV && a.record_only == 0) v = o_add(v, o_mul(a.mu, a.state[i]));
        a.theta_eval[i] = v;
    }
}

CPDP_GLOBAL void __launch_bounds__(32) k_optim_post(OptimArgs a) {
    CPDP_SHARED double s_norm[1];
    const int tid = threadIdx.x;
    if (a.it[1] != 0 || a.it[0] >= a.cap) return;              // stopped: the remaining launches of a captured run do nothing
    const int j = a.it[0];
    const double loss = a.red[0];
    if (tid == 0) {
        double n2 = 0.0;
        for (int i = 0; i < NP; ++i) n2 += a.red[1 + i] * a.red[1 + i];
        s_norm[0] = sqrt(n2);
    }
    __syncthreads();
    if (a.record_only <= 0) {
        const double idx = (double)(j + 1);
        for (int i = tid; i < NP; i += 32) {
            const double g = a.red[1 + i];
            double th = a.theta[i];
            double* s0 = a.state + i; double* s1 = a.state + NP + i; double* s2 = a.state + 2 * NP + i;
            if (a.method == OPT_VANILLA) {
                th = o_sub(th, o_mul(a.lr, g));
            } else if (a.method == OPT_NESTEROV) {
                const double v = o_sub(o_mul(a.mu, *s0), o_mul(a.lr, g));
                *s0 = v;
                th = o_add(th, v);
            } else {
                const double m = o_add(o_mul(a.beta1, *s0), o_mul(1.0 - a.beta1, g));
                const double v = o_add(o_mul(a.beta2, *s1), o_mul(1.0 - a.beta2, o_mul(g, g)));
                *s0 = m; *s1 = v;
                if (a.method == OPT_AMSGRAD) {
                    const double vh = fmax(*s2, v);
                    *s2 = vh;
                    th = o_sub(th, o_mul(a.lr, m) / o_add(sqrt(vh), a.eps));
                } else {
                    const double b1p = pow(a.beta1, idx), b2p = pow(a.beta2, idx);
                    const double mh = m / (1.0 - b1p), vh = v / (1.0 - b2p);
                    if (a.method == OPT_ADAM) th = o_sub(th, o_mul(a.lr, mh) / o_add(sqrt(vh), a.eps));
                    else th = o_sub(th, o_mul(a.lr, o_add(o_mul(a.beta1, mh), o_mul((1.0 - a.beta1) / (1.0 - b1p), g))) / o_add(sqrt(vh), a.eps));
                }
            }
            if (i == 0) th = fmax(th, 1e-8);                       // projection (QuadAlgorithm.py:250)
            a.theta[i] = th;
            a.param_trace[(size_t)(j + 1) * NP + i] = th;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // with a second evaluation pending (Nesterov true_loss_print_flag) the iteration is closed by the record-only call
        if (a.record_only >= 0) {
            a.loss_trace[j] = loss;
            const bool go_on = (loss > a.loss_stop) && (s_norm[0] > a.grad_stop);
            a.it[0] = j + 1;
            if (!go_on) a.it[1] = 1;
        }
    }
}

// rows [loss | dL/dtheta | bad] of every problem, for the all-gather + fixed-tree sum; bad = 1 when the forward solve did not
// converge or an auxiliary sweep failed (such a row's loss / gradient may be zero or off-optimum: the count travels with the sum)
CPDP_GLOBAL void __launch_bounds__(256) k_pack_rows(const double* loss, const double* dtheta, const int* solve_status,
                                                     const int* aux_status, int B, double* rows) {
    const int C = NP + 2;
    const size_t gs = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < (size_t)B * C; q += gs) {
        const int b = (int)(q / C), c = (int)(q % C);
        double v;
        if (c == 0) v = loss[b];
        else if (c <= NP) v = dtheta[(size_t)b * NP + c - 1];
        else v = ((solve_status && solve_status[b] != ST_CONVERGED) || (aux_status && aux_status[b] != 0)) ? 1.0 : 0.0;
        rows[q] = v;
    }
}

// canonical pairwise (binary tree over the row index) sum of B rows of C columns; same tree as k_reduce_tree
CPDP_GLOBAL void __launch_bounds__(256) k_reduce_rows(const double* rows, int B, int C, double* scratch, double* out) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int P2 = 1; while (P2 < B) P2 <<= 1;
    for (int q = tid; q < P2 * C; q += nt) scratch[q] = (q / C < B) ? rows[q] : 0.0;
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
