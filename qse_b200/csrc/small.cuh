// small.cuh — whole-schedule kernel for launch-bound sizes (N <= 12 qubits, one GPU).
//
// BASELINE config 1 (square 3x3, 9 qubits, thousands of 1 ns steps) and the reference's notebook
// workloads (8-9 qubits, 5000 steps: docs/use-cases/adiabatic.ipynb) are latency bound: a sweep step
// is ~9 H psi of a 4-64 KiB vector.  Here ONE thread block keeps psi and the two Taylor work vectors
// in shared memory and runs the entire loop of Qutip.calculate (qse/calc/qutip.py:40-47) -- every
// step, sub-step, Taylor term, norm test and energy e_op -- in a single launch: no launch latency,
// no host round trip per term.  Same series, same termination rule as the multi-kernel stepper.
#pragma once
#include "common.cuh"

namespace qsb {

constexpr int kSmallMaxQubits = 12;
constexpr int kSmallThreads = 1024;

struct SmallArgs {
    double2* psi;          // in/out (global)
    const double* eint;    // interaction diagonal (global)
    const double* omega;   // [n_steps] (site_stride 0) or [n_steps * n] (site_stride n): per-site schedule rows
    const double* delta;
    const double* dt;
    const double* w_omega; // [n] per-site weights and static X coefficients (README.md:93-95 Hamiltonian)
    const double* w_delta;
    const double* x0;
    int site_stride;
    long long n_steps;
    int n;
    double tol, theta_max;
    int max_terms, fixed_terms;
    double diag_min, diag_max;
    int n_eops;                 // energy e_ops only
    const double* eop_omega;
    const double* eop_delta;
    double* expect_out;         // [n_steps * n_eops]
    unsigned long long* info;   // [0] H psi count, [1] sub-steps, [2] last term count, [3] error, [4] failing step
};

QSB_DEV double block_allreduce(double v, double* red, double* bcast) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
        w = warp_sum(w);
        if (lane == 0) *bcast = w;
    }
    __syncthreads();
    return *bcast;
}

__global__ void __launch_bounds__(kSmallThreads, 1) evolve_small_kernel(const SmallArgs a) {
    extern __shared__ __align__(128) unsigned char small_smem_raw[];
    const int dim = 1 << a.n;
    double2* psi = reinterpret_cast<double2*>(small_smem_raw);
    double2* ya = psi + dim;
    double2* yb = ya + dim;
    __shared__ double red[32];
    __shared__ double bcast;
    __shared__ double hw_s[kSmallMaxQubits], d_s[kSmallMaxQubits];  // per index bit: Omega_q / 2, delta_q
    const int tid = threadIdx.x;

    for (int k = tid; k < dim; k += kSmallThreads) psi[k] = a.psi[k];
    __syncthreads();

    unsigned long long n_hpsi = 0, n_sub = 0, last_terms = 0, err = 0, err_step = 0;
    for (long long s = 0; s < a.n_steps && !err; ++s) {
        const double dt = a.dt[s];
        __syncthreads();
        if (tid < a.n) {
            const int q = a.n - 1 - tid;  // index bit tid <-> qubit q
            const long long row = a.site_stride ? s * a.site_stride + q : s;
            hw_s[tid] = 0.5 * a.omega[row] * a.w_omega[q] + a.x0[q];
            d_s[tid] = a.delta[row] * a.w_delta[q];
        }
        __syncthreads();
        double hw_abs = 0.0, d_pos = 0.0, d_neg = 0.0;
        for (int b = 0; b < a.n; ++b) {
            hw_abs += fabs(hw_s[b]);
            d_pos += fmax(d_s[b], 0.0);
            d_neg += fmax(-d_s[b], 0.0);
        }
        const double hi = a.diag_max + d_neg, lo = a.diag_min - d_pos;
        const double theta = (fmax(fabs(hi), fabs(lo)) + hw_abs) * fabs(dt);
        int nsub = (int)ceil(theta / a.theta_max);
        if (nsub < 1) nsub = 1;
        const double h = dt / nsub, theta_sub = theta / nsub;
        for (int sub = 0; sub < nsub && !err && dt != 0.0; ++sub) {
            ++n_sub;
            const double2* x = psi;
            bool converged = false;
            int j = 1;
            for (; j <= a.max_terms; ++j) {
                double2* y = (j & 1) ? ya : yb;
                const double2 c = make_double2(0.0, -h / (double)j);  // -i h / j
                double nrm = 0.0;
                for (int k = tid; k < dim; k += kSmallThreads) {
                    const double2 xk = x[k];
                    double d = __ldg(a.eint + k);
                    for (int q = 0; q < a.n; ++q)
                        if ((k >> q) & 1) d -= d_s[q];
                    double2 v = make_double2(d * xk.x, d * xk.y);
                    for (int q = 0; q < a.n; ++q) {
                        const double2 w = x[k ^ (1 << q)];
                        v.x = fma(hw_s[q], w.x, v.x);
                        v.y = fma(hw_s[q], w.y, v.y);
                    }
                    v = cmul(c, v);
                    nrm += cnorm2(v);
                    y[k] = v;
                }
                nrm = block_allreduce(nrm, red, &bcast);  // also orders: every read of x precedes the psi update
                for (int k = tid; k < dim; k += kSmallThreads) cacc(psi[k], y[k]);
                ++n_hpsi;
                x = y;
                if (a.fixed_terms > 0) {
                    if (j == a.fixed_terms) { converged = true; break; }
                    continue;
                }
                if (!(nrm == nrm)) { err = 2; err_step = (unsigned long long)s; break; }
                if (sqrt(nrm) <= a.tol && (double)j >= theta_sub) { converged = true; break; }
            }
            if (!converged && !err) { err = 1; err_step = (unsigned long long)s; }
            last_terms = (unsigned long long)j;
            __syncthreads();  // psi complete before it is read as the next x
        }
        // energy e_ops after the step: <psi|H(omega_e, delta_e)|psi>
        for (int o = 0; o < a.n_eops && !err; ++o) {
            const double oe = a.eop_omega[o], de = a.eop_delta[o];
            double acc = 0.0;
            for (int k = tid; k < dim; k += kSmallThreads) {
                const double2 xk = psi[k];
                double d = __ldg(a.eint + k);
                double2 v = make_double2(0.0, 0.0);
                for (int b = 0; b < a.n; ++b) {
                    const int q = a.n - 1 - b;
                    if ((k >> b) & 1) d -= de * a.w_delta[q];
                    const double hq = 0.5 * oe * a.w_omega[q] + a.x0[q];
                    const double2 w = psi[k ^ (1 << b)];
                    v.x = fma(hq, w.x, v.x);
                    v.y = fma(hq, w.y, v.y);
                }
                acc += xk.x * (d * xk.x + v.x) + xk.y * (d * xk.y + v.y);  // Re conj(psi) (H psi)
            }
            acc = block_allreduce(acc, red, &bcast);
            if (tid == 0) a.expect_out[s * a.n_eops + o] = acc;
        }
    }
    __syncthreads();
    for (int k = tid; k < dim; k += kSmallThreads) a.psi[k] = psi[k];
    if (tid == 0) {
        a.info[0] = n_hpsi;
        a.info[1] = n_sub;
        a.info[2] = last_terms;
        a.info[3] = err;
        a.info[4] = err_step;
    }
}

}  // namespace qsb
