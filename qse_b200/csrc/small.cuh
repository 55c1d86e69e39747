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
    const double* omega;
    const double* delta;
    const double* dt;
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
    const int tid = threadIdx.x;

    for (int k = tid; k < dim; k += kSmallThreads) psi[k] = a.psi[k];
    __syncthreads();

    unsigned long long n_hpsi = 0, n_sub = 0, last_terms = 0, err = 0, err_step = 0;
    for (long long s = 0; s < a.n_steps && !err; ++s) {
        const double omega = a.omega[s], delta = a.delta[s], dt = a.dt[s];
        const double half_omega = 0.5 * omega;
        const double hi = a.diag_max + fmax(-delta, 0.0) * a.n, lo = a.diag_min - fmax(delta, 0.0) * a.n;
        const double theta = (fmax(fabs(hi), fabs(lo)) + 0.5 * fabs(omega) * a.n) * fabs(dt);
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
                    double2 sum = make_double2(0.0, 0.0);
                    for (int q = 0; q < a.n; ++q) cacc(sum, x[k ^ (1 << q)]);
                    const double d = __ldg(a.eint + k) - delta * (double)__popc(k);
                    double2 v = make_double2(d * xk.x + half_omega * sum.x, d * xk.y + half_omega * sum.y);
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
            const double ho = 0.5 * a.eop_omega[o], de = a.eop_delta[o];
            double acc = 0.0;
            for (int k = tid; k < dim; k += kSmallThreads) {
                const double2 xk = psi[k];
                double2 sum = make_double2(0.0, 0.0);
                for (int q = 0; q < a.n; ++q) cacc(sum, psi[k ^ (1 << q)]);
                const double d = __ldg(a.eint + k) - de * (double)__popc(k);
                acc += xk.x * (d * xk.x + ho * sum.x) + xk.y * (d * xk.y + ho * sum.y);  // Re conj(psi) (H psi)
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
