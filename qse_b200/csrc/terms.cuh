// terms.cuh — operator-sum Hamiltonians beyond the Rydberg fast form.
//
// The reference builds H from `qse.Operators` (sums of coef * Pauli strings over "X", "Y", "Z", "N";
// qse/operator/operator.py:38-60, 80-92) and evolves it with qp.sesolve: Qutip.get_hamiltonian
// (qse/calc/qutip.py:55-58) for the Rydberg sweep, docs/use-cases/time_evo.ipynb for a generic spin
// model (TFIM), the XY / microwave channel of qse/calc/pulser.py:174-177 (sigma+ sigma- + h.c.).
//
// libqsb splits such an H(t) = amp(t) H_amp + det(t) H_det + H_static into
//   * the FAST FORM handled by the tiled H psi kernels: any diagonal static string (folded into the
//     E array by diag_terms_kernel), single-X strings (per-site drive) and single-N strings scaled by
//     the detuning (per-site detuning);
//   * everything else (Y, XX + YY, X Z, time-scaled diagonal strings, ...): a list of generic terms
//     applied by terms_kernel, one gather per term:
//         y[k] += s * i^ny * (-1)^popc(k' & zy) * [k' & nm == nm] * x[k'],   k' = k ^ flip
//     (Y = i X Z, Z and N act on the source index k'; the masks of one string are disjoint).
#pragma once
#include "common.cuh"

namespace qsb {

struct GenTerm {
    double coef;          // real coefficient (sign of sharded Z / Y bits folded in per rank)
    u64 flip;             // x | y, LOCAL index bits
    u64 zy;               // y | z, LOCAL index bits
    u64 nm;               // n, LOCAL index bits
    int ny;               // number of Y factors mod 4
    int chan;             // 0 static, 1 x amplitude(t), 2 x detuning(t)
    int src;              // rank whose shard holds x[k ^ flip] (this rank when the flip is local)
    int pad;
};

struct TermsArgs {
    const GenTerm* terms;
    int n_terms;
    const double2* xsrc[8];  // x of every rank (peer-mapped); xsrc[rank] is the local one
    double2* y;
    double chan[3];
    int nl;
};

__global__ void __launch_bounds__(256) terms_kernel(const TermsArgs a) {
    const u64 dim = 1ull << a.nl;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        double2 acc = a.y[k];
        for (int t = 0; t < a.n_terms; ++t) {
            const GenTerm g = a.terms[t];
            const u64 ks = k ^ g.flip;
            if ((ks & g.nm) != g.nm) continue;
            double s = g.coef * a.chan[g.chan];
            if (__popcll(ks & g.zy) & 1) s = -s;
            const double2 xv = a.xsrc[g.src][ks];
            switch (g.ny & 3) {  // i^ny
                case 0: acc.x = fma(s, xv.x, acc.x); acc.y = fma(s, xv.y, acc.y); break;
                case 1: acc.x = fma(-s, xv.y, acc.x); acc.y = fma(s, xv.x, acc.y); break;
                case 2: acc.x = fma(-s, xv.x, acc.x); acc.y = fma(-s, xv.y, acc.y); break;
                default: acc.x = fma(s, xv.y, acc.x); acc.y = fma(-s, xv.x, acc.y); break;
            }
        }
        a.y[k] = acc;
    }
}

// E[k] += sum_t coef_t (-1)^popc(K & z_t) [K & n_t == n_t] over static diagonal strings, K the GLOBAL index
struct DiagTerm {
    double coef;
    u64 z, nm;  // GLOBAL index bits
};
__global__ void __launch_bounds__(256) diag_terms_kernel(double* __restrict__ e, const DiagTerm* __restrict__ terms, int n_terms,
                                                         int nl, u64 rank_bits) {
    const u64 dim = 1ull << nl;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const u64 full = rank_bits | k;
        double acc = e[k];
        for (int t = 0; t < n_terms; ++t) {
            const DiagTerm g = terms[t];
            if ((full & g.nm) != g.nm) continue;
            acc += (__popcll(full & g.z) & 1) ? -g.coef : g.coef;
        }
        e[k] = acc;
    }
}

// The Taylor epilogue as a kernel of its own (operator terms follow the fast H psi, so it cannot be
// fused into the last tile pass): out = scale * y;  red = |out|^2 or Re conj(x) out;  psi += out [+ x]
__global__ void __launch_bounds__(256) epilogue_kernel(const double2* __restrict__ x, double2* __restrict__ y, double2* acc,
                                                       double2 scale, int acc_mode, int red_mode, u64 dim, double* part) {
    __shared__ double red[32];
    double r = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const double2 o = cmul(scale, y[k]);
        const double2 xk = x[k];
        r += (red_mode == 1) ? (xk.x * o.x + xk.y * o.y) : cnorm2(o);
        if (acc_mode != 0) {
            double2 v = cadd(acc[k], o);
            if (acc_mode == 2) cacc(v, xk);
            acc[k] = v;
        }
        y[k] = o;
    }
    const double tot = block_sum(r, red);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) axpy_kernel(double2* __restrict__ acc, const double2* __restrict__ x, u64 dim) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        double2 v = acc[k];
        cacc(v, x[k]);
        acc[k] = v;
    }
}

}  // namespace qsb
