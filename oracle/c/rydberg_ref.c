/*
 * rydberg_ref.c — CPU ORACLE (C restatement), TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, matrix-free restatement of the reference hot path for sizes where the numpy/scipy
 * oracle (oracle/rydberg_oracle.py) is too slow, and the CPU baseline bench.py times beside the
 * GPU ("kind": "port").  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  It is validated against the numpy/scipy oracle (which is
 * itself pinned to the reference's golden vectors) by tests/test_oracle_c_port.py.
 *
 * What it restates (ICHEC/qse 1.1.21):
 *   ref_build_diagonal   ham_int's diagonal: sum_{i<j} V_ij n_i n_j in the order
 *                        Operators.to_qutip adds the terms (qse/calc/qutip.py:61-64,
 *                        qse/operator/operators.py:41-44); qubit q <-> bit N-1-q
 *                        (qse/magnetic.py:51).
 *   ref_apply_h          get_hamiltonian(amp, det) * state (qse/calc/qutip.py:55-58): the sparse
 *                        mat-vec inside qp.sesolve, here without forming the matrix.
 *   ref_evolve_schedule  the loop of Qutip.calculate (qse/calc/qutip.py:40-45); each
 *                        qp.sesolve(H, psi, [0, dt]) is the exact propagator exp(-i H dt), summed
 *                        as a norm-terminated Taylor series with sub-stepping (about as many H psi
 *                        per segment, ~9, as QuTiP's zvode/adams takes RHS evaluations, ~10:
 *                        SURVEY.md App. D.3).
 *   ref_spins / ref_number   qse/magnetic.py:150-176, 322-326.
 *
 * Threads: OpenMP over tiles/blocks; call ref_set_threads to bound them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

int ref_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* E[k] = sum_{i<j} V_ij b_i(k) b_j(k), pairs added in (i, j) lexicographic order from 0.0 */
void ref_build_diagonal(int n, const double* v, double* e) {
    const uint64_t dim = 1ull << n;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)dim; ++k) {
        double acc = 0.0;
        for (int i = 0; i < n - 1; ++i) {
            if (!((k >> (n - 1 - i)) & 1)) continue;
            for (int j = i + 1; j < n; ++j)
                if (((k >> (n - 1 - j)) & 1) && v[i * n + j] != 0.0) acc += v[i * n + j];
        }
        e[k] = acc;
    }
}

#define LOW_BITS 11   /* 2^11 amplitudes = 32 KiB: the cache tile for the low-qubit flips */
#define GROUP_BITS 4  /* high qubits handled 4 at a time: 16 rows of CHUNK amplitudes */
#define CHUNK 256

/* y = (E - delta*popcount) x + (omega/2) sum_q x[k ^ (1<<q)] */
void ref_apply_h(int n, const double* e, double omega, double delta, const cplx* x, cplx* y) {
    const uint64_t dim = 1ull << n;
    const double h = 0.5 * omega;
    const int tb = n < LOW_BITS ? n : LOW_BITS;
    const uint64_t tile = 1ull << tb;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)(dim >> tb); ++t) {
        const uint64_t base = (uint64_t)t << tb;
        for (uint64_t k = 0; k < tile; ++k) {
            double sr = 0.0, si = 0.0;
            for (int q = 0; q < tb; ++q) {
                const cplx p = x[base + (k ^ (1ull << q))];
                sr += p.re;
                si += p.im;
            }
            const double d = e[base + k] - delta * (double)__builtin_popcountll(base + k);
            y[base + k].re = d * x[base + k].re + h * sr;
            y[base + k].im = d * x[base + k].im + h * si;
        }
    }
    for (int q0 = tb; q0 < n; q0 += GROUP_BITS) {
        const int gb = (n - q0 < GROUP_BITS) ? n - q0 : GROUP_BITS;
        const int rows = 1 << gb;
        const uint64_t stride = 1ull << q0;                 /* distance between rows */
        const uint64_t chunk = stride < CHUNK ? stride : CHUNK;
        const uint64_t chunks_per_row = stride / chunk;
        const uint64_t outer = dim >> (q0 + gb);
#pragma omp parallel for schedule(static)
        for (int64_t w = 0; w < (int64_t)(outer * chunks_per_row); ++w) {
            const uint64_t o = (uint64_t)w / chunks_per_row, c = (uint64_t)w % chunks_per_row;
            const uint64_t base = (o << (q0 + gb)) + c * chunk;
            for (int r = 0; r < rows; ++r) {
                cplx* yr = y + base + (uint64_t)r * stride;
                for (int b = 0; b < gb; ++b) {
                    const cplx* xr = x + base + (uint64_t)(r ^ (1 << b)) * stride;
                    for (uint64_t l = 0; l < chunk; ++l) {
                        yr[l].re += h * xr[l].re;
                        yr[l].im += h * xr[l].im;
                    }
                }
            }
        }
    }
}


/* psi <- exp(-i H dt) psi; w0, w1: work vectors of 2^n amplitudes. Returns #H psi, <0 on failure */
int ref_evolve_segment(int n, const double* e, double emin, double emax, double omega, double delta, double dt,
                       cplx* psi, cplx* w0, cplx* w1, double tol, double theta_max) {
    const uint64_t dim = 1ull << n;
    const double hi = emax + (delta < 0 ? -delta : 0.0) * n, lo = emin - (delta > 0 ? delta : 0.0) * n;
    const double bound = fmax(fabs(hi), fabs(lo)) + 0.5 * fabs(omega) * n;
    int nsub = (int)ceil(bound * fabs(dt) / theta_max);
    if (nsub < 1) nsub = 1;
    const double h = dt / nsub;
    int count = 0;
    for (int s = 0; s < nsub; ++s) {
        const cplx* x = psi;
        cplx* ybuf[2] = {w0, w1};
        int done = 0;
        for (int j = 1; j <= 200; ++j) {
            cplx* y = ybuf[j & 1];
            ref_apply_h(n, e, omega, delta, x, y);
            ++count;
            const double cr = 0.0, ci = -h / (double)j; /* -i h / j */
            double nn = 0.0;
#pragma omp parallel for reduction(+ : nn) schedule(static)
            for (int64_t k = 0; k < (int64_t)dim; ++k) {
                const double re = cr * y[k].re - ci * y[k].im, im = cr * y[k].im + ci * y[k].re;
                y[k].re = re;
                y[k].im = im;
                psi[k].re += re;
                psi[k].im += im;
                nn += re * re + im * im;
            }
            x = y;
            if (sqrt(nn) <= tol && (double)j >= bound * fabs(h)) {
                done = 1;
                break;
            }
        }
        if (!done) return -1;
    }
    return count;
}

/* the loop of Qutip.calculate from the state in psi; returns total #H psi or <0 */
long ref_evolve_schedule(int n, const double* e, const double* omega, const double* delta, const double* dt,
                         long n_steps, cplx* psi, double tol, double theta_max) {
    const uint64_t dim = 1ull << n;
    double emin = e[0], emax = e[0];
    for (uint64_t k = 1; k < dim; ++k) {
        if (e[k] < emin) emin = e[k];
        if (e[k] > emax) emax = e[k];
    }
    cplx* w0 = (cplx*)malloc(dim * sizeof(cplx));
    cplx* w1 = (cplx*)malloc(dim * sizeof(cplx));
    if (!w0 || !w1) { free(w0); free(w1); return -2; }
    long total = 0;
    for (long s = 0; s < n_steps; ++s) {
        const int c = ref_evolve_segment(n, e, emin, emax, omega[s], delta[s], dt[s], psi, w0, w1, tol, theta_max);
        if (c < 0) { total = -1; break; }
        total += c;
    }
    free(w0);
    free(w1);
    return total;
}

/* (N, 3) <X>, <Y>, <Z>: qse/magnetic.py:150-176 */
void ref_spins(int n, const cplx* psi, double* out) {
    const uint64_t dim = 1ull << n;
    for (int q = 0; q < n; ++q) {
        const uint64_t m = 1ull << (n - 1 - q);
        double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma omp parallel for reduction(+ : sx, sy, sz) schedule(static)
        for (int64_t k = 0; k < (int64_t)dim; ++k) {
            const cplx a = psi[k], b = psi[k ^ m];
            const double z = (k & m) ? -1.0 : 1.0;
            /* c_k conj(c_k') = (a.re b.re + a.im b.im) + i (a.im b.re - a.re b.im) */
            sx += a.re * b.re + a.im * b.im;
            sy += -z * (a.im * b.re - a.re * b.im); /* Re(c_k conj(c_k') * i z) */
            sz += z * (a.re * a.re + a.im * a.im);
        }
        out[3 * q] = sx;
        out[3 * q + 1] = sy;
        out[3 * q + 2] = sz;
    }
}

void ref_number(int n, const cplx* psi, double* out) {
    const uint64_t dim = 1ull << n;
    for (int q = 0; q < n; ++q) {
        const uint64_t m = 1ull << (n - 1 - q);
        double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t k = 0; k < (int64_t)dim; ++k)
            if (k & m) s += psi[k].re * psi[k].re + psi[k].im * psi[k].im;
        out[q] = s;
    }
}
