// observe.cuh — diagonal builder (K1), Pauli-string reductions (K6), probabilities / top-k (K7).
#pragma once
#include "common.cuh"
#include "hpsi.cuh"

namespace qsb {

// ---- K1: E_int[k] = sum_{i<j} V_ij b_i(k) b_j(k) ---------------------------------------------
// Replaces ham_int = compute_interaction_hamiltonian(...).to_qutip() (qse/calc/qutip.py:61-64).
// The pair terms are added in (i, j) lexicographic order starting from 0.0, the order
// Operators.to_qutip (qse/operator/operators.py:41-44) adds them, without FMA contraction, so
// the result is bit-identical to the sequential host sum (oracle: diagonal_energies).
// vtab[i * n + j] (i < j) holds V_ij (0 when dropped by the 1e-8 cut-off).
__global__ void __launch_bounds__(256) diag_build_kernel(double* __restrict__ eint, const double* __restrict__ vtab,
                                                         int n, int nl, u64 rank_bits) {
    extern __shared__ double vs[];
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) vs[i] = vtab[i];
    __syncthreads();
    const u64 dim = 1ull << nl;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const u64 full = rank_bits | k;  // global basis index; qubit q <-> bit n-1-q
        double e = 0.0;
        u64 rest_i = full;
        // iterate set bits from the most significant (qubit 0) downwards = i ascending
        while (rest_i) {
            const int bi = 63 - __clzll(rest_i);
            rest_i &= ~(1ull << bi);
            const int i = n - 1 - bi;
            u64 rest_j = rest_i;  // bits below bi = qubits j > i
            while (rest_j) {
                const int bj = 63 - __clzll(rest_j);
                rest_j &= ~(1ull << bj);
                const double v = vs[i * n + (n - 1 - bj)];
                if (v != 0.0) e = __dadd_rn(e, v);
            }
        }
        eint[k] = e;
    }
}

// min / max of a real vector: per-block partials {min, max}
__global__ void __launch_bounds__(256) minmax_kernel(const double* __restrict__ v, u64 n, double* part) {
    __shared__ double red[32];
    double lo = 1.0e308, hi = -1.0e308;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) {
        const double x = v[k];
        lo = fmin(lo, x);
        hi = fmax(hi, x);
    }
    const double bhi = block_max(hi, red);
    const double blo = -block_max(-lo, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = blo;
        part[2 * blockIdx.x + 1] = bhi;
    }
}
__global__ void minmax_final_kernel(const double* part, int n, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double lo = 1.0e308, hi = -1.0e308;
        for (int i = 0; i < n; ++i) {
            lo = fmin(lo, part[2 * i]);
            hi = fmax(hi, part[2 * i + 1]);
        }
        out[0] = lo;
        out[1] = hi;
    }
}

__global__ void __launch_bounds__(256) basis_state_kernel(double2* psi, u64 dim, u64 one_at) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x)
        psi[k] = make_double2(k == one_at ? 1.0 : 0.0, 0.0);
}

__global__ void __launch_bounds__(256) copy_kernel(double2* __restrict__ dst, const double2* __restrict__ src, u64 dim) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x)
        dst[k] = src[k];
}

// write-only sweep used to evict L2 between timed launches
__global__ void __launch_bounds__(256) flush_kernel(double* buf, u64 n, double v) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (u64)gridDim.x * blockDim.x) buf[k] = v;
}

// ---- K6: generic Pauli-string reduction ---------------------------------------------------------
//   r = sum_{k local, (k & nmask) == nmask} conj(a[k ^ flip]) * sgn(k) * b[k]
// with sgn(k) = (-1)^popc(k & zmask)   (mode 0)     or     1 - (-1)^popc(k & zmask)   (mode 1).
// <psi|P|psi> for P = X(xm) Y(ym) Z(zm) N(nm): flip = xm|ym, zmask = ym|zm, nmask = nm, times the
// host-side factor i^popc(ym) (and the rank-dependent sign / selection of sharded bits).
// `a` may point into a peer GPU's shard when the flip touches a sharded qubit.
struct PauliArgs {
    const double2* a;  // bra side: read at k ^ flip
    const double2* b;  // ket side: read at k
    u64 flip, zmask, nmask;
    int mode;
    int nl;
    double* part;  // 2 doubles per block
};
__global__ void __launch_bounds__(256) pauli_kernel(const PauliArgs p) {
    __shared__ double red[32];
    const u64 dim = 1ull << p.nl;
    double sr = 0.0, si = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        if ((k & p.nmask) != p.nmask) continue;
        const double sg = (__popcll(k & p.zmask) & 1) ? -1.0 : 1.0;
        const double w = p.mode ? (1.0 - sg) : sg;
        if (w == 0.0) continue;
        const double2 bk = p.b[k];
        const double2 ak = (p.flip == 0 && p.a == p.b) ? bk : p.a[k ^ p.flip];
        const double2 c = cconj_mul(ak, bk);
        sr += w * c.x;
        si += w * c.y;
    }
    const double tr = block_sum(sr, red);
    const double ti = block_sum(si, red);
    if (threadIdx.x == 0) {
        p.part[2 * blockIdx.x] = tr;
        p.part[2 * blockIdx.x + 1] = ti;
    }
}
// reduce the per-block {re, im} pairs into out[0..1]
// <X_iX_j + Y_iY_j> for two LOCAL qubits (qse.magnetic.get_sisj, qse/magnetic.py:224-238: the basis states with
// b_i != b_j paired by the double flip): only a QUARTER of the index space is visited -- every index with
// (b_i, b_j) = (1, 0) together with its partner (0, 1) -- and every amplitude that contributes is read exactly once
// (8 B per amplitude of the state instead of the 16 B of the masked full sweep):
//     sum_k conj(c_{k^f}) (1 - z_i z_j) c_k = 4 sum_{pairs} Re conj(c_a) c_b
__global__ void __launch_bounds__(256) xxyy_pair_kernel(const double2* __restrict__ psi, int nl, int bit_i, int bit_j, double* part) {
    __shared__ double red[32];
    const int lo = min(bit_i, bit_j), hi = max(bit_i, bit_j);
    const u64 n = 1ull << (nl - 2);
    const u64 lo_mask = (1ull << lo) - 1ull, hi_mask = (1ull << hi) - 1ull;
    double sr = 0.0;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (u64)gridDim.x * blockDim.x) {
        u64 k = ((idx & ~lo_mask) << 1) | (idx & lo_mask);  // a zero at bit lo ...
        k = ((k & ~hi_mask) << 1) | (k & hi_mask);          // ... and at bit hi
        const double2 a = psi[k | (1ull << bit_i)], b = psi[k | (1ull << bit_j)];
        sr += a.x * b.x + a.y * b.y;
    }
    const double tr = block_sum(4.0 * sr, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = tr;
        part[2 * blockIdx.x + 1] = 0.0;
    }
}

__global__ void __launch_bounds__(256) finalize_pair_kernel(const double* part, int n, double* out) {
    __shared__ double red[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        a += part[2 * i];
        b += part[2 * i + 1];
    }
    a = block_sum(a, red);
    b = block_sum(b, red);
    if (threadIdx.x == 0) {
        out[0] = a;
        out[1] = b;
    }
}

// ---- <Z_q>, <n_q> for ALL local qubits and <psi|psi> in one pass ---------------------------------
// out partials per block: [0] = sum p_k, [1 + q] = sum p_k * bit_q(k)  (q = local bit position)
// cond: only amplitudes with (k & cond) == cond contribute (cond = 1 << b gives the joint counts n_bq)
__global__ void __launch_bounds__(256) marginals_kernel(const double2* __restrict__ psi, int nl, u64 cond, double* part) {
    __shared__ double red[32];
    const u64 dim = 1ull << nl;
    double tot = 0.0;
    double cnt[QSB_MAX_QUBITS_DEV];
#pragma unroll
    for (int q = 0; q < QSB_MAX_QUBITS_DEV; ++q) cnt[q] = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        if ((k & cond) != cond) continue;
        const double pk = cnorm2(psi[k]);
        tot += pk;
#pragma unroll
        for (int q = 0; q < QSB_MAX_QUBITS_DEV; ++q)
            if ((k >> q) & 1ull) cnt[q] += pk;
    }
    const int stride = nl + 1;
    const double t = block_sum(tot, red);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * stride] = t;
    // fully unrolled with a uniform guard: a run-time index would put cnt[] in local memory (it did: 272 bytes of
    // stack, 0.73 ms per pass at 25 qubits instead of the 0.1 ms the 16 B/amplitude read takes)
#pragma unroll
    for (int q = 0; q < QSB_MAX_QUBITS_DEV; ++q) {
        if (q < nl) {
            const double c = block_sum(cnt[q], red);
            if (threadIdx.x == 0) part[(size_t)blockIdx.x * stride + 1 + q] = c;
        }
    }
}
__global__ void __launch_bounds__(256) finalize_rows_kernel(const double* part, int n_blocks, int stride, double* out) {
    // one block per column
    __shared__ double red[32];
    const int col = blockIdx.x;
    double v = 0.0;
    for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) v += part[(size_t)i * stride + col];
    v = block_sum(v, red);
    if (threadIdx.x == 0) out[col] = v;
}

// ---- K6 tiled: <X_q>, <Y_q> of every qubit whose index bit is active in a tile pass, all in ONE read ---------
// qse.magnetic.get_spins (qse/magnetic.py:157-173) pairs every amplitude with its bit-q partner.  With the tile
// machinery of H psi (hpsi.cuh: a tile of 2^12 amplitudes in shared memory, 8 per thread, register partners for
// the top three tile bits, XOR-permuted LDS.128 for the others) one pass over the state serves EVERY qubit whose
// index bit lies in the tile: the contiguous pass covers index bits 0..11 -- and, seeing each amplitude exactly
// once, the Z-basis marginals <n_q> of ALL qubits and the norm too --, each strided pass up to nine more bits.
// 25 qubits: 3 reads of the state instead of 1 + 2 * 25.
//   sx_b = sum_k Re conj(c_k) c_{k^m}       sy_b = sum_k (-1)^{b(k)} (c.x p.y - c.y p.x)     (p = c_{k^m})
// Output per CTA (row of kSpinsRow doubles): sx[12], sy[12] by TILE bit, then (FIRST) norm, n[tile bits 0..11],
// n[index bits 12..]; summed over CTAs by finalize_rows_kernel in a fixed order (deterministic).
constexpr int kSpinsRow = 24 + 1 + 12 + 24;
template <bool FIRST>
__global__ void __launch_bounds__(kThreads, 1) spins_tile_kernel(const HpsiPass p, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ double red[32];
    __shared__ double nh_s[kThreads / 32][24];  // per warp: probability mass by set index bit >= 12 of the tile base

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const u64 first_tile = blockIdx.x;
    const u64 stride = gridDim.x;
    const u64 n_my = (p.n_tiles > first_tile) ? (p.n_tiles - first_tile + stride - 1) / stride : 0;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    if (lane < 24) nh_s[warp][lane] = 0.0;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            if ((u64)s < n_my) tile_issue(p, tiles + (size_t)s * kTile, &full_bar[s], first_tile + s * stride);
    }
    const int act_lo = FIRST ? 0 : p.lb;
    double sx[kTileBits], sy[kTileBits];
#pragma unroll
    for (int b = 0; b < kTileBits; ++b) sx[b] = sy[b] = 0.0;
    double ptot = 0.0, pj[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) pj[j] = 0.0;

    for (u64 i = 0; i < n_my; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t parity = (uint32_t)((i / kStages) & 1ull);
        const double2* ts = tiles + (size_t)s * kTile;
        mbar_wait(&full_bar[s], parity);
        double2 xv[kPerThread];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) xv[j] = ts[t + j * kThreads];
        // tile bits 9..11: partners are this thread's own registers; the own bit is bit b of j
#pragma unroll
        for (int b = 0; b < kRegBits; ++b)
            if (kSmemBits + b >= act_lo) {
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) {
                    const double2 c = xv[j], q = xv[j ^ (1 << b)];
                    sx[kSmemBits + b] += c.x * q.x + c.y * q.y;
                    const double im = c.x * q.y - c.y * q.x;
                    sy[kSmemBits + b] += ((j >> b) & 1) ? -im : im;
                }
            }
        // tile bits act_lo..8: XOR-permuted shared-memory partners; the own bit is bit b of t
        for (int b = act_lo; b < kSmemBits; ++b) {
            const int tb = t ^ (1 << b);
            const double sg = ((t >> b) & 1) ? -1.0 : 1.0;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const double2 c = xv[j], q = ts[tb + j * kThreads];
                ax += c.x * q.x + c.y * q.y;
                ay += c.x * q.y - c.y * q.x;
            }
            // dynamic register index would go to local memory: select with a compile-time loop
#pragma unroll
            for (int bb = 0; bb < kSmemBits; ++bb)
                if (bb == b) {
                    sx[bb] += ax;
                    sy[bb] += sg * ay;
                }
        }
        if (FIRST) {
            double pt = 0.0;
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const double pk = cnorm2(xv[j]);
                pj[j] += pk;
                pt += pk;
            }
            ptot += pt;
            // index bits >= 12 are fixed by the tile: its mass goes to every set one (one lane per bit)
            const u64 hi = (first_tile + i * stride);
            pt = warp_sum(pt);
            if (lane < 24 && ((hi >> lane) & 1ull)) nh_s[warp][lane] += pt;
        }
        __syncthreads();  // every thread is done reading stage s
        if (warp == 0 && i + kStages < n_my)
            tile_issue(p, tiles + (size_t)s * kTile, &full_bar[s], first_tile + (i + kStages) * stride);
    }
    double* row = part + (size_t)blockIdx.x * kSpinsRow;
    for (int b = 0; b < kTileBits; ++b) {
        double v = 0.0, w = 0.0;
#pragma unroll
        for (int bb = 0; bb < kTileBits; ++bb)
            if (bb == b) {
                v = sx[bb];
                w = sy[bb];
            }
        v = block_sum(v, red);
        w = block_sum(w, red);
        if (t == 0) {
            row[b] = v;
            row[kTileBits + b] = w;
        }
    }
    if (FIRST) {
        double v = block_sum(ptot, red);
        if (t == 0) row[24] = v;
        for (int b = 0; b < kSmemBits; ++b) {
            v = block_sum(((t >> b) & 1) ? ptot : 0.0, red);
            if (t == 0) row[25 + b] = v;
        }
        for (int b = 0; b < kRegBits; ++b) {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < kPerThread; ++j)
                if ((j >> b) & 1) a += pj[j];
            v = block_sum(a, red);
            if (t == 0) row[25 + kSmemBits + b] = v;
        }
        __syncthreads();
        if (t < 24) {
            double a = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) a += nh_s[w][t];
            row[37 + t] = a;
        }
    } else if (t < 1 + 12 + 24) {
        row[24 + t] = 0.0;
    }
}

// <a|b> = sum conj(a_k) b_k
__global__ void __launch_bounds__(256) dot_kernel(const double2* __restrict__ a, const double2* __restrict__ b, u64 dim,
                                                  double* part) {
    __shared__ double red[32];
    double sr = 0.0, si = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const double2 c = cconj_mul(a[k], b[k]);
        sr += c.x;
        si += c.y;
    }
    const double tr = block_sum(sr, red);
    const double ti = block_sum(si, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = tr;
        part[2 * blockIdx.x + 1] = ti;
    }
}

__global__ void __launch_bounds__(256) prob_kernel(const double2* __restrict__ psi, double* __restrict__ out, u64 dim) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x)
        out[k] = cnorm2(psi[k]);
}

// ---- K7: top-k by radix select on the composite key (probability desc, index asc) ---------------
// A probability is a non-negative double: its bit pattern orders like the value.  One pass
// histograms 16 bits of the key among the entries matching the prefix found so far.
__global__ void __launch_bounds__(256) radix_hist_kernel(const double2* __restrict__ psi, u64 dim, u64 prefix,
                                                         u64 prefix_mask, int shift, unsigned int* hist) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const u64 key = (u64)__double_as_longlong(cnorm2(psi[k]));
        if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffffull], 1u);
    }
}
// gather every entry with key STRICTLY above the threshold into out (unordered)
__global__ void __launch_bounds__(256) gather_gt_kernel(const double2* __restrict__ psi, u64 dim, u64 threshold,
                                                        u64 index_base, u64* out_idx, double* out_p,
                                                        unsigned int* counter, unsigned int cap) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const double pk = cnorm2(psi[k]);
        if ((u64)__double_as_longlong(pk) > threshold) {
            const unsigned int slot = atomicAdd(counter, 1u);
            if (slot < cap) {
                out_idx[slot] = index_base + k;
                out_p[slot] = pk;
            }
        }
    }
}
// ties[i] = number of entries of chunk c0 + i (2^cb amplitudes each) whose key equals the threshold
__global__ void __launch_bounds__(256) count_eq_kernel(const double2* __restrict__ psi, int cb, u64 c0, u64 n_chunks,
                                                       u64 threshold, unsigned int* __restrict__ ties) {
    __shared__ double red[32];
    const u64 len = 1ull << cb;
    for (u64 c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        double s = 0.0;  // exact for counts <= 2^16
        for (u64 k = threadIdx.x; k < len; k += blockDim.x)
            if ((u64)__double_as_longlong(cnorm2(psi[((c0 + c) << cb) + k])) == threshold) s += 1.0;
        s = block_sum(s, red);
        if (threadIdx.x == 0) ties[c] = (unsigned int)s;
    }
}
// one warp: the first `take` entries of one chunk whose key equals the threshold, in index order
__global__ void take_eq_kernel(const double2* __restrict__ psi, int cb, u64 chunk, u64 threshold, u64 index_base,
                               unsigned int take, u64* __restrict__ out_idx, double* __restrict__ out_p) {
    const int lane = threadIdx.x & 31;
    const u64 base = chunk << cb, len = 1ull << cb;
    unsigned int found = 0;
    for (u64 k0 = 0; k0 < len && found < take; k0 += 32) {
        const u64 k = k0 + lane;
        const double pk = (k < len) ? cnorm2(psi[base + k]) : -1.0;
        const bool hit = (k < len) && (u64)__double_as_longlong(pk) == threshold;
        const unsigned int m = __ballot_sync(0xffffffffu, hit);
        const unsigned int pos = found + __popc(m & ((1u << lane) - 1u));
        if (hit && pos < take) {
            out_idx[pos] = index_base + base + k;
            out_p[pos] = pk;
        }
        found += __popc(m);
    }
}

// ---- K7: sampling basis states from |psi|^2 ---------------------------------------------------------
// chunk_sums[c] = sum of |psi_k|^2 over chunk c (2^cb amplitudes); one block per chunk (grid-stride)
__global__ void __launch_bounds__(256) chunk_prob_kernel(const double2* __restrict__ psi, int cb, u64 n_chunks,
                                                         double* __restrict__ chunk_sums) {
    __shared__ double red[32];
    const u64 len = 1ull << cb;
    for (u64 c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        double s = 0.0;
        for (u64 k = threadIdx.x; k < len; k += blockDim.x) s += cnorm2(psi[(c << cb) + k]);
        s = block_sum(s, red);
        if (threadIdx.x == 0) chunk_sums[c] = s;
    }
}
// one warp per shot: walk the chunk until the running sum of |psi|^2 passes the residual
__global__ void __launch_bounds__(256) sample_resolve_kernel(const double2* __restrict__ psi, int cb, long long n,
                                                             const u64* __restrict__ chunk, const double* __restrict__ resid,
                                                             u64* __restrict__ out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const u64 base = chunk[w] << cb, len = 1ull << cb;
    const double r = resid[w];
    double run = 0.0;
    u64 found = ~0ull, last_nonzero = 0;
    for (u64 k0 = 0; k0 < len && found == ~0ull; k0 += 32) {
        const u64 k = k0 + lane;
        const double pk = (k < len) ? cnorm2(psi[base + k]) : 0.0;
        double inc = pk;  // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += up;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, pk > 0.0 && run + inc > r);
        const unsigned nz = __ballot_sync(0xffffffffu, pk > 0.0);
        if (nz) last_nonzero = k0 + (31 - __clz(nz));
        if (hit) found = k0 + (__ffs(hit) - 1);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (found == ~0ull) found = last_nonzero;  // rounding: the residual sat on the chunk's upper edge
    if (lane == 0) out[w] = base + found;
}

}  // namespace qsb
