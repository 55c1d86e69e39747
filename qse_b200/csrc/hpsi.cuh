// hpsi.cuh — matrix-free H psi for the Rydberg Hamiltonian (kernels K2/K3/K5 of SURVEY.md §2.2):
// tile arithmetic shared with hpsi_ws.cuh, plus the round-1 v1 kernels that are kept as independent
// cross-checks (qsb_set_option "hpsi_path" = 2 / 1) and for shards smaller than 2^13 amplitudes.
// The production kernel is hpsi_ws_kernel (hpsi_ws.cuh).
//
//   y[k] = (E[k] - sum_q delta_q b_q(k)) x[k] + sum_q hw_q x[k ^ (1 << q)]        (hw_q = Omega_q / 2)
//
// which is `get_hamiltonian(amp, det) * state` of the reference (qse/calc/qutip.py:43-44, 55-58):
// the CSR mat-vec QuTiP performs ~10-30 times per sesolve call -- with the per-site drive and
// detuning of README.md:93-95 (the global pulse is hw_q = Omega/2, delta_q = delta for every q).
// No matrix is ever formed.
//
// The sum over qubits is a hypercube stencil: amplitude k needs N partners at distances 2^q.
// A "tile" is 2^12 amplitudes held in shared memory; the bits of the basis index a tile spans
// are the only qubits whose partners are on chip, so one H psi is a short sequence of passes:
//
//   pass 0 (FIRST): tile = 2^12 CONTIGUOUS amplitudes (index bits 0..11).  Computes the diagonal
//           term and the 12 low-qubit flips, adds the sharded-qubit partners read straight from
//           peer GPUs' memory (NVLink), writes the partial y.       HBM: 16 + 8 + 16 = 40 B/amp
//   pass i>0      : tile = 2^lb contiguous amplitudes x 2^hb rows strided by 2^hb0 (lb + hb = 12):
//           index bits [hb0, hb0+hb) are the active ones, the low lb bits only buy coalescing.
//           Adds those flips to y.                                   HBM: 16 + 16 + 16 = 48 B/amp
//   LAST pass also applies the Taylor coefficient, accumulates into psi and reduces |term|^2,
//           so the stepper needs no separate axpy / norm kernels (fused K4).
//
// Tiles are filled by the TMA engine with 1-D bulk copies (cp.async.bulk -> UBLKCP) completing on
// an mbarrier, double-buffered in a persistent CTA (one per SM); each thread keeps 8 amplitudes
// whose indices differ in the top 3 tile bits, so 3 of the 12 flips are register-only and the
// other 9 are conflict-free 16-byte shared-memory reads at XOR-permuted addresses.
#pragma once
#include "common.cuh"

namespace qsb {

constexpr int kTileBits = 12;
constexpr int kTile = 1 << kTileBits;  // amplitudes per tile (64 KiB)
constexpr int kThreads = 512;
constexpr int kPerThread = kTile / kThreads;  // 8
constexpr int kRegBits = 3;                   // log2(kPerThread)
constexpr int kSmemBits = kTileBits - kRegBits;  // tile bits 0..8 go through shared memory
constexpr int kStages = 2;
constexpr int kMaxPeers = 3;
constexpr int kMaxLocalBits = 34;

// Per-site coefficients of one H(t), indexed by LOCAL index bit b (qubit n-1-b) / rank bit g.
struct SiteParams {
    double hw[kMaxLocalBits];   // Omega_q / 2 (+ static X coefficient) of the qubit at index bit b
    double dl[kMaxLocalBits];   // delta_q of the qubit at index bit b
    double d_rank;              // sum of delta_q over this rank's set sharded bits
    double hw_peer[kMaxPeers];  // Omega_q / 2 of the sharded qubit behind rank bit g
};
// what the LAST pass adds to the accumulator (Taylor epilogue) and what it reduces
enum { kAccNone = 0, kAccTerm = 1, kAccPair = 2 };  // acc += out  /  acc += x + out (two terms per RMW)
enum { kRedNorm = 0, kRedDot = 1 };                 // |out|^2  /  Re conj(x) out

// sum_q delta_q b_q(k) over the local index bits of k (generic form: one add per set bit)
QSB_DEV double site_detuning(const SiteParams& sp, u64 k) {
    double d = sp.d_rank;
    for (u64 rest = k; rest; rest &= rest - 1ull) d += sp.dl[__ffsll((long long)rest) - 1];
    return d;
}
constexpr size_t kTileSmemBytes = (size_t)kStages * kTile * sizeof(double2);

struct HpsiPass {
    const double2* x;      // input vector (never written by this pass unless x == acc, see stepper)
    double2* y;            // FIRST: written; else read-modify-write
    const double* eint;    // FIRST only: interaction diagonal of this shard
    double2* acc;          // LAST only, may be null: acc += out
    double* norm_part;     // LAST only, may be null: per-CTA sum of |out|^2
    const double2* peer[kMaxPeers];  // FIRST only: x of the partner rank for each sharded qubit
    int n_peer;
    SiteParams sp;
    double2 scale;  // LAST: out = scale * (H x)
    int acc_mode, red_mode;  // LAST: kAcc*, kRed*
    int nl;         // local qubits
    int lb;         // contiguous low tile bits
    int hb;         // high tile bits
    int hb0;        // position of the first high tile bit
    u64 n_tiles;
};

QSB_DEV u64 tile_base(const HpsiPass& p, u64 tau) {
    const int mid_bits = p.hb0 - p.lb;
    const u64 mid = tau & ((1ull << mid_bits) - 1ull);
    const u64 outer = tau >> mid_bits;
    return (outer << (p.hb0 + p.hb)) | (mid << p.lb);
}

QSB_DEV void tile_issue(const HpsiPass& p, double2* stage, uint64_t* bar, u64 tau) {
    // called by warp 0 only
    const int lane = threadIdx.x & 31;
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(kTile * sizeof(double2)));
    __syncwarp();
    const u64 base = tile_base(p, tau);
    if (p.hb == 0) {
        // one contiguous 64 KiB tile: 32 bulk copies of 2 KiB, one per lane
        constexpr int chunk = kTile / 32;
        bulk_g2s(stage + lane * chunk, p.x + base + (u64)lane * chunk, chunk * sizeof(double2), bar);
    } else {
        const int rows = 1 << p.hb;
        const uint32_t row_bytes = (uint32_t)(sizeof(double2) << p.lb);
        for (int r = lane; r < rows; r += 32)
            bulk_g2s(stage + ((size_t)r << p.lb), p.x + base + ((u64)r << p.hb0), row_bytes, bar);
    }
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kThreads, 1) hpsi_tile_kernel(const HpsiPass p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ double red[32];

    const int t = threadIdx.x;
    const int warp = t >> 5;
    const u64 first_tile = blockIdx.x;
    const u64 stride = gridDim.x;
    const u64 n_my = (p.n_tiles > first_tile) ? (p.n_tiles - first_tile + stride - 1) / stride : 0;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            if ((u64)s < n_my) tile_issue(p, tiles + (size_t)s * kTile, &full_bar[s], first_tile + s * stride);
    }

    const u64 low_mask = (1ull << p.lb) - 1ull;
    const int act_lo = FIRST ? 0 : p.lb;  // lowest ACTIVE tile bit of this pass
    double nrm = 0.0;

    for (u64 i = 0; i < n_my; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t parity = (uint32_t)((i / kStages) & 1ull);
        const u64 base = tile_base(p, first_tile + i * stride);
        const double2* ts = tiles + (size_t)s * kTile;

        // global addresses of this thread's 8 amplitudes; operand prefetch before the wait
        u64 g[kPerThread];
        double e[kPerThread];
        double2 yv[kPerThread];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const u64 idx = (u64)(t + j * kThreads);
            g[j] = base | ((idx >> p.lb) << p.hb0) | (idx & low_mask);
            if (FIRST) e[j] = __ldg(p.eint + g[j]);
            else yv[j] = p.y[g[j]];
        }

        mbar_wait(&full_bar[s], parity);

        double2 xv[kPerThread], sum[kPerThread];
        const int gshift = FIRST ? 0 : (p.hb0 - p.lb);  // tile bit b >= lb is index bit b + gshift
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) xv[j] = ts[t + j * kThreads];
        // flips of tile bits 9..11: partners are this thread's own registers
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            double2 a = make_double2(0.0, 0.0);
#pragma unroll
            for (int b = 0; b < kRegBits; ++b)
                if (kSmemBits + b >= act_lo) {
                    const double h = p.sp.hw[kSmemBits + b + gshift];
                    a.x = fma(h, xv[j ^ (1 << b)].x, a.x);
                    a.y = fma(h, xv[j ^ (1 << b)].y, a.y);
                }
            sum[j] = a;
        }
        double2 accv[kPerThread];
        if (LAST && p.acc_mode != kAccNone) {
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) accv[j] = p.acc[g[j]];
        }
        // flips of tile bits act_lo..8: XOR-permuted shared-memory reads
        for (int b = act_lo; b < kSmemBits; ++b) {
            const int tb = t ^ (1 << b);
            const double h = p.sp.hw[b + gshift];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const double2 q = ts[tb + j * kThreads];
                sum[j].x = fma(h, q.x, sum[j].x);
                sum[j].y = fma(h, q.y, sum[j].y);
            }
        }
        if (FIRST) {
            for (int q = 0; q < p.n_peer; ++q) {
                const double2* __restrict__ px = p.peer[q];
                const double h = p.sp.hw_peer[q];
#pragma unroll
                for (int j = 0; j < kPerThread; ++j) {
                    const double2 w = px[g[j]];
                    sum[j].x = fma(h, w.x, sum[j].x);
                    sum[j].y = fma(h, w.y, sum[j].y);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            double2 v;
            if (FIRST) {
                const double d = e[j] - site_detuning(p.sp, g[j]);
                v = make_double2(fma(d, xv[j].x, sum[j].x), fma(d, xv[j].y, sum[j].y));
            } else {
                v = cadd(yv[j], sum[j]);
            }
            if (LAST) {
                v = cmul(p.scale, v);
                nrm += (p.red_mode == kRedDot) ? (xv[j].x * v.x + xv[j].y * v.y) : cnorm2(v);
                if (p.acc_mode != kAccNone) {
                    double2 a = cadd(accv[j], v);
                    if (p.acc_mode == kAccPair) cacc(a, xv[j]);
                    p.acc[g[j]] = a;
                }
            }
            p.y[g[j]] = v;
        }

        __syncthreads();  // every thread is done reading stage s
        if (warp == 0 && i + kStages < n_my)
            tile_issue(p, tiles + (size_t)s * kTile, &full_bar[s], first_tile + (i + kStages) * stride);
    }

    if (LAST && p.norm_part != nullptr) {
        const double tot = block_sum(nrm, red);
        if (t == 0) p.norm_part[blockIdx.x] = tot;
    }
}

// Untiled kernel: one thread per amplitude, partners read through L1/L2.  Used when the shard
// is smaller than a tile (N_local < 12, launch-bound sizes) and as an independent cross-check
// of the tiled path in the tests (qsb_set_option "hpsi_path" = 1).
struct HpsiFlat {
    const double2* x;
    double2* y;
    const double* eint;
    double2* acc;
    double* norm_part;
    const double2* peer[kMaxPeers];
    int n_peer;
    SiteParams sp;
    double2 scale;
    int nl;
    int fuse;  // apply scale / acc / reduction
    int acc_mode, red_mode;
};

__global__ void __launch_bounds__(256) hpsi_flat_kernel(const HpsiFlat p) {
    __shared__ double red[32];
    const u64 dim = 1ull << p.nl;
    double nrm = 0.0;
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (u64)gridDim.x * blockDim.x) {
        const double2 xk = p.x[k];
        const double d = p.eint[k] - site_detuning(p.sp, k);
        double2 v = make_double2(d * xk.x, d * xk.y);
        for (int q = 0; q < p.nl; ++q) {
            const double2 w = p.x[k ^ (1ull << q)];
            v.x = fma(p.sp.hw[q], w.x, v.x);
            v.y = fma(p.sp.hw[q], w.y, v.y);
        }
        for (int q = 0; q < p.n_peer; ++q) {
            const double2 w = p.peer[q][k];
            v.x = fma(p.sp.hw_peer[q], w.x, v.x);
            v.y = fma(p.sp.hw_peer[q], w.y, v.y);
        }
        if (p.fuse) {
            v = cmul(p.scale, v);
            nrm += (p.red_mode == kRedDot) ? (xk.x * v.x + xk.y * v.y) : cnorm2(v);
            if (p.acc_mode != kAccNone) {
                double2 a = cadd(p.acc[k], v);
                if (p.acc_mode == kAccPair) cacc(a, xk);
                p.acc[k] = a;
            }
        }
        p.y[k] = v;
    }
    if (p.fuse && p.norm_part != nullptr) {
        const double tot = block_sum(nrm, red);
        if (threadIdx.x == 0) p.norm_part[blockIdx.x] = tot;
    }
}

// sum n partials (possibly from several part arrays of the same launch) into out[slot]
__global__ void __launch_bounds__(256) finalize_sum_kernel(const double* part, int n, double* out) {
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) *out = v;
}

// ---- pass planning (host) -------------------------------------------------------------------
struct PassGeom {
    int lb, hb, hb0;
};
struct HpsiPlan {
    bool tiled;
    int n_pass;
    PassGeom pass[8];
    double bytes_per_amp;
};

inline HpsiPlan make_plan(int nl, bool force_flat) {
    HpsiPlan plan{};
    if (nl < kTileBits || force_flat) {
        plan.tiled = false;
        plan.n_pass = 1;
        plan.bytes_per_amp = 40.0;
        return plan;
    }
    plan.tiled = true;
    plan.pass[0] = PassGeom{kTileBits, 0, kTileBits};
    plan.n_pass = 1;
    plan.bytes_per_amp = 40.0;
    int rem = nl - kTileBits;
    const int max_hb = kTileBits - 3;  // keep >= 128-byte rows
    const int extra = (rem + max_hb - 1) / max_hb;
    int pos = kTileBits;
    for (int i = 0; i < extra; ++i) {
        const int hb = (rem + (extra - i) - 1) / (extra - i);  // balanced split
        plan.pass[plan.n_pass++] = PassGeom{kTileBits - hb, hb, pos};
        plan.bytes_per_amp += 48.0;
        pos += hb;
        rem -= hb;
    }
    return plan;
}

}  // namespace qsb
