// hpsi_ws.cuh — warp-specialised H psi with L2-resident super-blocks (round-1 v2 of K2/K3).
//
// Same tile arithmetic as hpsi.cuh, different orchestration:
//   * one producer warp per CTA takes work items from a global ticket counter, fills a 3-stage
//     ring of 64 KiB tiles with TMA bulk copies (cp.async.bulk -> UBLKCP) and publishes each stage
//     on a "full" mbarrier; 16 consumer warps process tiles and release stages through "empty"
//     mbarriers.  No __syncthreads in the steady state: consumer warps drift apart, so the
//     LDS-heavy flip phase of one warp overlaps the global loads/stores of another.
//   * a FUSED launch covers the low `sbits` index bits in two sub-passes that share one trip to
//     HBM: a super-block of 2^sbits amplitudes (x 16 B + y 16 B + E 8 B per amplitude) fits in the
//     126 MB L2, so sub-pass B (index bits 12..sbits-1, strided rows) re-reads x and the partial
//     y that sub-pass A (bits 0..11 + diagonal) just produced from L2.  Work items are ordered
//     A(0) A(1) B(0) A(2) B(1) ... so that B(s) is handed out one whole phase after A(s); a B item
//     still checks a per-super-block completion counter (release/acquire through L2) before its y
//     is read.  A CTA only ever waits for tickets that were taken earlier by running CTAs, so the
//     scheme cannot deadlock whatever the residency.
//   * the remaining high bits (>= sbits) are handled by PLAIN launches of the same kernel
//     (strided-row tiles, no dependencies), the last of which applies the fused Taylor epilogue.
// HBM traffic per amplitude: 40 B for the fused launch + 48 B per plain launch (25 qubits: 88 B
// instead of 136 B with one launch per tile-bit group).
#pragma once
#include "common.cuh"
#include "hpsi.cuh"

namespace qsb {

constexpr int kWsStages = 3;
constexpr int kConsumerWarps = 16;
// warps 0..15 consume, warp 16 produces, warps 17..19 only pad the producer to a full warpgroup:
// registers are allocated per warpgroup, and setmaxnreg moves the producer group's share
// (96 -> 24 per thread) to the consumers (96 -> 112).
constexpr int kWsThreads = (kConsumerWarps + 4) * 32;  // 640
constexpr size_t kWsSmemBytes = (size_t)kWsStages * kTile * sizeof(double2);

enum { kItemFirst = 0, kItemInner = 1, kItemPlain = 2, kItemDone = 3 };

struct WsItem {
    u64 base;
    int type;
    int sb;
};

struct HpsiWs {
    const double2* x;
    double2* y;
    const double* eint;
    double2* acc;
    double* norm_part;
    const double2* peer[kMaxPeers];
    int n_peer;
    double half_omega, delta;
    double2 scale;
    int nl;
    int pop_rank;
    int fused;       // 1: fused A+B launch over super-blocks; 0: plain launch
    int sbits;       // fused: log2 of the super-block
    int lb, hb, hb0; // geometry of the strided tiles (inner sub-pass or plain pass)
    int last;        // the strided tiles of this launch finish H psi (apply scale / acc / norm)
    unsigned int* ticket;
    unsigned long long* done;   // per super-block: A tiles completed (cumulative over launches)
    unsigned long long done_target;
    u64 n_tiles;     // tiles per sub-pass (= dim / 4096)
};

QSB_DEV u64 ws_tile_base(int lb, int hb, int hb0, u64 tau) {
    const int mid_bits = hb0 - lb;
    const u64 mid = tau & ((1ull << mid_bits) - 1ull);
    const u64 outer = tau >> mid_bits;
    return (outer << (hb0 + hb)) | (mid << lb);
}

// FIRST = contiguous tile (bits 0..11 + diagonal + peers); otherwise strided rows (bits >= lb)
// Global operands (E_int or partial y, and the Taylor accumulator) are requested up front with
// loads the compiler may not sink, so their latency hides behind the shared-memory flip phase.
// PEER (FIRST only): sharded-qubit partners are read over NVLink; x is then re-read from the tile
// at the end instead of being kept live, to make room for the partner amplitudes in flight.
template <bool FIRST, bool PEER>
QSB_DEV void ws_process(const HpsiWs& p, const double2* __restrict__ ts, u64 base, bool last, int t, double& nrm) {
    const int lb = FIRST ? kTileBits : p.lb;
    const int hb0 = FIRST ? kTileBits : p.hb0;
    const u64 low_mask = (1ull << lb) - 1ull;
    const int act_lo = FIRST ? 0 : lb;
    const bool with_acc = last && p.acc != nullptr;
#define QSB_GADDR(j) (base | ((((u64)(t + (j)*512)) >> lb) << hb0) | (((u64)(t + (j)*512)) & low_mask))

    double e[kPerThread];
    double2 yv[kPerThread];  // non-FIRST: partial y; FIRST+PEER: first partner's amplitudes; FIRST: x
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        if (FIRST) {
            e[j] = ldnc1(p.eint + QSB_GADDR(j));
            if (PEER) yv[j] = ldg2_volatile(p.peer[0] + QSB_GADDR(j));  // over NVLink
        } else {
            yv[j] = ldcg2(p.y + QSB_GADDR(j));
        }
    }
    double2 sum[kPerThread];
    {
        double2 xv[kPerThread];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) xv[j] = ts[t + j * 512];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            double2 a = make_double2(0.0, 0.0);
#pragma unroll
            for (int b = 0; b < kRegBits; ++b)
                if (kSmemBits + b >= act_lo) cacc(a, xv[j ^ (1 << b)]);
            sum[j] = a;
        }
        if (FIRST && !PEER) {
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) yv[j] = xv[j];
        }
    }
    double2 accv[kPerThread];
    if (with_acc) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) accv[j] = ldg2_volatile(p.acc + QSB_GADDR(j));
    }
    if (FIRST) {
#pragma unroll
        for (int b = 0; b < kSmemBits; ++b) {
            const int tb = t ^ (1 << b);
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cacc(sum[j], ts[tb + j * 512]);
        }
        if (PEER) {
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cacc(sum[j], yv[j]);
        }
        for (int q = 1; PEER && q < p.n_peer; ++q) {
            const double2* __restrict__ px = p.peer[q];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) yv[j] = ldg2_volatile(px + QSB_GADDR(j));
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cacc(sum[j], yv[j]);
        }
    } else {
        for (int b = act_lo; b < kSmemBits; ++b) {
            const int tb = t ^ (1 << b);
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cacc(sum[j], ts[tb + j * 512]);
        }
    }
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
        const u64 g = QSB_GADDR(j);
        double2 v;
        if (FIRST) {
            const double d = e[j] - p.delta * (double)(__popcll(g) + p.pop_rank);
            const double2 xj = PEER ? ts[t + j * 512] : yv[j];
            v = make_double2(d * xj.x + p.half_omega * sum[j].x, d * xj.y + p.half_omega * sum[j].y);
        } else {
            v = make_double2(yv[j].x + p.half_omega * sum[j].x, yv[j].y + p.half_omega * sum[j].y);
        }
        if (last) {
            v = cmul(p.scale, v);
            nrm += cnorm2(v);
            if (with_acc) p.acc[g] = cadd(accv[j], v);
        }
        p.y[g] = v;
    }
#undef QSB_GADDR
}

__global__ void __launch_bounds__(kWsThreads, 1) hpsi_ws_kernel(const HpsiWs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kWsStages];
    __shared__ __align__(8) uint64_t empty_bar[kWsStages];
    __shared__ WsItem desc[kWsStages];
    __shared__ int warps_done[kWsStages];
    __shared__ double red[32];

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < kWsStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
            warps_done[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();

    double nrm = 0.0;
    if (warp >= kConsumerWarps) {
        // ================= producer warpgroup (only its first warp works) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == kConsumerWarps) {
        const u64 n_a = p.fused ? (1ull << (p.sbits - kTileBits)) : 0;  // tiles per super-block sub-pass
        const u64 n_sb = p.fused ? (p.n_tiles / n_a) : 0;
        const u64 total = p.fused ? 2ull * p.n_tiles : p.n_tiles;
        for (u64 it = 0;; ++it) {
            const int s = (int)(it % kWsStages);
            const u64 use = it / kWsStages;
            if (use > 0) mbar_wait(&empty_bar[s], (uint32_t)((use - 1) & 1ull));
            unsigned int ticket = 0;
            if (lane == 0) ticket = atomicAdd(p.ticket, 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            if ((u64)ticket >= total) {
                if (lane == 0) {
                    desc[s].type = kItemDone;
                    mbar_arrive(&full_bar[s]);
                }
                break;
            }
            int type, sb = 0;
            u64 tau;
            if (p.fused) {
                const u64 ph = ticket / n_a, w = ticket % n_a;
                if (ph == 0) { type = kItemFirst; sb = 0; }
                else if (ph == 2 * n_sb - 1) { type = kItemInner; sb = (int)(n_sb - 1); }
                else if (ph & 1ull) { type = kItemFirst; sb = (int)((ph + 1) / 2); }
                else { type = kItemInner; sb = (int)(ph / 2 - 1); }
                tau = (u64)sb * n_a + w;
            } else {
                type = kItemPlain;
                tau = ticket;
            }
            double2* stage = tiles + (size_t)s * kTile;
            u64 base;
            if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(kTile * sizeof(double2)));
            __syncwarp();
            if (type == kItemFirst) {
                base = tau << kTileBits;
                constexpr int chunk = kTile / 32;
                bulk_g2s(stage + lane * chunk, p.x + base + (u64)lane * chunk, chunk * sizeof(double2), &full_bar[s]);
            } else {
                base = ws_tile_base(p.lb, p.hb, p.hb0, tau);
                const int rows = 1 << p.hb;
                const uint32_t row_bytes = (uint32_t)(sizeof(double2) << p.lb);
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s(stage + ((size_t)r << p.lb), p.x + base + ((u64)r << p.hb0), row_bytes, &full_bar[s]);
            }
            if (lane == 0) {
                if (type == kItemInner) {
                    // the partial y of this super-block must be complete (all its A tiles stored)
                    while (ld_acquire_u64(p.done + sb) < p.done_target) __nanosleep(32);
                    __threadfence();
                }
                desc[s].base = base;
                desc[s].type = type;
                desc[s].sb = sb;
                mbar_arrive(&full_bar[s]);  // release: descriptor + dependency visible to consumers
            }
            __syncwarp();
        }
        }
    } else {
        // ================= consumer warps =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        for (u64 it = 0;; ++it) {
            const int s = (int)(it % kWsStages);
            const u64 use = it / kWsStages;
            mbar_wait(&full_bar[s], (uint32_t)(use & 1ull));
            const int type = desc[s].type;
            if (type == kItemDone) break;
            const u64 base = desc[s].base;
            const int sb = desc[s].sb;
            const double2* ts = tiles + (size_t)s * kTile;
            if (type == kItemFirst) {
                if (p.n_peer > 0) ws_process<true, true>(p, ts, base, false, t, nrm);
                else ws_process<true, false>(p, ts, base, false, t, nrm);
                // publish this tile of partial y: warp barrier -> cta-scope release on the shared
                // counter; the 16th warp to finish issues ONE gpu-scope fence (cumulative over the
                // writes it observed through the counter) and bumps the super-block counter.
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const int c = atomicAdd(&warps_done[s], 1);
                    if (c == kConsumerWarps - 1) {
                        warps_done[s] = 0;
                        __threadfence();
                        atomicAdd(p.done + sb, 1ull);
                    }
                }
            } else {
                ws_process<false, false>(p, ts, base, p.last != 0, t, nrm);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }
    // per-CTA norm partial (only meaningful when p.last)
    nrm = warp_sum(nrm);
    if (lane == 0) red[warp] = (warp < kConsumerWarps) ? nrm : 0.0;
    __syncthreads();
    if (t == 0 && p.norm_part != nullptr) {
        double tot = 0.0;
        for (int w = 0; w < kConsumerWarps; ++w) tot += red[w];
        p.norm_part[blockIdx.x] = tot;
    }
}

}  // namespace qsb
