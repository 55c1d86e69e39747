// hpsi_ws.cuh — warp-specialised H psi with L2-resident super-blocks: the production K2/K3 kernel.
//
//   y[k] = (E[k] - sum_q delta_q b_q(k)) x[k] + sum_q hw_q x[k ^ (1 << q)]          (hw_q = Omega_q / 2)
//
// i.e. `get_hamiltonian(amp, det) * state` of the reference (qse/calc/qutip.py:43-44, 55-58) with the
// per-site drive and detuning of the reference's headline Hamiltonian (README.md:93-95); the global
// pulse of qse.calc.Qutip is the special case hw_q = Omega/2, delta_q = delta for every site.
//
// Same tile arithmetic as hpsi.cuh, different orchestration:
//   * one producer warp per CTA takes work items from a global ticket counter, fills a 3-stage
//     ring of 64 KiB tiles through the TMA engine -- 1-D bulk copies (cp.async.bulk -> UBLKCP) for
//     contiguous tiles, ONE 2-D tensor-map copy per <= 256 rows (cp.async.bulk.tensor.2d ->
//     UTMALDG) for strided tiles -- issues L2 prefetches for the operands the consumers read with
//     plain loads, and publishes each stage on a "full" mbarrier; 16 consumer warps process tiles
//     and release stages through "empty" mbarriers.  No __syncthreads in the steady state.
//   * sharded state (template parameter SH): behind a tile the same producer pulls the SAME amplitudes of the
//     partners' x (or, from 4 ranks up, of the remap buffers) over NVLink into ordinary stages (kItemPeer:
//     bulk copies for contiguous tiles, tensor-map boxes on the peer pointers for strided ones) and the
//     consumers fold them into the accumulators from shared memory after the flip phase.  Which tiles pull
//     in which launch is the host's choice (peer_sel): the NVLink traffic runs beside ALL local work.
//   * diagonal on the fly (template parameter OTF): no E array is read; see HpsiWs::otf.
//   * a FUSED launch covers the low `sbits` index bits in two sub-passes that share one trip to
//     HBM: a super-block of 2^sbits amplitudes (x 16 B + y 16 B + E 8 B per amplitude) fits in the
//     126 MB L2, so sub-pass B (index bits 12..sbits-1, strided rows) re-reads x and the partial
//     y that sub-pass A (bits 0..11 + diagonal) just produced from L2.  Work items are ordered
//     A(0) .. A(L-1) A(L) B(0) A(L+1) B(1) ... so that B(s) is handed out L blocks after A(s); a B item
//     still checks a per-super-block completion counter (release/acquire through L2) before its y
//     is read.  A CTA only ever waits for tickets that were taken earlier by running CTAs, so the
//     scheme cannot deadlock whatever the residency.
//   * the remaining high bits (>= sbits) are handled by PLAIN launches of the same kernel
//     (strided-row tiles, no dependencies), the last of which applies the fused Taylor epilogue
//     (scale, accumulate -- one term or a PAIR of terms per read-modify-write of psi --, |term|^2 or
//     Re<x|Hx>), and whose last CTA reduces the per-CTA partials in a fixed order: no finalize
//     kernel, no memset, no host round trip per term.
// HBM traffic per amplitude: 40 B for the fused launch + 48 B per plain launch (25 qubits: 88 B
// instead of 136 B with one launch per tile-bit group).
//
// Arithmetic: every flip is ONE fused multiply-add per component, v = fma(hw_q, x[k ^ m_q], v), the
// coefficient coming straight from the constant bank -- a per-site drive costs exactly what the
// global one does.  The accumulators (8 amplitudes = 32 registers) are the only state a thread
// carries through the shared-memory flip phase, which leaves room for eight 16-byte loads in
// flight per thread (round 1 carried sum + x + y = 96 registers and spilled y).  The kernel is bound by
// the L1 / shared-memory data pipe (ncu: 74 % busy), so anything that adds traffic to that pipe --
// spills, a second copy of y, more static shared memory (a smaller L1 carve-out) -- costs time
// one for one: see the tuning log in DESIGN.md before changing the consumer loop.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "hpsi.cuh"

namespace qsb {

constexpr int kConsumerWarps = 16;
// warps 0..15 consume, warp 16 produces, warps 17..19 only pad the producer to a full warpgroup:
// registers are allocated per warpgroup, and setmaxnreg moves the producer group's share
// (96 -> 32 per thread) to the consumers (96 -> 112).
constexpr int kWsThreads = (kConsumerWarps + 4) * 32;  // 640
// setmaxnreg.inc BLOCKS until the CTA's register pool (kWsRegsLaunch per thread at launch) holds what it
// asks for: the producer group must release at least what the consumers take, or the kernel deadlocks.
constexpr int kWsRegsLaunch = 96, kWsRegsProducer = 32, kWsRegsConsumer = 112;
static_assert(4 * (kWsRegsLaunch - kWsRegsProducer) >= kConsumerWarps * (kWsRegsConsumer - kWsRegsLaunch), "setmaxnreg budget");
#define QSB_STR2(x) #x
#define QSB_STR(x) QSB_STR2(x)

constexpr size_t kWsSmemBytes = 192 * 1024;            // the tile ring

// Tile geometry.  A tile holds 2^TB amplitudes; a consumer thread owns 8 of them, so a tile is
// processed by a GROUP of 2^(TB-3) threads and the 16 consumer warps form 16*32 / 2^(TB-3)
// independent groups (group g takes work items g, g + kGroups, ...).  TB = 12: one group, 3 stages
// of 64 KiB.  TB = 11: two groups, 6 stages of 32 KiB.
template <int TB>
struct WsCfg {
    static constexpr int kTileBitsW = TB;
    static constexpr int kTileW = 1 << TB;
    static constexpr int kGroupThreads = kTileW / kPerThread;
    static constexpr int kGroupWarps = kGroupThreads / 32;
    static constexpr int kGroups = kConsumerWarps / kGroupWarps;
    static constexpr int kRegLo = TB - kRegBits;  // tile bits [kRegLo, TB) are register-only flips
    static constexpr int kStages = (int)(kWsSmemBytes / (kTileW * sizeof(double2)));
    static_assert(kGroups >= 1 && kStages >= 2 && kStages <= 12, "tile size");
};
constexpr int kMaxStages = 12;

enum { kItemFirst = 0, kItemInner = 1, kItemPlain = 2, kItemDone = 3, kItemPeer = 4 };

struct WsItem {
    // first 128 bytes: the tile's row of the static diagonal table, copied in by the TMA engine (kItemFirst):
    //   w[b]    = sum over the set index bits h >= TB of the tile base of V(h, b), for tile bit b < TB
    //   w[12]   = sum over pairs h < h' of set bits of the base of V(h, h')
    double w[16];
    u64 base;
    double dhi;    // FIRST: c0 + sum of lin[h] over the set bits of the base (the time-dependent single-bit part)
    int type;
    int sb;        // super-block of the tile; kItemPeer: which sharded qubit's coefficient applies
    int n_follow;  // sharded state: partner stages (kItemPeer) that follow this tile in the ring
    int pad;
};
constexpr int kTileTabDoubles = 16;  // doubles per tile in the static diagonal table (128 bytes)

// Static part of the on-the-fly diagonal per contiguous tile (see WsItem::w); built once per interaction.
__global__ void __launch_bounds__(256) tile_table_kernel(double* __restrict__ tab, const double* __restrict__ vbits, int nl, int tb) {
    const u64 n_tiles = 1ull << (nl - tb);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_tiles * kTileTabDoubles; i += (u64)gridDim.x * blockDim.x) {
        const u64 tau = i / kTileTabDoubles;
        const int col = (int)(i % kTileTabDoubles);
        double acc = 0.0;
        if (col < tb) {
            for (u64 rest = tau; rest; rest &= rest - 1ull) acc += vbits[(tb + __ffsll((long long)rest) - 1) * kMaxLocalBits + col];
        } else if (col == 12) {
            for (u64 r1 = tau; r1; r1 &= r1 - 1ull) {
                const int h = tb + __ffsll((long long)r1) - 1;
                for (u64 r2 = r1 & (r1 - 1ull); r2; r2 &= r2 - 1ull) acc += vbits[h * kMaxLocalBits + tb + __ffsll((long long)r2) - 1];
            }
        }
        tab[i] = acc;
    }
}

// ring state a consumer carries into ws_process when partner stages follow the tile it is working on
struct WsRing {
    double2* tiles;
    uint64_t* full_bar;
    uint64_t* empty_bar;
    const WsItem* desc;
};
struct PeerMaps {
    CUtensorMap m[8];  // strided (plain-launch) tiles of the partners' x (direct exchange) or Z chunks (remap)
};

struct HpsiWs {
    const double2* x;
    double2* y;
    const double* eint;
    double2* acc;
    double* norm_part;
    const double2* peer[kMaxPeers];
    int n_peer;      // partner stages per tile that takes part in the exchange in THIS launch (0: none)
    // Which tiles pull their partners in this launch (the NVLink traffic of one H psi is spread over the fused
    // and the last plain launch): 0 = every tile, 1 = tiles with index bit peer_bit CLEAR, 2 = tiles with it SET.
    // peer_bit lies inside the super-block index above the contiguous tile, so it is constant over a contiguous
    // tile and over a strided tile of the plain launch alike.
    int peer_sel;
    int peer_bit;
    SiteParams sp;   // per-site Omega_q / 2 and delta_q by index bit (hw_peer = 1.0 with the remap exchange)
    // Diagonal on the fly (otf = 1): the diagonal of H is 2-local,
    //     d(k) = c0 + sum_b lin[b] b_b(k) + sum_{b<b'} vbits[b][b'] b_b(k) b_b'(k)       (LOCAL index bits b),
    // with the rank's bits, the detuning and any static 1- / 2-local diagonal strings folded into c0 and lin
    // by the host.  A thread's 8 amplitudes keep their tile-index part in 8 registers for the whole launch;
    // the producer adds the tile-base part per tile: no E array is read (-8 B/amplitude of HBM traffic, no
    // global load in front of the flip phase).  otf = 0 reads E from eint (3+-local static strings).
    int otf;
    const double* vbits;          // [kMaxLocalBits][kMaxLocalBits], symmetric, zero diagonal (device)
    const double* tile_tab;       // [n_tiles][kTileTabDoubles]: static per-tile part (tile_table_kernel)
    u64 tab_stride;               // kTileTabDoubles, or 0 when every tile shares one all-zero row (stored-E mode)
    double lin[kMaxLocalBits];
    double c0;
    double2 scale;
    int nl;
    int fused;       // 1: fused FIRST+INNER launch over super-blocks; 0: plain launch
    int sbits;       // fused: log2 of the super-block
    int lb, hb, hb0; // geometry of the strided tiles (inner sub-pass or plain pass): lb + hb = TB
    int last;        // the strided tiles of this launch finish H psi (apply scale / acc / reduction)
    int acc_mode;    // kAccNone / kAccTerm (acc += out) / kAccPair (acc += x + out)
    int red_mode;    // kRedNorm / kRedDot
    int lag;         // fused: lead of the FIRST tiles over the INNER tiles, in super-blocks
    unsigned int* ticket;
    unsigned int ticket_base;   // the counter is never reset: tickets of this launch start here
    unsigned int* tail;         // last launch: CTAs that have stored their partial (reset by the last one)
    double* red_out;            // last launch: sum of the per-CTA partials, fixed order
    unsigned long long* done;   // per super-block: FIRST tiles completed (cumulative over launches)
    unsigned long long done_target;
    u64 n_tiles;     // tiles per sub-pass (= dim >> TB)
    int use_tmap;    // strided tiles are fetched with ONE 2-D tensor-map TMA per <= 256 rows
    u64 zero;        // always 0, but only the host knows: (bits & zero) makes an address DEPEND on a value, which
                     // pins a load after the phase that produced the value whatever ptxas would like to hoist
    // remap exchange (P >= 4): zpeer[c] is rank c's Z buffer (remap_z_kernel); the sharded-qubit term of
    // tile tau lives in chunk `rank` of the Z buffer of rank c = chunk index of tau -- ONE pull per tile
    int remap;
    int rank;
    int cbits;       // log2 of the chunk (= shard / P)
    int pbits;       // log2 P
    unsigned int sb_xor;  // remap: super-block order staggered per rank (rank r starts in chunk r) so that at any
                     // moment the ranks pull from DIFFERENT sources instead of all hammering one GPU's links
    const double2* zpeer[8];
};

// Qubit-remap exchange, step 1.  With P = 2^p ranks, view every shard as P chunks
// (chunk index = top p LOCAL index bits).  This rank owns chunk column `rank`: it pulls chunk
// `rank` of every rank's x ONCE (7/8 of a shard inbound at P = 8, instead of 3 full shards) and
// forms, in registers, the hypercube-neighbour sums over the RANK index for all P of them:
//     z[c][w] = sum_g hw_g x_{c ^ 2^g}[rank][w]
// which is exactly the sharded-qubit term that rank c needs for its chunk `rank`; rank c pulls it
// back inside the fused H psi launch (kItemPeer).  Inbound per amplitude: 2 * 16 * (P-1)/P bytes.
struct RemapArgs {
    const double2* xs[8];
    double2* z;
    double hw[kMaxPeers];  // Omega/2 of the sharded qubit behind rank bit g
    int rank;
    int cbits;
};
template <int P>
__global__ void __launch_bounds__(1024, 1) remap_z_kernel(const RemapArgs a) {
    // U amplitudes per thread and trip: 8 peer loads of 16 bytes in flight per thread (128 KiB per SM), which is
    // what it takes to keep an NVLink at ~3 us of latency busy from the few SMs this kernel is given
    constexpr int U = 8 / P;
    const u64 clen = 1ull << a.cbits;
    const u64 src0 = (u64)a.rank << a.cbits;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 w0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; w0 < clen; w0 += stride * U) {
        double2 v[U][P];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u64 w = w0 + (u64)u * stride;
#pragma unroll
            for (int c = 0; c < P; ++c) v[u][c] = (w < clen) ? ldg2_volatile(a.xs[c] + src0 + w) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const u64 w = w0 + (u64)u * stride;
            if (w >= clen) continue;
#pragma unroll
            for (int c = 0; c < P; ++c) {
                double2 s = make_double2(0.0, 0.0);
                int gi = 0;
#pragma unroll
                for (int g = 1; g < P; g <<= 1, ++gi) {
                    s.x = fma(a.hw[gi], v[u][c ^ g].x, s.x);
                    s.y = fma(a.hw[gi], v[u][c ^ g].y, s.y);
                }
                a.z[((u64)c << a.cbits) + w] = s;
            }
        }
    }
}

// Static schedule of the fused launch: ticket -> (item type, tile index).  Tickets enumerate phases
// of n_a tiles:  A(0) .. A(L-1),  then  A(s), B(s-L)  for s = L .. n_sb-1,  then  B(n_sb-L) .. B(n_sb-1)
// (A = FIRST tiles of super-block s, B = its INNER tiles, L = lag).  sb_xor permutes the super-block
// order (remap exchange: every rank starts in a different chunk).  Host-callable so that the CPU
// tests can check the invariants the kernel relies on (tests/test_host_logic.py).
__host__ __device__ inline int ws_decode_ticket(u64 ticket, u64 n_a, u64 n_sb, u64 lag, u64 sb_xor, u64* tau) {
    const u64 ph = ticket / n_a, wi = ticket % n_a;
    int type;
    u64 sb;
    if (ph < lag) {
        type = 0;  // kItemFirst
        sb = ph;
    } else if (ph < 2 * n_sb - lag) {
        const u64 q = ph - lag;
        if (q & 1ull) {
            type = 1;  // kItemInner
            sb = q >> 1;
        } else {
            type = 0;
            sb = lag + (q >> 1);
        }
    } else {
        type = 1;
        sb = n_sb - lag + (ph - (2 * n_sb - lag));
    }
    sb ^= sb_xor;
    *tau = sb * n_a + wi;
    return type;
}

QSB_DEV u64 ws_tile_base(int lb, int hb, int hb0, u64 tau) {
    const int mid_bits = hb0 - lb;
    const u64 mid = tau & ((1ull << mid_bits) - 1ull);
    const u64 outer = tau >> mid_bits;
    return (outer << (hb0 + hb)) | (mid << lb);
}

QSB_DEV void cfma(double2& v, double c, double2 x) {
    v.x = fma(c, x.x, v.x);
    v.y = fma(c, x.y, v.y);
}

// FIRST = contiguous tile (index bits 0..TB-1 + diagonal + peers); otherwise strided rows (bits >= lb).
// A thread owns tile indices idx = t + j * kGroupThreads (j = 0..7): tile bits [TB-3, TB) are
// register-only flips, every other active tile bit is a conflict-free LDS.128 at an XOR-permuted
// address.  Global operands (E or partial y, the first NVLink partner) are requested up front with
// loads the compiler may not sink, so their latency hides behind the shared-memory flip phase.
// PEER (FIRST only, sharded state): the producer has already pulled the partner tiles over NVLink
// (kItemPeer stages, ws_peer below) and y holds sum_g hw_g * partner_g.
// dsum: FIRST only -- sum of delta_q over the set bits of this thread's index t (bits 0..TB-4) plus
// the tile's high part (WsItem::dhi).
template <int TB, bool FIRST, bool LAST, bool SH, bool OTF>
QSB_DEV void ws_process(const HpsiWs& p, const double2* __restrict__ ts, u64 base, int t, double dsum,
                        const double (&diag)[kPerThread], double& red, uint64_t* empty, const WsRing& ring, int n_follow,
                        u64& it) {
    using C = WsCfg<TB>;
    const int lb = FIRST ? TB : p.lb;
    const int hb0 = FIRST ? TB : p.hb0;
    const u64 low_mask = (1ull << lb) - 1ull;
    const int act_lo = FIRST ? 0 : lb;
    const int gshift = FIRST ? 0 : (hb0 - lb);  // tile bit b >= lb is index bit b + gshift
    // Global index of amplitude j of this thread: the tile -> index map is a bit permutation, so
    // g(t + j * 512) = (base | g(t)) + g(j * 512): ONE per-thread 64-bit base and eight warp-uniform
    // offsets (compile-time constants for a contiguous tile) instead of eight live 64-bit addresses.
    const u64 g0 = base | (FIRST ? (u64)t : (((((u64)t) >> lb) << hb0) | (((u64)t) & low_mask)));
#define QSB_IDX(j) (t + (j)*C::kGroupThreads)
#define QSB_GOFF(j) \
    (FIRST ? (u64)((j)*C::kGroupThreads) : (((((u64)((j)*C::kGroupThreads)) >> lb) << hb0) | (((u64)((j)*C::kGroupThreads)) & low_mask)))
#define QSB_GADDR(j) (g0 + QSB_GOFF(j))

    double e[FIRST ? kPerThread : 1];
    double2 v[kPerThread];
    if constexpr (FIRST && !OTF) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) e[j] = ldnc1(p.eint + QSB_GADDR(j));
    }
    // The accumulators START as the partial y (non-FIRST; L2-resident: written by the FIRST tiles of this
    // super-block, or prefetched by the producer) or as the partners' term (FIRST + PEER): the loads
    // land directly in v, there is no second copy to keep (or to spill) through the flip phases.
    if constexpr (!FIRST) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) v[j] = ldcg2(p.y + QSB_GADDR(j));
    } else {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) v[j] = make_double2(0.0, 0.0);
    }
    {
        double2 xv[kPerThread];
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) xv[j] = ts[QSB_IDX(j)];
        // flips of the top three tile bits: partners are this thread's own registers
        double hr[kRegBits];
#pragma unroll
        for (int b = 0; b < kRegBits; ++b) hr[b] = (C::kRegLo + b >= act_lo) ? p.sp.hw[C::kRegLo + b + gshift] : 0.0;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            // the three partner products first, y / zero last: the chain does not wait on the global load
            double2 a = make_double2(hr[0] * xv[j ^ 1].x, hr[0] * xv[j ^ 1].y);
            cfma(a, hr[1], xv[j ^ 2]);
            cfma(a, hr[2], xv[j ^ 4]);
            if constexpr (FIRST && OTF) {
                cfma(a, diag[j], xv[j]);  // the 2-local diagonal, evaluated on the fly by the caller
            } else if constexpr (FIRST) {
                // diagonal: d = E - sum_q delta_q b_q(k); bits TB-3..TB-1 of the index are j's bits
                const double dj = ((j & 1) ? p.sp.dl[C::kRegLo] : 0.0) + ((j & 2) ? p.sp.dl[C::kRegLo + 1] : 0.0) +
                                  ((j & 4) ? p.sp.dl[C::kRegLo + 2] : 0.0);
                cfma(a, e[j] - (dsum + dj), xv[j]);
            }
            if constexpr (!FIRST) cacc(a, v[j]);
            v[j] = a;
        }
    }
    if constexpr (FIRST) {
#pragma unroll
        for (int b = 0; b < C::kRegLo; ++b) {
            const int tb = t ^ (1 << b);
            const double h = p.sp.hw[b];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cfma(v[j], h, ts[tb + j * C::kGroupThreads]);
        }
    } else {
        for (int b = act_lo; b < C::kRegLo; ++b) {
            const int tb = t ^ (1 << b);
            const double h = p.sp.hw[b + gshift];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) cfma(v[j], h, ts[tb + j * C::kGroupThreads]);
        }
    }
    // SH (sharded state) only: the single-GPU instantiation carries none of this through the flip phase
    if constexpr (SH && !LAST) {
        // the tile has been read for the last time: hand the stage back to the producer before the stores
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty);
    }
    // partner stages: v += hw_g * partner tile, same tile geometry, straight from shared memory
    for (int f = 0; SH && f < n_follow; ++f) {
        it += C::kGroups;
        const int s2 = (int)(it % C::kStages);
        mbar_wait(&ring.full_bar[s2], (uint32_t)((it / C::kStages) & 1ull));
        const double h = p.sp.hw_peer[ring.desc[s2].sb];
        const double2* zs = ring.tiles + (size_t)s2 * C::kTileW;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) cfma(v[j], h, zs[QSB_IDX(j)]);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&ring.empty_bar[s2]);
    }
    // output addresses are formed only now (same trick): no 64-bit address lives through the flip phase
    const u64 ge = g0 + ((u64)__double_as_longlong(v[0].x) & p.zero);
    if constexpr (!LAST) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) p.y[ge + QSB_GOFF(j)] = v[j];
    } else {
        // Taylor epilogue, in two batches of four amplitudes so that few addresses and accumulator
        // values are in flight (the second batch's addresses depend on the first batch's results).
        // The accumulator was prefetched into L2 by the producer: it is loaded only now.
        const bool with_acc = p.acc_mode != kAccNone;
        const double pairw = (p.acc_mode == kAccPair) ? 1.0 : 0.0;
        const bool dot = p.red_mode == kRedDot;
        const bool need_x = p.acc_mode == kAccPair || dot;
        u64 chain = 0;
#pragma unroll
        for (int h = 0; h < kPerThread / 4; ++h) {
            const u64 gb = g0 + (((u64)__double_as_longlong(v[4 * h].x) | chain) & p.zero);
            double2 accv[4];
            if (with_acc) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) accv[jj] = ldg2_volatile(p.acc + gb + QSB_GOFF(4 * h + jj));
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * h + jj;
                const u64 g = gb + QSB_GOFF(j);
                const double2 o = cmul(p.scale, v[j]);
                double2 xo = make_double2(0.0, 0.0);
                if (need_x) xo = ts[QSB_IDX(j)];
                red += dot ? (xo.x * o.x + xo.y * o.y) : cnorm2(o);
                if (with_acc) {
                    double2 a = cadd(accv[jj], o);
                    cfma(a, pairw, xo);
                    p.acc[g] = a;
                    chain |= (u64)__double_as_longlong(a.x);
                }
                p.y[g] = o;
            }
        }
    }
    if constexpr (!SH || LAST) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty);
    }
#undef QSB_GADDR
#undef QSB_GOFF
#undef QSB_IDX
}

template <int TB, bool SH, bool OTF>
__global__ void __launch_bounds__(kWsThreads, 1) hpsi_ws_kernel(const HpsiWs p, const __grid_constant__ CUtensorMap tmap,
                                                                const __grid_constant__ CUtensorMap tmap_y,
                                                                const __grid_constant__ CUtensorMap tmap_acc,
                                                                const __grid_constant__ PeerMaps pmaps) {
    using C = WsCfg<TB>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[C::kStages];
    __shared__ __align__(8) uint64_t empty_bar[C::kStages];
    __shared__ __align__(16) WsItem desc[C::kStages];
    __shared__ int warps_done[2 * C::kStages];  // by item, not by stage: a stage is released before its tile is published
    __shared__ double red[32];
    __shared__ int is_last_cta;
    __shared__ double pjj[4];  // OTF: pairs among the three register bits of a contiguous tile: (0,1), (0,2), (1,2)

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;

    if (OTF && p.fused && t < 3) {
        const int a = (t == 2) ? 1 : 0, b = (t == 0) ? 1 : 2;
        pjj[t] = p.vbits[(C::kRegLo + a) * kMaxLocalBits + C::kRegLo + b];
    }
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], C::kGroupWarps);
            warps_done[s] = warps_done[s + C::kStages] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();

    double part = 0.0;
    if (warp >= kConsumerWarps) {
        // ================= producer warpgroup (only its first warp works) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");  // kWsRegsProducer
        if (warp == kConsumerWarps) {
            // Static schedule from ONE global ticket counter.  Tickets enumerate phases of n_a tiles:
            //   A(0) .. A(L-1),  then  A(s), B(s-L)  for s = L .. n_sb-1,  then  B(n_sb-L) .. B(n_sb-1)
            // (A = FIRST tiles of a super-block, B = its INNER tiles, L = p.lag super-blocks of lead).
            // B(s) is handed out L phases after A(s), so its dependency is normally already satisfied
            // when the ticket is taken; the producer still checks done[s] and waits if it is not.  It
            // only ever waits for tickets taken earlier by running CTAs: no deadlock whatever the
            // residency.  (A dynamic two-queue scheduler was measured slower: its extra L2 round
            // trips per item put the producer on the critical path.)
            const u64 n_a = p.fused ? (1ull << (p.sbits - TB)) : 1;  // tiles per super-block
            const u64 n_sb = p.fused ? (p.n_tiles / n_a) : 0;
            const u64 lag = p.fused ? ((u64)(p.lag < 0 ? 0 : p.lag) < n_sb ? (u64)(p.lag < 0 ? 0 : p.lag) : n_sb) : 0;  // 0: B(s) right after A(s)
            const u64 total = p.fused ? 2ull * p.n_tiles : p.n_tiles;
            int n_done = 0;
            int peer_left = 0;         // partner stages still to publish for the pending FIRST tile
            bool pend_first = false;   // the tile just issued was a contiguous (FIRST) one / a strided one
            u64 pend_tau = 0;
            for (u64 it = 0;; ++it) {
                const int s = (int)(it % C::kStages);
                const u64 use = it / C::kStages;
                if (use > 0) mbar_wait(&empty_bar[s], (uint32_t)((use - 1) & 1ull));
                int type = kItemDone;
                u64 tau = 0;
                if (SH && peer_left > 0) {
                    type = kItemPeer;
                    tau = pend_tau;
                } else {
                    if (lane == 0 && n_done == 0) {
                        // the counter is never reset: this launch's tickets start at ticket_base (mod 2^32)
                        const u64 ticket = (u64)(unsigned int)(atomicAdd(p.ticket, 1u) - p.ticket_base);
                        if (ticket < total) {
                            if (!p.fused) {
                                type = kItemPlain;
                                tau = ticket;
                            } else {
                                type = ws_decode_ticket(ticket, n_a, n_sb, lag, (u64)p.sb_xor, &tau);
                            }
                        }
                    }
                    type = __shfl_sync(0xffffffffu, type, 0);
                    tau = __shfl_sync(0xffffffffu, tau, 0);
                }
                if (SH && type == kItemPeer) {
                    // partner stage: the SAME amplitudes as the tile just issued, from a partner GPU's x (direct
                    // exchange: one stage per sharded qubit) or from the remap buffers (one stage), pulled over
                    // NVLink by the TMA engine into an ordinary stage of the ring
                    const int q = p.n_peer - peer_left;
                    --peer_left;
                    double2* pstage = tiles + (size_t)s * C::kTileW;
                    if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(C::kTileW * sizeof(double2)));
                    __syncwarp();
                    if (pend_first) {
                        const u64 pbase = tau << TB;
                        constexpr int kCopies = 4;
                        constexpr int chunk = C::kTileW / kCopies;
                        const double2* psrc = p.peer[q] + pbase;
                        if (p.remap) {
                            const int tcb = p.cbits - TB;  // tiles per chunk (log2)
                            psrc = p.zpeer[tau >> tcb] + ((u64)p.rank << p.cbits) + ((tau & ((1ull << tcb) - 1ull)) << TB);
                        }
                        if (lane < kCopies)
                            bulk_g2s(pstage + lane * chunk, psrc + (u64)lane * chunk, chunk * sizeof(double2), &full_bar[s]);
                    } else {
                        const int mid_bits = p.hb0 - p.lb;
                        const int c0 = (int)((tau & ((1ull << mid_bits) - 1ull)) << (p.lb + 1));
                        const int c1 = (int)((tau >> mid_bits) << p.hb);
                        const int rows = 1 << p.hb;
                        if (p.remap) {
                            // the rows of a strided tile fall into P groups by their top index bits = the rank whose
                            // remap buffer holds them: one box of 2^(hb - pbits) rows per source
                            const int rpc = p.hb - p.pbits;
                            if (lane < (1 << p.pbits)) tma_load_2d(pstage + ((size_t)(lane << rpc) << p.lb), &pmaps.m[lane], c0, 0, &full_bar[s]);
                        } else if (p.use_tmap) {
                            const int n_box = (rows + 255) >> 8;
                            if (lane < n_box) tma_load_2d(pstage + ((size_t)(lane * 256) << p.lb), &pmaps.m[q], c0, c1 + lane * 256, &full_bar[s]);
                        } else {
                            const u64 pbase = ws_tile_base(p.lb, p.hb, p.hb0, tau);
                            const uint32_t row_bytes = (uint32_t)(sizeof(double2) << p.lb);
                            for (int r = lane; r < rows; r += 32)
                                bulk_g2s(pstage + ((size_t)r << p.lb), p.peer[q] + pbase + ((u64)r << p.hb0), row_bytes, &full_bar[s]);
                        }
                    }
                    if (lane == 0) {
                        desc[s].type = kItemPeer;
                        desc[s].sb = q;
                        mbar_arrive(&full_bar[s]);
                    }
                    __syncwarp();
                    continue;
                }
                if (type == kItemDone) {
                    // out of work: one sentinel per consumer group (consecutive items go to distinct groups)
                    if (lane == 0) {
                        desc[s].type = kItemDone;
                        mbar_arrive(&full_bar[s]);
                    }
                    if (++n_done == C::kGroups) break;
                    continue;
                }
                const int sb = (int)(tau / n_a);
                double2* stage = tiles + (size_t)s * C::kTileW;
                u64 base;
                if (lane == 0)
                    mbar_expect_tx(&full_bar[s], (uint32_t)(C::kTileW * sizeof(double2)) +
                                                     ((OTF && type == kItemFirst) ? (uint32_t)(kTileTabDoubles * sizeof(double)) : 0u));
                __syncwarp();
                if (type == kItemFirst) {
                    base = tau << TB;
                    // OTF: the static pair part of the tile's diagonal comes from the per-tile table with the tile
                    if (OTF && lane == 5) bulk_g2s(&desc[s].w[0], p.tile_tab + tau * p.tab_stride, kTileTabDoubles * sizeof(double), &full_bar[s]);
#ifndef QSB_WS_COPIES
#define QSB_WS_COPIES 4
#endif
                    constexpr int kCopies = QSB_WS_COPIES;  // per-copy overhead of the TMA engine favours few large copies
                    constexpr int chunk = C::kTileW / kCopies;
                    if (lane < kCopies)
                        bulk_g2s(stage + lane * chunk, p.x + base + (u64)lane * chunk, chunk * sizeof(double2), &full_bar[s]);
                    if (!OTF && lane == kCopies) prefetch_l2(p.eint + base, (uint32_t)(C::kTileW * sizeof(double)));
                } else {
                    base = ws_tile_base(p.lb, p.hb, p.hb0, tau);
                    const int rows = 1 << p.hb;
                    if (p.use_tmap) {
                        // x viewed as [rows of 2^hb0 amplitudes][2 * 2^hb0 doubles]: the tile is the box
                        // {2 * 2^lb doubles} x {2^hb rows} at (2 * (mid << lb), outer << hb)
                        const int mid_bits = p.hb0 - p.lb;
                        const int c0 = (int)((tau & ((1ull << mid_bits) - 1ull)) << (p.lb + 1));
                        const int c1 = (int)((tau >> mid_bits) << p.hb);
                        const int n_box = (rows + 255) >> 8;
                        if (lane < n_box)
                            tma_load_2d(stage + ((size_t)(lane * 256) << p.lb), &tmap, c0, c1 + lane * 256, &full_bar[s]);
                        // the plain launch reads its partial y (and the accumulator) from HBM: start them now
                        if (type == kItemPlain && lane >= 8 && lane < 8 + n_box) tma_prefetch_2d(&tmap_y, c0, c1 + (lane - 8) * 256);
                        if (type == kItemPlain && p.last && p.acc_mode != kAccNone && lane >= 16 && lane < 16 + n_box)
                            tma_prefetch_2d(&tmap_acc, c0, c1 + (lane - 16) * 256);
                    } else {
                        const uint32_t row_bytes = (uint32_t)(sizeof(double2) << p.lb);
                        for (int r = lane; r < rows; r += 32)
                            bulk_g2s(stage + ((size_t)r << p.lb), p.x + base + ((u64)r << p.hb0), row_bytes, &full_bar[s]);
                    }
                }
                // does this tile pull its partners in THIS launch?
                const int n_follow = (SH && p.n_peer > 0 && type != kItemInner &&
                                      (p.peer_sel == 0 || (int)((base >> p.peer_bit) & 1ull) == (p.peer_sel == 2 ? 1 : 0)))
                                         ? p.n_peer : 0;
                double eh = 0.0;
                if (OTF && type == kItemFirst) {
                    // time-dependent single-bit part of the tile's diagonal: c0 + sum of lin[h] over the set bits of
                    // its base, one lane per bit
                    if ((tau >> lane) & 1ull) eh = p.lin[TB + lane];
                    eh = warp_sum(eh) + p.c0;
                }
                if (lane == 0) {
                    double dhi = eh;
                    if (!OTF && type == kItemFirst) {
                        // detuning of the index bits above the tile (identical for its 2^TB amplitudes)
                        dhi = p.sp.d_rank;
                        for (u64 rest = base >> TB; rest; rest &= rest - 1ull) dhi += p.sp.dl[TB + __ffsll((long long)rest) - 1];
                    }
                    if (type == kItemInner) {
                        // the partial y of this super-block must be complete (all its FIRST tiles stored)
                        while (ld_acquire_u64(p.done + sb) < p.done_target) __nanosleep(32);  // acquire side of done[]
                    }
                    desc[s].base = base;
                    desc[s].dhi = dhi;
                    desc[s].type = type;
                    desc[s].sb = sb;
                    desc[s].n_follow = n_follow;
                    mbar_arrive(&full_bar[s]);  // release: descriptor (and the INNER dependency) visible to consumers
                }
                if (n_follow > 0) {
                    peer_left = n_follow;
                    pend_tau = tau;
                    pend_first = (type == kItemFirst);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= consumer warps =================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");  // kWsRegsConsumer
        const int group = warp / C::kGroupWarps;
        const int tg = t - group * C::kGroupThreads;  // thread index within the group
        // detuning of the index bits this thread's tile index fixes (bits 0..TB-4 of a FIRST tile)
        double dlow = 0.0;
#pragma unroll
        for (int b = 0; b < C::kRegLo; ++b)
            if ((tg >> b) & 1) dlow += p.sp.dl[b];
        // OTF: tile-index part of the diagonal of the amplitudes this thread owns in EVERY contiguous tile: ett =
        // single-bit terms and pairs among the bits of tg; cross[jb] = single-bit term of register bit jb and its
        // pairs with the bits of tg.  Four values per thread for the whole launch.
        double ett = 0.0, cross[kRegBits] = {0.0, 0.0, 0.0};
        if (OTF && p.fused) {
#pragma unroll
            for (int b = 0; b < C::kRegLo; ++b)
                if ((tg >> b) & 1) {
                    ett += p.lin[b];
                    for (int b2 = b + 1; b2 < C::kRegLo; ++b2)
                        if ((tg >> b2) & 1) ett += __ldg(p.vbits + b * kMaxLocalBits + b2);
                }
#pragma unroll
            for (int jb = 0; jb < kRegBits; ++jb) {
                double c = p.lin[C::kRegLo + jb];
                for (int b = 0; b < C::kRegLo; ++b)
                    if ((tg >> b) & 1) c += __ldg(p.vbits + b * kMaxLocalBits + C::kRegLo + jb);
                cross[jb] = c;
            }
        }
        const WsRing ring{tiles, full_bar, empty_bar, desc};
        for (u64 it = (u64)group;; it += C::kGroups) {
            const int s = (int)(it % C::kStages);
            const u64 use = it / C::kStages;
            mbar_wait(&full_bar[s], (uint32_t)(use & 1ull));
            const int type = desc[s].type;
            if (type == kItemDone) break;
            const u64 base = desc[s].base;
            const int sb = desc[s].sb;
            const int n_follow = SH ? desc[s].n_follow : 0;
            const double2* ts = tiles + (size_t)s * C::kTileW;
            if (type == kItemFirst) {
                const double dsum = dlow + desc[s].dhi;
                double diag[kPerThread];
                if constexpr (OTF) {
                    // + the tile-base part: its own terms (dhi, w[12]) and its coupling to this thread's tile bits (w)
                    double eh = (desc[s].dhi + desc[s].w[12]) + ett;
#pragma unroll
                    for (int b = 0; b < C::kRegLo; ++b)
                        if ((tg >> b) & 1) eh += desc[s].w[b];
                    const double c0 = cross[0] + desc[s].w[C::kRegLo], c1 = cross[1] + desc[s].w[C::kRegLo + 1],
                                 c2 = cross[2] + desc[s].w[C::kRegLo + 2];
                    const double p01 = pjj[0], p02 = pjj[1], p12 = pjj[2];
                    diag[0] = eh;
                    diag[1] = eh + c0;
                    diag[2] = eh + c1;
                    diag[3] = diag[1] + (c1 + p01);
                    diag[4] = eh + c2;
                    diag[5] = diag[4] + (c0 + p02);
                    diag[6] = diag[4] + (c1 + p12);
                    diag[7] = diag[6] + (c0 + (p01 + p02));
                }
                ws_process<TB, true, false, SH, OTF>(p, ts, base, tg, dsum, diag, part, &empty_bar[s], ring, n_follow, it);
                // publish this tile of partial y: warp barrier -> cta-scope release on the shared
                // counter; the last warp of the group to finish issues ONE gpu-scope fence (cumulative
                // over the writes it observed through the counter) and bumps the super-block counter.
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    // a warp cannot run more than kStages items ahead of the slowest one: 2 * kStages slots never alias
                    const int slot = (int)(it % (2 * C::kStages));
                    const int c = atomicAdd(&warps_done[slot], 1);
                    if (c == C::kGroupWarps - 1) {
                        warps_done[slot] = 0;
                        __threadfence();
                        atomicAdd(p.done + sb, 1ull);
                    }
                }
            } else {
                const double nodiag[kPerThread] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (p.last) ws_process<TB, false, true, SH, OTF>(p, ts, base, tg, 0.0, nodiag, part, &empty_bar[s], ring, n_follow, it);
                else ws_process<TB, false, false, SH, OTF>(p, ts, base, tg, 0.0, nodiag, part, &empty_bar[s], ring, n_follow, it);
            }
        }
    }
    if (p.norm_part == nullptr) return;  // uniform: only the launch that finishes a fused H psi reduces
    // per-CTA partial, then the LAST CTA to arrive sums all partials in a fixed order (deterministic)
    part = warp_sum(part);
    if (lane == 0) red[warp] = (warp < kConsumerWarps) ? part : 0.0;
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        for (int w = 0; w < kConsumerWarps; ++w) tot += red[w];
        p.norm_part[blockIdx.x] = tot;
        __threadfence();
        const unsigned int arrived = atomicAdd(p.tail, 1u);
        is_last_cta = (arrived == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (is_last_cta && warp == 0) {
        __threadfence();
        double s = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) s += ld_volatile_f64(p.norm_part + i);
        s = warp_sum(s);
        if (lane == 0) {
            *p.red_out = s;
            *p.tail = 0u;  // ready for the next launch on this stream
        }
    }
}

}  // namespace qsb
