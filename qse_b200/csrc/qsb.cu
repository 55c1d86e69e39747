// qsb.cu — libqsb: context, kernel launchers, Taylor stepper, observables and the C ABI of
// include/qsb.h.  One context = one GPU (one shard of the state vector).
#define QSB_MAX_QUBITS_DEV 34
#include "../../include/qsb.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "hpsi.cuh"
#include "hpsi_ws.cuh"
#include "nccl_dl.h"
#include "observe.cuh"
#include "small.cuh"
#include "terms.cuh"

#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define QSB_HAVE_NVTX 1
#endif

using namespace qsb;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

struct qsb_ctx {
    int n = 0, p = 0, nl = 0, rank = 0, nranks = 1, device = 0;
    u64 dim = 0;        // amplitudes in this shard
    u64 rank_bits = 0;  // rank << nl
    int pop_rank = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double2 *psi = nullptr, *wa = nullptr, *wb = nullptr;
    double* eint = nullptr;
    double* vtab = nullptr;
    bool have_v = false;
    double diag_min = 0.0, diag_max = 0.0;
    double* part = nullptr;  // per-block partials
    size_t part_n = 0;
    double* slots = nullptr;  // device result slots
    double* h_slots = nullptr;
    size_t slots_n = 0;
    HpsiPlan plan;
    // warp-specialised super-block path (hpsi_ws.cuh)
    int path = 0;       // 0 auto (ws when N_local >= 13), 1 untiled, 2 one launch per tile-bit group
    int sb_lag = 0;     // lead of the FIRST tiles over the INNER tiles of the fused launch, in super-blocks
    int ws_tile_bits = 12;  // 12 (64 KiB tiles, 3 stages; measured best) or 11 (32 KiB tiles, 6 stages, two groups); 0 = fewest launches
    int sb_bits = 0;   // log2 of the L2-resident super-block (10 MiB each; 18 / lead 3 measured best on B200)
    unsigned int* tickets = nullptr;        // one work counter per launch of an H psi (never reset: see ticket_base)
    unsigned int ticket_base[16] = {0};     // host mirror: where the next launch's tickets start on each counter
    unsigned int* tail = nullptr;           // CTAs that have stored their reduction partial (self-resetting)
    unsigned long long* sb_done = nullptr;  // per super-block completion counters (cumulative)
    unsigned long long ws_epoch = 0;
    int ws_sbits_active = -1;
    int small_kernel = 1;  // option "small_kernel": whole-schedule single-CTA kernel for N <= 12
    int obs_tiled = 1;     // option "tiled_observables": spins from tile passes (3 reads at 25 qubits) instead of 1 + 2N
    double* sched_dev = nullptr;  // device copy of a schedule (small kernel)
    size_t sched_cap = 0;
    int use_tmap = 1;  // option "tensor_map": 2-D tensor-map TMA for strided tiles (0 = one bulk copy per row)
    struct TmapEntry {
        const void* ptr;
        int lb, hb, hb0, rows_log2, box_rows_log2;
        CUtensorMap map;
    };
    std::vector<TmapEntry> tmaps;  // super-block size the completion counters were accumulated with
    // per-site structure of H (README.md:93-95):
    //   H = sum_q (Omega/2 w_omega[q] + x0[q]) X_q - sum_q delta w_delta[q] n_q + E + operator terms
    //   w_omega / w_delta: qsb_set_site_weights (default 1);  t_omega / t_delta / x0: single-X / single-N strings
    //   of qsb_set_operator_terms that fit the fast form (default 0)
    double w_omega[QSB_MAX_QUBITS_DEV + 4], w_delta[QSB_MAX_QUBITS_DEV + 4], x0[QSB_MAX_QUBITS_DEV + 4];
    double t_omega[QSB_MAX_QUBITS_DEV + 4], t_delta[QSB_MAX_QUBITS_DEV + 4];
    std::vector<DiagTerm> diag_terms;       // static diagonal strings, added to E after the interaction
    DiagTerm* diag_dev = nullptr;
    // operator terms outside the fast form (terms.cuh), per rank
    GenTerm* gen_dev = nullptr;
    int n_gen = 0;
    double gen_bound[3] = {0.0, 0.0, 0.0};  // sum |coef| per channel (spectral bound)
    bool have_terms = false;                // qsb_set_operator_terms was called (E holds their static diagonal)
    std::vector<double> v_host;             // last interaction matrix (E is rebuilt when the terms change)
    // 2-local model of the static diagonal, by QUBIT: E(k) = dm_c0 + sum_q dm_lin[q] b_q + sum_{q<q'} dm_pair[q][q'] b_q b_q'
    // (interaction + static diagonal strings of at most two factors); dm_ok = false when a longer string exists
    bool dm_ok = false;
    double dm_c0 = 0.0;
    std::vector<double> dm_lin, dm_pair;
    double* vbits_dev = nullptr;            // dm_pair over the LOCAL index bits, [kMaxLocalBits][kMaxLocalBits]
    double* tile_tab[2] = {nullptr, nullptr};  // static per-tile part of the diagonal for 2^12 / 2^11-amplitude tiles
    double* tab_zero = nullptr;             // one all-zero table row (stored-E mode)
    int diag_otf = -1;                      // option "diag_on_the_fly": -1 auto (from 25 local qubits), 0 stored E, 1 on the fly
    // options
    double tol = 1e-14, theta_max = 3.0;
    int max_terms = 64, fixed_terms = 0, force_flat = 0;
    int m_prev = 0;
    // sharding
    nccl_comm_t comm = nullptr;
    bool ipc = false;
    double2* peer_psi[64] = {nullptr};
    double2* peer_wa[64] = {nullptr};
    double2* peer_wb[64] = {nullptr};
    double2* stage[kMaxPeers] = {nullptr};  // NCCL-staged partner shards when IPC is unavailable
    double2* zbuf = nullptr;                // remap exchange: neighbour sums in the transposed layout
    double2* peer_z[64] = {nullptr};
    int exchange = -1;                      // option "exchange": -1 auto, 0 direct partner pulls, 1 remap
    int exchange_split = -1;                // option "exchange_split": where the partner tiles are pulled: -1 auto, 0 fused
                                            // launch, 1 half in the fused and half in the last plain launch, 2 plain launch
    int remap_ctas = 32;                    // option "remap_ctas": SMs the remap kernel takes beside the fused launch
    cudaStream_t stream2 = nullptr;         // the remap kernel runs beside the fused launch (sharded contexts)
    cudaEvent_t ev_x = nullptr, ev_z = nullptr;
    // measurement
    double* flush_buf = nullptr;
    size_t flush_n = 0;
    unsigned char* topk_buf = nullptr;  // radix histogram + candidate lists (lazily allocated)
    size_t topk_bytes = 0;
    qsb_metrics met;
    char err[512] = "";
};

static int fail(qsb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    snprintf(g_err, sizeof(g_err), "%s", buf);
    if (c) snprintf(c->err, sizeof(c->err), "%s", buf);
    return code;
}

static inline void nvtx_push(const char* name) {
#ifdef QSB_HAVE_NVTX
    nvtxRangePushA(name);
#else
    (void)name;
#endif
}
static inline void nvtx_pop() {
#ifdef QSB_HAVE_NVTX
    nvtxRangePop();
#endif
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(c, QSB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)
#define NC(call)                                                                                      \
    do {                                                                                              \
        int r_ = (call);                                                                              \
        if (r_ != kNcclSuccess)                                                                       \
            return fail(c, QSB_ERR_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), \
                        __FILE__, __LINE__);                                                          \
    } while (0)
#define TRY(call)              \
    do {                       \
        int r_ = (call);       \
        if (r_ != QSB_OK) return r_; \
    } while (0)
#define LAUNCHED(c) ((c)->met.kernel_launches++)

static int check_launch(qsb_ctx* c, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, QSB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return QSB_OK;
}

static int grid_for(const qsb_ctx* c, u64 work, int per_block) {
    const u64 need = (work + per_block - 1) / per_block;
    const u64 cap = (u64)c->n_sm * 8;
    return (int)std::max<u64>(1, std::min<u64>(need, cap));
}

// ------------------------------------------------------------------------------------------------
// collectives on small device arrays
// ------------------------------------------------------------------------------------------------
static int allreduce(qsb_ctx* c, double* dev, size_t count, int op) {
    if (c->nranks == 1) return QSB_OK;
    NC(nccl_api()->AllReduce(dev, dev, count, kNcclFloat64, op, c->comm, c->stream));
    return QSB_OK;
}
// stream-ordered barrier across ranks (a 1-double all-reduce on a scratch slot)
static int rank_barrier(qsb_ctx* c) {
    if (c->nranks == 1) return QSB_OK;
    return allreduce(c, c->slots + c->slots_n - 1, 1, kNcclSum);
}
// asynchronous NCCL failures (a peer died, a transport error) surface here instead of as a hang
static int nccl_async_check(qsb_ctx* c) {
    if (c->nranks == 1 || !c->comm || !nccl_api()->CommGetAsyncError) return QSB_OK;
    int async_err = kNcclSuccess;
    const int r = nccl_api()->CommGetAsyncError(c->comm, &async_err);
    if (r != kNcclSuccess) return fail(c, QSB_ERR_NCCL, "ncclCommGetAsyncError failed: %s", nccl_api()->GetErrorString(r));
    if (async_err != kNcclSuccess && async_err != kNcclInProgress)
        return fail(c, QSB_ERR_NCCL, "asynchronous NCCL error on rank %d: %s", c->rank, nccl_api()->GetErrorString(async_err));
    return QSB_OK;
}
static int fetch_slots(qsb_ctx* c, size_t first, size_t count) {
    CU(cudaMemcpyAsync(c->h_slots + first, c->slots + first, count * sizeof(double), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return nccl_async_check(c);
}

// ------------------------------------------------------------------------------------------------
// H psi launcher
// ------------------------------------------------------------------------------------------------
static const double2* peer_of(qsb_ctx* c, const double2* x, int partner) {
    if (x == c->psi) return c->peer_psi[partner];
    if (x == c->wa) return c->peer_wa[partner];
    if (x == c->wb) return c->peer_wb[partner];
    return nullptr;
}

struct WsPlan {
    int tb;     // tile bits (11 or 12)
    int sbits;
    PassGeom inner;
    int n_plain;
    PassGeom plain[8];
    double bytes_per_amp;
};
static WsPlan make_ws_plan_tb(int nl, int sb_bits, int tb) {
    WsPlan w{};
    w.tb = tb;
    const int max_hb = tb - 3;  // keep >= 128-byte rows
    w.sbits = std::max(tb + 1, std::min(nl, std::min(sb_bits, tb + max_hb)));
    if (w.sbits > nl) w.sbits = nl;
    const int hb = w.sbits - tb;
    w.inner = PassGeom{tb - hb, hb, tb};
    w.bytes_per_amp = 40.0;
    int rem = nl - w.sbits;
    const int extra = (rem + max_hb - 1) / max_hb;
    int pos = w.sbits;
    for (int i = 0; i < extra; ++i) {
        const int h = (rem + (extra - i) - 1) / (extra - i);
        w.plain[w.n_plain++] = PassGeom{tb - h, h, pos};
        w.bytes_per_amp += 48.0;
        pos += h;
        rem -= h;
    }
    return w;
}
// Measured on B200 (tools/ab_hpsi.py): the super-block must leave at most 9 index bits to the
// plain launch (else a third trip to HBM), 2^18 amplitudes (10 MiB of x+y+E) with a lead of 3 blocks
// is best while that holds (N_local <= 27); larger blocks want a shorter lead to stay inside L2.
// Up to 21 local qubits the whole shard is ONE super-block (a single fused launch, measured faster:
// 19/20/21 qubits 22/31/51 us against 28/39/63 us with a plain launch on top).
// With the diagonal evaluated on the fly no E stream passes through the L2 and larger blocks stay resident; measured
// (same box, tools/ab_opts.py, profiles/r02_ab_sb_*.txt): 25 q 2^19 / lead 2: -4 %; 27 q 2^19 / 2: -15 %; 28 q 2^20 / 1:
// -14 %; 29 q 2^21 / 1: -7 % (sweep step -12 %); 30 q 2^21 with no lead (one 64 MiB block in flight instead of two): -4 %.
static int auto_sb_bits(int nl, bool otf) {
    if (nl <= 21) return nl;
    if (!otf) return std::max(18, std::min(21, nl - 9));
    return nl <= 24 ? 18 : (nl <= 27 ? 19 : (nl == 28 ? 20 : 21));
}
// (no lead only on one GPU: sharded, 30 local qubits are NVLink-bound and the lead-1 schedule is the one validated
// at 33 qubits on 8 GPUs)
static int auto_sb_lag(int sbits, int nl, int nranks) {
    return sbits <= 18 ? 3 : (sbits == 19 ? 2 : ((sbits == 21 && nl >= 30 && nranks == 1) ? 0 : 1));
}

static WsPlan make_ws_plan(int nl, int sb_bits, int tile_bits, bool otf = true) {
    if (sb_bits == 0) sb_bits = auto_sb_bits(nl, otf);
    if (tile_bits == 11 || tile_bits == 12) return make_ws_plan_tb(nl, sb_bits, tile_bits);
    const WsPlan a = make_ws_plan_tb(nl, sb_bits, 11), b = make_ws_plan_tb(nl, sb_bits, 12);
    return (a.n_plain <= b.n_plain) ? a : b;
}

// 2-D tensor map over a state-sized vector for strided tiles {2^lb amplitudes} x {2^hb rows, stride 2^hb0}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// rows_log2 / box_rows_log2 < 0: the whole shard ({2^(nl - hb0) rows}, boxes of min(2^hb, 256) rows)
static int get_tmap(qsb_ctx* c, const double2* x, int lb, int hb, int hb0, const CUtensorMap** out, int rows_log2 = -1,
                    int box_rows_log2 = -1) {
    if (rows_log2 < 0) rows_log2 = c->nl - hb0;
    if (box_rows_log2 < 0) box_rows_log2 = std::min(hb, 8);
    for (auto& e : c->tmaps)
        if (e.ptr == x && e.lb == lb && e.hb == hb && e.hb0 == hb0 && e.rows_log2 == rows_log2 && e.box_rows_log2 == box_rows_log2) {
            *out = &e.map;
            return QSB_OK;
        }
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(c, QSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = (EncodeTiledFn)fn;
    }
    qsb_ctx::TmapEntry e;
    e.ptr = x;
    e.lb = lb;
    e.hb = hb;
    e.hb0 = hb0;
    e.rows_log2 = rows_log2;
    e.box_rows_log2 = box_rows_log2;
    const cuuint64_t gdim[2] = {2ull << hb0, 1ull << rows_log2};
    const cuuint64_t gstride[1] = {16ull << hb0};
    const cuuint32_t box[2] = {2u << lb, 1u << box_rows_log2};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)x, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, QSB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for lb=%d hb=%d hb0=%d", (int)r, lb, hb, hb0);
    c->tmaps.push_back(e);
    *out = &c->tmaps.back().map;
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// H(t) parameters: per-site drive / detuning in the index-bit layout the kernels use
// ------------------------------------------------------------------------------------------------
struct HParams {
    SiteParams sp;
    double hw_abs = 0.0;            // sum_q |hw_q|                      (spectral bound)
    double d_pos = 0.0, d_neg = 0.0;  // sum_q max(delta_q, 0), sum_q max(-delta_q, 0)
    double chan[3] = {1.0, 0.0, 0.0};  // scale of the operator terms of channel 0 (static), 1 (amplitude), 2 (detuning)
    // diagonal of H over the LOCAL index bits with the rank's bits folded in (hpsi_ws.cuh, "diagonal on the fly"):
    // d(k) = c0 + sum_b lin[b] b_b(k) + sum_{b<b'} vbits[b][b'] b_b b_b'; with a stored E array the pair table is
    // zero and (c0, lin) carry only the detuning
    double lin[kMaxLocalBits];
    double c0 = 0.0;
    bool otf = false;
};

// hw_site[q], d_site[q] for q = 0..n-1 (qubit q <-> index bit n-1-q; bits >= nl are rank bits)
static HParams hparams_from_sites(const qsb_ctx* c, const double* hw_site, const double* d_site, double amp, double det) {
    HParams h;
    memset(&h.sp, 0, sizeof(h.sp));
    for (int q = 0; q < c->n; ++q) {
        const int bit = c->n - 1 - q;
        if (bit < c->nl) {
            h.sp.hw[bit] = hw_site[q];
            h.sp.dl[bit] = d_site[q];
        } else {
            const int g = bit - c->nl;
            h.sp.hw_peer[g] = hw_site[q];
            if ((c->rank >> g) & 1) h.sp.d_rank += d_site[q];
        }
        h.hw_abs += fabs(hw_site[q]);
        h.d_pos += std::max(d_site[q], 0.0);
        h.d_neg += std::max(-d_site[q], 0.0);
    }
    h.chan[0] = 1.0;
    h.chan[1] = amp;
    h.chan[2] = det;
    // diagonal model: detuning, plus (on the fly) the static 1- / 2-local part with this rank's bits folded in
    // measured (same box): with the same plan on the fly is 1.3 % slower at 25 qubits and 2.8 % faster at 28; but without
    // an E stream through the L2 a larger super-block stays resident (auto_sb_bits), which wins 4 % at 25 and 14 % at 28
    h.otf = c->dm_ok && (c->diag_otf < 0 ? c->nl >= 25 : c->diag_otf != 0);
    memset(h.lin, 0, sizeof(h.lin));
    h.c0 = -h.sp.d_rank;
    for (int b = 0; b < c->nl; ++b) h.lin[b] = -h.sp.dl[b];
    if (h.otf) {
        const int n = c->n;
        h.c0 += c->dm_c0;
        for (int q = 0; q < n; ++q) {
            const int bit = n - 1 - q;
            if (bit < c->nl) {
                h.lin[bit] += c->dm_lin[q];
                for (int g = 0; g < c->p; ++g)  // sharded qubit g' = n-1-(nl+g) set in this rank
                    if ((c->rank >> g) & 1) {
                        const int qg = n - 1 - (c->nl + g);
                        h.lin[bit] += c->dm_pair[(size_t)std::min(q, qg) * n + std::max(q, qg)];
                    }
            } else if ((c->rank >> (bit - c->nl)) & 1) {
                h.c0 += c->dm_lin[q];
                for (int q2 = q + 1; q2 < n; ++q2) {
                    const int bit2 = n - 1 - q2;
                    if (bit2 >= c->nl && ((c->rank >> (bit2 - c->nl)) & 1)) h.c0 += c->dm_pair[(size_t)q * n + q2];
                }
            }
        }
    }
    return h;
}
// the global pulse of qse.calc.Qutip (qse/calc/qutip.py:55-58) through the per-site weights
static HParams hparams_global(const qsb_ctx* c, double omega, double delta) {
    double hw[QSB_MAX_QUBITS_DEV + 4], d[QSB_MAX_QUBITS_DEV + 4];
    for (int q = 0; q < c->n; ++q) {
        hw[q] = 0.5 * omega * (c->w_omega[q] + c->t_omega[q]) + c->x0[q];
        d[q] = delta * (c->w_delta[q] + c->t_delta[q]);
    }
    return hparams_from_sites(c, hw, d, omega, delta);
}
// a per-site schedule row: omega_q and delta_q given for every qubit (weights still apply)
static HParams hparams_rows(const qsb_ctx* c, const double* omega_q, const double* delta_q, double amp, double det) {
    double hw[QSB_MAX_QUBITS_DEV + 4], d[QSB_MAX_QUBITS_DEV + 4];
    for (int q = 0; q < c->n; ++q) {
        hw[q] = 0.5 * (omega_q[q] * c->w_omega[q] + amp * c->t_omega[q]) + c->x0[q];
        d[q] = delta_q[q] * c->w_delta[q] + det * c->t_delta[q];
    }
    return hparams_from_sites(c, hw, d, amp, det);
}

// Tickets are claimed ONE at a time, right when a stage is free.  Measured (same box, tools/ab_hpsi.py, 25 qubits):
// claiming in batches of 2 / 4 / 8 costs +6 / +34 / +77 %, and claiming the next ticket ahead of the stage wait +6 %:
// both widen the window of super-blocks in flight (148 CTAs x tickets held) beyond what the L2 keeps resident and the
// A -> B lead covers.
static unsigned int take_tickets(qsb_ctx* c, int counter, u64 total, int grid) {
    // every launch performs exactly total + grid atomicAdds on its counter (one over-the-end ticket per CTA)
    const unsigned int base = c->ticket_base[counter];
    c->ticket_base[counter] = (unsigned int)(base + (unsigned int)total + (unsigned int)grid);
    return base;
}

// warp-specialised path: one fused launch over L2-resident super-blocks + plain launches for the
// index bits above the super-block (hpsi_ws.cuh)
static int launch_hpsi_ws(qsb_ctx* c, const double2* x, double2* y, const HParams& hp, bool fuse, double2 scale,
                          int acc_mode, int red_mode, double* red_slot, const double2* const* peers) {
    // partner stages and their FIRST stage must be consumed by the same threads: one consumer group
    const WsPlan w = make_ws_plan(c->nl, c->sb_bits, c->nranks > 1 ? 12 : c->ws_tile_bits, hp.otf);
    const u64 n_tiles = c->dim >> w.tb;
    const int grid = (int)std::min<u64>(n_tiles, (u64)c->n_sm);
    HpsiWs a{};
    a.sp = hp.sp;
    a.otf = hp.otf ? 1 : 0;
    a.vbits = c->vbits_dev + (hp.otf ? 0 : kMaxLocalBits * kMaxLocalBits);  // second table: all zero
    {
        double*& tab = c->tile_tab[w.tb == 12 ? 0 : 1];
        if (!tab) {  // built on first use for this tile size, from the current pair table
            const u64 n = (c->dim >> w.tb) * kTileTabDoubles;
            CU(cudaMalloc(&tab, n * sizeof(double)));
            tile_table_kernel<<<grid_for(c, n, 256), 256, 0, c->stream>>>(tab, c->vbits_dev, c->nl, w.tb);
            LAUNCHED(c);
            TRY(check_launch(c, "tile_table_kernel"));
        }
        a.tile_tab = hp.otf ? tab : c->tab_zero;
        a.tab_stride = hp.otf ? (u64)kTileTabDoubles : 0ull;
    }
    memcpy(a.lin, hp.lin, sizeof(a.lin));
    a.c0 = hp.c0;
    // remap exchange: from 4 ranks up, when peer memory is mapped and a chunk holds whole tiles
    const bool remap = c->nranks >= 4 && c->nranks <= 8 && c->ipc && c->zbuf && c->exchange != 0 && (c->nl - c->p) >= w.tb;
    const int np = c->p == 0 ? 0 : (remap ? 1 : c->p);  // partner stages per tile
    // Where the partner tiles are pulled (hpsi_ws.cuh, peer_sel): the NVLink traffic should run beside ALL the local
    // work of an H psi, not beside one launch only.
    //   direct exchange: half of the tiles pull in the fused launch, the other half in the last plain launch;
    //   remap exchange : the remap kernel runs BESIDE the fused launch (own stream, its own SMs), every tile
    //                    pulls its Z rows in the last plain launch.
    const PassGeom gl = w.n_plain > 0 ? w.plain[w.n_plain - 1] : w.inner;
    const bool plain_ok = w.n_plain > 0 && np > 0 && (!remap || (gl.hb >= c->p && gl.hb0 + gl.hb == c->nl && c->use_tmap && gl.lb <= 7));
    int split = c->exchange_split;
    if (split < 0) split = plain_ok ? (remap ? 2 : 1) : 0;
    if (!plain_ok) split = 0;
    // the plain launch that ends a Taylor term keeps its own tile while the partner stages pass through the ring:
    // with 3 stages that leaves room for two of them
    if (!remap && split != 0 && np > 2) split = 0;
    const bool concurrent_remap = remap && split == 2 && c->stream2 && c->remap_ctas > 0 && c->remap_ctas < c->n_sm;
    const int grid_fused = concurrent_remap ? std::max(1, std::min(grid, c->n_sm - c->remap_ctas)) : grid;
    if (remap) {
        RemapArgs ra{};
        for (int r = 0; r < c->nranks; ++r) {
            ra.xs[r] = peer_of(c, x, r);
            if (!ra.xs[r]) return fail(c, QSB_ERR_STATE, "H psi input is not an exported buffer");
            a.zpeer[r] = c->peer_z[r];
        }
        for (int g = 0; g < kMaxPeers; ++g) ra.hw[g] = hp.sp.hw_peer[g];
        ra.z = c->zbuf;
        ra.rank = c->rank;
        ra.cbits = c->nl - c->p;
        nvtx_push("qsb:remap_exchange");
        if (concurrent_remap) {
            // x is final once everything queued on the main stream so far has run; the remap kernel then owns
            // remap_ctas SMs (one 1024-thread CTA each) while the fused launch runs on the others
            CU(cudaEventRecord(c->ev_x, c->stream));
            CU(cudaStreamWaitEvent(c->stream2, c->ev_x, 0));
            if (c->nranks == 4) remap_z_kernel<4><<<c->remap_ctas, 1024, 0, c->stream2>>>(ra);
            else remap_z_kernel<8><<<c->remap_ctas, 1024, 0, c->stream2>>>(ra);
            LAUNCHED(c);
            TRY(check_launch(c, "remap_z_kernel"));
            CU(cudaEventRecord(c->ev_z, c->stream2));
        } else {
            const int rgrid = std::max(1, grid_for(c, 1ull << ra.cbits, 1024) / 4);
            if (c->nranks == 4) remap_z_kernel<4><<<rgrid, 1024, 0, c->stream>>>(ra);
            else remap_z_kernel<8><<<rgrid, 1024, 0, c->stream>>>(ra);
            LAUNCHED(c);
            TRY(check_launch(c, "remap_z_kernel"));
            TRY(rank_barrier(c));  // every rank's Z must be complete before anyone pulls from it
        }
        nvtx_pop();
        a.remap = 1;
        a.rank = c->rank;
        a.cbits = ra.cbits;
        a.pbits = c->p;
        // contiguous tiles pull from ONE source each: stagger the super-block order per rank (no NVLink hot spot)
        a.sb_xor = (split != 2 && ra.cbits >= w.sbits) ? ((unsigned int)c->rank << (ra.cbits - w.sbits)) : 0u;
        for (int g = 0; g < kMaxPeers; ++g) a.sp.hw_peer[g] = 1.0;  // the coefficients are already in Z
    }
    a.x = x;
    a.y = y;
    a.eint = c->eint;
    for (int g = 0; g < kMaxPeers; ++g) a.peer[g] = peers[g];
    a.peer_bit = w.sbits - 1;
    a.scale = fuse ? scale : make_double2(1.0, 0.0);
    a.nl = c->nl;
    a.done = c->sb_done;
    a.n_tiles = n_tiles;
    for (int i = 0; i <= w.n_plain; ++i) {
        const bool last = (i == w.n_plain);
        a.fused = (i == 0);
        const PassGeom g = (i == 0) ? w.inner : w.plain[i - 1];
        a.sbits = w.sbits;
        a.lb = g.lb;
        a.hb = g.hb;
        a.hb0 = g.hb0;
        a.last = (last && fuse) ? 1 : 0;
        a.lag = c->sb_lag > 0 ? c->sb_lag : (c->sb_lag < 0 ? 0 : auto_sb_lag(w.sbits, c->nl, c->nranks));
        // partner stages of this launch
        if (i == 0) {
            a.n_peer = (split == 2) ? 0 : np;
            a.peer_sel = (split == 1) ? 1 : 0;
        } else if (last) {
            a.n_peer = (split == 0) ? 0 : np;
            a.peer_sel = (split == 1) ? 2 : 0;
        } else {
            a.n_peer = 0;
            a.peer_sel = 0;
        }
        const int lgrid = (i == 0) ? grid_fused : grid;
        if (i == 1 && concurrent_remap) {
            // the fused launch is queued; Z must be complete on EVERY rank before the plain launch pulls from it
            CU(cudaStreamWaitEvent(c->stream, c->ev_z, 0));
            TRY(rank_barrier(c));
        }

        a.acc_mode = (fuse && last) ? acc_mode : kAccNone;
        a.red_mode = red_mode;
        a.acc = (a.acc_mode != kAccNone) ? c->psi : nullptr;
        const bool reduce = fuse && last && red_slot;
        a.norm_part = reduce ? c->part : nullptr;
        a.red_out = reduce ? red_slot : nullptr;
        a.tail = c->tail;
        a.ticket = c->tickets + i;
        a.ticket_base = take_tickets(c, i, a.fused ? 2 * n_tiles : n_tiles, lgrid);
        if (i == 0) {
            if (c->ws_sbits_active != w.sbits * 16 + w.tb) {
                // the counters are cumulative per super-block: a new block size starts a new count
                CU(cudaMemsetAsync(c->sb_done, 0, ((size_t)(c->dim >> 13) + 1) * sizeof(unsigned long long), c->stream));
                c->ws_epoch = 0;
                c->ws_sbits_active = w.sbits * 16 + w.tb;
            }
            c->ws_epoch++;
            a.done_target = c->ws_epoch << (w.sbits - w.tb);
        }
        // strided tiles: tensor-map TMA when a row fits one box (<= 256 doubles) and there are rows at all
        a.use_tmap = (c->use_tmap && g.hb > 0 && g.lb <= 7) ? 1 : 0;
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof(dummy));
        const CUtensorMap* tm = &dummy;
        const CUtensorMap *tm_y = &dummy, *tm_acc = &dummy;
        CUtensorMap tm_copy, tm_y_copy;  // get_tmap may grow the cache: copy before the next lookup
        if (a.use_tmap) {
            TRY(get_tmap(c, x, g.lb, g.hb, g.hb0, &tm));
            tm_copy = *tm;
            tm = &tm_copy;
            TRY(get_tmap(c, y, g.lb, g.hb, g.hb0, &tm_y));
            tm_y_copy = *tm_y;
            tm_y = &tm_y_copy;
            if (a.acc) TRY(get_tmap(c, a.acc, g.lb, g.hb, g.hb0, &tm_acc));
        }
        // strided partner tiles: the same boxes out of the partners' x (direct), or one box of rows per source rank
        // out of chunk `rank` of its remap buffer
        PeerMaps pm;
        memset(&pm, 0, sizeof(pm));
        if (i > 0 && a.n_peer > 0 && a.use_tmap) {
            const CUtensorMap* t = nullptr;
            if (remap) {
                const int rpc = g.hb - c->p;
                for (int r = 0; r < c->nranks; ++r) {
                    TRY(get_tmap(c, c->peer_z[r] + ((u64)c->rank << a.cbits), g.lb, g.hb, g.hb0, &t, rpc, rpc));
                    pm.m[r] = *t;
                }
            } else {
                for (int q = 0; q < c->p; ++q) {
                    TRY(get_tmap(c, peers[q], g.lb, g.hb, g.hb0, &t));
                    pm.m[q] = *t;
                }
            }
        }
#define QSB_WS_LAUNCH(TB_, SH_, OTF_) hpsi_ws_kernel<TB_, SH_, OTF_><<<lgrid, kWsThreads, kWsSmemBytes, c->stream>>>(a, *tm, *tm_y, *tm_acc, pm)
        if (w.tb == 11) {
            if (hp.otf) QSB_WS_LAUNCH(11, false, true);
            else QSB_WS_LAUNCH(11, false, false);
        } else if (c->nranks == 1) {
            if (hp.otf) QSB_WS_LAUNCH(12, false, true);
            else QSB_WS_LAUNCH(12, false, false);
        } else {
            if (hp.otf) QSB_WS_LAUNCH(12, true, true);
            else QSB_WS_LAUNCH(12, true, false);
        }
#undef QSB_WS_LAUNCH
        LAUNCHED(c);
        TRY(check_launch(c, "hpsi_ws_kernel"));
    }
    return QSB_OK;
}

// y (+)= sum_t chan[t] coef_t P_t x for the operator terms that are not in the fast form (terms.cuh)
static int launch_terms(qsb_ctx* c, const double2* x, double2* y, const HParams& hp);
// out = scale * y;  reduction;  accumulator update -- the Taylor epilogue as a kernel of its own
// (only needed when operator terms follow the fast H psi)
static int launch_epilogue(qsb_ctx* c, const double2* x, double2* y, double2 scale, int acc_mode, int red_mode, double* red_slot);

// y <- scale * (fast form of H) x   [fuse: + Taylor epilogue: acc_mode (kAcc*) into psi, *red_slot <- reduction
// (kRed*) of this shard]
static int launch_hpsi_core(qsb_ctx* c, const double2* x, double2* y, const HParams& hp, bool fuse, double2 scale,
                            int acc_mode, int red_mode, double* red_slot) {
    const bool fuse_here = fuse;
    const double2* peers[kMaxPeers] = {nullptr, nullptr, nullptr};
    if (c->p > kMaxPeers) return fail(c, QSB_ERR_UNSUPPORTED, "more than %d sharded qubits", kMaxPeers);
    nvtx_push("qsb:hpsi");
    for (int g = 0; g < c->p; ++g) {
        const int partner = c->rank ^ (1 << g);
        if (c->ipc) {
            peers[g] = peer_of(c, x, partner);
            if (!peers[g]) return fail(c, QSB_ERR_STATE, "H psi input is not an exported buffer");
        } else {
            NC(nccl_api()->GroupStart());
            NC(nccl_api()->Send(x, c->dim * 2, kNcclFloat64, partner, c->comm, c->stream));
            NC(nccl_api()->Recv(c->stage[g], c->dim * 2, kNcclFloat64, partner, c->comm, c->stream));
            NC(nccl_api()->GroupEnd());
            TRY(nccl_async_check(c));
            peers[g] = c->stage[g];
        }
    }
    c->met.hpsi_calls++;
    int rc = QSB_OK;
    if (!c->plan.tiled) {
        HpsiFlat a{};
        a.x = x;
        a.y = y;
        a.eint = c->eint;
        a.acc = (fuse_here && acc_mode != kAccNone) ? c->psi : nullptr;
        a.norm_part = (fuse_here && red_slot) ? c->part : nullptr;
        for (int g = 0; g < kMaxPeers; ++g) a.peer[g] = peers[g];
        a.n_peer = c->p;
        a.sp = hp.sp;
        a.scale = scale;
        a.nl = c->nl;
        a.fuse = fuse_here ? 1 : 0;
        a.acc_mode = fuse_here ? acc_mode : kAccNone;
        a.red_mode = red_mode;
        const int grid = grid_for(c, c->dim, 256);
        hpsi_flat_kernel<<<grid, 256, 0, c->stream>>>(a);
        LAUNCHED(c);
        rc = check_launch(c, "hpsi_flat_kernel");
        if (rc == QSB_OK && fuse_here && red_slot) {
            finalize_sum_kernel<<<1, 256, 0, c->stream>>>(c->part, grid, red_slot);
            LAUNCHED(c);
            rc = check_launch(c, "finalize_sum_kernel");
        }
    } else if (c->path == 0 && c->nl > kTileBits) {
        rc = launch_hpsi_ws(c, x, y, hp, fuse_here, scale, acc_mode, red_mode, red_slot, peers);
    } else {
        const u64 n_tiles = c->dim >> kTileBits;
        const int grid = (int)std::min<u64>(n_tiles, (u64)c->n_sm);
        for (int i = 0; i < c->plan.n_pass && rc == QSB_OK; ++i) {
            const bool first = (i == 0), last = (i == c->plan.n_pass - 1);
            HpsiPass a{};
            a.x = x;
            a.y = y;
            a.eint = c->eint;
            a.acc_mode = (fuse_here && last) ? acc_mode : kAccNone;
            a.red_mode = red_mode;
            a.acc = (a.acc_mode != kAccNone) ? c->psi : nullptr;
            a.norm_part = (fuse_here && last && red_slot) ? c->part : nullptr;
            for (int g = 0; g < kMaxPeers; ++g) a.peer[g] = peers[g];
            a.n_peer = first ? c->p : 0;
            a.sp = hp.sp;
            a.scale = fuse_here ? scale : make_double2(1.0, 0.0);
            a.nl = c->nl;
            a.lb = c->plan.pass[i].lb;
            a.hb = c->plan.pass[i].hb;
            a.hb0 = c->plan.pass[i].hb0;
            a.n_tiles = n_tiles;
            // LAST without fuse is a plain pass: scale = 1, no acc, no reduction
            if (first && last) hpsi_tile_kernel<true, true><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a);
            else if (first) hpsi_tile_kernel<true, false><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a);
            else if (last) hpsi_tile_kernel<false, true><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a);
            else hpsi_tile_kernel<false, false><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a);
            LAUNCHED(c);
            rc = check_launch(c, "hpsi_tile_kernel");
        }
        if (rc == QSB_OK && fuse_here && red_slot) {
            finalize_sum_kernel<<<1, 256, 0, c->stream>>>(c->part, grid, red_slot);
            LAUNCHED(c);
            rc = check_launch(c, "finalize_sum_kernel");
        }
    }
    nvtx_pop();
    return rc;
}

// y <- scale * H x with H = fast form (E, per-site X and n) + operator terms, optionally followed by the
// Taylor epilogue.  Without operator terms everything is ONE launch group with the epilogue fused in.
static int launch_hpsi(qsb_ctx* c, const double2* x, double2* y, const HParams& hp, bool fuse, double2 scale,
                       int acc_mode, int red_mode, double* red_slot) {
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    if (c->n_gen == 0) return launch_hpsi_core(c, x, y, hp, fuse, scale, acc_mode, red_mode, red_slot);
    TRY(launch_hpsi_core(c, x, y, hp, false, make_double2(1.0, 0.0), kAccNone, kRedNorm, nullptr));
    TRY(launch_terms(c, x, y, hp));
    if (fuse) TRY(launch_epilogue(c, x, y, scale, acc_mode, red_mode, red_slot));
    return QSB_OK;
}

static int launch_terms(qsb_ctx* c, const double2* x, double2* y, const HParams& hp) {
    TermsArgs a{};
    a.terms = c->gen_dev;
    a.n_terms = c->n_gen;
    for (int r = 0; r < c->nranks && r < 8; ++r) a.xsrc[r] = (c->nranks == 1) ? x : peer_of(c, x, r);
    a.xsrc[c->rank] = x;
    a.y = y;
    for (int i = 0; i < 3; ++i) a.chan[i] = hp.chan[i];
    a.nl = c->nl;
    terms_kernel<<<grid_for(c, c->dim, 256), 256, 0, c->stream>>>(a);
    LAUNCHED(c);
    return check_launch(c, "terms_kernel");
}

static int launch_epilogue(qsb_ctx* c, const double2* x, double2* y, double2 scale, int acc_mode, int red_mode,
                           double* red_slot) {
    const int grid = grid_for(c, c->dim, 256);
    epilogue_kernel<<<grid, 256, 0, c->stream>>>(x, y, c->psi, scale, acc_mode, red_mode, c->dim, c->part);
    LAUNCHED(c);
    TRY(check_launch(c, "epilogue_kernel"));
    if (red_slot) {
        finalize_sum_kernel<<<1, 256, 0, c->stream>>>(c->part, grid, red_slot);
        LAUNCHED(c);
        TRY(check_launch(c, "finalize_sum_kernel"));
    }
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// Taylor stepper (fused K4): psi <- sum_j (-i H dt)^j / j! psi, norm-terminated, sub-stepped
// ------------------------------------------------------------------------------------------------
static double spectral_bound(const qsb_ctx* c, const HParams& hp) {
    const double hi = c->diag_max + hp.d_neg;
    const double lo = c->diag_min - hp.d_pos;
    return std::max(fabs(hi), fabs(lo)) + hp.hw_abs + c->gen_bound[0] + fabs(hp.chan[1]) * c->gen_bound[1] +
           fabs(hp.chan[2]) * c->gen_bound[2];
}

static int axpy_vec(qsb_ctx* c, double2* acc, const double2* x) {
    axpy_kernel<<<grid_for(c, c->dim, 256), 256, 0, c->stream>>>(acc, x, c->dim);
    LAUNCHED(c);
    return check_launch(c, "axpy_kernel");
}

// One term x_j = (-i h / j) H x_{j-1} of the series, accumulated into psi as acc_mode says.
static int taylor_term(qsb_ctx* c, const HParams& hp, double h, int j, int acc_mode, const double2** x_io) {
    double2* y = (j & 1) ? c->wa : c->wb;
    const double2 scale = make_double2(0.0, -h / (double)j);  // -i h / j
    TRY(launch_hpsi(c, *x_io, y, hp, true, scale, acc_mode, kRedNorm, c->slots + j));
    // the all-reduce doubles as the cross-rank barrier that orders peer reads of y
    TRY(allreduce(c, c->slots + j, 1, kNcclSum));
    *x_io = y;
    return QSB_OK;
}

static int evolve_segment_impl(qsb_ctx* c, const HParams& hp, double dt, int* n_hpsi) {
    const double theta = spectral_bound(c, hp) * fabs(dt);
    const int nsub = std::max(1, (int)ceil(theta / c->theta_max));
    const double h = dt / nsub;
    const double theta_sub = theta / nsub;
    c->met.segments++;
    c->met.last_theta = theta;
    int applied = 0;
    if (dt == 0.0) {
        if (n_hpsi) *n_hpsi = 0;
        return QSB_OK;
    }
    nvtx_push("qsb:segment");
    // term 1 reads the partners' psi over NVLink: whatever wrote psi before this segment (qsb_set_state, a
    // checkpoint load, the axpy of a one-term step) must be complete on EVERY rank, not just on this one
    TRY(rank_barrier(c));
    for (int sub = 0; sub < nsub; ++sub) {
        c->met.substeps++;
        // Term 1 reads psi directly.  Terms are accumulated into psi in PAIRS -- the even term adds its
        // input x_{j-1}, which sits in the tile it has just loaded, together with its output: ONE
        // read-modify-write of psi per two terms -- so no term updates psi while it is an input, and
        // neither the untiled kernel nor peers reading x over NVLink need a copy of psi.
        const double2* x = c->psi;
        // Launch the number of terms the previous sub-step needed WITHOUT looking at the norms (no host
        // round trip per term); a term beyond convergence is a legitimate (tiny) term of the series.
        const int planned = c->fixed_terms > 0 ? c->fixed_terms : std::min(c->max_terms, std::max(1, c->m_prev));
        for (int j = 1; j <= planned; ++j) {
            int acc_mode = kAccPair;                     // even: psi += x_{j-1} + x_j
            if (j & 1) acc_mode = (j < planned || j == 1) ? kAccNone : kAccTerm;  // odd: deferred, or the last one
            TRY(taylor_term(c, hp, h, j, acc_mode, &x));
            ++applied;
        }
        if (planned == 1) {  // x_0 IS psi: term 1 could not update it in place
            TRY(axpy_vec(c, c->psi, x));
            TRY(rank_barrier(c));  // a partner's next term must not read psi before this update
        }
        int m = planned;
        if (c->fixed_terms == 0) {
            TRY(fetch_slots(c, 1, (size_t)planned));  // ONE synchronisation per sub-step
            int conv = 0;
            for (int j = 1; j <= planned && conv == 0; ++j) {
                const double nrm = sqrt(c->h_slots[j]);
                if (!(nrm == nrm)) return fail(c, QSB_ERR_NOCONV, "Taylor term %d is NaN (|H|dt bound %.3g)", j, theta_sub);
                if (nrm <= c->tol && (double)j >= theta_sub) conv = j;
            }
            // not there yet: one more term at a time (psi is complete after every term from here on)
            for (int j = planned + 1; conv == 0 && j <= c->max_terms; ++j) {
                TRY(taylor_term(c, hp, h, j, kAccTerm, &x));
                ++applied;
                m = j;
                TRY(fetch_slots(c, (size_t)j, 1));
                const double nrm = sqrt(c->h_slots[j]);
                if (!(nrm == nrm)) return fail(c, QSB_ERR_NOCONV, "Taylor term %d is NaN (|H|dt bound %.3g)", j, theta_sub);
                if (nrm <= c->tol && (double)j >= theta_sub) conv = j;
            }
            if (conv == 0)
                return fail(c, QSB_ERR_NOCONV, "Taylor series did not reach tol=%.1e in %d terms (|H|dt bound %.3g)",
                            c->tol, c->max_terms, theta_sub);
            c->m_prev = conv;
        }
        c->met.last_terms = m;
    }
    nvtx_pop();
    if (n_hpsi) *n_hpsi = applied;
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// launch-bound sizes: the whole schedule in one single-CTA kernel (small.cuh)
// ------------------------------------------------------------------------------------------------
static bool small_schedule_ok(const qsb_ctx* c, const qsb_eops* eops) {
    if (!c->small_kernel || c->nranks != 1 || c->nl > kSmallMaxQubits || c->n_gen > 0) return false;
    if (eops)
        for (int o = 0; o < eops->n_ops; ++o)
            if (eops->ops[o].kind != QSB_EOP_ENERGY) return false;
    return true;
}

// site_stride: 0 = omega / delta hold one value per step, n = one row of n per-site values per step
static int evolve_schedule_small(qsb_ctx* c, const double* omega, const double* delta, const double* dt_us,
                                 int64_t n_steps, int site_stride, const qsb_eops* eops, double* expectations_out) {
    const int n_ops = eops ? eops->n_ops : 0;
    const size_t row = site_stride ? (size_t)site_stride : 1;
    // one device buffer: omega | delta | dt | eop_omega | eop_delta | w_omega | w_delta | x0 | expectations | info
    const size_t n_od = (size_t)n_steps * row, n_sched = 2 * n_od + (size_t)n_steps, n_eop = 2 * (size_t)n_ops,
                 n_site = 3 * (size_t)c->n, n_exp = (size_t)n_steps * n_ops, n_info = 8;
    const size_t need = n_sched + n_eop + n_site + n_exp + n_info;
    if (need > c->sched_cap) {
        cudaFree(c->sched_dev);
        c->sched_dev = nullptr;
        c->sched_cap = 0;
        CU(cudaMalloc(&c->sched_dev, need * sizeof(double)));
        c->sched_cap = need;
    }
    std::vector<double> host(n_sched + n_eop + n_site);
    memcpy(host.data(), omega, n_od * sizeof(double));
    memcpy(host.data() + n_od, delta, n_od * sizeof(double));
    memcpy(host.data() + 2 * n_od, dt_us, n_steps * sizeof(double));
    for (int o = 0; o < n_ops; ++o) {
        host[n_sched + o] = eops->ops[o].omega;
        host[n_sched + n_ops + o] = eops->ops[o].delta;
    }
    for (int q = 0; q < c->n; ++q) {
        host[n_sched + n_eop + q] = c->w_omega[q] + (site_stride ? 0.0 : c->t_omega[q]);
        host[n_sched + n_eop + c->n + q] = c->w_delta[q] + (site_stride ? 0.0 : c->t_delta[q]);
        host[n_sched + n_eop + 2 * c->n + q] = c->x0[q];
    }
    CU(cudaMemcpyAsync(c->sched_dev, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SmallArgs a{};
    a.psi = c->psi;
    a.eint = c->eint;
    a.omega = c->sched_dev;
    a.delta = c->sched_dev + n_od;
    a.dt = c->sched_dev + 2 * n_od;
    a.site_stride = site_stride;
    a.n_steps = n_steps;
    a.n = c->nl;
    a.tol = c->tol;
    a.theta_max = c->theta_max;
    a.max_terms = c->max_terms;
    a.fixed_terms = c->fixed_terms;
    a.diag_min = c->diag_min;
    a.diag_max = c->diag_max;
    a.n_eops = n_ops;
    a.eop_omega = c->sched_dev + n_sched;
    a.eop_delta = c->sched_dev + n_sched + n_ops;
    a.w_omega = c->sched_dev + n_sched + n_eop;
    a.w_delta = a.w_omega + c->n;
    a.x0 = a.w_delta + c->n;
    a.expect_out = c->sched_dev + n_sched + n_eop + n_site;
    a.info = reinterpret_cast<unsigned long long*>(c->sched_dev + n_sched + n_eop + n_site + n_exp);
    const size_t smem = 3 * sizeof(double2) << c->nl;
    evolve_small_kernel<<<1, kSmallThreads, smem, c->stream>>>(a);
    LAUNCHED(c);
    TRY(check_launch(c, "evolve_small_kernel"));
    std::vector<double> back(n_exp + n_info);
    CU(cudaMemcpyAsync(back.data(), a.expect_out, back.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long info[8];
    memcpy(info, back.data() + n_exp, sizeof(info));
    c->met.hpsi_calls += info[0];
    c->met.substeps += info[1];
    c->met.segments += (uint64_t)n_steps;
    c->met.last_terms = (int32_t)info[2];
    c->m_prev = (int)info[2];
    if (info[3] == 2) return fail(c, QSB_ERR_NOCONV, "Taylor term is NaN at step %llu", info[4]);
    if (info[3] == 1)
        return fail(c, QSB_ERR_NOCONV, "Taylor series did not reach tol=%.1e in %d terms at step %llu", c->tol, c->max_terms, info[4]);
    if (n_exp) memcpy(expectations_out, back.data(), n_exp * sizeof(double));
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// observables
// ------------------------------------------------------------------------------------------------
struct TermResult {
    double re, im;
};

// evaluate <psi|term_t|psi> (coef excluded) for a batch of Pauli terms into host results
static int pauli_batch(qsb_ctx* c, int64_t n_terms, const u64* xm, const u64* ym, const u64* zm, const u64* nm,
                       TermResult* out, const int* modes = nullptr) {
    const u64 lmask = c->dim - 1ull;
    const size_t cap = (c->slots_n - 2) / 2;
    const int grid = grid_for(c, c->dim, 256);
    std::vector<double> sign;
    u64 staged_flip = 0;  // sharded flip whose partner shard currently sits in stage[0] (no-IPC mode)
    for (int64_t t0 = 0; t0 < n_terms; t0 += (int64_t)cap) {
        const int64_t nb = std::min<int64_t>((int64_t)cap, n_terms - t0);
        sign.assign((size_t)nb, 0.0);  // 0: this shard contributes nothing to the term
        CU(cudaMemsetAsync(c->slots, 0, (size_t)nb * 2 * sizeof(double), c->stream));
        for (int64_t t = 0; t < nb; ++t) {
            const u64 x = xm[t0 + t], y = ym[t0 + t], z = zm[t0 + t], nmk = nm[t0 + t];
            if ((x & y) | (x & z) | (x & nmk) | (y & z) | (y & nmk) | (z & nmk))
                return fail(c, QSB_ERR_INVALID, "Pauli term %lld: masks overlap", (long long)(t0 + t));
            if ((x | y | z | nmk) >> c->n)
                return fail(c, QSB_ERR_INVALID, "Pauli term %lld: mask beyond qubit count", (long long)(t0 + t));
            const int mode = modes ? modes[t0 + t] : 0;  // 1: weight 1 - sgn with sgn over the flipped bits
            const u64 flip = x | y, zz = mode ? flip : (y | z);
            const u64 rank = (u64)c->rank;
            const u64 flip_g = flip >> c->nl, zz_g = zz >> c->nl, n_g = nmk >> c->nl;
            const double2* bra = c->psi;
            if (flip_g) {
                if (c->ipc) {
                    bra = c->peer_psi[rank ^ flip_g];
                } else {
                    // no peer mapping: stage the partner shard with ncclSend/Recv, as H psi does.  Every
                    // rank takes part (before the N-projector test below, which is rank dependent); the
                    // staged shard is reused by consecutive terms with the same sharded flip.
                    if (staged_flip != flip_g) {
                        const int partner = (int)(rank ^ flip_g);
                        NC(nccl_api()->GroupStart());
                        NC(nccl_api()->Send(c->psi, c->dim * 2, kNcclFloat64, partner, c->comm, c->stream));
                        NC(nccl_api()->Recv(c->stage[0], c->dim * 2, kNcclFloat64, partner, c->comm, c->stream));
                        NC(nccl_api()->GroupEnd());
                        staged_flip = flip_g;
                    }
                    bra = c->stage[0];
                }
            }
            if ((rank & n_g) != n_g) continue;  // this shard has a 0 where N projects on 1
            PauliArgs a{};
            a.b = c->psi;
            a.a = bra;
            a.flip = flip & lmask;
            a.zmask = zz & lmask;
            a.nmask = nmk & lmask;
            a.mode = mode;  // mode 1 is only used with shard-local masks (see qsb_sisj)
            a.nl = c->nl;
            a.part = c->part;
            sign[(size_t)t] = (__builtin_popcountll(rank & zz_g) & 1) ? -1.0 : 1.0;  // sharded Z / Y bits
            if (mode == 1 && __builtin_popcountll(a.flip) == 2 && c->nl >= 2) {
                // XX + YY of a local pair: the quarter-space pair kernel (each contributing amplitude read once)
                const int bi = __builtin_ctzll(a.flip), bj = 63 - __builtin_clzll(a.flip);
                xxyy_pair_kernel<<<std::min(grid, grid_for(c, c->dim >> 2, 256)), 256, 0, c->stream>>>(c->psi, c->nl, bi, bj, c->part);
                LAUNCHED(c);
                TRY(check_launch(c, "xxyy_pair_kernel"));
                finalize_pair_kernel<<<1, 256, 0, c->stream>>>(c->part, std::min(grid, grid_for(c, c->dim >> 2, 256)), c->slots + 2 * t);
                LAUNCHED(c);
                TRY(check_launch(c, "finalize_pair_kernel"));
                continue;
            }
            pauli_kernel<<<grid, 256, 0, c->stream>>>(a);
            LAUNCHED(c);
            TRY(check_launch(c, "pauli_kernel"));
            finalize_pair_kernel<<<1, 256, 0, c->stream>>>(c->part, grid, c->slots + 2 * t);
            LAUNCHED(c);
            TRY(check_launch(c, "finalize_pair_kernel"));
        }
        TRY(fetch_slots(c, 0, (size_t)nb * 2));
        for (int64_t t = 0; t < nb; ++t) {
            c->h_slots[2 * t] *= sign[(size_t)t];
            c->h_slots[2 * t + 1] *= sign[(size_t)t];
        }
        if (c->nranks > 1) {
            CU(cudaMemcpyAsync(c->slots, c->h_slots, (size_t)nb * 2 * sizeof(double), cudaMemcpyHostToDevice,
                               c->stream));
            TRY(allreduce(c, c->slots, (size_t)nb * 2, kNcclSum));
            TRY(fetch_slots(c, 0, (size_t)nb * 2));
        }
        for (int64_t t = 0; t < nb; ++t) {
            const int ny = __builtin_popcountll(ym[t0 + t]) & 3;  // global factor i^ny of the Y's
            const double re = c->h_slots[2 * t], im = c->h_slots[2 * t + 1];
            TermResult r;
            switch (ny) {
                case 0: r = {re, im}; break;
                case 1: r = {-im, re}; break;
                case 2: r = {-re, -im}; break;
                default: r = {im, -re}; break;
            }
            out[t0 + t] = r;
        }
    }
    return QSB_OK;
}

// probabilities' marginals: tot = <psi|psi>, cnt[q] = <n_q> for global qubit q (whole register)
// cond_qubit >= 0: restrict to basis states with that qubit excited -> tot = <n_c>, number[q] = <n_c n_q>
static int marginals(qsb_ctx* c, double* tot, double* number /* n */, int cond_qubit = -1) {
    const int grid = std::min(grid_for(c, c->dim, 256), c->n_sm * 2);
    const int stride = c->nl + 1;
    if ((size_t)grid * stride > c->part_n) return fail(c, QSB_ERR_STATE, "partials buffer too small");
    u64 cond = 0;
    bool shard_in = true;  // a sharded condition qubit selects whole shards
    if (cond_qubit >= 0) {
        const int bit = c->n - 1 - cond_qubit;
        if (bit < c->nl) cond = 1ull << bit;
        else shard_in = ((c->rank >> (bit - c->nl)) & 1) != 0;
    }
    marginals_kernel<<<grid, 256, 0, c->stream>>>(c->psi, c->nl, cond, c->part);
    LAUNCHED(c);
    TRY(check_launch(c, "marginals_kernel"));
    finalize_rows_kernel<<<stride, 256, 0, c->stream>>>(c->part, grid, stride, c->slots);
    LAUNCHED(c);
    TRY(check_launch(c, "finalize_rows_kernel"));
    TRY(fetch_slots(c, 0, (size_t)stride));
    // lay out per global qubit: qubit q <-> bit n-1-q; bits >= nl are rank bits
    std::vector<double> v(c->n + 1, 0.0);
    if (!shard_in)
        for (int i = 0; i <= c->nl; ++i) c->h_slots[i] = 0.0;
    v[0] = c->h_slots[0];
    for (int q = 0; q < c->n; ++q) {
        const int bit = c->n - 1 - q;
        if (bit < c->nl) v[1 + q] = c->h_slots[1 + bit];
        else v[1 + q] = ((c->rank >> (bit - c->nl)) & 1) ? c->h_slots[0] : 0.0;
    }
    if (c->nranks > 1) {
        CU(cudaMemcpyAsync(c->slots, v.data(), (size_t)(c->n + 1) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(allreduce(c, c->slots, (size_t)(c->n + 1), kNcclSum));
        TRY(fetch_slots(c, 0, (size_t)(c->n + 1)));
        for (int i = 0; i <= c->n; ++i) v[i] = c->h_slots[i];
    }
    if (tot) *tot = v[0];
    if (number)
        for (int q = 0; q < c->n; ++q) number[q] = v[1 + q];
    return QSB_OK;
}

// <psi|H|psi>: ONE H psi launch group whose last pass reduces Re conj(psi) (H psi) (no dot kernel)
static int energy_impl(qsb_ctx* c, const HParams& hp, double* out) {
    TRY(rank_barrier(c));
    TRY(launch_hpsi(c, c->psi, c->wa, hp, true, make_double2(1.0, 0.0), kAccNone, kRedDot, c->slots));
    TRY(allreduce(c, c->slots, 1, kNcclSum));
    TRY(fetch_slots(c, 0, 1));
    *out = c->h_slots[0];
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// sharding set-up: NCCL communicator + CUDA-IPC mapping of every peer's psi / wa / wb
// ------------------------------------------------------------------------------------------------
static int setup_comm(qsb_ctx* c, const void* uid_bytes) {
    NcclApi* api = nccl_api();
    if (!api->ok) return fail(c, QSB_ERR_NCCL, "%s", api->why);
    nccl_uid_t uid;
    memcpy(&uid, uid_bytes, sizeof(uid));
    NC(api->CommInitRank(&c->comm, c->nranks, uid, c->rank));

    // exchange IPC handles of the state-sized buffers (psi, wa, wb and, from 4 ranks up, the Z buffer
    // of the remap exchange)
    const int nbuf = c->nranks >= 4 ? 4 : 3;
    if (nbuf == 4) CU(cudaMalloc(&c->zbuf, c->dim * sizeof(double2)));
    const size_t hsz = sizeof(cudaIpcMemHandle_t);
    const size_t rec = (size_t)nbuf * hsz;
    std::vector<unsigned char> mine(rec), all(rec * c->nranks);
    bool ok = getenv("QSB_NO_IPC") == nullptr;
    cudaIpcMemHandle_t h[4];
    double2* bufs[4] = {c->psi, c->wa, c->wb, c->zbuf};
    for (int i = 0; i < nbuf && ok; ++i) {
        if (cudaIpcGetMemHandle(&h[i], bufs[i]) != cudaSuccess) {
            ok = false;
            cudaGetLastError();
        } else {
            memcpy(mine.data() + i * hsz, &h[i], hsz);
        }
    }
    unsigned char *d_mine = nullptr, *d_all = nullptr;
    CU(cudaMalloc(&d_mine, rec + 8));
    CU(cudaMalloc(&d_all, (rec + 8) * c->nranks));
    // record = handles + an "ok" flag so that every rank takes the same decision
    std::vector<unsigned char> rec_mine(rec + 8, 0), rec_all((rec + 8) * c->nranks, 0);
    memcpy(rec_mine.data(), mine.data(), rec);
    rec_mine[rec] = ok ? 1 : 0;
    CU(cudaMemcpyAsync(d_mine, rec_mine.data(), rec + 8, cudaMemcpyHostToDevice, c->stream));
    NC(api->AllGather(d_mine, d_all, rec + 8, kNcclUint8, c->comm, c->stream));
    CU(cudaMemcpyAsync(rec_all.data(), d_all, (rec + 8) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_mine);
    cudaFree(d_all);
    for (int r = 0; r < c->nranks; ++r) ok = ok && rec_all[(rec + 8) * r + rec] == 1;
    if (ok) {
        for (int r = 0; r < c->nranks && ok; ++r) {
            if (r == c->rank) {
                c->peer_psi[r] = c->psi;
                c->peer_wa[r] = c->wa;
                c->peer_wb[r] = c->wb;
                c->peer_z[r] = c->zbuf;
                continue;
            }
            double2** dst[4] = {&c->peer_psi[r], &c->peer_wa[r], &c->peer_wb[r], &c->peer_z[r]};
            for (int i = 0; i < nbuf; ++i) {
                cudaIpcMemHandle_t hh;
                memcpy(&hh, rec_all.data() + (rec + 8) * r + i * hsz, hsz);
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, hh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                *dst[i] = (double2*)ptr;
            }
        }
    }
    // agree on the outcome (an open can fail on one rank only)
    c->h_slots[0] = ok ? 0.0 : 1.0;
    CU(cudaMemcpyAsync(c->slots, c->h_slots, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    TRY(allreduce(c, c->slots, 1, kNcclSum));
    TRY(fetch_slots(c, 0, 1));
    c->ipc = (c->h_slots[0] == 0.0);
    if (c->ipc && c->zbuf) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_z, cudaEventDisableTiming));
    }
    if (!c->ipc) {
        for (int g = 0; g < c->p; ++g) CU(cudaMalloc(&c->stage[g], c->dim * sizeof(double2)));
    }
    return QSB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int qsb_abi_version(void) { return QSB_ABI_VERSION; }

const char* qsb_build_info(void) {
    return "libqsb sm_100a; warp-specialised H psi (TMA producer + 16 consumer warps, 2^12-amplitude tiles, "
           "L2-resident super-blocks, tensor-map TMA); built " __DATE__;
}

int qsb_device_count(int* count) {
    qsb_ctx* c = nullptr;
    if (!count) return fail(c, QSB_ERR_INVALID, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(c, QSB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return QSB_OK;
}

const char* qsb_last_error(const qsb_ctx* ctx) { return ctx ? ctx->err : g_err; }

int qsb_nccl_unique_id(void* out_uid) {
    qsb_ctx* c = nullptr;
    if (!out_uid) return fail(c, QSB_ERR_INVALID, "out_uid is NULL");
    NcclApi* api = nccl_api();
    if (!api->ok) return fail(c, QSB_ERR_NCCL, "%s", api->why);
    nccl_uid_t uid;
    NC(api->GetUniqueId(&uid));
    memcpy(out_uid, &uid, sizeof(uid));
    return QSB_OK;
}

int qsb_create_sharded(int n_qubits, int device, int rank, int nranks, const void* nccl_uid, qsb_ctx** out) {
    qsb_ctx* c = nullptr;
    if (!out) return fail(c, QSB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > QSB_MAX_QUBITS) return fail(c, QSB_ERR_INVALID, "n_qubits=%d out of range", n_qubits);
    if (nranks < 1 || (nranks & (nranks - 1)) || nranks > (1 << kMaxPeers))
        return fail(c, QSB_ERR_INVALID, "nranks=%d must be a power of two <= %d (at most %d sharded qubits)", nranks,
                    1 << kMaxPeers, kMaxPeers);
    if (rank < 0 || rank >= nranks) return fail(c, QSB_ERR_INVALID, "rank=%d out of range", rank);
    int p = 0;
    while ((1 << p) < nranks) ++p;
    if (p > n_qubits - 1 && nranks > 1) return fail(c, QSB_ERR_INVALID, "more ranks than half the basis");
    if (n_qubits - p > QSB_MAX_QUBITS_DEV - 1) return fail(c, QSB_ERR_INVALID, "shard of 2^%d amplitudes is too large", n_qubits - p);
    if (nranks > 1 && !nccl_uid) return fail(c, QSB_ERR_INVALID, "nccl_uid is NULL");

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(c, QSB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    c = new qsb_ctx();
    memset(&c->met, 0, sizeof(c->met));
    c->n = n_qubits;
    c->p = p;
    c->nl = n_qubits - p;
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    c->dim = 1ull << c->nl;
    c->rank_bits = (u64)rank << c->nl;
    c->pop_rank = __builtin_popcount((unsigned)rank);
    for (int q = 0; q < QSB_MAX_QUBITS_DEV + 4; ++q) {
        c->w_omega[q] = 1.0;
        c->w_delta[q] = 1.0;
        c->x0[q] = c->t_omega[q] = c->t_delta[q] = 0.0;
    }
    int rc = QSB_OK;
    do {
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { rc = fail(c, QSB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); break; }
        if (prop.major < 10) { rc = fail(c, QSB_ERR_UNSUPPORTED, "device %d is sm_%d%d; libqsb is built for sm_100a only", device, prop.major, prop.minor); break; }
        c->n_sm = prop.multiProcessorCount;
#define CUB(call) if ((e = (call)) != cudaSuccess) { rc = fail(c, QSB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e)); break; }
        CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUB(cudaEventCreate(&c->ev0));
        CUB(cudaEventCreate(&c->ev1));
        CUB(cudaMalloc(&c->psi, c->dim * sizeof(double2)));
        CUB(cudaMalloc(&c->wa, c->dim * sizeof(double2)));
        CUB(cudaMalloc(&c->wb, c->dim * sizeof(double2)));
        CUB(cudaMalloc(&c->eint, c->dim * sizeof(double)));
        CUB(cudaMalloc(&c->vtab, (size_t)n_qubits * n_qubits * sizeof(double)));
        c->part_n = (size_t)c->n_sm * 8 * (QSB_MAX_QUBITS_DEV + 2);
        CUB(cudaMalloc(&c->part, c->part_n * sizeof(double)));
        c->slots_n = 8192;
        CUB(cudaMalloc(&c->slots, c->slots_n * sizeof(double)));
        CUB(cudaMemset(c->slots, 0, c->slots_n * sizeof(double)));
        CUB(cudaMallocHost(&c->h_slots, c->slots_n * sizeof(double)));
        CUB(cudaFuncSetAttribute(hpsi_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<11, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<11, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<12, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<12, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<12, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaFuncSetAttribute(hpsi_ws_kernel<12, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes));
        CUB(cudaMalloc(&c->tickets, 17 * sizeof(unsigned int)));
        CUB(cudaMemset(c->tickets, 0, 17 * sizeof(unsigned int)));
        c->tail = c->tickets + 16;
        CUB(cudaFuncSetAttribute(evolve_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (int)sizeof(double2) << kSmallMaxQubits));
        {
            const size_t n_done = (size_t)(c->dim >> 13) + 1;
            CUB(cudaMalloc(&c->sb_done, n_done * sizeof(unsigned long long)));
            CUB(cudaMemset(c->sb_done, 0, n_done * sizeof(unsigned long long)));
        }
#undef CUB
        c->plan = make_plan(c->nl, false);
        if (nranks > 1) rc = setup_comm(c, nccl_uid);
    } while (0);
    if (rc != QSB_OK) {
        snprintf(g_err, sizeof(g_err), "%s", c->err);
        qsb_destroy(c);
        return rc;
    }
    rc = qsb_reset_state(c);
    if (rc != QSB_OK) {
        qsb_destroy(c);
        return rc;
    }
    *out = c;
    return QSB_OK;
}

int qsb_create(int n_qubits, int device, qsb_ctx** out) {
    return qsb_create_sharded(n_qubits, device, 0, 1, nullptr, out);
}

int qsb_destroy(qsb_ctx* c) {
    if (!c) return QSB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->ipc) {
        for (int r = 0; r < c->nranks; ++r) {
            if (r == c->rank) continue;
            if (c->peer_psi[r]) cudaIpcCloseMemHandle(c->peer_psi[r]);
            if (c->peer_wa[r]) cudaIpcCloseMemHandle(c->peer_wa[r]);
            if (c->peer_wb[r]) cudaIpcCloseMemHandle(c->peer_wb[r]);
            if (c->peer_z[r]) cudaIpcCloseMemHandle(c->peer_z[r]);
        }
    }
    if (c->comm && nccl_api()->ok) nccl_api()->CommDestroy(c->comm);
    for (int g = 0; g < kMaxPeers; ++g) cudaFree(c->stage[g]);
    cudaFree(c->zbuf);
    cudaFree(c->psi);
    cudaFree(c->wa);
    cudaFree(c->wb);
    cudaFree(c->eint);
    cudaFree(c->vtab);
    cudaFree(c->part);
    cudaFree(c->slots);
    cudaFree(c->flush_buf);
    cudaFree(c->topk_buf);
    cudaFree(c->tickets);
    cudaFree(c->gen_dev);
    cudaFree(c->vbits_dev);
    cudaFree(c->tile_tab[0]);
    cudaFree(c->tile_tab[1]);
    cudaFree(c->tab_zero);
    cudaFree(c->diag_dev);
    cudaFree(c->sched_dev);
    cudaFree(c->sb_done);
    if (c->h_slots) cudaFreeHost(c->h_slots);
    if (c->stream2) {
        cudaStreamSynchronize(c->stream2);
        cudaStreamDestroy(c->stream2);
    }
    if (c->ev_x) cudaEventDestroy(c->ev_x);
    if (c->ev_z) cudaEventDestroy(c->ev_z);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return QSB_OK;
}

int qsb_host_alloc(uint64_t bytes, void** out) {
    qsb_ctx* c = nullptr;
    if (!out || bytes == 0) return fail(c, QSB_ERR_INVALID, "bad argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return QSB_OK;
}

int qsb_host_free(void* ptr) {
    qsb_ctx* c = nullptr;
    if (!ptr) return QSB_OK;
    CU(cudaFreeHost(ptr));
    return QSB_OK;
}

int qsb_local_dim(const qsb_ctx* ctx, uint64_t* dim) {
    if (!ctx || !dim) return fail(nullptr, QSB_ERR_INVALID, "NULL argument");
    *dim = ctx->dim;
    return QSB_OK;
}

int qsb_set_option(qsb_ctx* c, const char* key, double value) {
    if (!c || !key) return fail(c, QSB_ERR_INVALID, "NULL argument");
    if (!strcmp(key, "tol")) {
        if (!(value > 0.0)) return fail(c, QSB_ERR_INVALID, "tol must be positive");
        c->tol = value;
    } else if (!strcmp(key, "theta_max")) {
        if (!(value > 0.0)) return fail(c, QSB_ERR_INVALID, "theta_max must be positive");
        c->theta_max = value;
    } else if (!strcmp(key, "max_terms")) {
        if (value < 1 || value > 4000) return fail(c, QSB_ERR_INVALID, "max_terms out of range");
        c->max_terms = (int)value;
    } else if (!strcmp(key, "fixed_terms")) {
        if (value < 0 || value > 4000) return fail(c, QSB_ERR_INVALID, "fixed_terms out of range");
        c->fixed_terms = (int)value;
    } else if (!strcmp(key, "hpsi_path")) {
        if (value != 0.0 && value != 1.0 && value != 2.0) return fail(c, QSB_ERR_INVALID, "hpsi_path must be 0, 1 or 2");
        c->path = (int)value;
        c->force_flat = c->path == 1;
        c->plan = make_plan(c->nl, c->force_flat);
    } else if (!strcmp(key, "exchange")) {
        if (value != -1.0 && value != 0.0 && value != 1.0) return fail(c, QSB_ERR_INVALID, "exchange must be -1, 0 or 1");
        c->exchange = (int)value;
    } else if (!strcmp(key, "exchange_split")) {
        if (value != -1.0 && value != 0.0 && value != 1.0 && value != 2.0) return fail(c, QSB_ERR_INVALID, "exchange_split must be -1, 0, 1 or 2");
        c->exchange_split = (int)value;
    } else if (!strcmp(key, "remap_ctas")) {
        if (value < 0 || value > 128) return fail(c, QSB_ERR_INVALID, "remap_ctas must be in 0..128");
        c->remap_ctas = (int)value;
    } else if (!strcmp(key, "ticket_batch")) {
        if (value != 0.0 && value != 1.0) return fail(c, QSB_ERR_INVALID, "ticket_batch: only 1 (or 0 = auto) is supported; batches were measured slower");
    } else if (!strcmp(key, "diag_on_the_fly")) {
        if (value != -1.0 && value != 0.0 && value != 1.0) return fail(c, QSB_ERR_INVALID, "diag_on_the_fly must be -1, 0 or 1");
        c->diag_otf = (int)value;
    } else if (!strcmp(key, "tiled_observables")) {
        c->obs_tiled = value != 0.0;
    } else if (!strcmp(key, "small_kernel")) {
        c->small_kernel = value != 0.0;
    } else if (!strcmp(key, "tensor_map")) {
        c->use_tmap = value != 0.0;
    } else if (!strcmp(key, "tile_bits")) {
        if (value != 0.0 && value != 11.0 && value != 12.0) return fail(c, QSB_ERR_INVALID, "tile_bits must be 0, 11 or 12");
        c->ws_tile_bits = (int)value;
    } else if (!strcmp(key, "sb_lag")) {
        if (value < -1 || value > 64) return fail(c, QSB_ERR_INVALID, "sb_lag must be -1 (no lead), 0 (auto) or in 1..64");
        c->sb_lag = (int)value;
    } else if (!strcmp(key, "sb_bits")) {
        if (value != 0.0 && (value < 13 || value > 21)) return fail(c, QSB_ERR_INVALID, "sb_bits must be 0 (auto) or in 13..21");
        c->sb_bits = (int)value;
    } else {
        return fail(c, QSB_ERR_INVALID, "unknown option '%s'", key);
    }
    if ((size_t)c->max_terms + 2 > c->slots_n) return fail(c, QSB_ERR_INVALID, "max_terms too large");
    return QSB_OK;
}

// The static diagonal as a 2-local polynomial in the occupation bits (Z = 1 - 2 n): what the H psi kernel
// evaluates on the fly instead of reading E.  Strings of three or more factors do not fit (dm_ok = false).
static int build_diag_model(qsb_ctx* c) {
    const int n = c->n;
    c->dm_c0 = 0.0;
    c->dm_lin.assign((size_t)n, 0.0);
    c->dm_pair.assign((size_t)n * n, 0.0);
    c->dm_ok = true;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) c->dm_pair[(size_t)i * n + j] = c->v_host[(size_t)i * n + j];
    for (const DiagTerm& g : c->diag_terms) {
        int q[2], isz[2], k = 0;
        bool fits = true;
        for (int bit = 0; bit < n; ++bit) {
            const bool z = (g.z >> bit) & 1ull, nn = (g.nm >> bit) & 1ull;
            if (!z && !nn) continue;
            if (k == 2) {
                fits = false;
                break;
            }
            q[k] = n - 1 - bit;
            isz[k] = z ? 1 : 0;
            ++k;
        }
        if (!fits) {
            c->dm_ok = false;
            continue;
        }
        if (k == 0) {
            c->dm_c0 += g.coef;
        } else if (k == 1) {
            if (isz[0]) {
                c->dm_c0 += g.coef;
                c->dm_lin[q[0]] -= 2.0 * g.coef;
            } else {
                c->dm_lin[q[0]] += g.coef;
            }
        } else {
            const int a = std::min(q[0], q[1]), b = std::max(q[0], q[1]);
            double& pr = c->dm_pair[(size_t)a * n + b];
            if (isz[0] && isz[1]) {  // (1 - 2 n_a)(1 - 2 n_b)
                c->dm_c0 += g.coef;
                c->dm_lin[q[0]] -= 2.0 * g.coef;
                c->dm_lin[q[1]] -= 2.0 * g.coef;
                pr += 4.0 * g.coef;
            } else if (!isz[0] && !isz[1]) {
                pr += g.coef;
            } else {  // Z_z n_m = n_m - 2 n_z n_m
                const int m = isz[0] ? q[1] : q[0];
                c->dm_lin[m] += g.coef;
                pr -= 2.0 * g.coef;
            }
        }
    }
    // pair table over the LOCAL index bits (+ an all-zero copy for the stored-E mode), symmetric
    std::vector<double> tab(2 * (size_t)kMaxLocalBits * kMaxLocalBits, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const int bi = n - 1 - i, bj = n - 1 - j;
            if (bi >= c->nl || bj >= c->nl) continue;
            tab[(size_t)bi * kMaxLocalBits + bj] = tab[(size_t)bj * kMaxLocalBits + bi] = c->dm_pair[(size_t)i * n + j];
        }
    for (int i = 0; i < 2; ++i) {  // the per-tile tables follow the pair table: rebuilt on next use
        cudaFree(c->tile_tab[i]);
        c->tile_tab[i] = nullptr;
    }
    if (!c->tab_zero) {
        CU(cudaMalloc(&c->tab_zero, kTileTabDoubles * sizeof(double)));
        CU(cudaMemset(c->tab_zero, 0, kTileTabDoubles * sizeof(double)));
    }
    if (!c->vbits_dev) CU(cudaMalloc(&c->vbits_dev, tab.size() * sizeof(double)));
    CU(cudaMemcpyAsync(c->vbits_dev, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // tab is a local
    return QSB_OK;
}

// E <- interaction diagonal (K1) + static diagonal operator strings; refresh its range
static int rebuild_diagonal(qsb_ctx* c) {
    const int n = c->n;
    TRY(build_diag_model(c));
    CU(cudaMemcpyAsync(c->vtab, c->v_host.data(), c->v_host.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int grid = grid_for(c, c->dim, 256);
    diag_build_kernel<<<grid, 256, (size_t)n * n * sizeof(double), c->stream>>>(c->eint, c->vtab, n, c->nl, c->rank_bits);
    LAUNCHED(c);
    TRY(check_launch(c, "diag_build_kernel"));
    if (!c->diag_terms.empty()) {
        cudaFree(c->diag_dev);
        c->diag_dev = nullptr;
        CU(cudaMalloc(&c->diag_dev, c->diag_terms.size() * sizeof(DiagTerm)));
        CU(cudaMemcpyAsync(c->diag_dev, c->diag_terms.data(), c->diag_terms.size() * sizeof(DiagTerm), cudaMemcpyHostToDevice,
                           c->stream));
        diag_terms_kernel<<<grid, 256, 0, c->stream>>>(c->eint, c->diag_dev, (int)c->diag_terms.size(), c->nl, c->rank_bits);
        LAUNCHED(c);
        TRY(check_launch(c, "diag_terms_kernel"));
    }
    minmax_kernel<<<grid, 256, 0, c->stream>>>(c->eint, c->dim, c->part);
    LAUNCHED(c);
    TRY(check_launch(c, "minmax_kernel"));
    minmax_final_kernel<<<1, 32, 0, c->stream>>>(c->part, grid, c->slots);
    LAUNCHED(c);
    TRY(check_launch(c, "minmax_final_kernel"));
    TRY(fetch_slots(c, 0, 2));
    double lo = c->h_slots[0], hi = c->h_slots[1];
    if (c->nranks > 1) {
        c->h_slots[0] = -lo;
        c->h_slots[1] = hi;
        CU(cudaMemcpyAsync(c->slots, c->h_slots, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(allreduce(c, c->slots, 2, kNcclMax));
        TRY(fetch_slots(c, 0, 2));
        lo = -c->h_slots[0];
        hi = c->h_slots[1];
    }
    c->diag_min = lo;
    c->diag_max = hi;
    c->have_v = true;
    c->m_prev = 0;
    return QSB_OK;
}

int qsb_set_interaction(qsb_ctx* c, const double* v) {
    if (!c || !v) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    const int n = c->n;
    std::vector<double> tab((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const double x = v[(size_t)i * n + j];
            if (!(x == x) || isinf(x)) return fail(c, QSB_ERR_INVALID, "V[%d,%d] is not finite", i, j);
            tab[(size_t)i * n + j] = x;
        }
    c->v_host.swap(tab);
    return rebuild_diagonal(c);
}

int qsb_set_site_weights(qsb_ctx* c, const double* w_omega, const double* w_delta) {
    if (!c) return fail(c, QSB_ERR_INVALID, "NULL context");
    for (int q = 0; q < c->n; ++q) {
        const double a = w_omega ? w_omega[q] : 1.0, b = w_delta ? w_delta[q] : 1.0;
        if (!(a == a) || isinf(a) || !(b == b) || isinf(b)) return fail(c, QSB_ERR_INVALID, "site weight %d is not finite", q);
        c->w_omega[q] = a;
        c->w_delta[q] = b;
    }
    c->m_prev = 0;
    return QSB_OK;
}

int qsb_set_operator_terms(qsb_ctx* c, int64_t n_terms, const double* coef, const uint64_t* xmask, const uint64_t* ymask,
                           const uint64_t* zmask, const uint64_t* nmask, const int32_t* channel) {
    if (!c || n_terms < 0 || (n_terms > 0 && (!coef || !xmask || !ymask || !zmask || !nmask)))
        return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    const u64 lmask = c->dim - 1ull;
    std::vector<GenTerm> gen;
    std::vector<DiagTerm> diag;
    double t_omega[QSB_MAX_QUBITS_DEV + 4] = {0}, t_delta[QSB_MAX_QUBITS_DEV + 4] = {0}, x0[QSB_MAX_QUBITS_DEV + 4] = {0};
    double bound[3] = {0.0, 0.0, 0.0};
    for (int64_t t = 0; t < n_terms; ++t) {
        const u64 x = xmask[t], y = ymask[t], z = zmask[t], nm = nmask[t];
        const int ch = channel ? channel[t] : 0;
        if (ch < 0 || ch > 2) return fail(c, QSB_ERR_INVALID, "operator term %lld: channel must be 0, 1 or 2", (long long)t);
        if ((x & y) | (x & z) | (x & nm) | (y & z) | (y & nm) | (z & nm))
            return fail(c, QSB_ERR_INVALID, "operator term %lld: masks overlap", (long long)t);
        if ((x | y | z | nm) >> c->n) return fail(c, QSB_ERR_INVALID, "operator term %lld: mask beyond qubit count", (long long)t);
        if (!(coef[t] == coef[t]) || isinf(coef[t])) return fail(c, QSB_ERR_INVALID, "operator term %lld: coefficient is not finite", (long long)t);
        if (coef[t] == 0.0) continue;
        const u64 flip = x | y;
        const bool single_x = y == 0 && z == 0 && nm == 0 && __builtin_popcountll(x) == 1;
        const bool single_n = x == 0 && y == 0 && z == 0 && __builtin_popcountll(nm) == 1;
        if (single_x && ch != 2) {  // per-site drive: static, or scaled by the amplitude (H_amp = 0.5 X is coef 0.5)
            const int q = c->n - 1 - (63 - __builtin_clzll(x));
            if (ch == 0) x0[q] += coef[t];
            else t_omega[q] += 2.0 * coef[t];
            continue;
        }
        if (single_n && ch == 2) {  // per-site detuning: H_det = -1.0 N is coef -1
            const int q = c->n - 1 - (63 - __builtin_clzll(nm));
            t_delta[q] += -coef[t];
            continue;
        }
        if (flip == 0 && ch == 0) {  // static diagonal string: folded into E
            diag.push_back(DiagTerm{coef[t], z, nm});
            continue;
        }
        // generic term, per rank: source shard, sign of the sharded Z / Y bits, sharded N projector
        const u64 flip_g = flip >> c->nl, zy_g = (y | z) >> c->nl, n_g = nm >> c->nl;
        bound[ch] += fabs(coef[t]);
        if (flip_g && c->nranks > 1 && !c->ipc)
            return fail(c, QSB_ERR_UNSUPPORTED, "operator term %lld flips a sharded qubit and peer memory is not mapped", (long long)t);
        if (((u64)c->rank & n_g) != n_g) continue;
        const u64 src = (u64)c->rank ^ flip_g;
        GenTerm g{};
        g.coef = (__builtin_popcountll(src & zy_g) & 1) ? -coef[t] : coef[t];
        g.flip = flip & lmask;
        g.zy = (y | z) & lmask;
        g.nm = nm & lmask;
        g.ny = __builtin_popcountll(y) & 3;
        g.chan = ch;
        g.src = (int)src;
        gen.push_back(g);
    }
    cudaFree(c->gen_dev);
    c->gen_dev = nullptr;
    c->n_gen = 0;
    if (!gen.empty()) {
        CU(cudaMalloc(&c->gen_dev, gen.size() * sizeof(GenTerm)));
        CU(cudaMemcpyAsync(c->gen_dev, gen.data(), gen.size() * sizeof(GenTerm), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    // every rank must take the same code path (the terms kernel is not a collective, but the stepper's
    // launch sequence is): the count that matters is "any generic term in the register", not per rank
    c->n_gen = (bound[0] + bound[1] + bound[2] > 0.0) ? std::max<int>(1, (int)gen.size()) : 0;
    if (c->n_gen && gen.empty()) {
        GenTerm zero{};  // this rank holds none of them: a single null term keeps the launch sequence uniform
        zero.src = c->rank;
        CU(cudaMalloc(&c->gen_dev, sizeof(GenTerm)));
        CU(cudaMemcpy(c->gen_dev, &zero, sizeof(GenTerm), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 3; ++i) c->gen_bound[i] = bound[i];
    for (int q = 0; q < c->n; ++q) {
        c->t_omega[q] = t_omega[q];
        c->t_delta[q] = t_delta[q];
        c->x0[q] = x0[q];
    }
    c->diag_terms.swap(diag);
    c->have_terms = n_terms > 0;
    if (c->v_host.empty()) c->v_host.assign((size_t)c->n * c->n, 0.0);
    return rebuild_diagonal(c);
}

int qsb_get_diagonal(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->eint, c->dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_reset_state(qsb_ctx* c) {
    if (!c) return fail(c, QSB_ERR_INVALID, "NULL context");
    CU(cudaSetDevice(c->device));
    const u64 one_at = (c->rank == 0) ? 0ull : ~0ull;
    basis_state_kernel<<<grid_for(c, c->dim, 256), 256, 0, c->stream>>>(c->psi, c->dim, one_at);
    LAUNCHED(c);
    TRY(check_launch(c, "basis_state_kernel"));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_set_state(qsb_ctx* c, const double* re_im) {
    if (!c || !re_im) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(c->psi, re_im, c->dim * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_get_state(qsb_ctx* c, double* re_im) {
    if (!c || !re_im) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(re_im, c->psi, c->dim * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_apply_h(qsb_ctx* c, double omega, double delta) {
    if (!c) return fail(c, QSB_ERR_INVALID, "NULL context");
    CU(cudaSetDevice(c->device));
    TRY(rank_barrier(c));
    TRY(launch_hpsi(c, c->psi, c->wa, hparams_global(c, omega, delta), false, make_double2(1.0, 0.0), kAccNone, kRedNorm, nullptr));
    TRY(rank_barrier(c));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_apply_h_to(qsb_ctx* c, double omega, double delta, const double* in_re_im, double* out_re_im) {
    if (!c || !in_re_im || !out_re_im) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    // through the work vectors only: the evolved state (psi) is left untouched
    CU(cudaMemcpyAsync(c->wb, in_re_im, c->dim * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    TRY(rank_barrier(c));  // partners read wb over NVLink
    TRY(launch_hpsi(c, c->wb, c->wa, hparams_global(c, omega, delta), false, make_double2(1.0, 0.0), kAccNone, kRedNorm, nullptr));
    TRY(rank_barrier(c));
    CU(cudaMemcpyAsync(out_re_im, c->wa, c->dim * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_get_work(qsb_ctx* c, double* re_im) {
    if (!c || !re_im) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(re_im, c->wa, c->dim * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_evolve_segment(qsb_ctx* c, double omega, double delta, double dt_us, int* n_hpsi) {
    if (!c) return fail(c, QSB_ERR_INVALID, "NULL context");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    CU(cudaSetDevice(c->device));
    TRY(evolve_segment_impl(c, hparams_global(c, omega, delta), dt_us, n_hpsi));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_expect_pauli(qsb_ctx* c, int64_t n_terms, const double* coef, const uint64_t* xmask, const uint64_t* ymask,
                     const uint64_t* zmask, const uint64_t* nmask, double* out_re_im) {
    if (!c || (n_terms > 0 && (!coef || !xmask || !ymask || !zmask || !nmask || !out_re_im)))
        return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    TRY(rank_barrier(c));
    std::vector<TermResult> res((size_t)n_terms, TermResult{0.0, 0.0});
    TRY(pauli_batch(c, n_terms, (const u64*)xmask, (const u64*)ymask, (const u64*)zmask, (const u64*)nmask, res.data()));
    for (int64_t t = 0; t < n_terms; ++t) {
        out_re_im[2 * t] = coef[t] * res[t].re;
        out_re_im[2 * t + 1] = coef[t] * res[t].im;
    }
    return QSB_OK;
}

// the loop of Qutip.calculate; site_stride 0: one (omega, delta) per step, n: one row of per-site values per step
static int evolve_schedule_impl(qsb_ctx* c, const double* omega, const double* delta, const double* dt_us, int64_t n_steps,
                                int site_stride, const qsb_eops* eops, double* expectations_out) {
    if (!c || (n_steps > 0 && (!omega || !delta || !dt_us))) return fail(c, QSB_ERR_INVALID, "NULL argument");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    const int n_ops = eops ? eops->n_ops : 0;
    if (n_ops > 0 && !expectations_out) return fail(c, QSB_ERR_INVALID, "expectations_out is NULL");
    CU(cudaSetDevice(c->device));
    if (n_steps > 0 && small_schedule_ok(c, eops))
        return evolve_schedule_small(c, omega, delta, dt_us, n_steps, site_stride, eops, expectations_out);
    std::vector<TermResult> res;
    for (int64_t s = 0; s < n_steps; ++s) {
        if (site_stride) {
            // channel scales of the operator terms under a per-site schedule: the site average of the row
            double amp = 0.0, det = 0.0;
            for (int q = 0; q < c->n; ++q) {
                amp += omega[s * site_stride + q];
                det += delta[s * site_stride + q];
            }
            TRY(evolve_segment_impl(c, hparams_rows(c, omega + s * site_stride, delta + s * site_stride, amp / c->n, det / c->n),
                                    dt_us[s], nullptr));
        } else {
            TRY(evolve_segment_impl(c, hparams_global(c, omega[s], delta[s]), dt_us[s], nullptr));
        }
        for (int o = 0; o < n_ops; ++o) {
            const qsb_eop& op = eops->ops[o];
            double val = 0.0;
            if (op.kind == QSB_EOP_ENERGY) {
                TRY(energy_impl(c, hparams_global(c, op.omega, op.delta), &val));
            } else if (op.kind == QSB_EOP_PAULI) {
                if (op.term_begin < 0 || op.term_end > eops->n_terms || op.term_begin > op.term_end)
                    return fail(c, QSB_ERR_INVALID, "e_op %d: bad term range", o);
                const int64_t nt = op.term_end - op.term_begin;
                res.assign((size_t)nt, TermResult{0.0, 0.0});
                TRY(rank_barrier(c));
                TRY(pauli_batch(c, nt, (const u64*)eops->xmask + op.term_begin, (const u64*)eops->ymask + op.term_begin,
                                (const u64*)eops->zmask + op.term_begin, (const u64*)eops->nmask + op.term_begin,
                                res.data()));
                for (int64_t t = 0; t < nt; ++t) val += eops->coef[op.term_begin + t] * res[t].re;
            } else {
                return fail(c, QSB_ERR_INVALID, "e_op %d: unknown kind %d", o, op.kind);
            }
            expectations_out[s * n_ops + o] = val;
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_evolve_schedule(qsb_ctx* c, const double* omega, const double* delta, const double* dt_us, int64_t n_steps,
                        const qsb_eops* eops, double* expectations_out) {
    return evolve_schedule_impl(c, omega, delta, dt_us, n_steps, 0, eops, expectations_out);
}

int qsb_evolve_schedule_sites(qsb_ctx* c, const double* omega_sites, const double* delta_sites, const double* dt_us,
                              int64_t n_steps, const qsb_eops* eops, double* expectations_out) {
    if (!c) return fail(c, QSB_ERR_INVALID, "NULL context");
    return evolve_schedule_impl(c, omega_sites, delta_sites, dt_us, n_steps, c->n, eops, expectations_out);
}

int qsb_norm(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    return marginals(c, out, nullptr);
}

int qsb_energy(qsb_ctx* c, double omega, double delta, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    CU(cudaSetDevice(c->device));
    return energy_impl(c, hparams_global(c, omega, delta), out);
}

int qsb_number(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    return marginals(c, nullptr, out);
}

// <X_q>, <Y_q>, <n_q> of every LOCAL qubit and the norm from one tile pass per group of index bits (observe.cuh):
// 3 reads of the state at 25 qubits.  loc: [3 * n + 1] partial sums of this shard (sx, sy, n per qubit, then norm).
static int spins_tiled(qsb_ctx* c, double* loc) {
    const HpsiPlan plan = make_plan(c->nl, false);
    const u64 n_tiles = c->dim >> kTileBits;
    const int grid = (int)std::min<u64>(n_tiles, (u64)c->n_sm);
    if ((size_t)grid * kSpinsRow > c->part_n || (size_t)plan.n_pass * 64 > c->slots_n) return fail(c, QSB_ERR_STATE, "partials buffer too small");
    static bool attr_set = false;
    if (!attr_set) {
        CU(cudaFuncSetAttribute(spins_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        CU(cudaFuncSetAttribute(spins_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmemBytes));
        attr_set = true;
    }
    for (int i = 0; i < plan.n_pass; ++i) {
        HpsiPass a{};
        a.x = c->psi;
        a.nl = c->nl;
        a.lb = plan.pass[i].lb;
        a.hb = plan.pass[i].hb;
        a.hb0 = plan.pass[i].hb0;
        a.n_tiles = n_tiles;
        if (i == 0) spins_tile_kernel<true><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a, c->part);
        else spins_tile_kernel<false><<<grid, kThreads, kTileSmemBytes, c->stream>>>(a, c->part);
        LAUNCHED(c);
        TRY(check_launch(c, "spins_tile_kernel"));
        finalize_rows_kernel<<<kSpinsRow, 256, 0, c->stream>>>(c->part, grid, kSpinsRow, c->slots + (size_t)i * 64);
        LAUNCHED(c);
        TRY(check_launch(c, "finalize_rows_kernel"));
    }
    TRY(fetch_slots(c, 0, (size_t)plan.n_pass * 64));
    const int n = c->n;
    for (int i = 0; i < 3 * n + 1; ++i) loc[i] = 0.0;
    const double* first = c->h_slots;
    loc[3 * n] = first[24];
    for (int bit = 0; bit < c->nl; ++bit) {
        const int q = n - 1 - bit;
        loc[3 * q + 2] = bit < kTileBits ? first[25 + bit] : first[37 + (bit - kTileBits)];
    }
    for (int i = 0; i < plan.n_pass; ++i) {
        const double* row = c->h_slots + (size_t)i * 64;
        const int lo = (i == 0) ? 0 : plan.pass[i].lb, shift = (i == 0) ? 0 : plan.pass[i].hb0 - plan.pass[i].lb;
        for (int b = lo; b < kTileBits; ++b) {
            const int q = n - 1 - (b + shift);
            loc[3 * q + 0] = row[b];
            loc[3 * q + 1] = row[kTileBits + b];
        }
    }
    for (int g = 0; g < c->p; ++g)  // a sharded qubit is set in whole shards
        if ((c->rank >> g) & 1) loc[3 * (n - 1 - (c->nl + g)) + 2] = loc[3 * n];
    return QSB_OK;
}

int qsb_spins(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    const int n = c->n;
    if (c->plan.tiled && c->obs_tiled && 3 * (size_t)n + 1 <= c->slots_n) {
        std::vector<double> loc(3 * (size_t)n + 1);
        TRY(rank_barrier(c));
        TRY(spins_tiled(c, loc.data()));
        if (c->nranks > 1) {
            CU(cudaMemcpyAsync(c->slots, loc.data(), loc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            TRY(allreduce(c, c->slots, loc.size(), kNcclSum));
            TRY(fetch_slots(c, 0, loc.size()));
            for (size_t i = 0; i < loc.size(); ++i) loc[i] = c->h_slots[i];
        }
        for (int q = 0; q < n; ++q) {
            out[3 * q + 0] = loc[3 * q + 0];
            out[3 * q + 1] = loc[3 * q + 1];
            out[3 * q + 2] = loc[3 * n] - 2.0 * loc[3 * q + 2];  // sum p (1 - 2 b), qse/magnetic.py:150-155
        }
        if (c->p > 0) {
            // X / Y of the sharded qubits pair amplitudes across shards: the Pauli-string path (peer memory)
            const int np = c->p;
            std::vector<u64> xm(2 * np, 0), ym(2 * np, 0), zm(2 * np, 0), nm(2 * np, 0);
            for (int q = 0; q < np; ++q) {
                xm[q] = 1ull << (n - 1 - q);
                ym[np + q] = 1ull << (n - 1 - q);
            }
            std::vector<TermResult> res(2 * np, TermResult{0.0, 0.0});
            TRY(pauli_batch(c, 2 * np, xm.data(), ym.data(), zm.data(), nm.data(), res.data()));
            for (int q = 0; q < np; ++q) {
                out[3 * q + 0] = res[q].re;
                out[3 * q + 1] = res[np + q].re;
            }
        }
        return QSB_OK;
    }
    double tot = 0.0;
    std::vector<double> num(n);
    TRY(marginals(c, &tot, num.data()));
    // <X_q>, <Y_q> as Pauli strings (qse/magnetic.py:157-173)
    std::vector<u64> xm(2 * n, 0), ym(2 * n, 0), zm(2 * n, 0), nm(2 * n, 0);
    for (int q = 0; q < n; ++q) {
        const u64 m = 1ull << (n - 1 - q);
        xm[q] = m;
        ym[n + q] = m;
    }
    std::vector<TermResult> res(2 * n, TermResult{0.0, 0.0});
    TRY(rank_barrier(c));
    TRY(pauli_batch(c, 2 * n, xm.data(), ym.data(), zm.data(), nm.data(), res.data()));
    for (int q = 0; q < n; ++q) {
        out[3 * q + 0] = res[q].re;
        out[3 * q + 1] = res[n + q].re;
        out[3 * q + 2] = tot - 2.0 * num[q];  // sum p (1 - 2 b), qse/magnetic.py:150-155
    }
    return QSB_OK;
}

int qsb_sisj(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    const int n = c->n;
    // s_ij = <Z_iZ_j> + (1/3)<X_iX_j + Y_iY_j>        (qse/magnetic.py:213-243)
    // <Z_iZ_j> = <1> - 2<n_i> - 2<n_j> + 4<n_i n_j>: one conditional-marginal pass per qubit i gives
    // <n_i n_j> for every j (N passes instead of N(N-1)/2).
    double tot = 0.0;
    std::vector<double> num(n), row(n), nij((size_t)n * n, 0.0);
    TRY(marginals(c, &tot, num.data()));
    for (int i = 0; i < n; ++i) {
        double ni = 0.0;
        TRY(marginals(c, &ni, row.data(), i));
        for (int j = 0; j < n; ++j) nij[(size_t)i * n + j] = row[j];
    }
    // <XX + YY> = sum_k conj(psi[k^f]) (1 - z_i z_j) psi[k], f = m_i | m_j: ONE pass per pair when both
    // qubits are shard-local (mode 1); pairs touching a sharded qubit take the XX and YY strings
    // separately because their shard-dependent sign cannot be folded into the weight.
    std::vector<u64> xm, ym, zm, nm;
    std::vector<int> modes, pair_of;
    int pidx = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j, ++pidx) {
            const u64 m = (1ull << (n - 1 - i)) | (1ull << (n - 1 - j));
            if ((m >> c->nl) == 0) {
                xm.push_back(m); ym.push_back(0); zm.push_back(0); nm.push_back(0); modes.push_back(1); pair_of.push_back(pidx);
            } else {
                xm.push_back(m); ym.push_back(0); zm.push_back(0); nm.push_back(0); modes.push_back(0); pair_of.push_back(pidx);
                xm.push_back(0); ym.push_back(m); zm.push_back(0); nm.push_back(0); modes.push_back(0); pair_of.push_back(pidx);
            }
        }
    std::vector<TermResult> res(xm.size(), TermResult{0.0, 0.0});
    TRY(rank_barrier(c));
    TRY(pauli_batch(c, (int64_t)xm.size(), xm.data(), ym.data(), zm.data(), nm.data(), res.data(), modes.data()));
    std::vector<double> xxyy((size_t)n * (n - 1) / 2 + 1, 0.0);
    for (size_t t = 0; t < res.size(); ++t) xxyy[pair_of[t]] += res[t].re;
    pidx = 0;
    for (int i = 0; i < n; ++i) {
        out[(size_t)i * n + i] = 1.0;
        for (int j = i + 1; j < n; ++j, ++pidx) {
            const double zz = tot - 2.0 * num[i] - 2.0 * num[j] + 4.0 * nij[(size_t)i * n + j];
            const double sij = zz + xxyy[pidx] / 3.0;
            out[(size_t)i * n + j] = sij;
            out[(size_t)j * n + i] = sij;
        }
    }
    return QSB_OK;
}

int qsb_probabilities(qsb_ctx* c, double* out) {
    if (!c || !out) return fail(c, QSB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    // wa is free between calls: use it as the device staging buffer for |psi|^2
    double* d = reinterpret_cast<double*>(c->wa);
    prob_kernel<<<grid_for(c, c->dim, 256), 256, 0, c->stream>>>(c->psi, d, c->dim);
    LAUNCHED(c);
    TRY(check_launch(c, "prob_kernel"));
    CU(cudaMemcpyAsync(out, d, c->dim * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_topk_probs(qsb_ctx* c, int k, uint64_t* idx, double* prob) {
    if (!c || !idx || !prob || k < 1) return fail(c, QSB_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    const u64 total = c->dim;
    const u64 kk = std::min<u64>((u64)k, total);
    // scratch: 65536 counters | cap x (idx, p) candidates | all-gather staging of 2*kk doubles per rank
    const unsigned int cap = (unsigned int)std::min<u64>(total, std::max<u64>(4 * kk, 1u << 16));
    const size_t hist_bytes = 65536 * sizeof(unsigned int);
    const size_t need = hist_bytes + (size_t)cap * 16 + (size_t)2 * kk * (c->nranks + 1) * sizeof(double);
    if (need > c->topk_bytes) {
        cudaFree(c->topk_buf);
        c->topk_buf = nullptr;
        c->topk_bytes = 0;
        CU(cudaMalloc(&c->topk_buf, need));
        c->topk_bytes = need;
    }
    // radix select (4 x 16 bits) of the kk-th largest probability of this shard
    unsigned int* hist = reinterpret_cast<unsigned int*>(c->topk_buf);
    std::vector<unsigned int> h(65536);
    const int grid = grid_for(c, c->dim, 256);
    u64 prefix = 0, pmask = 0, want = kk;
    for (int shift = 48; shift >= 0; shift -= 16) {
        CU(cudaMemsetAsync(hist, 0, 65536 * sizeof(unsigned int), c->stream));
        radix_hist_kernel<<<grid, 256, 0, c->stream>>>(c->psi, c->dim, prefix, pmask, shift, hist);
        LAUNCHED(c);
        TRY(check_launch(c, "radix_hist_kernel"));
        CU(cudaMemcpyAsync(h.data(), hist, 65536 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        int b = 65535;
        for (; b >= 0; --b) {
            if ((u64)h[b] >= want) break;
            want -= h[b];
        }
        if (b < 0) return fail(c, QSB_ERR_STATE, "radix select lost count");
        prefix |= (u64)b << shift;
        pmask |= 0xffffull << shift;
    }
    // `prefix` is now the key of the kk-th largest probability and `want` the number of entries EQUAL to
    // it that still belong to the top kk (the others are strictly larger).  Gather the strictly larger
    // ones (at most kk - want <= cap), then the `want` ties of LOWEST index in index order: a sparse
    // state (all ties at probability 0, e.g. |0...0> or an empty shard) costs nothing extra.
    u64* d_idx = reinterpret_cast<u64*>(c->topk_buf + hist_bytes);
    double* d_p = reinterpret_cast<double*>(c->topk_buf + hist_bytes) + cap;
    unsigned int* d_count = reinterpret_cast<unsigned int*>(c->slots);
    CU(cudaMemsetAsync(d_count, 0, sizeof(double), c->stream));
    gather_gt_kernel<<<grid, 256, 0, c->stream>>>(c->psi, c->dim, prefix, c->rank_bits, d_idx, d_p, d_count, cap);
    LAUNCHED(c);
    TRY(check_launch(c, "gather_gt_kernel"));
    unsigned int count = 0;
    CU(cudaMemcpyAsync(&count, d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((u64)count + want != kk || count > cap)  // cannot happen: the radix select counted exactly these
        return fail(c, QSB_ERR_STATE, "top-k: %u entries above the threshold, %llu ties wanted, k = %llu", count,
                    (unsigned long long)want, (unsigned long long)kk);
    {
        // ties in index order: per-chunk tie counts -> host prefix -> the chunks that hold the first `want`
        const int cb = std::min(c->nl, 16);
        const u64 n_chunks = c->dim >> cb;
        unsigned int* d_tie = hist;  // the histogram is free again (n_chunks <= 65536 ... else several rounds)
        std::vector<unsigned int> tie(65536);
        u64 found = 0;
        for (u64 c0 = 0; c0 < n_chunks && found < want; c0 += 65536) {
            const u64 nc = std::min<u64>(65536, n_chunks - c0);
            count_eq_kernel<<<(int)std::min<u64>(nc, (u64)c->n_sm * 8), 256, 0, c->stream>>>(c->psi, cb, c0, nc, prefix, d_tie);
            LAUNCHED(c);
            TRY(check_launch(c, "count_eq_kernel"));
            CU(cudaMemcpyAsync(tie.data(), d_tie, nc * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            for (u64 i = 0; i < nc && found < want; ++i) {
                if (tie[i] == 0) continue;
                const unsigned int take = (unsigned int)std::min<u64>(tie[i], want - found);
                take_eq_kernel<<<1, 32, 0, c->stream>>>(c->psi, cb, c0 + i, prefix, c->rank_bits, take, d_idx + count + found,
                                                        d_p + count + found);
                LAUNCHED(c);
                TRY(check_launch(c, "take_eq_kernel"));
                found += take;
            }
        }
        if (found != want) return fail(c, QSB_ERR_STATE, "top-k: tie scan found %llu of %llu", (unsigned long long)found, (unsigned long long)want);
        count += (unsigned int)want;
    }
    std::vector<u64> hi(count);
    std::vector<double> hp(count);
    CU(cudaMemcpyAsync(hi.data(), d_idx, count * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hp.data(), d_p, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemsetAsync(c->slots, 0, sizeof(double), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<std::pair<double, u64>> cand(count);
    for (unsigned int i = 0; i < count; ++i) cand[i] = {hp[i], hi[i]};
    auto better = [](const std::pair<double, u64>& a, const std::pair<double, u64>& b) {
        return a.first > b.first || (a.first == b.first && a.second < b.second);
    };
    std::sort(cand.begin(), cand.end(), better);
    if (cand.size() > kk) cand.resize(kk);
    if (c->nranks > 1) {
        // merge the per-shard winners: all-gather kk (p, idx) pairs per rank
        std::vector<double> mine(2 * kk, -1.0), all(2 * kk * c->nranks);
        for (size_t i = 0; i < cand.size(); ++i) {
            mine[2 * i] = cand[i].first;
            u64 v = cand[i].second;
            memcpy(&mine[2 * i + 1], &v, sizeof(double));
        }
        double* d_m = reinterpret_cast<double*>(c->topk_buf + hist_bytes + (size_t)cap * 16);
        double* d_a = d_m + 2 * kk;
        CU(cudaMemcpyAsync(d_m, mine.data(), mine.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        NC(nccl_api()->AllGather(d_m, d_a, 2 * kk, kNcclFloat64, c->comm, c->stream));
        CU(cudaMemcpyAsync(all.data(), d_a, all.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cand.clear();
        for (size_t i = 0; i < (size_t)kk * c->nranks; ++i) {
            if (all[2 * i] < 0.0) continue;
            u64 v;
            memcpy(&v, &all[2 * i + 1], sizeof(double));
            cand.push_back({all[2 * i], v});
        }
        std::sort(cand.begin(), cand.end(), better);
    }
    for (int i = 0; i < k; ++i) {
        if ((size_t)i < cand.size()) {
            prob[i] = cand[i].first;
            idx[i] = cand[i].second;
        } else {
            prob[i] = 0.0;
            idx[i] = ~0ull;
        }
    }
    return QSB_OK;
}

static inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int qsb_sample(qsb_ctx* c, uint64_t seed, int64_t shots, uint64_t* out) {
    if (!c || !out || shots < 0) return fail(c, QSB_ERR_INVALID, "bad argument");
    if (shots == 0) return QSB_OK;
    CU(cudaSetDevice(c->device));
    const int cb = std::min(c->nl, 12);
    const u64 n_chunks = c->dim >> cb;
    // scratch: chunk sums | per-shot (chunk, residual, result) -- carved from a dedicated allocation
    const size_t need = n_chunks * sizeof(double) + (size_t)shots * 24 + (size_t)(c->nranks + shots) * sizeof(double);
    if (need > c->topk_bytes) {
        cudaFree(c->topk_buf);
        c->topk_buf = nullptr;
        c->topk_bytes = 0;
        CU(cudaMalloc(&c->topk_buf, need));
        c->topk_bytes = need;
    }
    double* d_sums = reinterpret_cast<double*>(c->topk_buf);
    u64* d_chunk = reinterpret_cast<u64*>(d_sums + n_chunks);
    double* d_resid = reinterpret_cast<double*>(d_chunk + shots);
    u64* d_out = reinterpret_cast<u64*>(d_resid + shots);
    double* d_comm = reinterpret_cast<double*>(d_out + shots);  // nranks + shots doubles
    chunk_prob_kernel<<<(int)std::min<u64>(n_chunks, (u64)c->n_sm * 8), 256, 0, c->stream>>>(c->psi, cb, n_chunks, d_sums);
    LAUNCHED(c);
    TRY(check_launch(c, "chunk_prob_kernel"));
    std::vector<double> sums(n_chunks);
    CU(cudaMemcpyAsync(sums.data(), d_sums, n_chunks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<double> cdf(n_chunks);
    double mine = 0.0;
    for (u64 i = 0; i < n_chunks; ++i) {
        mine += sums[i];
        cdf[i] = mine;
    }
    // every rank needs every shard's probability mass
    std::vector<double> mass(c->nranks, 0.0);
    mass[c->rank] = mine;
    if (c->nranks > 1) {
        CU(cudaMemcpyAsync(d_comm, mass.data(), c->nranks * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(allreduce(c, d_comm, (size_t)c->nranks, kNcclSum));
        CU(cudaMemcpyAsync(mass.data(), d_comm, c->nranks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    double total = 0.0;
    std::vector<double> rank_lo(c->nranks + 1, 0.0);
    for (int r = 0; r < c->nranks; ++r) {
        total += mass[r];
        rank_lo[r + 1] = total;
    }
    if (!(total > 0.0)) return fail(c, QSB_ERR_STATE, "cannot sample from a state of zero norm");
    // same uniforms on every rank; each rank resolves the shots that land in its shard
    std::vector<u64> h_chunk;
    std::vector<double> h_resid;
    std::vector<int64_t> h_slot;
    uint64_t state = seed;
    for (int64_t sidx = 0; sidx < shots; ++sidx) {
        const double u = (double)(splitmix64(state) >> 11) * (1.0 / 9007199254740992.0) * total;
        int r = (int)(std::upper_bound(rank_lo.begin() + 1, rank_lo.end(), u) - (rank_lo.begin() + 1));
        if (r >= c->nranks) r = c->nranks - 1;
        if (r != c->rank) continue;
        const double local = u - rank_lo[r];
        u64 ch = (u64)(std::upper_bound(cdf.begin(), cdf.end(), local) - cdf.begin());
        if (ch >= n_chunks) ch = n_chunks - 1;
        h_chunk.push_back(ch);
        h_resid.push_back(local - (ch ? cdf[ch - 1] : 0.0));
        h_slot.push_back(sidx);
    }
    const long long n_mine = (long long)h_chunk.size();
    std::vector<u64> h_out((size_t)n_mine);
    if (n_mine > 0) {
        CU(cudaMemcpyAsync(d_chunk, h_chunk.data(), n_mine * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_resid, h_resid.data(), n_mine * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        const long long threads = n_mine * 32;
        sample_resolve_kernel<<<(int)((threads + 255) / 256), 256, 0, c->stream>>>(c->psi, cb, n_mine, d_chunk, d_resid, d_out);
        LAUNCHED(c);
        TRY(check_launch(c, "sample_resolve_kernel"));
        CU(cudaMemcpyAsync(h_out.data(), d_out, n_mine * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (c->nranks == 1) {
        for (long long i = 0; i < n_mine; ++i) out[h_slot[i]] = h_out[i];
        return QSB_OK;
    }
    // merge: basis indices < 2^53 are exact in float64, every shot is owned by exactly one rank
    std::vector<double> merged((size_t)shots, 0.0);
    for (long long i = 0; i < n_mine; ++i) merged[h_slot[i]] = (double)(c->rank_bits + h_out[i]);
    CU(cudaMemcpyAsync(d_comm, merged.data(), (size_t)shots * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    TRY(allreduce(c, d_comm, (size_t)shots, kNcclSum));
    CU(cudaMemcpyAsync(merged.data(), d_comm, (size_t)shots * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < shots; ++i) out[i] = (uint64_t)merged[i];
    return QSB_OK;
}

int qsb_describe_plan(int n_local, int sb_bits, int tile_bits, int* out) {
    qsb_ctx* c = nullptr;
    if (!out || n_local <= kTileBits || n_local >= QSB_MAX_QUBITS_DEV) return fail(c, QSB_ERR_INVALID, "n_local out of range");
    if (sb_bits != 0 && (sb_bits < 13 || sb_bits > 21)) return fail(c, QSB_ERR_INVALID, "sb_bits must be 0 or in 13..21");
    if (tile_bits != 0 && tile_bits != 11 && tile_bits != 12) return fail(c, QSB_ERR_INVALID, "tile_bits must be 0, 11 or 12");
    const WsPlan w = make_ws_plan(n_local, sb_bits, tile_bits);
    out[0] = w.tb;
    out[1] = w.sbits;
    out[2] = w.n_plain;
    out[3] = w.inner.lb;
    out[4] = w.inner.hb;
    out[5] = w.inner.hb0;
    for (int i = 0; i < w.n_plain; ++i) {
        out[6 + 3 * i] = w.plain[i].lb;
        out[7 + 3 * i] = w.plain[i].hb;
        out[8 + 3 * i] = w.plain[i].hb0;
    }
    return QSB_OK;
}

int qsb_describe_ticket(uint64_t ticket, uint64_t n_a, uint64_t n_sb, uint64_t lag, uint64_t sb_xor, int* type,
                        uint64_t* tile) {
    qsb_ctx* c = nullptr;
    if (!type || !tile || n_a == 0 || n_sb == 0 || lag > n_sb || ticket >= 2 * n_a * n_sb || sb_xor >= n_sb)
        return fail(c, QSB_ERR_INVALID, "bad schedule parameters");
    u64 tau = 0;
    *type = ws_decode_ticket(ticket, n_a, n_sb, lag, sb_xor, &tau);
    *tile = tau;
    return QSB_OK;
}

int qsb_get_metrics(const qsb_ctx* c, qsb_metrics* out) {
    if (!c || !out) return fail(nullptr, QSB_ERR_INVALID, "NULL argument");
    *out = c->met;
    out->hpsi_passes = c->plan.n_pass;
    out->hpsi_bytes_per_amp = c->plan.bytes_per_amp;
    if (c->path == 0 && c->nl > kTileBits) {
        const WsPlan w = make_ws_plan(c->nl, c->sb_bits, c->nranks > 1 ? 12 : c->ws_tile_bits);
        out->hpsi_passes = 1 + w.n_plain;  // launches; the first one fuses two sub-passes through L2
        out->hpsi_bytes_per_amp = w.bytes_per_amp;
    }
    out->diag_min = c->diag_min;
    out->diag_max = c->diag_max;
    return QSB_OK;
}

static int flush_l2(qsb_ctx* c) {
    if (!c->flush_buf) {
        c->flush_n = (size_t)(192u << 20) / sizeof(double);  // 192 MiB > 126 MB L2
        CU(cudaMalloc(&c->flush_buf, c->flush_n * sizeof(double)));
    }
    flush_kernel<<<c->n_sm * 8, 256, 0, c->stream>>>(c->flush_buf, c->flush_n, 1.0);
    return check_launch(c, "flush_kernel");
}

int qsb_time_apply_h(qsb_ctx* c, double omega, double delta, int iters, int do_flush, float* ms_each) {
    if (!c || !ms_each || iters < 1) return fail(c, QSB_ERR_INVALID, "bad argument");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < iters; ++i) {
        if (do_flush) TRY(flush_l2(c));
        TRY(rank_barrier(c));
        CU(cudaEventRecord(c->ev0, c->stream));
        TRY(launch_hpsi(c, c->psi, c->wa, hparams_global(c, omega, delta), false, make_double2(1.0, 0.0), kAccNone, kRedNorm, nullptr));
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaEventSynchronize(c->ev1));
        CU(cudaEventElapsedTime(&ms_each[i], c->ev0, c->ev1));
    }
    TRY(rank_barrier(c));
    CU(cudaStreamSynchronize(c->stream));
    return QSB_OK;
}

int qsb_time_schedule(qsb_ctx* c, const double* omega, const double* delta, const double* dt_us, int64_t n_steps,
                      float* ms_each) {
    if (!c || !omega || !delta || !dt_us || !ms_each) return fail(c, QSB_ERR_INVALID, "NULL argument");
    if (!c->have_v) return fail(c, QSB_ERR_STATE, "qsb_set_interaction has not been called");
    CU(cudaSetDevice(c->device));
    for (int64_t s = 0; s < n_steps; ++s) {
        CU(cudaEventRecord(c->ev0, c->stream));
        TRY(evolve_segment_impl(c, hparams_global(c, omega[s], delta[s]), dt_us[s], nullptr));
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaEventSynchronize(c->ev1));
        CU(cudaEventElapsedTime(&ms_each[s], c->ev0, c->ev1));
    }
    return QSB_OK;
}

}  // extern "C"
