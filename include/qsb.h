/*
 * qsb.h — C ABI of libqsb, the B200-native state-vector engine behind qse_b200.
 *
 * The reference (ICHEC/qse 1.1.21) has no FFI: its plugin interface is the Python class
 * protocol of qse/calc/calculator.py:70-187, and the arithmetic of the hot path is a
 * third-party Python package (QuTiP).  This header is therefore the boundary a maintainer
 * of the reference would bind with ctypes from a new `qse/calc/b200.py` (INTEGRATION.md
 * shows that stub).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success, <0 on error; qsb_last_error() gives the text.
 *   - host arrays are caller-owned, C-contiguous; complex128 = interleaved double[2].
 *     The library never keeps a host pointer after the call returns.
 *   - device memory is library-owned and freed by qsb_destroy().
 *   - one context = one GPU = one host thread (not re-entrant).  Multi-GPU = one process
 *     per GPU, every process creating a context with the same nccl_uid (SPMD): functions
 *     marked [collective] must be called by all ranks in the same order.
 *   - basis index k is uint64; qubit q is bit (N-1-q) of k (qubit 0 = most significant:
 *     qse/magnetic.py:51, qse/operator/operator.py:89-92).  With P = 2^p ranks, rank r owns
 *     the contiguous block [r*2^(N-p), (r+1)*2^(N-p)): the sharded qubits are 0..p-1, and
 *     concatenating the shards in rank order reproduces the reference ordering.
 *   - units: omega, delta in rad/us; dt in us; V in rad/us (qse/calc/qutip.py:7, 91-92).
 */
#ifndef QSB_H
#define QSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSB_ABI_VERSION 1
#define QSB_NCCL_UID_BYTES 128
#define QSB_MAX_QUBITS 36   /* 33 local qubits + 3 sharded ones (8 ranks) */

typedef struct qsb_ctx qsb_ctx;

/* error codes */
#define QSB_OK 0
#define QSB_ERR_INVALID (-1)   /* bad argument */
#define QSB_ERR_CUDA (-2)      /* CUDA runtime failure */
#define QSB_ERR_NCCL (-3)      /* NCCL failure / not loadable */
#define QSB_ERR_STATE (-4)     /* call out of order (e.g. no interaction set) */
#define QSB_ERR_NOCONV (-5)    /* time stepper did not converge */
#define QSB_ERR_UNSUPPORTED (-6)

/* ---- library ---------------------------------------------------------------------- */
int qsb_abi_version(void);
const char* qsb_build_info(void);
int qsb_device_count(int* count);
/* text of the last error on ctx (or of the last failed call without a context if NULL) */
const char* qsb_last_error(const qsb_ctx* ctx);
/* rank 0 creates the id and the host broadcasts it (torch.distributed, MPI, a file...) */
int qsb_nccl_unique_id(void* out_uid /* QSB_NCCL_UID_BYTES */);

/* ---- context ---------------------------------------------------------------------- */
/* replaces: Qutip.__init__ state set-up (qse/calc/qutip.py:13-32) */
int qsb_create(int n_qubits, int device, qsb_ctx** out);
/* [collective] nranks must be 1, 2, 4 or 8 (at most 3 sharded qubits); nccl_uid from qsb_nccl_unique_id */
int qsb_create_sharded(int n_qubits, int device, int rank, int nranks, const void* nccl_uid,
                       qsb_ctx** out);
int qsb_destroy(qsb_ctx* ctx);
int qsb_local_dim(const qsb_ctx* ctx, uint64_t* dim);   /* 2^(N-p) amplitudes in this shard */
/* options: "tol" (Taylor term tolerance, default 1e-14), "theta_max" (sub-step bound on
 * |H| dt, default 3), "max_terms" (default 64), "fixed_terms" (0 = norm-terminated;
 * 4 = classical RK4), "hpsi_path" (0 auto = warp-specialised super-block kernel, 1 = untiled
 * one-thread-per-amplitude kernel, 2 = one launch per tile-bit group), "sb_bits" (log2 of the
 * L2-resident super-block, 13..21; 0 = auto: 18..21 by shard size), "sb_lag" (lead, in
 * super-blocks, of the first sub-pass over the second inside the fused launch; 0 = auto, -1 = no lead),
 * "tile_bits" (12 or 11),
 * "tensor_map" (1 = strided tiles fetched with 2-D tensor-map TMA, 0 = one bulk copy per row),
 * "small_kernel" (1 = schedules of <= 12 qubits run in one single-block launch),
 * "diag_on_the_fly" (1 = the H psi kernel evaluates the 2-local diagonal itself instead of reading E; -1 = auto:
 * from 26 local qubits), "exchange" (-1 auto, 0 = direct partner pulls,
 * 1 = qubit-remap exchange), "exchange_split" (where partner tiles are pulled: -1 auto, 0 fused launch,
 * 1 split between the fused and the last plain launch, 2 last plain launch), "remap_ctas" (SMs the remap kernel
 * takes beside the fused launch; 0 = run it before the fused launch) */
int qsb_set_option(qsb_ctx* ctx, const char* key, double value);

/* page-locked host memory for state-sized transfers (qsb_get_state / qsb_set_state run at PCIe
 * speed into such a buffer; into pageable memory the driver stages the copy).  Free with
 * qsb_host_free.  The reference keeps its state in ordinary numpy memory (qse/calc/qutip.py:49). */
int qsb_host_alloc(uint64_t bytes, void** out);
int qsb_host_free(void* ptr);

/* ---- Hamiltonian ------------------------------------------------------------------ */
/* replaces: Qutip._build_hamiltonian_operators -> ham_int (qse/calc/qutip.py:60-64) and
 * Operators.to_qutip (qse/operator/operators.py:32-44).  v is N*N row-major; only the strict
 * upper triangle is read (v[i*N+j], i<j), already thresholded by the host exactly as
 * qse/qbits.py:1296-1304 does.  Builds E_int[k] = sum_{i<j} v_ij b_i(k) b_j(k) on the device,
 * summed in (i, j) lexicographic order (bit-identical to the sequential host sum). [collective] */
int qsb_set_interaction(qsb_ctx* ctx, const double* v);
/* E_int of this shard (local_dim doubles): the diagonal of ham_int */
int qsb_get_diagonal(qsb_ctx* ctx, double* out);
/* Per-site drive and detuning of the reference's headline Hamiltonian (README.md:93-95):
 *   H = sum_q (omega/2) w_omega[q] X_q - sum_q delta w_delta[q] n_q + interaction (+ operator terms).
 * N weights each, NULL = all 1 (the global pulse of qse.calc.Qutip, qse/calc/qutip.py:55-58).  They scale every
 * (omega, delta) passed afterwards: qsb_apply_h, qsb_evolve_*, qsb_energy, QSB_EOP_ENERGY. */
int qsb_set_site_weights(qsb_ctx* ctx, const double* w_omega, const double* w_delta);
/* Operator-sum Hamiltonians (qse.Operators: sums of coef * Pauli strings over X, Y, Z, N;
 * qse/operator/operator.py:38-60, 80-92): H gains sum_t s_t(t) coef_t P_t with s = 1 (channel 0, static),
 * omega(t) (channel 1) or delta(t) (channel 2) -- qse.calc.Qutip's H = amp*H_amp + det*H_det + H_int with
 * arbitrary H_amp / H_det, the TFIM of docs/use-cases/time_evo.ipynb, the XY channel of
 * qse/calc/pulser.py:174-177.  Masks as in qsb_eops (pairwise disjoint).  Static diagonal strings are folded into
 * the diagonal, single-X / single-N strings into the per-site drive / detuning (the tiled H psi kernels cost the
 * same); anything else is applied by a generic gather kernel, one pass over the state per H psi.  Replaces any
 * previous term list; n_terms = 0 clears it.  channel may be NULL (all static).  qsb_set_interaction (zeros for a
 * pure operator Hamiltonian) must have been called or is implied with V = 0.  [collective] */
int qsb_set_operator_terms(qsb_ctx* ctx, int64_t n_terms, const double* coef, const uint64_t* xmask,
                           const uint64_t* ymask, const uint64_t* zmask, const uint64_t* nmask,
                           const int32_t* channel);

/* ---- state ------------------------------------------------------------------------ */
/* |0...0>: qp.tensor([qp.fock(2)] * N), qse/calc/qutip.py:35 */
int qsb_reset_state(qsb_ctx* ctx);
int qsb_set_state(qsb_ctx* ctx, const double* re_im /* local_dim*2 */);
int qsb_get_state(qsb_ctx* ctx, double* re_im /* local_dim*2 */);

/* ---- H psi and time evolution ------------------------------------------------------- */
/* work <- H(omega, delta) psi with H of qse/calc/qutip.py:55-58 (0.5*omega*sum X - delta*sum n
 * + H_int); the sparse mat-vec inside qp.sesolve.  Exported for parity tests and profiling.
 * [collective] */
int qsb_apply_h(qsb_ctx* ctx, double omega, double delta);
int qsb_get_work(qsb_ctx* ctx, double* re_im /* local_dim*2 */);
/* out <- H(omega, delta) in for a caller-supplied shard (get_hamiltonian(amp, det) * state for any state,
 * qse/calc/qutip.py:55-58) WITHOUT touching the evolved state psi.  [collective] */
int qsb_apply_h_to(qsb_ctx* ctx, double omega, double delta, const double* in_re_im, double* out_re_im);
/* psi <- exp(-i H(omega, delta) dt) psi: one qp.sesolve(H, psi, [0, dt]) call
 * (qse/calc/qutip.py:43-45).  n_hpsi (may be NULL) receives the number of H psi applied.
 * [collective] */
int qsb_evolve_segment(qsb_ctx* ctx, double omega, double delta, double dt_us, int* n_hpsi);

/* expectation-value requests recorded after every step (e_ops of Qutip.calculate) */
#define QSB_EOP_ENERGY 0 /* <psi|H(omega, delta)|psi>  (notebook idiom: get_hamiltonian as e_op) */
#define QSB_EOP_PAULI 1  /* sum of Pauli strings, terms [term_begin, term_end) */
typedef struct {
    int32_t kind;
    int32_t reserved;
    double omega, delta;          /* QSB_EOP_ENERGY */
    int64_t term_begin, term_end; /* QSB_EOP_PAULI  */
} qsb_eop;
/* A Pauli term is coef * prod X(xmask) Y(ymask) Z(zmask) N(nmask); masks are in basis-index
 * bit positions (qubit q -> bit N-1-q) and pairwise disjoint.  Operator.to_qutip semantics
 * (qse/operator/operator.py:80-92, 131-142): N = diag(0, 1). */
typedef struct {
    int32_t n_ops;
    int32_t reserved;
    const qsb_eop* ops;
    int64_t n_terms;
    const double* coef;
    const uint64_t* xmask;
    const uint64_t* ymask;
    const uint64_t* zmask;
    const uint64_t* nmask;
} qsb_eops;
/* the loop of Qutip.calculate (qse/calc/qutip.py:40-47) from the CURRENT state: for each step s,
 * psi <- exp(-i H(omega[s], delta[s]) dt_us[s]) psi, then expectations_out[s*n_ops + o] =
 * Re <psi|op_o|psi>.  eops may be NULL.  [collective] */
int qsb_evolve_schedule(qsb_ctx* ctx, const double* omega, const double* delta, const double* dt_us,
                        int64_t n_steps, const qsb_eops* eops, double* expectations_out);

/* the same loop with PER-SITE signal values: omega_sites / delta_sites are n_steps x N row-major, row s holding
 * Omega_q and delta_q of step s for q = 0..N-1 (site weights still apply on top).  [collective] */
int qsb_evolve_schedule_sites(qsb_ctx* ctx, const double* omega_sites, const double* delta_sites,
                              const double* dt_us, int64_t n_steps, const qsb_eops* eops,
                              double* expectations_out);

/* ---- observables (qse/magnetic.py, qp.expect) --------------------------------------- */
int qsb_norm(qsb_ctx* ctx, double* out);                                   /* <psi|psi> [collective] */
int qsb_energy(qsb_ctx* ctx, double omega, double delta, double* out);     /* [collective] */
/* out_re_im[2*t], [2*t+1] = <psi|term_t|psi> for each term (coef included). [collective] */
int qsb_expect_pauli(qsb_ctx* ctx, int64_t n_terms, const double* coef, const uint64_t* xmask,
                     const uint64_t* ymask, const uint64_t* zmask, const uint64_t* nmask,
                     double* out_re_im);
/* (N, 3) <X_i>, <Y_i>, <Z_i>: magnetic.get_spins, qse/magnetic.py:119-176. [collective] */
int qsb_spins(qsb_ctx* ctx, double* out);
/* (N, N): magnetic.get_sisj, qse/magnetic.py:179-243 — <Z_iZ_j> + (1/3)<X_iX_j + Y_iY_j> off the
 * diagonal, exactly 1 on it (the CODE's normalisation). [collective] */
int qsb_sisj(qsb_ctx* ctx, double* out);
/* (N,): magnetic.get_number_operator, qse/magnetic.py:295-326. [collective] */
int qsb_number(qsb_ctx* ctx, double* out);
/* |psi_k|^2 of this shard (local_dim doubles): notebook idiom conj(state)*state */
int qsb_probabilities(qsb_ctx* ctx, double* out);
/* the k most probable basis states of the whole register, descending (ties: lower index first);
 * idx are global basis indices.  docs/use-cases/adiabatic.ipynb cell 8. [collective] */
int qsb_topk_probs(qsb_ctx* ctx, int k, uint64_t* idx, double* prob);

/* `shots` basis states drawn from |psi_k|^2 / <psi|psi> (global basis indices, in draw order);
 * deterministic in `seed`, identical on every rank.  The measurement read-out Pulser users get from
 * results.sample_final_state (docs/tutorials/pulser-calc-example.ipynb cell 10). [collective] */
int qsb_sample(qsb_ctx* ctx, uint64_t seed, int64_t shots, uint64_t* out);

/* The launch plan of one H psi for a shard of 2^n_local amplitudes (host-only, needs no GPU):
 * out[0] = tile bits, out[1] = super-block bits, out[2] = number of plain launches, then (lb, hb, hb0)
 * of the inner sub-pass and of each plain launch: out[3 + 3*i ...].  out must hold 3 + 3*9 ints.
 * sb_bits / tile_bits as in qsb_set_option (0 = auto). */
int qsb_describe_plan(int n_local, int sb_bits, int tile_bits, int* out);
/* The static work schedule of the fused launch (host-only): ticket -> *type (0 = first sub-pass,
 * 1 = inner sub-pass) and *tile, for n_sb super-blocks of n_a tiles, a lead of `lag` super-blocks and
 * the per-rank order permutation sb_xor.  The function the kernel's producer warp evaluates. */
int qsb_describe_ticket(uint64_t ticket, uint64_t n_a, uint64_t n_sb, uint64_t lag, uint64_t sb_xor,
                        int* type, uint64_t* tile);

/* ---- measurement ------------------------------------------------------------------- */
typedef struct {
    uint64_t hpsi_calls;        /* H psi applications since creation */
    uint64_t kernel_launches;   /* kernels launched by this library since creation */
    uint64_t segments;          /* evolve_segment calls (sub-steps not counted) */
    uint64_t substeps;
    double last_theta;          /* |H| dt bound of the last segment */
    int32_t last_terms;         /* Taylor terms of the last sub-step */
    int32_t hpsi_passes;        /* memory passes one H psi takes at this size */
    double hpsi_bytes_per_amp;  /* planned HBM bytes per amplitude per H psi (all passes) */
    double diag_min, diag_max;  /* range of E_int (whole register) */
} qsb_metrics;
int qsb_get_metrics(const qsb_ctx* ctx, qsb_metrics* out);
/* time `iters` back-to-back H psi with CUDA events on the library's stream; ms_each[i] is the
 * duration of the i-th application.  flush_l2 != 0 overwrites a >L2 scratch buffer between
 * applications.  For bench.py's roofline leg. [collective] */
int qsb_time_apply_h(qsb_ctx* ctx, double omega, double delta, int iters, int flush_l2,
                     float* ms_each);
/* same for whole sweep steps: ms_each[s] = device time of step s of the schedule */
int qsb_time_schedule(qsb_ctx* ctx, const double* omega, const double* delta, const double* dt_us,
                      int64_t n_steps, float* ms_each);

#ifdef __cplusplus
}
#endif
#endif /* QSB_H */
