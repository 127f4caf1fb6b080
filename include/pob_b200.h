/* pob_b200.h -- C ABI of the B200-native batched witness generator for worm-privacy/proof-of-burn.
 *
 * Drop-in boundary.  The reference has no library interface for this path: witness generation is the
 * process CLI the circom toolchain emits, `./<circuit> input.json witness.wtns`
 * (reference Makefile:5-6, tests/test.py:60-63), with the circuit identity fixed at compile time by
 * `component main = ...` (circuits/main_proof_of_burn.circom:27, circuits/main_spend.circom:6).  The entry
 * points below are what an FFI for that path binds: fix a circuit shape once (pob_create == `circom -c` +
 * `make`), then push batches of inputs through it (pob_run_batch == N runs of `./<circuit>`), and read back
 * per-instance accept/reject + output signals, optionally a full `.wtns`.  Plain pointers and sizes only.
 *
 * Field elements cross the ABI as 4 x uint64 little-endian limbs, canonical (< p, non-Montgomery): the same
 * 32-byte form the .wtns file stores.  JSON parsing (the schema tests/main.py:160-178 emits) lives in the
 * Python host (proof-of-burn_b200/pob_b200), which flattens inputs in declaration order
 * (circuits/proof_of_burn.circom:43-72, circuits/spend.circom:33-36).
 *
 * Errors: every function returns 0 on success and a negative POB_E_* code otherwise; pob_last_error()
 * returns a message for the calling thread.  There is NO CPU fallback: without a CUDA device pob_create
 * fails with POB_E_NO_DEVICE.  A failed circuit constraint is not an API error: it is reported per
 * instance in status[] (reference: any failed `===` aborts the calculator, tests/test.py:65-68).
 */
#ifndef POB_B200_H
#define POB_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pob_handle pob_handle;

enum {
    POB_OK = 0,
    POB_E_BAD_ARG = -1,       /* null pointer, unknown template, wrong parameter count, n == 0 ... */
    POB_E_NO_DEVICE = -2,     /* no CUDA device / device index out of range */
    POB_E_CUDA = -3,          /* a CUDA runtime call failed (message in pob_last_error) */
    POB_E_NO_MEMORY = -4,     /* not even one witness slot fits in free HBM */
    POB_E_RANGE = -5,         /* instance index not resident / offset out of range */
    POB_E_IO = -6,            /* file write failed */
    POB_E_COMPILE = -7,       /* the layout compiler rejected the circuit shape */
    POB_E_REJECTED = -8,      /* the instance failed a circuit constraint: it has no witness (reference tests/test.py:65-68) */
    POB_E_BUSY = -9,          /* every witness slot is held by the consumer: release one first / a batch is in flight */
    POB_DONE = 1              /* pob_acquire: no further witness in this batch (not an error) */
};

/* flags for pob_run_batch */
enum {
    POB_RUN_EXPAND = 1u,          /* materialise every instance's full witness vector in its HBM slot */
    POB_RUN_DIGEST = 2u,          /* also compute a 64-bit digest of each materialised witness (reads it back once) */
    POB_RUN_INPUTS_STAGED = 4u,   /* ignore `inputs`, use the device-resident batch set by pob_stage_inputs */
    POB_RUN_DISCARD = 8u          /* generation-only run: with n > n_slots earlier witnesses are overwritten UNREAD by later
                                     ones (throughput measurement of the path itself).  Without this flag, a digest or a
                                     retain list, pob_run_batch refuses n > n_slots: use pob_submit/pob_acquire/pob_release. */
};

typedef struct {
    uint64_t n_signals;        /* witness entries incl. witness[0] = 1 (== nWitness in the .wtns header) */
    uint32_t n_outputs;        /* main output signals = witness[1 .. n_outputs] */
    uint32_t n_inputs;         /* scalar input signals in declaration order */
    uint64_t witness_bytes;    /* 32 * n_signals */
    uint64_t wtns_file_bytes;  /* 76 + 32 * n_signals */
    uint64_t store_bytes;      /* compact per-instance evaluation store */
    uint64_t n_ops;            /* thread ops per instance */
    uint32_t n_absorbs;        /* Keccak absorbs (= Keccak-f permutations) per instance */
    uint32_t n_levels;         /* dependency levels of the eval program */
    uint32_t n_tiles;          /* expand tiles per instance */
    uint32_t n_slots;          /* witness slots resident in HBM (0 when no device handle) */
    uint32_t chunk;            /* instances evaluated per eval launch */
    uint32_t expand_group;     /* instances materialised per expand launch (<= n_slots) */
    uint32_t opt_level;        /* 0 = circom --O0 witness (every signal), 1 = reduced witness (POB_CREATE_O1) */
    uint64_t n_signals_o0;     /* witness entries of the --O0 layout (== n_signals when opt_level == 0) */
} pob_desc;

/* `hcreate` argument of pob_create / pob_layout_info / pob_constraint_info: bit 0 = creation-order sub-component numbering
 * (SURVEY.md App. C R3), POB_CREATE_O1 = produce the REDUCED witness the circom simplifier's `--O1` level implies (the reference
 * deploys through circom's default simplifier: .github/workflows/circuitscan.yml:29,36; its Makefile builds --O0): every signal
 * tied to an earlier signal or to a constant by a `signal = signal` / `signal = constant` constraint is dropped, main inputs and
 * outputs always stay, the rest keeps its --O0 order.  main_proof_of_burn: 215,907,954 -> see pob_desc.n_signals.  Which member
 * of an equality class circom itself keeps is not pinned by the reference: the reduced ORDER is "parity unpinned" like the
 * --O0 order; every retained value equals the --O0 witness through pob_witness_map (tested). */
enum { POB_CREATE_HCREATE = 1, POB_CREATE_O1 = 0x100 };

/* replaces: `circom -c <main>.circom --O0 && make` (reference Makefile:2-3, tests/test.py:32,55).
 * main_name/params = the `component main = Name(p0, p1, ...)` expression; params are nparams x 4 limbs.
 * hcreate: 0 = completion-order sub-component numbering (default), 1 = creation order.
 * max_slots: upper bound on resident witness slots (0 = as many as fit in 80 % of free HBM). */
int pob_create(const char *main_name, const uint64_t *params, int nparams, int hcreate, int device,
               uint32_t max_slots, pob_handle **out);
void pob_destroy(pob_handle *h);

/* host-only: run the layout compiler and report the shape; needs no GPU (used by the CPU test-suite) */
int pob_layout_info(const char *main_name, const uint64_t *params, int nparams, int hcreate, pob_desc *out);
/* input schema "name[d0][d1],name2,..." with dims as expressions over p0..p7; NULL if unknown template */
const char *pob_input_schema(const char *main_name, int *nparams);

/* host-only, order-pinning kit (the reference pins no witness ORDER: it holds no .sym / .wtns).  Writes
 * one line `first_signal,n_own_signals,template` per component instance of the --O0 layout, in numbering order, for
 * tools/diff_sym.py to compare with the `.sym` of a real `circom --O0 --sym` build.  *n_components receives the line count. */
int pob_write_components(const char *main_name, const uint64_t *params, int nparams, int hcreate, const char *path, uint64_t *n_components);

int pob_describe(const pob_handle *h, pob_desc *out);
/* reduced witness only: map[k] = --O0 signal index of reduced witness entry k (n_signals entries) */
int pob_witness_map(const pob_handle *h, uint32_t *map);

/* pinned host memory for the caller's input / output arrays (so H2D/D2H inside pob_run_batch are async DMA) */
void *pob_alloc_pinned(uint64_t bytes);
void pob_free_pinned(void *p);

/* copy a batch of inputs (n x n_inputs x 4 limbs) to HBM once; later runs with POB_RUN_INPUTS_STAGED reuse it */
int pob_stage_inputs(pob_handle *h, const uint64_t *inputs, uint32_t n);

/* replaces: n runs of `./<circuit> input.json witness.wtns` (reference Makefile:5-6, tests/test.py:60-63).
 * inputs : n x n_inputs x 4 limbs (host; pinned preferred), flattened in declaration order.
 * status : n x uint32 out; 0 = every constraint holds, else 1 + witness index of the first signal of the
 *          lowest-numbered component that owns a failing constraint.
 * outputs: n x n_outputs x 4 limbs out (may be NULL).
 * digests: n x uint64 out, only with POB_RUN_DIGEST (may be NULL otherwise); 0 for a rejected instance.
 * A rejected instance (status != 0) contributes NO witness: its slot is not written and every accessor below
 * answers POB_E_REJECTED for it (the reference calculator aborts without a usable witness, tests/test.py:65-68).
 * With POB_RUN_EXPAND instance i lives in slot i % n_slots.  n > n_slots is only accepted together with
 * POB_RUN_DIGEST (every witness is consumed by the digest kernel before its slot is reused) or POB_RUN_DISCARD;
 * afterwards the last n_slots instances are resident. */
int pob_run_batch(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags,
                  uint32_t *status, uint64_t *outputs, uint64_t *digests);
/* the same, but only the instances retain[0..n_retain) (strictly ascending, n_retain <= n_slots) are materialised:
 * all n instances are evaluated (status, outputs), the retained ones stay resident, the others cost no HBM traffic
 * (SURVEY.md 8(b) "which indices to retain"). */
int pob_run_batch_retain(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags,
                         const uint32_t *retain, uint32_t n_retain,
                         uint32_t *status, uint64_t *outputs, uint64_t *digests);

/* ---- consumer-paced hand-off: no witness is ever overwritten unread ----------------------
 * The reference's contract is one run -> one witness.wtns the prover reads (Makefile:5-6).  Here the consumer (an
 * on-GPU prover stage, an exporter) takes the witnesses of a batch one by one, in instance order, and hands each
 * slot back when it is done; generation stalls -- on the GPU, stream-ordered -- while all slots are held.
 *   pob_submit   start a batch (POB_RUN_EXPAND implied); returns at once, work is queued as slots allow.
 *   pob_acquire  next instance: POB_OK (*index, *dptr = its resident witness), POB_E_REJECTED (*index set, no witness,
 *                nothing to release), POB_DONE (batch exhausted), POB_E_BUSY (release a slot first).
 *                consumer_stream == NULL: returns when the witness is complete in HBM (host wait);
 *                else (a cudaStream_t): returns at once and makes that stream wait for the witness on the GPU.
 *   pob_release  the consumer is done with instance `index`; consumer_stream (or NULL = already finished on the
 *                host) orders the reuse of the slot after the consumer's queued work.
 *   pob_finish   drain the batch (instances never acquired are generated and dropped), return status/outputs/digests.
 * One batch at a time per handle; the last n_slots non-released instances stay resident after pob_finish. */
int pob_submit(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags);
int pob_acquire(pob_handle *h, uint32_t *index, void **dptr, void *consumer_stream);
int pob_release(pob_handle *h, uint32_t index, void *consumer_stream);
int pob_finish(pob_handle *h, uint32_t *status, uint64_t *outputs, uint64_t *digests);

/* replaces: n runs of `./<circuit> input_i.json witness_i.wtns` INCLUDING the files: every accepted instance is
 * exported through the consumer-paced path above -- D2H over several copy streams into a pinned staging ring, a writer
 * thread per batch doing the file I/O -- so generation, PCIe transfer and disk writes overlap and no witness is dropped.
 * paths: n entries; NULL entry (or paths == NULL) = transfer to host memory only (PCIe line-rate measurement). */
typedef struct {
    uint64_t witnesses;        /* exported instances */
    uint64_t bytes;            /* .wtns bytes moved to the host */
    float total_ms;            /* wall time of the call */
    float d2h_gbs;             /* bytes / total time */
} pob_export_stats;
int pob_export_batch(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags, const char *const *paths,
                     uint32_t *status, uint64_t *outputs, pob_export_stats *stats);

/* device timings of the last pob_run_batch (CUDA events on the library's own streams), and kernel count */
typedef struct {
    float total_ms;          /* first H2D / first kernel -> last kernel of the batch */
    float expand_ms;         /* sum of expand-kernel durations */
    float eval_ms;           /* sum of eval-kernel durations */
    uint32_t expand_launches, eval_launches, other_launches;
    uint64_t h2d_bytes, d2h_bytes;
} pob_timing;
int pob_last_timing(const pob_handle *h, pob_timing *out);

/* replaces: the `witness.wtns` the calculator writes (layout: SURVEY.md Appendix B).  `index` is the
 * instance index within the last batch; it must still be resident. */
int pob_copy_witness(pob_handle *h, uint32_t index, uint64_t first_signal, uint64_t n_signals, uint64_t *dst_host);
int pob_write_wtns(pob_handle *h, uint32_t index, const char *path);
/* device pointer of a resident witness (for an on-GPU consumer such as a prover's first stage) */
int pob_witness_device_ptr(pob_handle *h, uint32_t index, void **dptr);

/* ---- on-GPU self-check of a resident witness (Keccak part) -----------------------------------
 * Independent of the layout tables that produced the witness: for every KeccakfRound component (95.8 % of the entries)
 * the kernel reads the component's own `in[25][64]` and `out[25][64]` signals straight from the witness (positions fixed
 * by circuits/utils/keccak.circom:290-297: out first, then in), recomputes one textbook Keccak-f round (theta, rho-pi,
 * chi, iota with the round index the block has inside its Keccakf) and compares; it also checks that all 3200 entries
 * are 0 or 1.  *n_blocks = blocks examined, *n_bad = blocks that fail. */
int pob_selfcheck_keccak(pob_handle *h, uint32_t index, uint64_t *n_blocks, uint64_t *n_bad);

/* ---- full constraint self-check ------------------------------------------------------------------
 * Evaluates EVERY constraint of the circuit against resident witness `index`, on the GPU.  The constraint system is written
 * from the circom sources statement by statement -- every `<==` and `===` of the include closure: the gate equations of
 * circuits/utils/keccak.circom:58-297 (out = a + b - 2ab, ...), circomlib/circuits/poseidon.circom:5-65 (Sigma, Ark, Mix,
 * MixS, MixLast), bitify.circom:33,38 (Num2Bits bits and sum), comparators.circom:32-33 (IsZero), selector.circom:33-45,
 * substring_check.circom:46-99, ... -- over witness INDICES, independently of the program that produced the values.  It
 * reads 100 % of the witness entries.  `hint` records additionally pin the signals the circuit itself leaves free
 * (`inv <-- in != 0 ? 1/in : 0`; the never-assigned temp[] of merkle_patricia_trie_leaf.circom:76) to the values the
 * reference calculator writes.  The first call compiles and uploads the constraint system (seconds for the main shape). */
typedef struct {
    uint64_t n_constraints;     /* circuit constraints evaluated (the shared KeccakfRound set counted once per round block) */
    uint64_t n_nonlinear;       /* of those, A*B = C records with a non-empty A (an .r1cs would call them non-linear) */
    uint64_t n_hints;           /* hint records evaluated */
    uint64_t n_failed;          /* failing circuit constraints */
    uint64_t n_hint_failed;     /* failing hint records */
    uint64_t first_failed;      /* smallest failing record id, UINT64_MAX when everything holds */
    uint64_t signals_read;      /* distinct witness entries referenced by at least one record (== n_signals) */
    float ms;                   /* device time of the check */
} pob_check_report;
int pob_selfcheck(pob_handle *h, uint32_t index, pob_check_report *out);
/* host-only: compile the constraint system of a circuit shape and report its size (no GPU needed) */
int pob_constraint_info(const char *main_name, const uint64_t *params, int nparams, int hcreate, pob_check_report *out);

/* ---- the step just before the path ------------------------------------------------------
 * replaces: find_burn_key() of the reference input generator (tests/main.py:47-56): starting at start_key, find the
 * first burnKey >= start_key whose keccak256(burnKey[32 BE] | revealAmount[32 BE] | burnExtraCommitment[32 BE] |
 * "EIP-7503") begins with `zero_bytes` zero bytes (circuits/utils/proof_of_work.circom:54-81).  Searches at most
 * max_tries consecutive keys on the GPU; *tries receives the number of keys up to and including the hit.
 * Returns POB_E_RANGE when no key in the window satisfies the check. */
int pob_pow_grind(int device, const uint64_t start_key[4], const uint64_t reveal_amount[4], const uint64_t burn_extra_commitment[4],
                  uint32_t zero_bytes, uint64_t max_tries, uint64_t found_key[4], uint64_t *tries);

const char *pob_last_error(void);
const char *pob_version(void);

#ifdef __cplusplus
}
#endif
#endif
