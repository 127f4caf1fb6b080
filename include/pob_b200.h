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
    POB_E_COMPILE = -7        /* the layout compiler rejected the circuit shape */
};

/* flags for pob_run_batch */
enum {
    POB_RUN_EXPAND = 1u,          /* materialise every instance's full witness vector in its HBM slot */
    POB_RUN_DIGEST = 2u,          /* also compute a 64-bit digest of each materialised witness (reads it back once) */
    POB_RUN_INPUTS_STAGED = 4u    /* ignore `inputs`, use the device-resident batch set by pob_stage_inputs */
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
} pob_desc;

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

int pob_describe(const pob_handle *h, pob_desc *out);

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
 * digests: n x uint64 out, only with POB_RUN_DIGEST (may be NULL otherwise).
 * Instance i's witness lives in slot i % n_slots until overwritten by a later instance or batch. */
int pob_run_batch(pob_handle *h, const uint64_t *inputs, uint32_t n, uint32_t flags,
                  uint32_t *status, uint64_t *outputs, uint64_t *digests);

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
/* test hook for the self-check: overwrite ONE entry of a resident witness (fault injection) */
int pob_debug_poke_witness(pob_handle *h, uint32_t index, uint64_t signal, const uint64_t value[4]);

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
