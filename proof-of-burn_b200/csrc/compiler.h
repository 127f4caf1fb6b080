// compiler.h -- layout compiler: circuit shape -> witness program (host C++, runs once per shape at create time).
//
// This is the B200-native counterpart of what `circom -c ... --O0` does for the reference (Makefile:2-3):
// it fixes the witness numbering (SURVEY.md Appendix C rules R1-R4) and emits the program the GPU executes.
// It never sees instance data and computes no witness values.
#pragma once
#include <string>
#include <vector>
#include "program.h"

namespace pob {

// main_name / params name the `component main = Name(params...)` expression, e.g. ("ProofOfBurn",
// {16,4,16,50,31,2,10^19,10^20}) for circuits/main_proof_of_burn.circom:27 or ("Spend", {31}) for
// circuits/main_spend.circom:6; every gadget the reference's suites instantiate (tests/test.py:146-201) is
// accepted too.  hcreate selects creation-order numbering at the two sites where it differs from
// completion order (Num2Bits_strict, MultiAND n>=3).  Throws std::runtime_error on unknown template/shape.
// want_constraints additionally emits the circuit's constraint system (Program::cons_*; see program.h).
// opt_level 1 produces the reduced (`--O1`-style) witness program (program.h: Program::witness_map).
Program compile_circuit(const std::string &main_name, const std::vector<Fr> &params, bool hcreate, bool want_constraints = false, int opt_level = 0);

// order-pinning kit: writes `first_signal,n_own_signals,template` for every component instance in numbering order; returns the count
uint64_t write_components(const std::string &main_name, const std::vector<Fr> &params, bool hcreate, const std::string &path);

// Input schema of a main template ("name[d0][d1],name2,...", dims are expressions over p0..p7) or nullptr.
const char *main_input_schema(const std::string &main_name, int *nparams);

// inverses of 0..INV_TABLE_N-1 (entry 0 = 0), computed with one inversion (Montgomery's trick)
std::vector<Fr> build_inverse_table();

}  // namespace pob
