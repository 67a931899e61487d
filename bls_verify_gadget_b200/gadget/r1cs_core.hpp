// Host-side R1CS builder for the reference's verify circuit: constraint system, field / boolean / byte variables.
//
// This is the producer side of the K8 kernel (csrc/r1cs.cuh): it synthesises (A, B, C) and the assignment z of the circuits
// the reference writes against ark-r1cs-std -- src/hasher.rs (hash_to_g2_with_cons) and src/constraints.rs
// (BlsSignatureVerifyGadget::verify) -- so that `blsgpu_r1cs_check` can be run on verify-shaped systems without a Rust
// toolchain.  The reference's matrices themselves can only be produced by arkworks (un-vendored): this
// builder follows the reference's gadget code line by line and arkworks' gadget semantics (one constraint per variable
// product, free linear combinations, witness-hinted inversions and comparisons), so row counts, sparsity and coefficient
// statistics are those of the real circuit, but the variable numbering is NOT arkworks' -- parity of the matrices stays
// unpinned, while the VALUES the circuit computes are pinned (they must equal the native path: KAT src/bls.rs:645).
//
// Field arithmetic is the host build of csrc/fp.cuh (Montgomery form, R = 2^384).  Single-threaded per circuit, like
// ark-relations' Rc<RefCell<ConstraintSystem>>.
#pragma once
#include <vector>
#include <array>
#include <string>
#include <algorithm>
#include <stdexcept>
#include <cstring>
#include "../csrc/stages.cuh"

namespace gadget {
using namespace bls;

inline fp fp_from_u64(uint64_t v) { fp a = fp_zero(); a.l[0] = (uint32_t)v; a.l[1] = (uint32_t)(v >> 32); return fp_to_mont(a); }
inline fp fp_from_dec(const char* s) { fp r = fp_zero(), ten = fp_from_u64(10); for (; *s; s++) r = fp_add(fp_mul(r, ten), fp_from_u64((uint64_t)(*s - '0'))); return r; }
inline std::vector<uint8_t> bytes_from_hex(const std::string& h) {
    auto nib = [](char c) { return (uint8_t)(c <= '9' ? c - '0' : (c | 32) - 'a' + 10); };
    std::vector<uint8_t> o(h.size() / 2); for (size_t i = 0; i < o.size(); i++) o[i] = (uint8_t)(nib(h[2 * i]) << 4 | nib(h[2 * i + 1])); return o;
}

// Witness-only synthesis: the matrices do not depend on the inputs, so after the first synthesis only the assignment is
// needed.  With this flag set (per thread) linear combinations stay empty and no rows are emitted; variables are allocated
// in the same order with the same values, so z is identical (checked by tests/test_gadget_circuit.py).
inline bool& witness_only_mode() { static thread_local bool f = false; return f; }

struct Term { uint32_t v; fp c; };
// linear combination over the variables (index 0 is the constant ONE); kept unsorted, merged when a row is emitted
struct LC {
    std::vector<Term> t;
    static LC var(uint32_t v) { LC l; if (!witness_only_mode()) l.t.push_back({v, fp_one()}); return l; }
    static LC constant(const fp& c) { LC l; if (!witness_only_mode() && !fp_is_zero(c)) l.t.push_back({0u, c}); return l; }
    // duplicates are merged once a combination grows (x + x style doubling would otherwise double the list every step)
    LC& operator+=(const LC& o) { t.insert(t.end(), o.t.begin(), o.t.end()); if (t.size() > 12) compact(); return *this; }
    LC operator+(const LC& o) const { LC r = *this; r += o; return r; }
    LC scaled(const fp& k) const { LC r; if (t.empty() || fp_is_zero(k)) return r; r.t.reserve(t.size()); for (auto& x : t) r.t.push_back({x.v, fp_mul(x.c, k)}); return r; }
    LC neg() const { LC r; r.t.reserve(t.size()); for (auto& x : t) r.t.push_back({x.v, fp_neg(x.c)}); return r; }
    LC operator-(const LC& o) const { LC r = *this; r += o.neg(); return r; }
    inline void compact();
};
inline void LC::compact() {
        if (t.size() < 2) { if (t.size() == 1 && fp_is_zero(t[0].c)) t.clear(); return; }
        std::stable_sort(t.begin(), t.end(), [](const Term& a, const Term& b) { return a.v < b.v; });
        size_t w = 0;
        for (size_t i = 0; i < t.size();) {
            Term acc = t[i]; size_t j = i + 1;
            while (j < t.size() && t[j].v == acc.v) { acc.c = fp_add(acc.c, t[j].c); j++; }
            if (!fp_is_zero(acc.c)) t[w++] = acc;
            i = j;
        }
        t.resize(w);
}

// Witness program (GPU witness generation): with `record_rules` set, every allocated variable also records HOW its
// value follows from earlier variables -- the circuit's gadgets allocate witnesses through a dozen primitives only, each with
// one rule kind.  The rules, replayed in allocation order by csrc/witness.cuh, reproduce the assignment on the GPU for any
// input; LC ids index a flat table (0 = the empty combination).
enum RuleKind : uint8_t {
    RULE_MULADD = 0,     // (A z)(B z) + D z            products, AND / XOR, select, packing, constants (A empty)
    RULE_INV = 1,        // 1 / (A z)                   (0 -> 0)
    RULE_NEQ = 2,        // A z != 0 ? 1 : 0
    RULE_NEQMULT = 3,    // A z != 0 ? 1 / (A z) : 1
    RULE_BIT = 4,        // bit `aux` of the canonical integer A z       (addmany results, to_bits_le)
    RULE_FP2INV = 5,     // component `aux` of 1 / (A z + u B z)
    RULE_FP12INV = 6,    // coefficient `aux` (0..11, tower order) of the inverse of the Fp12 element given by LC ids a .. a+11
    RULE_INPUT = 7       // external input slot `aux`: 0/1 pk.x/y, 2..5 sig x.c0 x.c1 y.c0 y.c1, 6 + 8 i + b = bit b of message byte i
};
struct Rule { uint8_t kind; uint8_t pad; uint16_t aux; uint32_t a, b, d; };

// z = [1, instance.., witness..] (instances must be allocated before the first witness, as in ark-relations' layout)
struct ConstraintSystem {
    bool record_rules = false; std::vector<Rule> rules; std::vector<uint64_t> lc_ptr; std::vector<uint32_t> lc_col; std::vector<fp> lc_val;
    Rule pending; bool has_pending = false;
    uint32_t add_lc(LC l) {                          // id >= 1 covers terms [lc_ptr[id-1], lc_ptr[id])
        l.compact(); if (l.t.empty()) return 0;
        for (auto& x : l.t) { lc_col.push_back(x.v); lc_val.push_back(x.c); }
        lc_ptr.push_back(lc_col.size()); return (uint32_t)(lc_ptr.size() - 1);
    }
    // scratch values of the witness program: a long combination read by many rules (the 35 result bits of an addmany, the 381
    // bits of to_bits_le) is evaluated ONCE into a scratch column (index >= SCRATCH_BASE here, remapped behind the variables at
    // export); scratch columns are not circuit variables and never reach the matrices
    static constexpr uint32_t SCRATCH_BASE = 0x80000000u;
    std::vector<Rule> scratch_rules;
    uint32_t scratch_column(const LC& l) {            // scratch column holding the value of l; returns its (unmapped) column index
        uint32_t src = add_lc(l);
        scratch_rules.push_back(Rule{RULE_MULADD, 0, 0, 0, 0, src});
        return SCRATCH_BASE + (uint32_t)(scratch_rules.size() - 1);
    }
#ifndef BLS_WIT_CHUNK
#define BLS_WIT_CHUNK 8
#endif
    // scratch column holding the value of a (compacted) combination, as a tree of BLS_WIT_CHUNK-term partial sums: a rule is
    // evaluated by ONE warp, and a general coefficient on a field-sized value costs a full Montgomery product (~1 us at one warp
    // per task), so the depth of a level is set by its longest combination
    uint32_t scratch_tree(const LC& l, bool final_terms) {
        if (l.t.size() <= BLS_WIT_CHUNK) return final_terms ? scratch_column_raw(l) : scratch_column(l);
        LC total;
        for (size_t i = 0; i < l.t.size(); i += BLS_WIT_CHUNK) {
            LC part; part.t.assign(l.t.begin() + i, l.t.begin() + std::min(l.t.size(), i + (size_t)BLS_WIT_CHUNK));
            total.t.push_back({final_terms ? scratch_column_raw(part) : scratch_column(part), fp_one()});
        }
        return scratch_tree(total, true);
    }
    uint32_t scratch_of(LC l) {                       // returns the id of the one-term combination "1 * scratch"
        l.compact(); if (l.t.empty()) return 0;
        uint32_t col = scratch_tree(l, false);
        lc_col.push_back(col); lc_val.push_back(fp_one());
        lc_ptr.push_back(lc_col.size()); return (uint32_t)(lc_ptr.size() - 1);
    }
    uint32_t scratch_column_raw(const LC& l) {        // like scratch_column, but the terms are already final (may reference scratch columns; no sorting by variable needed)
        for (auto& x : l.t) { lc_col.push_back(x.v); lc_val.push_back(x.c); }
        lc_ptr.push_back(lc_col.size());
        scratch_rules.push_back(Rule{RULE_MULADD, 0, 0, 0, 0, (uint32_t)(lc_ptr.size() - 1)});
        return SCRATCH_BASE + (uint32_t)(scratch_rules.size() - 1);
    }
    void set_rule_ids(uint8_t kind, uint16_t aux, uint32_t a, uint32_t b, uint32_t d) { if (!record_rules) return; pending = Rule{kind, 0, aux, a, b, d}; has_pending = true; }
    void set_rule(uint8_t kind, uint16_t aux, const LC* a, const LC* b = nullptr, const LC* d = nullptr) {
        if (!record_rules) return;
        // long combinations go through scratch columns (partial sums in parallel) so that no single rule walks hundreds of terms
        auto id = [&](const LC* l) { if (!l) return 0u; LC t = *l; t.compact(); return t.t.size() > 2 * BLS_WIT_CHUNK ? scratch_of(t) : add_lc(t); };
        uint32_t ia = id(a), ib = id(b), idd = id(d);
        set_rule_ids(kind, aux, ia, ib, idd);
    }
    void take_rule() {
        if (!record_rules) return;
        if (!has_pending) throw std::logic_error("witness allocated without a rule while recording the witness program");
        rules.push_back(pending); has_pending = false;
    }
    std::vector<fp> z;
    size_t num_instance = 1;                          // including the constant ONE
    std::vector<uint64_t> rowptr[3]; std::vector<uint32_t> col[3]; std::vector<fp> val[3];
    bool witness_started = false;
    ConstraintSystem() { z.push_back(fp_one()); for (int m = 0; m < 3; m++) rowptr[m].push_back(0); lc_ptr.push_back(0); rules.push_back(Rule{RULE_INPUT, 0, 0xffff, 0, 0, 0}); }
    uint32_t new_input(const fp& v) { if (witness_started) throw std::logic_error("instance variable after a witness"); take_rule(); z.push_back(v); num_instance++; return (uint32_t)(z.size() - 1); }
    uint32_t new_witness(const fp& v) { witness_started = true; take_rule(); z.push_back(v); return (uint32_t)(z.size() - 1); }
    size_t num_constraints() const { return rowptr[0].size() - 1; }
    size_t num_variables() const { return z.size(); }
    fp eval(const LC& l) const { fp s = fp_zero(); for (auto& x : l.t) s = fp_add(s, fp_mul(x.c, z[x.v])); return s; }
    void push_row(int m, LC l) { l.compact(); for (auto& x : l.t) { col[m].push_back(x.v); val[m].push_back(x.c); } rowptr[m].push_back(col[m].size()); }
    void enforce(const LC& a, const LC& b, const LC& c) { if (witness_only_mode()) return; push_row(0, a); push_row(1, b); push_row(2, c); }
    // host self-check (builder tests): first unsatisfied row or -1
    long first_unsatisfied(const std::vector<fp>* zz = nullptr) const {
        const std::vector<fp>& w = zz ? *zz : z;
        for (size_t r = 0; r + 1 < rowptr[0].size(); r++) {
            fp s[3];
            for (int m = 0; m < 3; m++) { s[m] = fp_zero(); for (uint64_t k = rowptr[m][r]; k < rowptr[m][r + 1]; k++) s[m] = fp_add(s[m], fp_mul(val[m][k], w[col[m][k]])); }
            if (!fp_eq(fp_mul(s[0], s[1]), s[2])) return (long)r;
        }
        return -1;
    }
};

// ------------------------------------------------------------------------------------------------ FpVar
struct Boolean;
struct FpVar {
    LC lc; fp val; bool cst;
    FpVar() : val(fp_zero()), cst(true) {}
    static FpVar constant(const fp& v) { FpVar r; r.lc = LC::constant(v); r.val = v; r.cst = true; return r; }
    static FpVar zero() { return constant(fp_zero()); }
    static FpVar one() { return constant(fp_one()); }
    static FpVar witness(ConstraintSystem& cs, const fp& v) { FpVar r; r.lc = LC::var(cs.new_witness(v)); r.val = v; r.cst = false; return r; }
    static FpVar witness_const(ConstraintSystem& cs, const fp& v) { LC c = LC::constant(v); cs.set_rule(RULE_MULADD, 0, nullptr, nullptr, &c); return witness(cs, v); }     // a witness whose value is a constant (z = 1 of a point, lib_str)
    static FpVar witness_input(ConstraintSystem& cs, const fp& v, uint16_t slot) { cs.set_rule(RULE_INPUT, slot, nullptr); return witness(cs, v); }
    static FpVar input(ConstraintSystem& cs, const fp& v) { FpVar r; r.lc = LC::var(cs.new_input(v)); r.val = v; r.cst = false; return r; }
    FpVar operator+(const FpVar& o) const { FpVar r; r.lc = lc + o.lc; r.val = fp_add(val, o.val); r.cst = cst && o.cst; if (r.cst) r.lc = LC::constant(r.val); return r; }
    FpVar operator-(const FpVar& o) const { FpVar r; r.lc = lc - o.lc; r.val = fp_sub(val, o.val); r.cst = cst && o.cst; if (r.cst) r.lc = LC::constant(r.val); return r; }
    FpVar neg() const { FpVar r; r.lc = lc.neg(); r.val = fp_neg(val); r.cst = cst; return r; }
    FpVar scaled(const fp& k) const { FpVar r; r.lc = lc.scaled(k); r.val = fp_mul(val, k); r.cst = cst; return r; }
    FpVar dbl() const { return *this + *this; }
    // one constraint per product of two non-constant values (ark-r1cs-std AllocatedFp::mul); constants scale the LC
    FpVar mul(ConstraintSystem& cs, const FpVar& o) const {
        if (cst) return o.scaled(val);
        if (o.cst) return scaled(o.val);
        LC a = lc, b = o.lc; a.compact(); b.compact();
        cs.set_rule(RULE_MULADD, 0, &a, &b);
        FpVar r = witness(cs, fp_mul(val, o.val));
        cs.enforce(a, b, r.lc);
        return r;
    }
    FpVar square(ConstraintSystem& cs) const { return mul(cs, *this); }
    // witness-hinted inverse: a * inv = 1 (unsatisfiable for a = 0, like AllocatedFp::inverse)
    FpVar inverse(ConstraintSystem& cs) const {
        if (cst) return constant(fp_inv(val));
        cs.set_rule(RULE_INV, 0, &lc);
        FpVar r = witness(cs, fp_inv(val));
        cs.enforce(lc, r.lc, LC::constant(fp_one()));
        return r;
    }
    void enforce_equal(ConstraintSystem& cs, const FpVar& o) const { cs.enforce(lc - o.lc, LC::constant(fp_one()), LC()); }
};

// ------------------------------------------------------------------------------------------------ Boolean
struct Boolean {
    LC lc; bool val; bool cst;
    Boolean() : val(false), cst(true) {}
    static Boolean constant(bool b) { Boolean r; r.val = b; r.cst = true; if (b) r.lc = LC::constant(fp_one()); return r; }
    static Boolean from_var(uint32_t v, bool b) { Boolean r; r.lc = LC::var(v); r.val = b; r.cst = false; return r; }
    // allocate and constrain b (1 - b) = 0
    static Boolean witness(ConstraintSystem& cs, bool b) {
        Boolean r = from_var(cs.new_witness(b ? fp_one() : fp_zero()), b);
        cs.enforce(r.lc, LC::constant(fp_one()) - r.lc, LC());
        return r;
    }
    static Boolean input(ConstraintSystem& cs, bool b) {
        Boolean r = from_var(cs.new_input(b ? fp_one() : fp_zero()), b);
        cs.enforce(r.lc, LC::constant(fp_one()) - r.lc, LC());
        return r;
    }
    Boolean not_() const { Boolean r; r.lc = LC::constant(fp_one()) - lc; r.lc.compact(); r.val = !val; r.cst = cst; return r; }
    Boolean and_(ConstraintSystem& cs, const Boolean& o) const {
        if (cst) return val ? o : constant(false);
        if (o.cst) return o.val ? *this : constant(false);
        cs.set_rule(RULE_MULADD, 0, &lc, &o.lc);
        Boolean r = from_var(cs.new_witness((val && o.val) ? fp_one() : fp_zero()), val && o.val);
        cs.enforce(lc, o.lc, r.lc);
        return r;
    }
    Boolean or_(ConstraintSystem& cs, const Boolean& o) const { return not_().and_(cs, o.not_()).not_(); }
    // (a + a) b = a + b - c   (ark-r1cs-std AllocatedBool::xor)
    Boolean xor_(ConstraintSystem& cs, const Boolean& o) const {
        if (cst) return val ? o.not_() : o;
        if (o.cst) return o.val ? not_() : *this;
        if (cs.record_rules) { LC m2 = (lc + lc).neg(), sum = lc + o.lc; cs.set_rule(RULE_MULADD, 0, &m2, &o.lc, &sum); }          // a + b - 2ab
        Boolean r = from_var(cs.new_witness((val != o.val) ? fp_one() : fp_zero()), val != o.val);
        cs.enforce(lc + lc, o.lc, lc + o.lc - r.lc);
        return r;
    }
    Boolean is_eq(ConstraintSystem& cs, const Boolean& o) const { return xor_(cs, o).not_(); }
    FpVar to_fp() const { FpVar r; r.lc = lc; r.val = val ? fp_one() : fp_zero(); r.cst = cst; return r; }
    // cond ? t : f  -- one constraint per selected field element: cond (t - f) = r - f
    FpVar select(ConstraintSystem& cs, const FpVar& t, const FpVar& f) const {
        if (cst) return val ? t : f;
        if (t.cst && f.cst) { FpVar r; r.lc = f.lc + lc.scaled(fp_sub(t.val, f.val)); r.val = val ? t.val : f.val; r.cst = false; return r; }
        if (cs.record_rules) { LC df = t.lc - f.lc; cs.set_rule(RULE_MULADD, 0, &lc, &df, &f.lc); }                                 // f + cond (t - f)
        FpVar r = FpVar::witness(cs, val ? t.val : f.val);
        cs.enforce(lc, t.lc - f.lc, r.lc - f.lc);
        return r;
    }
};
inline Boolean kary_and(ConstraintSystem& cs, const std::vector<Boolean>& v) { Boolean r = v[0]; for (size_t i = 1; i < v.size(); i++) r = r.and_(cs, v[i]); return r; }

// a == b for field variables (ark-r1cs-std FpVar::is_eq via is_neq: 2 constraints + the Boolean)
inline Boolean fp_is_eq(ConstraintSystem& cs, const FpVar& a, const FpVar& b) {
    if (a.cst && b.cst) return Boolean::constant(fp_eq(a.val, b.val));
    fp d = fp_sub(a.val, b.val); bool neq = !fp_is_zero(d);
    LC diff = a.lc - b.lc;
    cs.set_rule(RULE_NEQ, 0, &diff);
    Boolean is_neq = Boolean::witness(cs, neq);
    cs.set_rule(RULE_NEQMULT, 0, &diff);
    FpVar mult = FpVar::witness(cs, neq ? fp_inv(d) : fp_one());
    cs.enforce(diff, mult.lc, is_neq.lc);                                     // (a - b) m = is_neq
    cs.enforce(diff, is_neq.not_().lc, LC());                                 // (a - b)(1 - is_neq) = 0
    return is_neq.not_();
}

// little-endian bits of a field element, with the strict range check bits < p (FpVar::to_bits_le = non-unique bits +
// enforce_in_field_le): 381 booleans, one packing row, and the run-length comparison against p - 1.
inline std::vector<Boolean> fp_to_bits_le(ConstraintSystem& cs, const FpVar& a) {
    fp canon = fp_from_mont(a.val);
    std::vector<Boolean> bits(381);
    if (a.cst) { for (int i = 0; i < 381; i++) bits[i] = Boolean::constant((canon.l[i >> 5] >> (i & 31)) & 1); return bits; }
    LC sum; fp pw = fp_one();
    uint32_t src_id = cs.record_rules ? cs.scratch_of(a.lc) : 0;              // the source combination is evaluated once, all 381 bit rules read the scratch value
    for (int i = 0; i < 381; i++) {
        cs.set_rule_ids(RULE_BIT, (uint16_t)i, src_id, 0, 0);
        bits[i] = Boolean::witness(cs, (canon.l[i >> 5] >> (i & 31)) & 1);
        sum += bits[i].lc.scaled(pw); pw = fp_add(pw, pw);
    }
    cs.enforce(sum, LC::constant(fp_one()), a.lc);
    // enforce bits <= p - 1 (Boolean::enforce_smaller_or_equal_than_le)
    fp pm1 = fp_modulus(); pm1.l[0] -= 1;
    Boolean last_run = Boolean::constant(true); std::vector<Boolean> run;
    for (int i = 380; i >= 0; i--) {
        bool b = (pm1.l[i >> 5] >> (i & 31)) & 1;
        if (b) run.push_back(bits[i]);
        else {
            if (!run.empty()) { run.push_back(last_run); last_run = kary_and(cs, run); run.clear(); }
            cs.enforce(bits[i].lc, last_run.lc, LC());                      // a_i = 1 while all higher bits matched p - 1: a > p - 1
        }
    }
    return bits;
}

// ------------------------------------------------------------------------------------------------ UInt8 / UInt32
struct UInt8 { std::array<Boolean, 8> b; uint8_t value() const { uint8_t v = 0; for (int i = 0; i < 8; i++) v |= (uint8_t)(b[i].val << i); return v; } };   // little-endian bits
inline UInt8 u8_constant(uint8_t v) { UInt8 r; for (int i = 0; i < 8; i++) r.b[i] = Boolean::constant((v >> i) & 1); return r; }
inline UInt8 u8_witness(ConstraintSystem& cs, uint8_t v) { UInt8 r; for (int i = 0; i < 8; i++) r.b[i] = Boolean::witness(cs, (v >> i) & 1); return r; }
inline UInt8 u8_witness_input(ConstraintSystem& cs, uint8_t v, uint16_t slot0) { UInt8 r; for (int i = 0; i < 8; i++) { cs.set_rule(RULE_INPUT, (uint16_t)(slot0 + i), nullptr); r.b[i] = Boolean::witness(cs, (v >> i) & 1); } return r; }
inline UInt8 u8_witness_const(ConstraintSystem& cs, uint8_t v) { UInt8 r; for (int i = 0; i < 8; i++) { LC c = LC::constant(((v >> i) & 1) ? fp_one() : fp_zero()); cs.set_rule(RULE_MULADD, 0, nullptr, nullptr, &c); r.b[i] = Boolean::witness(cs, (v >> i) & 1); } return r; }
inline UInt8 u8_input(ConstraintSystem& cs, uint8_t v) { UInt8 r; for (int i = 0; i < 8; i++) r.b[i] = Boolean::input(cs, (v >> i) & 1); return r; }
inline UInt8 u8_xor(ConstraintSystem& cs, const UInt8& a, const UInt8& c) { UInt8 r; for (int i = 0; i < 8; i++) r.b[i] = a.b[i].xor_(cs, c.b[i]); return r; }
inline std::vector<UInt8> u8_constant_vec(const uint8_t* p, size_t n) { std::vector<UInt8> v(n); for (size_t i = 0; i < n; i++) v[i] = u8_constant(p[i]); return v; }

struct UInt32 {
    std::array<Boolean, 32> b;                                            // little-endian bits
    uint32_t value() const { uint32_t v = 0; for (int i = 0; i < 32; i++) v |= (uint32_t)b[i].val << i; return v; }
    bool is_constant() const { for (auto& x : b) if (!x.cst) return false; return true; }
    static UInt32 constant(uint32_t v) { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = Boolean::constant((v >> i) & 1); return r; }
    UInt32 rotr(int n) const { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = b[(i + n) & 31]; return r; }
    UInt32 shr(int n) const { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = i + n < 32 ? b[i + n] : Boolean::constant(false); return r; }
    UInt32 xor_(ConstraintSystem& cs, const UInt32& o) const { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = b[i].xor_(cs, o.b[i]); return r; }
    UInt32 and_(ConstraintSystem& cs, const UInt32& o) const { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = b[i].and_(cs, o.b[i]); return r; }
    UInt32 not_() const { UInt32 r; for (int i = 0; i < 32; i++) r.b[i] = b[i].not_(); return r; }
    LC lc() const { LC s; fp pw = fp_one(); for (int i = 0; i < 32; i++) { if (!b[i].lc.t.empty()) s += b[i].lc.scaled(pw); pw = fp_add(pw, pw); } return s; }
};
// sum of n words modulo 2^32 (ark-r1cs-std UInt32::addmany): the result bits (32 + carry bits) are fresh booleans and one
// row ties sum(operands) to sum(result bits)
inline UInt32 u32_addmany(ConstraintSystem& cs, const std::vector<UInt32>& ops) {
    uint64_t total = 0; bool all_const = true;
    for (auto& o : ops) { total += o.value(); all_const &= o.is_constant(); }
    if (all_const) return UInt32::constant((uint32_t)total);
    int nbits = 32; { size_t n = ops.size(); uint64_t maxv = n * 0xffffffffull; while ((maxv >> nbits) != 0) nbits++; }
    LC sum; for (auto& o : ops) sum += o.lc();
    LC res; fp pw = fp_one(); UInt32 r;
    uint32_t sum_id = cs.record_rules ? cs.scratch_of(sum) : 0;
    for (int i = 0; i < nbits; i++) {
        cs.set_rule_ids(RULE_BIT, (uint16_t)i, sum_id, 0, 0);
        Boolean bit = Boolean::witness(cs, (total >> i) & 1);
        res += bit.lc.scaled(pw); pw = fp_add(pw, pw);
        if (i < 32) r.b[i] = bit;
    }
    cs.enforce(sum, LC::constant(fp_one()), res);
    return r;
}

}  // namespace gadget
