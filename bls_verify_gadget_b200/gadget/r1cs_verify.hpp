// The verify circuit: host C++ mirror of BlsSignatureVerifyGadget::verify (reference src/constraints.rs:90-128) on the
// builder of r1cs_core.hpp.  The pairing part restates ark-r1cs-std's bls12 PairingVar: G2 points are "prepared" into
// affine line coefficients (slope supplied as a witness-hinted inverse), the Miller loop multiplies the sparse lines into
// an Fp12Var with mul_by_014, and the final exponentiation is the chain of ark-ec (the one csrc/pairing.cuh runs
// natively).  The affine lines differ from the native projective lines by Fp2 factors, which the final exponentiation
// removes: the GT VALUE the circuit computes must equal the native one (checked by the tests), and the Boolean it returns
// must reproduce the reference's own expectations (constraints.rs:326-332: true, false, false).
#pragma once
#include "r1cs_hasher.hpp"

namespace gadget {

// ------------------------------------------------------------------------------------------------ Fp6Var / Fp12Var
struct Fp6Var {
    Fp2Var c0, c1, c2;
    static Fp6Var constant(const fp6& v) { Fp6Var r; r.c0 = Fp2Var::constant(v.c0); r.c1 = Fp2Var::constant(v.c1); r.c2 = Fp2Var::constant(v.c2); return r; }
    static Fp6Var witness(ConstraintSystem& cs, const fp6& v) { Fp6Var r; r.c0 = Fp2Var::witness(cs, v.c0); r.c1 = Fp2Var::witness(cs, v.c1); r.c2 = Fp2Var::witness(cs, v.c2); return r; }
    fp6 value() const { fp6 v; v.c0 = c0.value(); v.c1 = c1.value(); v.c2 = c2.value(); return v; }
    Fp6Var operator+(const Fp6Var& o) const { Fp6Var r; r.c0 = c0 + o.c0; r.c1 = c1 + o.c1; r.c2 = c2 + o.c2; return r; }
    Fp6Var operator-(const Fp6Var& o) const { Fp6Var r; r.c0 = c0 - o.c0; r.c1 = c1 - o.c1; r.c2 = c2 - o.c2; return r; }
    Fp6Var neg() const { Fp6Var r; r.c0 = c0.neg(); r.c1 = c1.neg(); r.c2 = c2.neg(); return r; }
    Fp6Var mul_v() const { Fp6Var r; r.c0 = c2.mul_xi(); r.c1 = c0; r.c2 = c1; return r; }
    Fp6Var mul(ConstraintSystem& cs, const Fp6Var& b) const {                      // Karatsuba over the cubic extension: 6 Fp2 products
        Fp2Var v0 = c0.mul(cs, b.c0), v1 = c1.mul(cs, b.c1), v2 = c2.mul(cs, b.c2);
        Fp2Var t0 = (c1 + c2).mul(cs, b.c1 + b.c2) - v1 - v2;
        Fp2Var t1 = (c0 + c1).mul(cs, b.c0 + b.c1) - v0 - v1;
        Fp2Var t2 = (c0 + c2).mul(cs, b.c0 + b.c2) - v0 - v2;
        Fp6Var r; r.c0 = v0 + t0.mul_xi(); r.c1 = t1 + v2.mul_xi(); r.c2 = t2 + v1; return r;
    }
    Fp6Var mul_by_01(ConstraintSystem& cs, const Fp2Var& b0, const Fp2Var& b1) const {
        Fp2Var v0 = c0.mul(cs, b0), v1 = c1.mul(cs, b1);
        Fp2Var t0 = (c1 + c2).mul(cs, b1) - v1;
        Fp2Var t1 = (c0 + c1).mul(cs, b0 + b1) - v0 - v1;
        Fp2Var t2 = (c0 + c2).mul(cs, b0) - v0 + v1;
        Fp6Var r; r.c0 = v0 + t0.mul_xi(); r.c1 = t1; r.c2 = t2; return r;
    }
    Fp6Var mul_by_1(ConstraintSystem& cs, const Fp2Var& b1) const { Fp6Var r; r.c0 = c2.mul(cs, b1).mul_xi(); r.c1 = c0.mul(cs, b1); r.c2 = c1.mul(cs, b1); return r; }
};
struct Fp12Var {
    Fp6Var c0, c1;
    static Fp12Var constant(const fp12& v) { Fp12Var r; r.c0 = Fp6Var::constant(v.c0); r.c1 = Fp6Var::constant(v.c1); return r; }
    static Fp12Var one() { fp12 o; fp12_one(o); return constant(o); }
    static Fp12Var witness(ConstraintSystem& cs, const fp12& v) { Fp12Var r; r.c0 = Fp6Var::witness(cs, v.c0); r.c1 = Fp6Var::witness(cs, v.c1); return r; }
    fp12 value() const { fp12 v; v.c0 = c0.value(); v.c1 = c1.value(); return v; }
    Fp2Var* coeff(int k) { Fp2Var* t[6] = {&c0.c0, &c0.c1, &c0.c2, &c1.c0, &c1.c1, &c1.c2}; return t[k]; }
    Fp12Var mul(ConstraintSystem& cs, const Fp12Var& b) const {
        Fp6Var t0 = c0.mul(cs, b.c0), t1 = c1.mul(cs, b.c1), t2 = (c0 + c1).mul(cs, b.c0 + b.c1);
        Fp12Var r; r.c1 = t2 - t0 - t1; r.c0 = t0 + t1.mul_v(); return r;
    }
    Fp12Var square(ConstraintSystem& cs) const {                                    // complex squaring: 2 Fp6 products
        Fp6Var ab = c0.mul(cs, c1), s = (c0 + c1).mul(cs, c0 + c1.mul_v());
        Fp12Var r; r.c0 = s - ab - ab.mul_v(); r.c1 = ab + ab; return r;
    }
    Fp12Var mul_by_014(ConstraintSystem& cs, const Fp2Var& d0, const Fp2Var& d1, const Fp2Var& d4) const {
        Fp6Var t0 = c0.mul_by_01(cs, d0, d1), t1 = c1.mul_by_1(cs, d4);
        Fp6Var s = (c0 + c1).mul_by_01(cs, d0, d1 + d4);
        Fp12Var r; r.c1 = s - t0 - t1; r.c0 = t0 + t1.mul_v(); return r;
    }
    Fp12Var unitary_inverse() const { Fp12Var r; r.c0 = c0; r.c1 = c1.neg(); return r; }                 // conjugation
    Fp12Var inverse(ConstraintSystem& cs) const {                                    // witness-hinted: a * inv = 1
        fp12 v = value(), iv; fp12_inv(iv, v);
        Fp12Var inv, me = *this; uint32_t first = 0;
        if (cs.record_rules) for (int k = 0; k < 6; k++) { LC l0 = me.coeff(k)->c0.lc, l1 = me.coeff(k)->c1.lc; l0.compact(); l1.compact();
            // ids must be consecutive and non-empty: an (impossible here) empty combination would break the 12-id window
            uint32_t i0 = cs.add_lc(l0), i1 = cs.add_lc(l1); if (!i0 || !i1) throw std::logic_error("Fp12 inverse of a value with a constant-zero coefficient"); if (!k) first = i0; }
        { const fp* ivf = &iv.c0.c0.c0; for (int k = 0; k < 6; k++) for (int j = 0; j < 2; j++) { cs.set_rule_ids(RULE_FP12INV, (uint16_t)(2 * k + j), first, 0, 0); (j ? inv.coeff(k)->c1 : inv.coeff(k)->c0) = FpVar::witness(cs, ivf[2 * k + j]); } }
        Fp12Var prod = mul(cs, inv), o = one();
        for (int k = 0; k < 6; k++) prod.coeff(k)->enforce_equal(cs, *o.coeff(k));
        return inv;
    }
    Fp12Var frobenius_map(int power) const {                                         // 1 or 2: constants only, no constraints
        Fp12Var r = *this; const Fp12Var& a = *this;
        const fp2 F1[6] = BLS_C_FROB1; const fp F2[6] = BLS_C_FROB2;
        auto m1 = [&](const Fp2Var& x, int k) { return x.conj().mul_cst(F1[k]); };
        auto m2 = [&](const Fp2Var& x, int k) { Fp2Var t; t.c0 = x.c0.scaled(F2[k]); t.c1 = x.c1.scaled(F2[k]); return t; };
        if (power == 1) { r.c0.c0 = a.c0.c0.conj(); r.c1.c0 = m1(a.c1.c0, 1); r.c0.c1 = m1(a.c0.c1, 2); r.c1.c1 = m1(a.c1.c1, 3); r.c0.c2 = m1(a.c0.c2, 4); r.c1.c2 = m1(a.c1.c2, 5); }
        else { r.c1.c0 = m2(a.c1.c0, 1); r.c0.c1 = m2(a.c0.c1, 2); r.c1.c1 = m2(a.c1.c1, 3); r.c0.c2 = m2(a.c0.c2, 4); r.c1.c2 = m2(a.c1.c2, 5); }
        return r;
    }
    // Granger-Scott squaring in the cyclotomic subgroup (Fp12Var::cyclotomic_square)
    Fp12Var cyclotomic_square(ConstraintSystem& cs) const {
        auto fp4_sqr = [&](Fp2Var& t0, Fp2Var& t1, const Fp2Var& a, const Fp2Var& b) {
            Fp2Var ab = a.mul(cs, b), s = (a + b).mul(cs, a + b.mul_xi());
            t0 = s - ab - ab.mul_xi(); t1 = ab.dbl();
        };
        Fp2Var t0, t1, t2, t3, t4, t5;
        fp4_sqr(t0, t1, c0.c0, c1.c1); fp4_sqr(t2, t3, c1.c0, c0.c2); fp4_sqr(t4, t5, c0.c1, c1.c2);
        auto comb = [](const Fp2Var& t, const Fp2Var& a, bool plus) { Fp2Var z = plus ? t + a : t - a; return z.dbl() + t; };
        Fp12Var r; Fp2Var x5 = t5.mul_xi();
        r.c0.c0 = comb(t0, c0.c0, false); r.c1.c1 = comb(t1, c1.c1, true); r.c1.c0 = comb(x5, c1.c0, true);
        r.c0.c2 = comb(t4, c0.c2, false); r.c0.c1 = comb(t2, c0.c1, false); r.c1.c2 = comb(t3, c1.c2, true);
        return r;
    }
    Boolean is_one(ConstraintSystem& cs) const {
        Fp12Var o = one(), me = *this; std::vector<Boolean> eq;
        for (int k = 0; k < 6; k++) eq.push_back(me.coeff(k)->is_eq(cs, *o.coeff(k)));
        return kary_and(cs, eq);
    }
};

// ------------------------------------------------------------------------------------------------ PairingVar (bls12)
struct G1AffineVar { FpVar x, y; };
struct LCoeff { Fp2Var c0, c1; };
struct G2PreparedVar { std::vector<LCoeff> ell_coeffs; };

struct PairingVar {
    // G2PreparedVar::double: slope 3x^2 / 2y with the inverse of y as a witness
    static LCoeff g2_double(ConstraintSystem& cs, Fp2Var& rx, Fp2Var& ry, const fp& two_inv) {
        Fp2Var a = ry.inverse(cs);
        Fp2Var b = rx.square(cs); Fp2Var bh; bh.c0 = b.c0.scaled(two_inv); bh.c1 = b.c1.scaled(two_inv); b = b + bh;
        Fp2Var c = a.mul(cs, b);
        Fp2Var x3 = c.square(cs) - rx.dbl();
        Fp2Var e = c.mul(cs, rx) - ry;
        Fp2Var y3 = e - c.mul(cs, x3);
        rx = x3; ry = y3;
        return {e, c.neg()};
    }
    static LCoeff g2_add(ConstraintSystem& cs, Fp2Var& rx, Fp2Var& ry, const Fp2Var& qx, const Fp2Var& qy) {
        Fp2Var a = (qx - rx).inverse(cs);
        Fp2Var c = a.mul(cs, qy - ry);
        Fp2Var x3 = c.square(cs) - (rx + qx);
        Fp2Var y3 = (rx - x3).mul(cs, c) - ry;
        Fp2Var g = c.mul(cs, rx) - ry;
        rx = x3; ry = y3;
        return {g, c.neg()};
    }
    // to_affine of the homogeneous point (the inverse of z proves it is not the identity) and the coefficient schedule
    static G2PreparedVar prepare_g2(ConstraintSystem& cs, const G2Var& q) {
        Fp2Var zi = q.z.inverse(cs), qx = q.x.mul(cs, zi), qy = q.y.mul(cs, zi);
        fp two_inv = fp_two_inv();
        G2PreparedVar p; Fp2Var rx = qx, ry = qy;
        const uint64_t x = BLS_X_ABS;
        for (int i = 62; i >= 0; i--) {
            p.ell_coeffs.push_back(g2_double(cs, rx, ry, two_inv));
            if ((x >> i) & 1) p.ell_coeffs.push_back(g2_add(cs, rx, ry, qx, qy));
        }
        return p;
    }
    static void ell(ConstraintSystem& cs, Fp12Var& f, const LCoeff& co, const G1AffineVar& p) {                    // M-type twist
        Fp2Var c1; c1.c0 = co.c1.c0.mul(cs, p.x); c1.c1 = co.c1.c1.mul(cs, p.x);
        Fp2Var c2; c2.c0 = p.y; c2.c1 = FpVar::zero();
        f = f.mul_by_014(cs, co.c0, c1, c2);
    }
    static Fp12Var miller_loop(ConstraintSystem& cs, const std::vector<G1AffineVar>& ps, const std::vector<G2PreparedVar>& qs) {
        std::vector<size_t> pos(ps.size(), 0);
        Fp12Var f = Fp12Var::one();
        const uint64_t x = BLS_X_ABS;
        for (int i = 62; i >= 0; i--) {
            f = f.square(cs);
            for (size_t k = 0; k < ps.size(); k++) ell(cs, f, qs[k].ell_coeffs[pos[k]++], ps[k]);
            if ((x >> i) & 1) for (size_t k = 0; k < ps.size(); k++) ell(cs, f, qs[k].ell_coeffs[pos[k]++], ps[k]);
        }
        return f.unitary_inverse();                                   // x < 0
    }
    static Fp12Var exp_by_x(ConstraintSystem& cs, const Fp12Var& a) {
        Fp12Var acc = a; const uint64_t x = BLS_X_ABS;
        for (int i = 62; i >= 0; i--) { acc = acc.cyclotomic_square(cs); if ((x >> i) & 1) acc = acc.mul(cs, a); }
        return acc.unitary_inverse();
    }
    static Fp12Var final_exponentiation(ConstraintSystem& cs, const Fp12Var& f) {                                    // the chain of csrc/pairing.cuh
        Fp12Var r = f.unitary_inverse().mul(cs, f.inverse(cs));
        r = r.frobenius_map(2).mul(cs, r);
        Fp12Var y0 = r.cyclotomic_square(cs);
        Fp12Var y1 = exp_by_x(cs, r);
        Fp12Var y2 = r.unitary_inverse();
        y1 = y1.mul(cs, y2);
        y2 = exp_by_x(cs, y1);
        y1 = y1.unitary_inverse();
        y1 = y1.mul(cs, y2);
        y2 = exp_by_x(cs, y1);
        y1 = y1.frobenius_map(1);
        y1 = y1.mul(cs, y2);
        r = r.mul(cs, y0);
        y0 = exp_by_x(cs, y1);
        y2 = exp_by_x(cs, y0);
        y0 = y1.frobenius_map(2);
        y1 = y1.unitary_inverse();
        y1 = y1.mul(cs, y2);
        y1 = y1.mul(cs, y0);
        return r.mul(cs, y1);
    }
};

// G1Var: homogeneous projective point on y^2 = x^3 + 4 over Fp, complete addition (Renes-Costello-Batina, a = 0), as
// ark-r1cs-std's short_weierstrass::ProjectiveVar.  (0 : 1 : 0) is the identity.
struct G1Var {
    FpVar x, y, z;
    static G1Var make(const FpVar& x, const FpVar& y, const FpVar& z) { G1Var r; r.x = x; r.y = y; r.z = z; return r; }
    static G1Var zero() { return make(FpVar::zero(), FpVar::one(), FpVar::zero()); }
    static G1Var witness(ConstraintSystem& cs, const g1_aff& p, uint16_t slot0 = 0) {  // input slots slot0 / slot0 + 1 (verify circuit: 0 / 1); z = 1
        FpVar x = FpVar::witness_input(cs, p.x, slot0), y = FpVar::witness_input(cs, p.y, (uint16_t)(slot0 + 1)); return make(x, y, FpVar::witness_const(cs, fp_one()));
    }
    G1Var add(ConstraintSystem& cs, const G1Var& q) const {
        fp b3 = fp_from_u64(12);
        FpVar t0 = x.mul(cs, q.x), t1 = y.mul(cs, q.y), t2 = z.mul(cs, q.z);
        FpVar t3 = (x + y).mul(cs, q.x + q.y) - t0 - t1;
        FpVar t4 = (y + z).mul(cs, q.y + q.z) - t1 - t2;
        FpVar y3 = (x + z).mul(cs, q.x + q.z) - t0 - t2;
        FpVar x3 = t0.dbl() + t0;
        FpVar bt2 = t2.scaled(b3), z3 = t1 + bt2, t1m = t1 - bt2, by3 = y3.scaled(b3);
        return make(t3.mul(cs, t1m) - t4.mul(cs, by3), t1m.mul(cs, z3) + by3.mul(cs, x3), z3.mul(cs, t4) + x3.mul(cs, t3));
    }
};
inline G1Var select_g1(ConstraintSystem& cs, const Boolean& c, const G1Var& t, const G1Var& f) { return G1Var::make(c.select(cs, t.x, f.x), c.select(cs, t.y, f.y), c.select(cs, t.z, f.z)); }

// BlsSignatureVerifyGadget::verify (constraints.rs:90-128) on an already allocated public key; message bytes and the
// signature are allocated as witnesses (the modes of the reference's tests, constraints.rs:335-366), parameters constant.
inline bool verify_gadget(ConstraintSystem& cs, const G1Var& pk, const std::vector<UInt8>& m, const g2_aff& sig, fp12* gt, uint16_t sig_slot0 = 2) {
    G2Var sg;                                                                          // input slots sig_slot0 .. +3 (verify circuit: 2..5); z = (1, 0)
    sg.x.c0 = FpVar::witness_input(cs, sig.x.c0, sig_slot0); sg.x.c1 = FpVar::witness_input(cs, sig.x.c1, (uint16_t)(sig_slot0 + 1));
    sg.y.c0 = FpVar::witness_input(cs, sig.y.c0, (uint16_t)(sig_slot0 + 2)); sg.y.c1 = FpVar::witness_input(cs, sig.y.c1, (uint16_t)(sig_slot0 + 3));
    sg.z.c0 = FpVar::witness_const(cs, fp_one()); sg.z.c1 = FpVar::witness_const(cs, fp_zero());
    // public_key.enforce_not_equal(zero) and prepare_g1: z has an inverse, affine coordinates by two products
    FpVar zi = pk.z.inverse(cs);
    G1AffineVar P1; P1.x = pk.x.mul(cs, zi); P1.y = pk.y.mul(cs, zi);
    G1AffineVar G; G.x = FpVar::constant(fp_const(C_G1X)); G.y = FpVar::constant(fp_const(C_G1Y_NEG));               // -g1, a constant (bls.rs:449-450)
    G2Var h = hash_to_g2_with_cons(cs, m);
    G2PreparedVar hp = PairingVar::prepare_g2(cs, h), sp = PairingVar::prepare_g2(cs, sg);
    Fp12Var f = PairingVar::miller_loop(cs, {G, P1}, {sp, hp});
    Fp12Var e = PairingVar::final_exponentiation(cs, f);
    if (gt) *gt = e.value();
    return e.is_one(cs).val;
}
// PublicKeyVar / SignatureVar are projective witnesses (x, y, z = 1); like the reference, no on-curve / subgroup rows
// (constraints.rs:101-106).  Returns the value of the output Boolean; *gt (nullable) receives the GT element.
inline bool synthesize_verify(ConstraintSystem& cs, const g1_aff& pk, const uint8_t* msg, size_t len, const g2_aff& sig, fp12* gt = nullptr) {
    G1Var pkv = G1Var::witness(cs, pk);
    std::vector<UInt8> m(len); for (size_t i = 0; i < len; i++) m[i] = u8_witness_input(cs, msg[i], (uint16_t)(6 + 8 * i));     // input slots 6 .. 6 + 8 len - 1: any message length (the gadget takes &[UInt8], constraints.rs:90-95)
    return verify_gadget(cs, pkv, m, sig, gt);
}
// aggregate_verify / mapped_aggregate (constraints.rs:153-191): keys masked by a witness bitmap (select key or zero), summed
// with the complete addition, the participant count accumulated with UInt32::addmany; then verify on the aggregate.
// Allocation order follows the reference's test (constraints.rs:393-440): keys, bitmap, message, signature.
inline bool synthesize_aggregate_verify(ConstraintSystem& cs, const std::vector<g1_aff>& pks, const std::vector<uint8_t>& bitmap, const uint8_t* msg, size_t len,
                                        const g2_aff& sig, uint32_t* count_out, fp12* gt = nullptr) {
    if (pks.size() != bitmap.size() || pks.empty()) throw std::invalid_argument("public_keys.len() != bitmap.len()");
    // witness-program input slots of this circuit (n keys, L message bytes): [0, 2n) key coordinates, [2n, 3n) bitmap bits, [3n, 3n + 4) the
    // signature, [3n + 4, 3n + 4 + 8 L) message bits -- csrc/witness.cuh k_witness_inputs_agg fills them in this order
    size_t n = pks.size(); if (3 * n + 4 + 8 * len > 65000) throw std::invalid_argument("too many input slots");
    std::vector<G1Var> keys; for (size_t i = 0; i < n; i++) keys.push_back(G1Var::witness(cs, pks[i], (uint16_t)(2 * i)));
    std::vector<Boolean> bits; for (size_t i = 0; i < n; i++) { cs.set_rule(RULE_INPUT, (uint16_t)(2 * n + i), nullptr); bits.push_back(Boolean::witness(cs, bitmap[i] != 0)); }
    std::vector<UInt8> m(len); for (size_t i = 0; i < len; i++) m[i] = u8_witness_input(cs, msg[i], (uint16_t)(3 * n + 4 + 8 * i));
    G1Var zero = G1Var::zero(), ret = zero;
    UInt32 count; for (int i = 0; i < 32; i++) { LC c0 = LC::constant(fp_zero()); cs.set_rule(RULE_MULADD, 0, nullptr, nullptr, &c0); count.b[i] = Boolean::witness(cs, false); }     // UInt32::new_variable(|| Ok(0), Witness)
    for (size_t i = 0; i < keys.size(); i++) {
        ret = ret.add(cs, select_g1(cs, bits[i], keys[i], zero));
        UInt32 inc = UInt32::constant(0); inc.b[0] = bits[i];                                                         // bit.select(&count_one, &count_zero)
        count = u32_addmany(cs, {count, inc});
    }
    if (count_out) *count_out = count.value();
    return verify_gadget(cs, ret, m, sig, gt, (uint16_t)(3 * n));
}

}  // namespace gadget
