// In-circuit hash-to-G2: the host C++ mirror of the reference's src/hasher.rs (same type and method names), on the builder
// of r1cs_core.hpp.  Sha256Gadget restates ark-crypto-primitives' sha256 constraints (UInt32 bit gadgets), Fp2Var /
// G2Var restate ark-r1cs-std's quadratic-extension and short-Weierstrass projective variables.
//   DefaultFieldHasherWithCons::{hash_to_field, expand}   hasher.rs:37-174
//   DensePolynomialVar::evaluate                          hasher.rs:176-207
//   CurveMapperWithCons::{map_to_curve, isogeny_map, map_to_curve_9mod16, cmov, is_zero, sgn0, pow}   hasher.rs:209-549
//   to_projective_short / to_affine_unchecked             hasher.rs:551-583
//   MapToCurveHasherWithCons::{hash, clear_cofactor2}, hash_to_g2_with_cons   hasher.rs:585-740
#pragma once
#include "r1cs_core.hpp"

namespace gadget {

// ------------------------------------------------------------------------------------------------ SHA-256
struct Sha256Gadget {
    static UInt32 ch(ConstraintSystem& cs, const UInt32& x, const UInt32& y, const UInt32& z) { return x.and_(cs, y).xor_(cs, x.not_().and_(cs, z)); }
    static UInt32 maj(ConstraintSystem& cs, const UInt32& x, const UInt32& y, const UInt32& z) { return x.and_(cs, y).xor_(cs, x.and_(cs, z)).xor_(cs, y.and_(cs, z)); }
    static UInt32 big_sigma(ConstraintSystem& cs, const UInt32& x, int a, int b, int c) { return x.rotr(a).xor_(cs, x.rotr(b)).xor_(cs, x.rotr(c)); }
    static UInt32 small_sigma(ConstraintSystem& cs, const UInt32& x, int a, int b, int s) { return x.rotr(a).xor_(cs, x.rotr(b)).xor_(cs, x.shr(s)); }
    static void compress(ConstraintSystem& cs, std::array<UInt32, 8>& h, const UInt8* block /*64 bytes*/) {
        std::vector<UInt32> w(64);
        for (int i = 0; i < 16; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 8; k++) w[i].b[8 * (3 - j) + k] = block[4 * i + j].b[k];      // big-endian words
        for (int i = 16; i < 64; i++) w[i] = u32_addmany(cs, {w[i - 16], small_sigma(cs, w[i - 15], 7, 18, 3), w[i - 7], small_sigma(cs, w[i - 2], 17, 19, 10)});
        UInt32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            UInt32 t1 = u32_addmany(cs, {hh, big_sigma(cs, e, 6, 11, 25), ch(cs, e, f, g), UInt32::constant(SHA_K[i]), w[i]});
            UInt32 t2 = u32_addmany(cs, {big_sigma(cs, a, 2, 13, 22), maj(cs, a, b, c)});
            hh = g; g = f; f = e; e = u32_addmany(cs, {d, t1}); d = c; c = b; b = a; a = u32_addmany(cs, {t1, t2});
        }
        const UInt32 v[8] = {a, b, c, d, e, f, g, hh};
        for (int i = 0; i < 8; i++) h[i] = u32_addmany(cs, {h[i], v[i]});
    }
    // digest of a byte string (padding bytes are constants)
    static std::vector<UInt8> digest(ConstraintSystem& cs, const std::vector<UInt8>& data) {
        std::vector<UInt8> m = data;
        uint64_t bits = (uint64_t)data.size() * 8;
        m.push_back(u8_constant(0x80));
        while (m.size() % 64 != 56) m.push_back(u8_constant(0));
        for (int i = 7; i >= 0; i--) m.push_back(u8_constant((uint8_t)(bits >> (8 * i))));
        const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
        std::array<UInt32, 8> h; for (int i = 0; i < 8; i++) h[i] = UInt32::constant(iv[i]);
        for (size_t off = 0; off < m.size(); off += 64) compress(cs, h, &m[off]);
        std::vector<UInt8> out(32);
        for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 8; k++) out[4 * i + j].b[k] = h[i].b[8 * (3 - j) + k];
        return out;
    }
};

// ------------------------------------------------------------------------------------------------ Fp2Var
struct Fp2Var {
    FpVar c0, c1;
    static Fp2Var constant(const fp2& v) { Fp2Var r; r.c0 = FpVar::constant(v.c0); r.c1 = FpVar::constant(v.c1); return r; }
    static Fp2Var zero() { return constant(fp2_zero()); }
    static Fp2Var one() { return constant(fp2_one()); }
    static Fp2Var witness(ConstraintSystem& cs, const fp2& v) { Fp2Var r; r.c0 = FpVar::witness(cs, v.c0); r.c1 = FpVar::witness(cs, v.c1); return r; }
    fp2 value() const { fp2 v; v.c0 = c0.val; v.c1 = c1.val; return v; }
    bool is_constant() const { return c0.cst && c1.cst; }
    Fp2Var operator+(const Fp2Var& o) const { Fp2Var r; r.c0 = c0 + o.c0; r.c1 = c1 + o.c1; return r; }
    Fp2Var operator-(const Fp2Var& o) const { Fp2Var r; r.c0 = c0 - o.c0; r.c1 = c1 - o.c1; return r; }
    Fp2Var neg() const { Fp2Var r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
    Fp2Var dbl() const { return *this + *this; }
    Fp2Var conj() const { Fp2Var r; r.c0 = c0; r.c1 = c1.neg(); return r; }
    Fp2Var mul_by_fp(ConstraintSystem& cs, const FpVar& s) const { Fp2Var r; r.c0 = c0.mul(cs, s); r.c1 = c1.mul(cs, s); return r; }
    Fp2Var mul_cst(const fp2& k) const { Fp2Var r; r.c0 = c0.scaled(k.c0) - c1.scaled(k.c1); r.c1 = c0.scaled(k.c1) + c1.scaled(k.c0); return r; }
    Fp2Var mul_xi() const { Fp2Var r; r.c0 = c0 - c1; r.c1 = c0 + c1; return r; }                      // * (1 + u)
    // Karatsuba, 3 constraints (ark-r1cs-std QuadExtVar::mul); a constant operand makes it linear
    Fp2Var mul(ConstraintSystem& cs, const Fp2Var& o) const {
        if (is_constant() || o.is_constant()) {
            const Fp2Var& k = is_constant() ? *this : o; const Fp2Var& x = is_constant() ? o : *this;
            Fp2Var r; r.c0 = x.c0.scaled(k.c0.val) - x.c1.scaled(k.c1.val); r.c1 = x.c0.scaled(k.c1.val) + x.c1.scaled(k.c0.val); return r;
        }
        FpVar v0 = c0.mul(cs, o.c0), v1 = c1.mul(cs, o.c1);
        FpVar s = (c0 + c1).mul(cs, o.c0 + o.c1);
        Fp2Var r; r.c0 = v0 - v1; r.c1 = s - v0 - v1; return r;
    }
    // 2 constraints: (a0 + a1)(a0 - a1), a0 a1
    Fp2Var square(ConstraintSystem& cs) const {
        if (is_constant()) return constant(fp2_sqr(value()));
        FpVar v = c0.mul(cs, c1), d = (c0 + c1).mul(cs, c0 - c1);
        Fp2Var r; r.c0 = d; r.c1 = v.dbl(); return r;
    }
    // witness-hinted inverse, a * inv = 1 (3 constraints); the inverse of zero is unsatisfiable, callers substitute
    Fp2Var inverse(ConstraintSystem& cs) const {
        if (is_constant()) return constant(fp2_inv(value()));
        fp2 iv = fp2_inv(value()); Fp2Var inv;
        cs.set_rule(RULE_FP2INV, 0, &c0.lc, &c1.lc); inv.c0 = FpVar::witness(cs, iv.c0);
        cs.set_rule(RULE_FP2INV, 1, &c0.lc, &c1.lc); inv.c1 = FpVar::witness(cs, iv.c1);
        Fp2Var prod = mul(cs, inv);
        prod.c0.enforce_equal(cs, FpVar::one()); prod.c1.enforce_equal(cs, FpVar::zero());
        return inv;
    }
    Boolean is_eq(ConstraintSystem& cs, const Fp2Var& o) const { return fp_is_eq(cs, c0, o.c0).and_(cs, fp_is_eq(cs, c1, o.c1)); }
    void enforce_equal(ConstraintSystem& cs, const Fp2Var& o) const { c0.enforce_equal(cs, o.c0); c1.enforce_equal(cs, o.c1); }
};
inline Fp2Var select2(ConstraintSystem& cs, const Boolean& c, const Fp2Var& t, const Fp2Var& f) { Fp2Var r; r.c0 = c.select(cs, t.c0, f.c0); r.c1 = c.select(cs, t.c1, f.c1); return r; }

// ------------------------------------------------------------------------------------------------ G2Var
// Homogeneous projective point on y^2 = x^3 + 4(1+u) with the complete formulas of Renes-Costello-Batina (a = 0), the
// arithmetic ark-r1cs-std's short_weierstrass::ProjectiveVar uses.  (0 : 1 : 0) is the identity.
struct G2Var {
    Fp2Var x, y, z;
    static G2Var make(const Fp2Var& x, const Fp2Var& y, const Fp2Var& z) { G2Var r; r.x = x; r.y = y; r.z = z; return r; }
    static G2Var identity() { return make(Fp2Var::zero(), Fp2Var::one(), Fp2Var::zero()); }
    static Fp2Var b3() { fp2 b = g2_b(); fp2 t = fp2_add(fp2_dbl(b), b); return Fp2Var::constant(t); }
    // RCB15 algorithm 7 (a = 0): 12 multiplications
    G2Var add(ConstraintSystem& cs, const G2Var& q) const {
        Fp2Var B3 = b3();
        Fp2Var t0 = x.mul(cs, q.x), t1 = y.mul(cs, q.y), t2 = z.mul(cs, q.z);
        Fp2Var t3 = (x + y).mul(cs, q.x + q.y) - t0 - t1;
        Fp2Var t4 = (y + z).mul(cs, q.y + q.z) - t1 - t2;
        Fp2Var y3 = (x + z).mul(cs, q.x + q.z) - t0 - t2;
        Fp2Var x3 = t0.dbl() + t0;                       // 3 t0
        Fp2Var bt2 = B3.mul(cs, t2);
        Fp2Var z3 = t1 + bt2, t1m = t1 - bt2;
        Fp2Var by3 = B3.mul(cs, y3);
        Fp2Var X3 = t3.mul(cs, t1m) - t4.mul(cs, by3);
        Fp2Var Y3 = t1m.mul(cs, z3) + by3.mul(cs, x3);
        Fp2Var Z3 = z3.mul(cs, t4) + x3.mul(cs, t3);
        return make(X3, Y3, Z3);
    }
    // RCB15 algorithm 9 (a = 0): 6 multiplications + 2 squarings
    G2Var dbl(ConstraintSystem& cs) const {
        Fp2Var B3 = b3();
        Fp2Var t0 = y.square(cs);
        Fp2Var z3 = t0.dbl().dbl().dbl();                // 8 y^2
        Fp2Var t1 = y.mul(cs, z), t2 = z.square(cs);
        t2 = B3.mul(cs, t2);
        Fp2Var x3 = t2.mul(cs, z3);
        Fp2Var y3 = t0 + t2;
        z3 = t1.mul(cs, z3);
        t1 = t2.dbl(); t2 = t1 + t2;                     // 3 b3 z^2
        t0 = t0 - t2;
        y3 = t0.mul(cs, y3); y3 = x3 + y3;
        t1 = x.mul(cs, y);
        x3 = t0.mul(cs, t1); x3 = x3.dbl();
        return make(x3, y3, z3);
    }
    G2Var negate() const { return make(x, y.neg(), z); }
    // double-and-add over constant little-endian scalar bits (the reference only multiplies by the constant h_eff)
    G2Var scalar_mul_le_const(ConstraintSystem& cs, const std::vector<bool>& bits) const {
        G2Var acc = identity(), base = *this; bool first = true;
        size_t top = bits.size(); while (top && !bits[top - 1]) top--;
        for (size_t i = 0; i < top; i++) {
            if (bits[i]) { if (first) { acc = base; first = false; } else acc = acc.add(cs, base); }
            if (i + 1 < top) base = base.dbl(cs);
        }
        return acc;
    }
    // affine value (x / z, y / z); false for the identity
    bool value_affine(g2_aff& a) const {
        fp2 zz = z.value(); if (fp2_is_zero(zz)) { a.x = fp2_zero(); a.y = fp2_zero(); return false; }
        fp2 zi = fp2_inv(zz); a.x = fp2_mul(x.value(), zi); a.y = fp2_mul(y.value(), zi); return true;
    }
};
inline G2Var select_g2(ConstraintSystem& cs, const Boolean& c, const G2Var& t, const G2Var& f) { return G2Var::make(select2(cs, c, t.x, f.x), select2(cs, c, t.y, f.y), select2(cs, c, t.z, f.z)); }

// ------------------------------------------------------------------------------------------------ hasher.rs mirror
static const char* const DST_POP = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";      // hasher.rs:734

struct DefaultFieldHasherWithCons {                                  // hasher.rs:37-174
    ConstraintSystem& cs; std::vector<UInt8> dst; size_t len_per_base_elem = 64;
    DefaultFieldHasherWithCons(ConstraintSystem& c, const std::vector<UInt8>& d) : cs(c), dst(d) { if (d.size() > 255) throw std::invalid_argument("DST too long"); }
    // expand_message_xmd, hasher.rs:110-173
    std::vector<UInt8> expand(const std::vector<UInt8>& message, size_t len_in_bytes) {
        size_t ell = (len_in_bytes + 31) / 32;
        if (ell > 255 || len_in_bytes > 65535) throw std::invalid_argument("expand: output too long");
        std::vector<UInt8> dst_prime = dst; dst_prime.push_back(u8_constant((uint8_t)dst.size()));
        std::vector<UInt8> msg_prime(64, u8_constant(0));                                             // z_pad
        msg_prime.insert(msg_prime.end(), message.begin(), message.end());
        msg_prime.push_back(u8_witness_const(cs, (uint8_t)(len_in_bytes >> 8)));                       // lib_str is a WITNESS in the reference (hasher.rs:131-132)
        msg_prime.push_back(u8_witness_const(cs, (uint8_t)len_in_bytes));
        msg_prime.push_back(u8_constant(0));
        msg_prime.insert(msg_prime.end(), dst_prime.begin(), dst_prime.end());
        std::vector<UInt8> b0 = Sha256Gadget::digest(cs, msg_prime);
        std::vector<UInt8> data = b0; data.push_back(u8_constant(1)); data.insert(data.end(), dst_prime.begin(), dst_prime.end());
        std::vector<UInt8> b1 = Sha256Gadget::digest(cs, data);
        std::vector<UInt8> ret = b1, last_b = b1;
        for (size_t i = 2; i <= ell; i++) {
            std::vector<UInt8> bx(32);
            for (int k = 0; k < 32; k++) bx[k] = u8_xor(cs, b0[k], last_b[k]);
            bx.push_back(u8_constant((uint8_t)i)); bx.insert(bx.end(), dst_prime.begin(), dst_prime.end());
            std::vector<UInt8> bi = Sha256Gadget::digest(cs, bx);
            ret.insert(ret.end(), bi.begin(), bi.end()); last_b = bi;
        }
        ret.resize(len_in_bytes);
        return ret;
    }
    // UInt8 slice (little-endian bytes, at most 47) -> one field variable: ToConstraintFieldGadget packs the bits into a fresh
    // variable tied to the bit combination by one row (Boolean::le_bits_to_fp_var), so later rows read ONE column per element
    FpVar to_constraint_field(const UInt8* bytes_le, size_t n) {
        LC sum; fp val = fp_zero(), pw = fp_one(); bool all_const = true;
        for (size_t i = 0; i < n; i++) for (int k = 0; k < 8; k++) {
            const Boolean& b = bytes_le[i].b[k]; all_const &= b.cst;
            if (!b.lc.t.empty()) sum += b.lc.scaled(pw);
            if (b.val) val = fp_add(val, pw);
            pw = fp_add(pw, pw);
        }
        if (all_const) return FpVar::constant(val);
        cs.set_rule(RULE_MULADD, 0, nullptr, nullptr, &sum);
        FpVar v = FpVar::witness(cs, val);
        cs.enforce(sum, LC::constant(fp_one()), v.lc);
        return v;
    }
    // 64 big-endian bytes -> field element modulo p (hasher.rs:71-104): reverse, split into the 47 high bytes ("head") and the
    // 17 low bytes ("tail"), pack each, f = head * 256^17 + tail
    FpVar bytes_be_to_fp(const UInt8* bytes, size_t n) {
        std::vector<UInt8> le(bytes, bytes + n); std::reverse(le.begin(), le.end());
        size_t pos = (381 - 1) / 8, ntail = n - pos;
        FpVar f_tail = to_constraint_field(le.data(), ntail), f_head = to_constraint_field(le.data() + ntail, pos);
        fp sh = fp_one(); for (size_t i = 0; i < 8 * ntail; i++) sh = fp_add(sh, sh);
        return f_head.scaled(sh) + f_tail;
    }
    std::vector<Fp2Var> hash_to_field(const std::vector<UInt8>& message, size_t count) {                // hasher.rs:58-107
        if (count != 2) throw std::invalid_argument("count must be 2");
        std::vector<UInt8> uni = expand(message, count * 2 * len_per_base_elem);
        std::vector<Fp2Var> out;
        for (size_t i = 0; i < count; i++) {
            Fp2Var u; u.c0 = bytes_be_to_fp(&uni[len_per_base_elem * (2 * i)], len_per_base_elem); u.c1 = bytes_be_to_fp(&uni[len_per_base_elem * (2 * i + 1)], len_per_base_elem);
            out.push_back(u);
        }
        return out;
    }
};

struct DensePolynomialVar {                                          // hasher.rs:176-207
    std::vector<Fp2Var> coeffs;
    Fp2Var evaluate(ConstraintSystem& cs, const Fp2Var& point) const {
        Fp2Var result = Fp2Var::zero(), curr = Fp2Var::one();
        for (size_t i = 0; i < coeffs.size(); i++) { result = result + curr.mul(cs, coeffs[i]); curr = curr.mul(cs, point); }
        return result;
    }
};

inline G2Var to_projective_short(ConstraintSystem& cs, const Fp2Var& xd, const Fp2Var& xn, const Fp2Var& y) {   // hasher.rs:551-559 (Jacobian X, Y, Z)
    Fp2Var xd3 = xd.square(cs).mul(cs, xd);
    return G2Var::make(xn.mul(cs, xd), y.mul(cs, xd3), xd);
}
inline void to_affine_unchecked(ConstraintSystem& cs, const G2Var& p, Fp2Var& x, Fp2Var& y) {                   // hasher.rs:569-583
    Fp2Var z_inv = fp2_is_zero(p.z.value()) ? Fp2Var::zero() : p.z.inverse(cs);                                 // inverse().unwrap_or_else(zero)
    Fp2Var z2 = z_inv.square(cs), z3 = z2.mul(cs, z_inv);
    x = p.x.mul(cs, z2); y = p.y.mul(cs, z3);
}

struct CurveMapperWithCons {                                         // hasher.rs:209-549
    ConstraintSystem& cs;
    Fp2Var COEFF_A, COEFF_B, ZETA, C2, C3, C4, C5; std::string C1;
    explicit CurveMapperWithCons(ConstraintSystem& c) : cs(c) {
        fp2 a; a.c0 = fp_zero(); a.c1 = fp_from_u64(240); COEFF_A = Fp2Var::constant(a);
        fp2 b; b.c0 = fp_from_u64(1012); b.c1 = b.c0; COEFF_B = Fp2Var::constant(b);
        fp2 zt; zt.c0 = fp_neg(fp_from_u64(2)); zt.c1 = fp_neg(fp_from_u64(1)); ZETA = Fp2Var::constant(zt);
        C1 = "2a437a4b8c35fc74bd278eaa22f25e9e2dc90e50e7046b466e59e49349e8bd050a62cfd16ddca6ef53149330978ef011d68619c86185c7b292e85a87091a04966bf91ed3e71b743162c338362113cfd7ced6b1d76382eab26aa00001c718e3";
        fp2 c2; c2.c0 = fp_zero(); c2.c1 = fp_one(); C2 = Fp2Var::constant(c2);
        fp2 c3; c3.c0 = fp_from_dec("2973677408986561043442465346520108879172042883009249989176415018091420807192182638567116318576472649347015917690530");
        c3.c1 = fp_from_dec("1028732146235106349975324479215795277384839936929757896155643118032610843298655225875571310552543014690878354869257"); C3 = Fp2Var::constant(c3);
        fp2 c4; c4.c0 = fp_from_dec("1015919005498129635886032702454337503112659152043614931979881174103627376789972962005013361970813319613593700736144");
        c4.c1 = fp_from_dec("1244231661155348484223428017511856347821538750986231559855759541903146219579071812422210818684355842447591283616181"); C4 = Fp2Var::constant(c4);
        fp2 c5; c5.c0 = fp_from_dec("1637752706019426886789797193293828301565549384974986623510918743054325021588194075665960171838131772227885159387073");
        c5.c1 = fp_from_dec("2356393562099837637521906572659114847248791943663835535137223682689832134851362912628461394915339516530489788841108"); C5 = Fp2Var::constant(c5);
    }
    Fp2Var cmov(const Fp2Var& f, const Fp2Var& t, const Boolean& cond) { return select2(cs, cond, t, f); }            // hasher.rs:506-513
    Boolean is_zero(const Fp2Var& v) { return v.is_eq(cs, Fp2Var::zero()); }                                          // hasher.rs:515-517
    Boolean sgn0(const Fp2Var& v) {                                                                                   // hasher.rs:520-530
        std::vector<Boolean> b0 = fp_to_bits_le(cs, v.c0), b1 = fp_to_bits_le(cs, v.c1);
        Boolean zero_0 = fp_is_eq(cs, v.c0, FpVar::zero());
        return b0[0].or_(cs, zero_0.and_(cs, b1[0]));
    }
    Fp2Var pow(const Fp2Var& v, const std::string& exp_hex) {                                                         // hasher.rs:532-548
        std::vector<uint8_t> e = bytes_from_hex(exp_hex);
        Fp2Var one = Fp2Var::one(), r = one;
        for (uint8_t byte : e) for (int k = 7; k >= 0; k--) {
            r = r.square(cs);
            Fp2Var tv = select2(cs, Boolean::constant((byte >> k) & 1), v, one);
            r = r.mul(cs, tv);
        }
        return r;
    }
    G2Var map_to_curve_9mod16(const Fp2Var& u) {                                                                      // hasher.rs:352-502, RFC 9380 G.2.3 steps 1-70
        const Fp2Var &Z = ZETA, &A = COEFF_A, &B = COEFF_B;
        Fp2Var tv1 = u.square(cs);
        Fp2Var tv3 = Z.mul(cs, tv1);
        Fp2Var tv5 = tv3.square(cs);
        Fp2Var xd = tv5 + tv3;
        Fp2Var x1n = (xd + Fp2Var::one()).mul(cs, B);
        xd = A.neg().mul(cs, xd);
        Boolean e1 = is_zero(xd);
        xd = cmov(xd, Z.mul(cs, A), e1);
        Fp2Var tv2 = xd.square(cs);
        Fp2Var gxd = tv2.mul(cs, xd);
        tv2 = A.mul(cs, tv2);
        Fp2Var gx1 = (x1n.square(cs) + tv2).mul(cs, x1n);
        tv2 = B.mul(cs, gxd);
        gx1 = gx1 + tv2;
        Fp2Var tv4 = gxd.square(cs);
        tv2 = tv4.mul(cs, gxd);
        tv4 = tv4.square(cs);
        tv2 = tv2.mul(cs, tv4);
        tv2 = tv2.mul(cs, gx1);
        tv4 = tv4.square(cs);
        tv4 = tv2.mul(cs, tv4);
        Fp2Var y = pow(tv4, C1);
        y = y.mul(cs, tv2);
        tv4 = y.mul(cs, C2);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e2 = tv2.is_eq(cs, gx1);
        y = cmov(y, tv4, e2);
        tv4 = y.mul(cs, C3);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e3 = tv2.is_eq(cs, gx1);
        y = cmov(y, tv4, e3);
        tv4 = tv4.mul(cs, C2);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e4 = tv2.is_eq(cs, gx1);
        y = cmov(y, tv4, e4);
        Fp2Var gx2 = gx1.mul(cs, tv5).mul(cs, tv3);
        tv5 = y.mul(cs, tv1);
        tv5 = tv5.mul(cs, u);
        tv1 = tv5.mul(cs, C4);
        tv4 = tv1.mul(cs, C2);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e5 = tv2.is_eq(cs, gx2);
        tv1 = cmov(tv1, tv4, e5);
        tv4 = tv5.mul(cs, C5);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e6 = tv2.is_eq(cs, gx2);
        tv1 = cmov(tv1, tv4, e6);
        tv4 = tv4.mul(cs, C2);
        tv2 = tv4.square(cs).mul(cs, gxd);
        Boolean e7 = tv2.is_eq(cs, gx2);
        tv1 = cmov(tv1, tv4, e7);
        tv2 = y.square(cs).mul(cs, gxd);
        Boolean e8 = tv2.is_eq(cs, gx1);
        y = cmov(tv1, y, e8);
        tv2 = tv3.mul(cs, x1n);
        Fp2Var xn = cmov(tv2, x1n, e8);
        Boolean e9 = sgn0(u).is_eq(cs, sgn0(y));
        y = cmov(y.neg(), y, e9);
        return to_projective_short(cs, xd, xn, y);
    }
    G2Var isogeny_map(const G2Var& point) {                                                                           // hasher.rs:294-348
        Boolean is_infinity = point.z.is_eq(cs, Fp2Var::zero());
        Fp2Var x, y; to_affine_unchecked(cs, point, x, y);
        auto poly = [](const fp2* k, int n) { DensePolynomialVar p; for (int i = 0; i < n; i++) p.coeffs.push_back(Fp2Var::constant(k[i])); return p; };
        const fp2 K1[4] = BLS_C_ISO_K1; const fp2 K2[3] = BLS_C_ISO_K2; const fp2 K3[4] = BLS_C_ISO_K3; const fp2 K4[4] = BLS_C_ISO_K4;      // WBConfig::ISOGENY_MAP (RFC 9380 E.3), leading 1 included
        DensePolynomialVar x_num = poly(K1, 4), x_den = poly(K2, 3), y_num = poly(K3, 4), y_den = poly(K4, 4);
        Fp2Var x_den_inv = x_den.evaluate(cs, x).inverse(cs);
        Fp2Var y_den_inv = y_den.evaluate(cs, x).inverse(cs);
        Fp2Var img_x = x_num.evaluate(cs, x).mul(cs, x_den_inv);
        Fp2Var img_y = y_num.evaluate(cs, x).mul(cs, y).mul(cs, y_den_inv);
        G2Var projective = G2Var::make(img_x, img_y, Fp2Var::one());
        G2Var zero = G2Var::make(Fp2Var::zero(), Fp2Var::zero(), Fp2Var::zero());
        return select_g2(cs, is_infinity, zero, projective);
    }
    G2Var map_to_curve(const Fp2Var& u) { return isogeny_map(map_to_curve_9mod16(u)); }                               // hasher.rs:273-276
};

struct MapToCurveHasherWithCons {                                    // hasher.rs:585-725
    ConstraintSystem& cs; DefaultFieldHasherWithCons field_hasher; CurveMapperWithCons curve_mapper;
    MapToCurveHasherWithCons(ConstraintSystem& c, const std::vector<UInt8>& domain) : cs(c), field_hasher(c, domain), curve_mapper(c) {}
    G2Var clear_cofactor2(const G2Var& point) {                                                                        // hasher.rs:664-673
        std::vector<uint8_t> h = bytes_from_hex("0bc69f08f2ee75b3584c6a0ea91b352888e2a8e9145ad7689986ff031508ffe1329c2f178731db956d82bf015d1212b02ec0ec69d7477c1ae954cbc06689f6a359894c0adebbf6b4e8020005aaa95551");
        std::vector<bool> bits;                                       // little-endian
        for (size_t i = h.size(); i-- > 0;) for (int k = 0; k < 8; k++) bits.push_back((h[i] >> k) & 1);
        return point.scalar_mul_le_const(cs, bits);
    }
    G2Var hash(const std::vector<UInt8>& msg) {                                                                        // hasher.rs:641-661
        std::vector<Fp2Var> u = field_hasher.hash_to_field(msg, 2);
        G2Var q0 = curve_mapper.map_to_curve(u[0]), q1 = curve_mapper.map_to_curve(u[1]);
        return clear_cofactor2(q0.add(cs, q1));
    }
};
inline G2Var hash_to_g2_with_cons(ConstraintSystem& cs, const std::vector<UInt8>& message) {                            // hasher.rs:727-740
    std::vector<UInt8> domain = u8_constant_vec((const uint8_t*)DST_POP, strlen(DST_POP));
    MapToCurveHasherWithCons h(cs, domain);
    return h.hash(message);
}

}  // namespace gadget
