// CPU oracle for the BLS12-381 verify path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain C++17 restatement (6 x 64-bit limbs, unsigned __int128) of what the reference computes
// on its hot path.  The reference (lightec-xyz/bls-verify-gadget) holds no arithmetic of its own:
// every field/curve/pairing operation is a call into un-vendored arkworks crates
// (ark-ff / ark-ec / ark-bls12-381 / ark-serialize ^0.4.0, ark-relations ^0.4.0, sha2 0.10;
// /root/reference/Cargo.toml:17-35 -- no Cargo.lock is shipped).  So this file restates the published
// algorithms and is anchored on the reference's call sites and fixtures:
//
//   ora_verify          src/bls.rs:427-458 (identity test, check() x2, 2-pair multi_pairing, is_one)
//   ora_hash_to_g2      src/bls.rs:477-493; algorithm spec src/hasher.rs:58-173 (xmd, hash_to_field),
//                       352-502 (SSWU), 294-348 (3-isogeny), 664-673 ([h_eff] cofactor clearing)
//   ora_sign/sk_to_pk   src/bls.rs:411-425, 210-216
//   ora_g1/g2_aggregate src/bls.rs:183-195, 288-300
//   ora_deser_*         src/bls.rs:219-223, 316-320 (deserialize_compressed = ZCash format + subgroup check)
//   ora_r1cs_check      ark-relations ConstraintSystem::is_satisfied applied to the circuit of
//                       src/constraints.rs:90-128 (all rows reported, no early exit)
//
// It deliberately uses the SIMPLE algorithms ([r]P subgroup tests, [h_eff] scalar-mul cofactor
// clearing, RFC 9380 6.6.2 SSWU with inversions, Jacobian generic addition) so that it is
// independent of the CUDA implementation, which uses the endomorphism-based fast forms.
//
// Parity status: PINNED for compressed G1/G2 bytes, verify booleans, deserialisation accept/reject,
// hash-to-G2 and xmd bytes by the reference's 78 JSON fixtures + inline KATs (tests/golden/
// eth_vectors.json; tests/test_oracle_golden.py).  GT bytes and R1CS satisfaction vectors are
// parity-UNPINNED: no fixture of the reference observes them.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <thread>
#include <atomic>
#include <string>

typedef uint64_t u64; typedef unsigned __int128 u128; typedef uint8_t u8;

// ---------------------------------------------------------------------------------------- Fp
struct Fp { u64 l[6]; };
static const Fp P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const u64 INV64 = 0x89f3fffcfffcfffdULL;            // -p^-1 mod 2^64
static Fp FP_ONE, FP_R2, FP_ZERO;                          // Montgomery 1, R^2; set by init()

static inline bool geq(const Fp& a, const Fp& b) {         // a >= b as integers
    for (int i = 5; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] > b.l[i]; }
    return true;
}
static inline u64 add_raw(Fp& r, const Fp& a, const Fp& b) {
    u128 c = 0; for (int i = 0; i < 6; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (u64)c; c >>= 64; } return (u64)c;
}
static inline u64 sub_raw(Fp& r, const Fp& a, const Fp& b) {
    u64 br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)a.l[i] - b.l[i] - br; r.l[i] = (u64)d; br = (u64)(d >> 64) & 1; }
    return br;
}
static inline Fp fadd(const Fp& a, const Fp& b) { Fp r; add_raw(r, a, b); if (geq(r, P)) sub_raw(r, r, P); return r; }
static inline Fp fsub(const Fp& a, const Fp& b) { Fp r; if (sub_raw(r, a, b)) add_raw(r, r, P); return r; }
static inline bool fis_zero(const Fp& a) { u64 o = 0; for (int i = 0; i < 6; i++) o |= a.l[i]; return o == 0; }
static inline bool feq(const Fp& a, const Fp& b) { return memcmp(&a, &b, sizeof(Fp)) == 0; }
static inline Fp fneg(const Fp& a) { if (fis_zero(a)) return a; Fp r; sub_raw(r, P, a); return r; }
static Fp fmul(const Fp& a, const Fp& b) {                 // CIOS Montgomery product a*b/R mod p
    u64 t[8] = {0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[6]; t[6] = (u64)c; t[7] = (u64)(c >> 64);
        u64 m = t[0] * INV64;
        c = ((u128)m * P.l[0] + t[0]) >> 64;
        for (int j = 1; j < 6; j++) { c += (u128)m * P.l[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[6]; t[5] = (u64)c; t[6] = t[7] + (u64)(c >> 64);
    }
    Fp r; memcpy(r.l, t, 48);
    if (t[6] || geq(r, P)) sub_raw(r, r, P);
    return r;
}
static inline Fp fsqr(const Fp& a) { return fmul(a, a); }
static inline Fp to_mont(const Fp& a) { return fmul(a, FP_R2); }
static inline Fp from_mont(const Fp& a) { Fp one = {{1, 0, 0, 0, 0, 0}}; return fmul(a, one); }
// exponent: little-endian u64 limbs
static Fp fpow(const Fp& a, const u64* e, int n) {
    Fp r = FP_ONE; bool started = false;
    for (int i = n * 64 - 1; i >= 0; i--) {
        if (started) r = fsqr(r);
        if ((e[i / 64] >> (i % 64)) & 1) { r = started ? fmul(r, a) : a; started = true; }
    }
    return r;
}
// tiny bignum helpers for exponents derived from p (6 limbs)
struct Big { u64 l[6]; };
static Big big_p() { Big b; memcpy(b.l, P.l, 48); return b; }
static void big_add_small(Big& b, long v) {
    if (v >= 0) { u128 c = (u64)v; for (int i = 0; i < 6; i++) { c += b.l[i]; b.l[i] = (u64)c; c >>= 64; } }
    else { u64 br = (u64)(-v); for (int i = 0; i < 6; i++) { u64 o = b.l[i]; b.l[i] = o - br; br = o < br; } }
}
static u64 big_div_small(Big& b, u64 d) {
    u128 rem = 0; for (int i = 5; i >= 0; i--) { u128 cur = (rem << 64) | b.l[i]; b.l[i] = (u64)(cur / d); rem = cur % d; } return (u64)rem;
}
static Big E_PM2, E_P1D4, E_PM3D4, E_PM1D2, E_PM1D6;
static inline Fp finv(const Fp& a) { return fpow(a, E_PM2.l, 6); }
static Fp fp_from_u64(u64 v) { Fp a = {{v, 0, 0, 0, 0, 0}}; return to_mont(a); }
static bool fp_from_be(Fp& out, const u8* b) {             // canonical BE 48 bytes -> Montgomery; false if >= p
    Fp a; for (int i = 0; i < 6; i++) { u64 w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[(5 - i) * 8 + k]; a.l[i] = w; }
    if (geq(a, P)) return false;
    out = to_mont(a); return true;
}
static void fp_to_be(u8* b, const Fp& m) {
    Fp a = from_mont(m);
    for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[(5 - i) * 8 + k] = (u8)(a.l[i] >> (56 - 8 * k));
}
static void fp_to_le(u8* b, const Fp& m) { Fp a = from_mont(m); memcpy(b, a.l, 48); }
static bool fp_from_le(Fp& out, const u8* b) { Fp a; memcpy(a.l, b, 48); if (geq(a, P)) return false; out = to_mont(a); return true; }
static Fp fp_from_hex(const char* h) {                     // big-endian hex, any length <= 96
    size_t n = strlen(h); u8 b[48] = {0};
    for (size_t i = 0; i < n; i++) {
        char c = h[n - 1 - i]; int v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10;
        b[47 - i / 2] |= (u8)(v << (4 * (i & 1)));
    }
    Fp r; fp_from_be(r, b); return r;
}
static bool fp_is_odd(const Fp& m) { return from_mont(m).l[0] & 1; }
static int fp_cmp(const Fp& ma, const Fp& mb) {            // compare canonical integers
    Fp a = from_mont(ma), b = from_mont(mb);
    for (int i = 5; i >= 0; i--) if (a.l[i] != b.l[i]) return a.l[i] > b.l[i] ? 1 : -1;
    return 0;
}

// ---------------------------------------------------------------------------------------- Fp2
struct Fp2 { Fp c0, c1; };
static Fp2 F2_ZERO, F2_ONE;
static inline Fp2 operator+(const Fp2& a, const Fp2& b) { return {fadd(a.c0, b.c0), fadd(a.c1, b.c1)}; }
static inline Fp2 operator-(const Fp2& a, const Fp2& b) { return {fsub(a.c0, b.c0), fsub(a.c1, b.c1)}; }
static inline Fp2 operator-(const Fp2& a) { return {fneg(a.c0), fneg(a.c1)}; }
static inline Fp2 operator*(const Fp2& a, const Fp2& b) {
    Fp t0 = fmul(a.c0, b.c0), t1 = fmul(a.c1, b.c1);
    Fp t2 = fmul(fadd(a.c0, a.c1), fadd(b.c0, b.c1));
    return {fsub(t0, t1), fsub(fsub(t2, t0), t1)};
}
static inline Fp2 sqr(const Fp2& a) { Fp t = fmul(a.c0, a.c1); return {fmul(fadd(a.c0, a.c1), fsub(a.c0, a.c1)), fadd(t, t)}; }
static inline Fp2 conj(const Fp2& a) { return {a.c0, fneg(a.c1)}; }
static inline Fp2 mulfp(const Fp2& a, const Fp& s) { return {fmul(a.c0, s), fmul(a.c1, s)}; }
static inline Fp2 mul_xi(const Fp2& a) { return {fsub(a.c0, a.c1), fadd(a.c0, a.c1)}; }    // * (1+u)
static inline bool is_zero(const Fp2& a) { return fis_zero(a.c0) && fis_zero(a.c1); }
static inline bool operator==(const Fp2& a, const Fp2& b) { return feq(a.c0, b.c0) && feq(a.c1, b.c1); }
static inline bool operator!=(const Fp2& a, const Fp2& b) { return !(a == b); }
static Fp2 inv(const Fp2& a) { Fp n = finv(fadd(fsqr(a.c0), fsqr(a.c1))); return {fmul(a.c0, n), fneg(fmul(a.c1, n))}; }
static Fp2 pow2(const Fp2& a, const u64* e, int n) {
    Fp2 r = F2_ONE;
    for (int i = n * 64 - 1; i >= 0; i--) { r = sqr(r); if ((e[i / 64] >> (i % 64)) & 1) r = r * a; }
    return r;
}
static bool is_square(const Fp2& a) {                      // Euler criterion on the norm
    Fp n = fadd(fsqr(a.c0), fsqr(a.c1)); Fp l = fpow(n, E_PM1D2.l, 6); return fis_zero(n) || feq(l, FP_ONE);
}
static bool sqrt2(Fp2& out, const Fp2& a) {                // Adj & Rodriguez-Henriquez alg. 9 (p = 3 mod 4)
    if (is_zero(a)) { out = a; return true; }
    Fp2 a1 = pow2(a, E_PM3D4.l, 6), alpha = sqr(a1) * a, x0 = a1 * a, x;
    Fp2 m1 = -F2_ONE;
    if (alpha == m1) x = Fp2{fneg(x0.c1), x0.c0};           // u * x0
    else x = pow2(F2_ONE + alpha, E_PM1D2.l, 6) * x0;
    out = x; return sqr(x) == a;
}
static int sgn0(const Fp2& a) {                            // hasher.rs:520-530
    int s0 = fp_is_odd(a.c0), z0 = fis_zero(a.c0), s1 = fp_is_odd(a.c1); return s0 | (z0 & s1);
}
static bool lex_largest(const Fp2& y) {                    // arkworks Fp2 ordering: c1 first, then c0
    Fp2 n = -y; int c = fp_cmp(y.c1, n.c1); if (c) return c > 0; return fp_cmp(y.c0, n.c0) > 0;
}

// ---------------------------------------------------------------------------------------- Fp6 / Fp12
struct Fp6 { Fp2 c0, c1, c2; };
struct Fp12 { Fp6 c0, c1; };
static inline Fp6 operator+(const Fp6& a, const Fp6& b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
static inline Fp6 operator-(const Fp6& a, const Fp6& b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
static inline Fp6 operator-(const Fp6& a) { return {-a.c0, -a.c1, -a.c2}; }
// ORA_FAST selects the algorithm set: 0 (default, used by every parity test) = the simple forms described in the header;
// 1 = the forms arkworks itself uses (Karatsuba Fp6, cyclotomic squaring, endomorphism subgroup tests, psi cofactor clearing),
// used only by the timed CPU-baseline legs so that the reported CPU figure is not handicapped.  tests/test_oracle_golden.py
// checks that both sets produce identical bytes.
static int ORA_FAST = 0;
static Fp6 operator*(const Fp6& a, const Fp6& b) {
    if (ORA_FAST) {
        Fp2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
        return {v0 + mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2), (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + mul_xi(v2), (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1};
    }
    return {a.c0 * b.c0 + mul_xi(a.c1 * b.c2 + a.c2 * b.c1),
            a.c0 * b.c1 + a.c1 * b.c0 + mul_xi(a.c2 * b.c2),
            a.c0 * b.c2 + a.c1 * b.c1 + a.c2 * b.c0};
}
static inline Fp6 mul_v(const Fp6& a) { return {mul_xi(a.c2), a.c0, a.c1}; }
static Fp6 mul_by_01(const Fp6& a, const Fp2& b0, const Fp2& b1) {
    return {a.c0 * b0 + mul_xi(a.c2 * b1), a.c0 * b1 + a.c1 * b0, a.c1 * b1 + a.c2 * b0};
}
static Fp6 mul_by_1(const Fp6& a, const Fp2& b1) { return {mul_xi(a.c2 * b1), a.c0 * b1, a.c1 * b1}; }
static Fp6 inv(const Fp6& a) {
    Fp2 t0 = sqr(a.c0) - mul_xi(a.c1 * a.c2), t1 = mul_xi(sqr(a.c2)) - a.c0 * a.c1, t2 = sqr(a.c1) - a.c0 * a.c2;
    Fp2 d = inv(a.c0 * t0 + mul_xi(a.c2 * t1 + a.c1 * t2));
    return {t0 * d, t1 * d, t2 * d};
}
static Fp12 F12_ONE;
static Fp12 operator*(const Fp12& a, const Fp12& b) {
    Fp6 t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
    return {t0 + mul_v(t1), (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1};
}
static inline Fp12 sqr(const Fp12& a) { return a * a; }
static inline Fp12 conj(const Fp12& a) { return {a.c0, -a.c1}; }
static Fp12 inv(const Fp12& a) { Fp6 d = inv(a.c0 * a.c0 - mul_v(a.c1 * a.c1)); return {a.c0 * d, -(a.c1 * d)}; }
static bool operator==(const Fp12& a, const Fp12& b) {
    return a.c0.c0 == b.c0.c0 && a.c0.c1 == b.c0.c1 && a.c0.c2 == b.c0.c2 && a.c1.c0 == b.c1.c0 && a.c1.c1 == b.c1.c1 && a.c1.c2 == b.c1.c2;
}
static Fp12 mul_by_014(const Fp12& f, const Fp2& c0, const Fp2& c1, const Fp2& c4) {
    Fp6 t0 = mul_by_01(f.c0, c0, c1), t1 = mul_by_1(f.c1, c4);
    return {t0 + mul_v(t1), mul_by_01(f.c0 + f.c1, c0, c1 + c4) - t0 - t1};
}
static Fp2 FROB_G[6];                                      // xi^(i(p-1)/6)
static Fp12 frob(const Fp12& a) {                          // a^p : coefficient of w^i -> conj * gamma_i
    Fp12 r;
    r.c0.c0 = conj(a.c0.c0);               r.c1.c0 = conj(a.c1.c0) * FROB_G[1];
    r.c0.c1 = conj(a.c0.c1) * FROB_G[2];   r.c1.c1 = conj(a.c1.c1) * FROB_G[3];
    r.c0.c2 = conj(a.c0.c2) * FROB_G[4];   r.c1.c2 = conj(a.c1.c2) * FROB_G[5];
    return r;
}
static void fp12_to_le(u8* out, const Fp12& a) {           // c0.c0.c0 ... c1.c2.c1, 48 B LE canonical
    const Fp2* c[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
    for (int i = 0; i < 6; i++) { fp_to_le(out + 96 * i, c[i]->c0); fp_to_le(out + 96 * i + 48, c[i]->c1); }
}
static bool fp12_from_le(Fp12& a, const u8* in) {
    Fp2* c[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
    for (int i = 0; i < 6; i++) if (!fp_from_le(c[i]->c0, in + 96 * i) || !fp_from_le(c[i]->c1, in + 96 * i + 48)) return false;
    return true;
}

// ---------------------------------------------------------------------------------------- curves (Jacobian, a = 0)
static inline Fp  f_add(const Fp& a, const Fp& b) { return fadd(a, b); }
static inline Fp2 f_add(const Fp2& a, const Fp2& b) { return a + b; }
static inline Fp  f_sub(const Fp& a, const Fp& b) { return fsub(a, b); }
static inline Fp2 f_sub(const Fp2& a, const Fp2& b) { return a - b; }
static inline Fp  f_mul(const Fp& a, const Fp& b) { return fmul(a, b); }
static inline Fp2 f_mul(const Fp2& a, const Fp2& b) { return a * b; }
static inline Fp  f_sqr(const Fp& a) { return fsqr(a); }
static inline Fp2 f_sqr(const Fp2& a) { return sqr(a); }
static inline Fp  f_inv(const Fp& a) { return finv(a); }
static inline Fp2 f_inv(const Fp2& a) { return inv(a); }
static inline bool f_zero(const Fp& a) { return fis_zero(a); }
static inline bool f_zero(const Fp2& a) { return is_zero(a); }
static inline bool f_eq(const Fp& a, const Fp& b) { return feq(a, b); }
static inline bool f_eq(const Fp2& a, const Fp2& b) { return a == b; }
static inline void f_one(Fp& a) { a = FP_ONE; }
static inline void f_one(Fp2& a) { a = F2_ONE; }

template <class F> struct Aff { F x, y; bool inf; };
template <class F> struct Jac { F X, Y, Z; };              // Z == 0 <=> identity
template <class F> static Jac<F> jac_identity() { Jac<F> r; f_one(r.X); f_one(r.Y); r.Z = f_sub(r.X, r.X); return r; }
template <class F> static Jac<F> to_jac(const Aff<F>& a) { if (a.inf) return jac_identity<F>(); Jac<F> r; r.X = a.x; r.Y = a.y; f_one(r.Z); return r; }
template <class F> static Jac<F> dbl(const Jac<F>& p) {
    if (f_zero(p.Z)) return p;
    F A = f_sqr(p.X), B = f_sqr(p.Y), C = f_sqr(B);
    F D = f_sub(f_sub(f_sqr(f_add(p.X, B)), A), C); D = f_add(D, D);
    F E = f_add(f_add(A, A), A), Fq = f_sqr(E);
    Jac<F> r; r.X = f_sub(Fq, f_add(D, D));
    F C8 = f_add(C, C); C8 = f_add(C8, C8); C8 = f_add(C8, C8);
    r.Y = f_sub(f_mul(E, f_sub(D, r.X)), C8);
    F YZ = f_mul(p.Y, p.Z); r.Z = f_add(YZ, YZ);
    return r;
}
template <class F> static Jac<F> add(const Jac<F>& p, const Jac<F>& q) {
    if (f_zero(p.Z)) return q;
    if (f_zero(q.Z)) return p;
    F Z1Z1 = f_sqr(p.Z), Z2Z2 = f_sqr(q.Z);
    F U1 = f_mul(p.X, Z2Z2), U2 = f_mul(q.X, Z1Z1);
    F S1 = f_mul(f_mul(p.Y, q.Z), Z2Z2), S2 = f_mul(f_mul(q.Y, p.Z), Z1Z1);
    if (f_eq(U1, U2)) { if (f_eq(S1, S2)) return dbl(p); return jac_identity<F>(); }
    F Hh = f_sub(U2, U1), Rr = f_sub(S2, S1);
    F H2 = f_sqr(Hh), H3 = f_mul(Hh, H2), V = f_mul(U1, H2);
    Jac<F> r; r.X = f_sub(f_sub(f_sqr(Rr), H3), f_add(V, V));
    r.Y = f_sub(f_mul(Rr, f_sub(V, r.X)), f_mul(S1, H3));
    r.Z = f_mul(f_mul(p.Z, q.Z), Hh);
    return r;
}
template <class F> static Jac<F> neg(const Jac<F>& p) { Jac<F> r = p; F z = f_sub(p.Y, p.Y); r.Y = f_sub(z, p.Y); return r; }
template <class F> static Jac<F> smul(const Jac<F>& p, const u8* k_be, int nbytes) {   // big-endian scalar
    Jac<F> r = jac_identity<F>();
    for (int i = 0; i < nbytes; i++) for (int b = 7; b >= 0; b--) { r = dbl(r); if ((k_be[i] >> b) & 1) r = add(r, p); }
    return r;
}
template <class F> static Aff<F> to_aff(const Jac<F>& p) {
    Aff<F> a; a.inf = f_zero(p.Z); if (a.inf) { a.x = p.Z; a.y = p.Z; return a; }
    F zi = f_inv(p.Z), zi2 = f_sqr(zi); a.x = f_mul(p.X, zi2); a.y = f_mul(p.Y, f_mul(zi2, zi)); return a;
}
typedef Aff<Fp> G1A; typedef Jac<Fp> G1J; typedef Aff<Fp2> G2A; typedef Jac<Fp2> G2J;
static Fp B1; static Fp2 B2C;                              // curve constants 4 and 4(1+u)
static u8 R_BE[32], HEFF_BE[80];
static G1A G1_GEN, G1_GEN_NEG;
static bool on_curve(const G1A& a) { return a.inf || feq(fsqr(a.y), fadd(fmul(fsqr(a.x), a.x), B1)); }
static bool on_curve(const G2A& a) { return a.inf || sqr(a.y) == sqr(a.x) * a.x + B2C; }
static Fp BETA; static Fp2 PSI_CX, PSI_CY; static u8 X_BE[8];
template <class F> static bool jac_eq_aff(const Jac<F>& p, const Aff<F>& q) {
    if (f_zero(p.Z)) return q.inf;
    if (q.inf) return false;
    F zz = f_sqr(p.Z); return f_eq(p.X, f_mul(q.x, zz)) && f_eq(p.Y, f_mul(q.y, f_mul(zz, p.Z)));
}
static G2J psi(const G2J& p) { return {conj(p.X) * PSI_CX, conj(p.Y) * PSI_CY, conj(p.Z)}; }
static bool in_subgroup_slow(const G1A& a) { return f_zero(smul(to_jac(a), R_BE, 32).Z); }   // [r]P == O
static bool in_subgroup_slow(const G2A& a) { return f_zero(smul(to_jac(a), R_BE, 32).Z); }
static bool in_subgroup(const G1A& a) {
    if (!ORA_FAST || a.inf) return in_subgroup_slow(a);
    G1J t = smul(smul(to_jac(a), X_BE, 8), X_BE, 8);                     // [x^2]P
    G1A s = {fmul(a.x, BETA), fneg(a.y), false};                         // -sigma(P)
    return jac_eq_aff(t, s);
}
static bool in_subgroup(const G2A& a) {
    if (!ORA_FAST || a.inf) return in_subgroup_slow(a);
    G2J t = smul(to_jac(a), X_BE, 8);                                    // [|x|]P = -[x]P
    G2J ps = psi(to_jac(a)); G2A s = {ps.X, -ps.Y, false};               // -psi(P)   (Z = 1)
    return jac_eq_aff(t, s);
}

// ---------------------------------------------------------------------------------------- ZCash compressed codec
enum { DE_OK = 0, DE_FLAGS = 1, DE_RANGE = 2, DE_CURVE = 3, DE_SUBGROUP = 4 };
static int deser_g1(G1A& out, const u8* b, bool check_subgroup = true) {
    if (!(b[0] & 0x80)) return DE_FLAGS;
    if (b[0] & 0x40) { out.inf = true; out.x = FP_ZERO; out.y = FP_ZERO; return DE_OK; }   // lenient: SURVEY B7
    u8 t[48]; memcpy(t, b, 48); t[0] &= 0x1f;
    Fp x; if (!fp_from_be(x, t)) return DE_RANGE;
    Fp y2 = fadd(fmul(fsqr(x), x), B1), y = fpow(y2, E_P1D4.l, 6);
    if (!feq(fsqr(y), y2)) return DE_CURVE;
    Fp ny = fneg(y); bool largest = fp_cmp(y, ny) > 0;
    if (largest != !!(b[0] & 0x20)) y = ny;
    out.x = x; out.y = y; out.inf = false;
    if (check_subgroup && !in_subgroup(out)) return DE_SUBGROUP;
    return DE_OK;
}
static void ser_g1(u8* b, const G1A& a) {
    if (a.inf) { memset(b, 0, 48); b[0] = 0xc0; return; }
    fp_to_be(b, a.x); b[0] |= 0x80; if (fp_cmp(a.y, fneg(a.y)) > 0) b[0] |= 0x20;
}
static int deser_g2(G2A& out, const u8* b, bool check_subgroup = true) {
    if (!(b[0] & 0x80)) return DE_FLAGS;
    if (b[0] & 0x40) { out.inf = true; out.x = F2_ZERO; out.y = F2_ZERO; return DE_OK; }
    u8 t[48]; memcpy(t, b, 48); t[0] &= 0x1f;
    Fp2 x; if (!fp_from_be(x.c1, t) || !fp_from_be(x.c0, b + 48)) return DE_RANGE;
    Fp2 y; if (!sqrt2(y, sqr(x) * x + B2C)) return DE_CURVE;
    if (lex_largest(y) != !!(b[0] & 0x20)) y = -y;
    out.x = x; out.y = y; out.inf = false;
    if (check_subgroup && !in_subgroup(out)) return DE_SUBGROUP;
    return DE_OK;
}
static void ser_g2(u8* b, const G2A& a) {
    if (a.inf) { memset(b, 0, 96); b[0] = 0xc0; return; }
    fp_to_be(b, a.x.c1); fp_to_be(b + 48, a.x.c0); b[0] |= 0x80; if (lex_largest(a.y)) b[0] |= 0x20;
}

// ZCash UNCOMPRESSED codec (96 / 192 bytes; the other wire format of the upstream bls12-381-tests suite, reference tests/readme.md:4-7):
// bit 7 of the first byte must be clear, bit 6 = infinity (lenient like the compressed form), bit 5 ignored (ark-bls12-381 0.4's
// read_g1_uncompressed); Validate::Yes = on curve + in the subgroup.  Unpinned: the reference vendors no uncompressed vectors.
static int deser_g1_unc(G1A& out, const u8* b) {
    if (b[0] & 0x80) return DE_FLAGS;
    if (b[0] & 0x40) { out.inf = true; out.x = FP_ZERO; out.y = FP_ZERO; return DE_OK; }
    u8 t[48]; memcpy(t, b, 48); t[0] &= 0x1f;
    if (!fp_from_be(out.x, t) || !fp_from_be(out.y, b + 48)) return DE_RANGE;
    out.inf = false;
    if (!on_curve(out)) return DE_CURVE;
    if (!in_subgroup(out)) return DE_SUBGROUP;
    return DE_OK;
}
static void ser_g1_unc(u8* b, const G1A& a) { if (a.inf) { memset(b, 0, 96); b[0] = 0x40; return; } fp_to_be(b, a.x); fp_to_be(b + 48, a.y); }
static int deser_g2_unc(G2A& out, const u8* b) {
    if (b[0] & 0x80) return DE_FLAGS;
    if (b[0] & 0x40) { out.inf = true; out.x = F2_ZERO; out.y = F2_ZERO; return DE_OK; }
    u8 t[48]; memcpy(t, b, 48); t[0] &= 0x1f;
    if (!fp_from_be(out.x.c1, t) || !fp_from_be(out.x.c0, b + 48) || !fp_from_be(out.y.c1, b + 96) || !fp_from_be(out.y.c0, b + 144)) return DE_RANGE;
    out.inf = false;
    if (!on_curve(out)) return DE_CURVE;
    if (!in_subgroup(out)) return DE_SUBGROUP;
    return DE_OK;
}
static void ser_g2_unc(u8* b, const G2A& a) {
    if (a.inf) { memset(b, 0, 192); b[0] = 0x40; return; }
    fp_to_be(b, a.x.c1); fp_to_be(b + 48, a.x.c0); fp_to_be(b + 96, a.y.c1); fp_to_be(b + 144, a.y.c0);
}

// ---------------------------------------------------------------------------------------- SHA-256, xmd, hash_to_field
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha256_block(uint32_t h[8], const u8* blk) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)blk[4 * i] << 24 | blk[4 * i + 1] << 16 | blk[4 * i + 2] << 8 | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256(u8 out[32], const std::vector<u8>& m) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<u8> p(m); u64 bits = (u64)m.size() * 8; p.push_back(0x80);
    while (p.size() % 64 != 56) p.push_back(0);
    for (int i = 7; i >= 0; i--) p.push_back((u8)(bits >> (8 * i)));
    for (size_t o = 0; o < p.size(); o += 64) sha256_block(h, &p[o]);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
static void expand_xmd(u8* out, size_t n, const u8* msg, size_t mlen, const u8* dst, size_t dlen) {   // hasher.rs:110-173
    size_t ell = (n + 31) / 32;
    std::vector<u8> dstp(dst, dst + dlen); dstp.push_back((u8)dlen);
    std::vector<u8> m(64, 0); m.insert(m.end(), msg, msg + mlen);
    m.push_back((u8)(n >> 8)); m.push_back((u8)n); m.push_back(0); m.insert(m.end(), dstp.begin(), dstp.end());
    u8 b0[32], bi[32]; sha256(b0, m);
    std::vector<u8> t(b0, b0 + 32); t.push_back(1); t.insert(t.end(), dstp.begin(), dstp.end()); sha256(bi, t);
    std::vector<u8> uni(bi, bi + 32);
    for (size_t i = 2; i <= ell; i++) {
        std::vector<u8> q(32); for (int k = 0; k < 32; k++) q[k] = b0[k] ^ bi[k];
        q.push_back((u8)i); q.insert(q.end(), dstp.begin(), dstp.end()); sha256(bi, q);
        uni.insert(uni.end(), bi, bi + 32);
    }
    memcpy(out, uni.data(), n);
}
static Fp fp_from_be64_mod(const u8* b) {                  // 64-byte BE integer mod p (hasher.rs:71-104)
    // value = hi(16 B) * 2^384 + lo(48 B): reduce as hi*R + lo with Montgomery arithmetic
    Fp lo, hi; memset(&hi, 0, sizeof hi);
    for (int i = 0; i < 6; i++) { u64 w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[16 + (5 - i) * 8 + k]; lo.l[i] = w; }
    for (int i = 0; i < 2; i++) { u64 w = 0; for (int k = 0; k < 8; k++) w = (w << 8) | b[(1 - i) * 8 + k]; hi.l[i] = w; }
    while (geq(lo, P)) sub_raw(lo, lo, P);
    // to_mont(lo) = lo*R ; hi*2^384 in Montgomery form = hi * R * R = mont(mont(hi, R2), R2)
    Fp lom = to_mont(lo), him = fmul(to_mont(hi), FP_R2);
    return fadd(lom, him);
}
static const char DST_POP[] = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";               // bls.rs:482
static void hash_to_field(Fp2 u[2], const u8* msg, size_t mlen) {
    u8 uni[256]; expand_xmd(uni, 256, msg, mlen, (const u8*)DST_POP, sizeof(DST_POP) - 1);
    u[0] = {fp_from_be64_mod(uni), fp_from_be64_mod(uni + 64)}; u[1] = {fp_from_be64_mod(uni + 128), fp_from_be64_mod(uni + 192)};
}

// ---------------------------------------------------------------------------------------- SSWU + isogeny + cofactor
static Fp2 ISO_A, ISO_B, SSWU_Z, K1[4], K2[3], K3[4], K4[4];
static Fp2 g_iso(const Fp2& x) { return sqr(x) * x + ISO_A * x + ISO_B; }
static void sswu(Fp2& x, Fp2& y, const Fp2& u) {           // RFC 9380 6.6.2 (simple form; same function as hasher.rs:352-502)
    Fp2 zu2 = SSWU_Z * sqr(u), ta = sqr(zu2) + zu2, x1;
    if (is_zero(ta)) x1 = ISO_B * inv(SSWU_Z * ISO_A);
    else x1 = (-ISO_B) * inv(ISO_A) * (F2_ONE + inv(ta));
    Fp2 gx1 = g_iso(x1);
    if (is_square(gx1)) { x = x1; sqrt2(y, gx1); }
    else { x = zu2 * x1; sqrt2(y, g_iso(x)); }
    if (sgn0(u) != sgn0(y)) y = -y;
}
static Fp2 horner(const Fp2* k, int n, const Fp2& x) { Fp2 acc = k[n - 1]; for (int i = n - 2; i >= 0; i--) acc = acc * x + k[i]; return acc; }
static G2A iso3(const Fp2& x, const Fp2& y) {              // hasher.rs:294-348
    Fp2 xd = horner(K2, 3, x), yd = horner(K4, 4, x);
    G2A r; if (is_zero(xd) || is_zero(yd)) { r.inf = true; r.x = F2_ZERO; r.y = F2_ZERO; return r; }
    r.inf = false; r.x = horner(K1, 4, x) * inv(xd); r.y = y * horner(K3, 4, x) * inv(yd); return r;
}
static G2J map_uncleared(const u8* msg, size_t mlen) {
    Fp2 u[2]; hash_to_field(u, msg, mlen);
    Fp2 x0, y0, x1, y1; sswu(x0, y0, u[0]); sswu(x1, y1, u[1]);
    return add(to_jac(iso3(x0, y0)), to_jac(iso3(x1, y1)));
}
static G2J clear_cofactor_psi(const G2J& P) {              // Budroni-Pintore: [x^2-x-1]P + [x-1]psi(P) + psi^2(2P)
    G2J t1 = neg(smul(P, X_BE, 8)), t2 = psi(P), t3 = psi(psi(dbl(P)));
    t3 = add(t3, neg(t2)); t2 = add(t1, t2); t2 = neg(smul(t2, X_BE, 8)); t3 = add(t3, t2); t3 = add(t3, neg(t1));
    return add(t3, neg(P));
}
static G2A hash_to_g2(const u8* msg, size_t mlen) {
    G2J m = map_uncleared(msg, mlen);
    return to_aff(ORA_FAST ? clear_cofactor_psi(m) : smul(m, HEFF_BE, 80));
}

// ---------------------------------------------------------------------------------------- pairing (ark-ec bls12 style)
static const u64 X_ABS = 0xd201000000010000ULL;
struct Coeff { Fp2 a, b, c; };
static Fp FP_TWO_INV;
static void prepare_g2(std::vector<Coeff>& co, const G2A& q) {
    Fp2 rx = q.x, ry = q.y, rz = F2_ONE;
    for (int i = 62; i >= 0; i--) {
        Fp2 a = mulfp(rx * ry, FP_TWO_INV), b = sqr(ry), c = sqr(rz);
        Fp2 e = B2C * (c + c + c), f = e + e + e, g = mulfp(b + f, FP_TWO_INV);
        Fp2 h = sqr(ry + rz) - (b + c), ii = e - b, j = sqr(rx), es = sqr(e);
        rx = a * (b - f); ry = sqr(g) - (es + es + es); rz = b * h;
        co.push_back({ii, j + j + j, -h});
        if ((X_ABS >> i) & 1) {
            Fp2 th = ry - q.y * rz, la = rx - q.x * rz, c2 = sqr(th), d = sqr(la), e2 = la * d;
            Fp2 f2 = rz * c2, g2 = rx * d, h2 = e2 + f2 - (g2 + g2);
            rx = la * h2; ry = th * (g2 - h2) - e2 * ry; rz = rz * e2;
            co.push_back({th * q.x - la * q.y, -th, la});
        }
    }
}
static Fp12 miller(const G1A* ps, const G2A* qs, int n) {
    std::vector<std::vector<Coeff>> co; std::vector<G1A> pp;
    for (int i = 0; i < n; i++) if (!ps[i].inf && !qs[i].inf) { co.emplace_back(); prepare_g2(co.back(), qs[i]); pp.push_back(ps[i]); }
    Fp12 f = F12_ONE; size_t idx = 0;
    for (int i = 62; i >= 0; i--) {
        f = sqr(f);
        for (size_t k = 0; k < pp.size(); k++) { const Coeff& c = co[k][idx]; f = mul_by_014(f, c.a, mulfp(c.b, pp[k].x), mulfp(c.c, pp[k].y)); }
        idx++;
        if ((X_ABS >> i) & 1) {
            for (size_t k = 0; k < pp.size(); k++) { const Coeff& c = co[k][idx]; f = mul_by_014(f, c.a, mulfp(c.b, pp[k].x), mulfp(c.c, pp[k].y)); }
            idx++;
        }
    }
    return conj(f);
}
static void fp4_sqr(Fp2& t0, Fp2& t1, const Fp2& a, const Fp2& b) { Fp2 ab = a * b; t0 = (a + b) * (a + mul_xi(b)) - ab - mul_xi(ab); t1 = ab + ab; }
static Fp12 cyclo_sqr(const Fp12& a) {                     // Granger-Scott, valid in the cyclotomic subgroup
    Fp2 t0, t1, t2, t3, t4, t5; fp4_sqr(t0, t1, a.c0.c0, a.c1.c1); fp4_sqr(t2, t3, a.c1.c0, a.c0.c2); fp4_sqr(t4, t5, a.c0.c1, a.c1.c2);
    auto f = [](const Fp2& t, const Fp2& z, bool plus) { Fp2 d = plus ? t + z : t - z; return d + d + t; };
    Fp12 r; Fp2 x5 = mul_xi(t5);
    r.c0.c0 = f(t0, a.c0.c0, false); r.c1.c1 = f(t1, a.c1.c1, true); r.c1.c0 = f(x5, a.c1.c0, true);
    r.c0.c2 = f(t4, a.c0.c2, false); r.c0.c1 = f(t2, a.c0.c1, false); r.c1.c2 = f(t3, a.c1.c2, true);
    return r;
}
static Fp12 exp_by_x(const Fp12& a) {                      // a^x, x negative; a in the cyclotomic subgroup
    Fp12 r = F12_ONE;
    for (int i = 63; i >= 0; i--) { r = ORA_FAST ? cyclo_sqr(r) : sqr(r); if ((X_ABS >> i) & 1) r = r * a; }
    return conj(r);
}
static Fp12 final_exp(const Fp12& f) {                     // ark-ec bls12 final_exponentiation: f^(3(p^12-1)/r)
    Fp12 r = conj(f) * inv(f);
    r = frob(frob(r)) * r;
    Fp12 y0 = sqr(r), y1 = exp_by_x(r), y2 = conj(r);
    y1 = y1 * y2; y2 = exp_by_x(y1); y1 = conj(y1); y1 = y1 * y2; y2 = exp_by_x(y1);
    y1 = frob(y1); y1 = y1 * y2; r = r * y0; y0 = exp_by_x(y1); y2 = exp_by_x(y0);
    y0 = frob(frob(y1)); y1 = conj(y1); y1 = y1 * y2; y1 = y1 * y0; r = r * y1;
    return r;
}

// ---------------------------------------------------------------------------------------- init
static bool INITED = false;
static void hex_to_be(u8* out, int n, const char* h) {
    size_t len = strlen(h); memset(out, 0, n);
    for (size_t i = 0; i < len; i++) { char c = h[len - 1 - i]; int v = c <= '9' ? c - '0' : (c | 32) - 'a' + 10; out[n - 1 - i / 2] |= (u8)(v << (4 * (i & 1))); }
}
static Fp2 f2hex(const char* a, const char* b) { return {fp_from_hex(a), fp_from_hex(b)}; }
static void init() {
    if (INITED) return;
    memset(&FP_ZERO, 0, sizeof FP_ZERO);
    // R = 2^384 mod p by repeated doubling of 1; R2 by 384 more doublings
    Fp t = {{1, 0, 0, 0, 0, 0}};
    for (int i = 0; i < 384; i++) t = fadd(t, t);
    FP_ONE = t;
    for (int i = 0; i < 384; i++) t = fadd(t, t);
    FP_R2 = t;
    E_PM2 = big_p(); big_add_small(E_PM2, -2);
    E_P1D4 = big_p(); big_add_small(E_P1D4, 1); big_div_small(E_P1D4, 4);
    E_PM3D4 = big_p(); big_add_small(E_PM3D4, -3); big_div_small(E_PM3D4, 4);
    E_PM1D2 = big_p(); big_add_small(E_PM1D2, -1); big_div_small(E_PM1D2, 2);
    E_PM1D6 = big_p(); big_add_small(E_PM1D6, -1); big_div_small(E_PM1D6, 6);
    F2_ZERO = {FP_ZERO, FP_ZERO}; F2_ONE = {FP_ONE, FP_ZERO};
    F12_ONE.c0 = {F2_ONE, F2_ZERO, F2_ZERO}; F12_ONE.c1 = {F2_ZERO, F2_ZERO, F2_ZERO};
    FP_TWO_INV = finv(fp_from_u64(2));
    B1 = fp_from_u64(4); B2C = {B1, B1};
    Fp2 xi = {FP_ONE, FP_ONE}, g = pow2(xi, E_PM1D6.l, 6);
    FROB_G[0] = F2_ONE; for (int i = 1; i < 6; i++) FROB_G[i] = FROB_G[i - 1] * g;
    hex_to_be(X_BE, 8, "d201000000010000");
    { Big e3 = big_p(); big_add_small(e3, -1); big_div_small(e3, 3); Big e2 = E_PM1D2;
      PSI_CX = inv(pow2(xi, e3.l, 6)); PSI_CY = inv(pow2(xi, e2.l, 6)); }
    hex_to_be(R_BE, 32, "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001");
    hex_to_be(HEFF_BE, 80, "0bc69f08f2ee75b3584c6a0ea91b352888e2a8e9145ad7689986ff031508ffe1329c2f178731db956d82bf015d1212b02ec0ec69d7477c1ae954cbc06689f6a359894c0adebbf6b4e8020005aaa95551");   // hasher.rs:666
    G1_GEN.inf = false;
    G1_GEN.x = fp_from_hex("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
    G1_GEN.y = fp_from_hex("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1");
    G1_GEN_NEG = G1_GEN; G1_GEN_NEG.y = fneg(G1_GEN.y);
    { // beta: the primitive cube root of unity with sigma(P) = (beta x, y) = -[x^2]P on G1 (found by testing both on the generator)
      Big e3 = big_p(); big_add_small(e3, -1); big_div_small(e3, 3);
      Fp b = fpow(fp_from_u64(2), e3.l, 6); if (feq(b, FP_ONE)) b = fpow(fp_from_u64(3), e3.l, 6);
      G1J t = smul(smul(to_jac(G1_GEN), X_BE, 8), X_BE, 8);
      G1A c1 = {fmul(G1_GEN.x, b), fneg(G1_GEN.y), false};
      BETA = jac_eq_aff(t, c1) ? b : fsqr(b); }
    ISO_A = {FP_ZERO, fp_from_u64(240)}; ISO_B = {fp_from_u64(1012), fp_from_u64(1012)};          // hasher.rs:229-236
    SSWU_Z = {fneg(fp_from_u64(2)), fneg(FP_ONE)};                                                // hasher.rs:237-240
    // 3-isogeny coefficients, RFC 9380 E.3 (ascending degree)
    const char* k10 = "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6";
    K1[0] = f2hex(k10, k10);
    K1[1] = f2hex("0", "11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a");
    K1[2] = f2hex("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e",
                  "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d");
    K1[3] = f2hex("171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1", "0");
    K2[0] = f2hex("0", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa63");
    K2[1] = f2hex("c", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa9f");
    K2[2] = F2_ONE;
    const char* k30 = "1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706";
    K3[0] = f2hex(k30, k30);
    K3[1] = f2hex("0", "5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be");
    K3[2] = f2hex("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c",
                  "8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f");
    K3[3] = f2hex("124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10", "0");
    const char* k40 = "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa8fb";
    K4[0] = f2hex(k40, k40);
    K4[1] = f2hex("0", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa9d3");
    K4[2] = f2hex("12", "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa99");
    K4[3] = F2_ONE;
    INITED = true;
}

// ---------------------------------------------------------------------------------------- scheme
enum { ST_OK = 0, ST_FALSE = 1, ST_BAD_PK = 2, ST_BAD_SIG = 3, ST_EMPTY = 4, ST_BAD_SK = 5 };
static int verify_points(const G1A& pk, const u8* msg, size_t mlen, const G2A& sig, Fp12* gt_out) {   // bls.rs:427-458
    if (pk.inf) return ST_BAD_PK;                                                    // :434
    if (!on_curve(pk) || !in_subgroup(pk)) return ST_BAD_PK;                         // :438
    if (!on_curve(sig) || !in_subgroup(sig)) return ST_BAD_SIG;                      // :443
    G2A h = hash_to_g2(msg, mlen);                                                   // :452
    G1A ps[2] = {G1_GEN_NEG, pk}; G2A qs[2] = {sig, h};
    Fp12 gt = final_exp(miller(ps, qs, 2));                                          // :454-455
    if (gt_out) *gt_out = gt;
    return gt == F12_ONE ? ST_OK : ST_FALSE;                                         // :457
}
static int verify_bytes(const u8* pk48, const u8* msg, size_t mlen, const u8* sig96, Fp12* gt_out) {
    G1A pk; if (deser_g1(pk, pk48) != DE_OK || pk.inf) return ST_BAD_PK;
    G2A sig; if (deser_g2(sig, sig96) != DE_OK) return ST_BAD_SIG;
    // points decoded with Validate::Yes are already checked; verify_points re-checks like bls.rs:438,443
    return verify_points(pk, msg, mlen, sig, gt_out);
}
static bool sk_valid(const u8* sk_le) {                    // canonical Fr: < r
    for (int i = 31; i >= 0; i--) { u8 rb = R_BE[31 - i]; if (sk_le[i] != rb) return sk_le[i] < rb; }
    return false;
}
static void sk_be(u8* be, const u8* le) { for (int i = 0; i < 32; i++) be[i] = le[31 - i]; }

template <class Fn> static void parallel_for(size_t n, int threads, Fn fn) {
    if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<size_t> next(0); std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&]() { for (;;) { size_t i = next.fetch_add(1); if (i >= n) break; fn(i); } });
    for (auto& t : th) t.join();
}
static inline const u8* msg_ptr(const u8* msg, const uint32_t* off, size_t i, size_t& len) {
    if (off) { len = off[i + 1] - off[i]; return msg + off[i]; } len = 32; return msg + 32 * i;
}

extern "C" {
int ora_init() { init(); return 0; }
int ora_set_fast(int on) { init(); ORA_FAST = on ? 1 : 0; return 0; }
// Fp Montgomery product on raw 48-byte LE limb images (for the K0 parity test)
void ora_fp_mul_raw(const u8* a, const u8* b, u8* out, size_t n) {
    init(); for (size_t i = 0; i < n; i++) { Fp x, y; memcpy(&x, a + 48 * i, 48); memcpy(&y, b + 48 * i, 48); Fp z = fmul(x, y); memcpy(out + 48 * i, &z, 48); }
}
void ora_expand_xmd(const u8* msg, size_t mlen, const u8* dst, size_t dlen, u8* out, size_t n) { init(); expand_xmd(out, n, msg, mlen, dst, dlen); }
int ora_deser_g1(const u8* in48, size_t n, u8* status) { init(); for (size_t i = 0; i < n; i++) { G1A a; status[i] = (u8)deser_g1(a, in48 + 48 * i); } return 0; }
int ora_deser_g2(const u8* in96, size_t n, u8* status) { init(); for (size_t i = 0; i < n; i++) { G2A a; status[i] = (u8)deser_g2(a, in96 + 96 * i); } return 0; }
int ora_hash_to_g2(const u8* msg, const uint32_t* off, size_t n, u8* out96, int cleared, int threads) {
    init();
    parallel_for(n, threads, [&](size_t i) { size_t len; const u8* m = msg_ptr(msg, off, i, len);
        G2A h = cleared ? hash_to_g2(m, len) : to_aff(map_uncleared(m, len)); ser_g2(out96 + 96 * i, h); });
    return 0;
}
int ora_verify(const u8* pk48, const u8* msg, const uint32_t* off, const u8* sig96, size_t n, u8* status, u8* gt_acc576, int threads) {
    init(); std::vector<Fp12> gts(gt_acc576 ? n : 0);
    parallel_for(n, threads, [&](size_t i) { size_t len; const u8* m = msg_ptr(msg, off, i, len);
        Fp12 gt = F12_ONE; status[i] = (u8)verify_bytes(pk48 + 48 * i, m, len, sig96 + 96 * i, &gt);
        if (gt_acc576) gts[i] = status[i] <= ST_FALSE ? gt : F12_ONE; });
    if (gt_acc576) { Fp12 acc = F12_ONE; for (size_t i = 0; i < n; i++) acc = acc * gts[i]; fp12_to_le(gt_acc576, acc); }
    return 0;
}
int ora_sk_to_pk(const u8* sk_le, size_t n, u8* pk48, int threads) {
    init(); parallel_for(n, threads, [&](size_t i) { u8 be[32]; sk_be(be, sk_le + 32 * i); ser_g1(pk48 + 48 * i, to_aff(smul(to_jac(G1_GEN), be, 32))); });
    return 0;
}
int ora_sign(const u8* sk_le, const u8* msg, const uint32_t* off, size_t n, u8* sig96, u8* status, int threads) {
    init();
    parallel_for(n, threads, [&](size_t i) { size_t len; const u8* m = msg_ptr(msg, off, i, len);
        const u8* sk = sk_le + 32 * i; bool zero = true; for (int k = 0; k < 32; k++) zero &= sk[k] == 0;
        if (zero || !sk_valid(sk)) { status[i] = ST_BAD_SK; memset(sig96 + 96 * i, 0, 96); return; }
        u8 be[32]; sk_be(be, sk); ser_g2(sig96 + 96 * i, to_aff(smul(to_jac(hash_to_g2(m, len)), be, 32))); status[i] = ST_OK; });
    return 0;
}
// aggregate: status 0 ok, 2/3 = an input failed to decode (G1/G2), 4 = empty segment (None)
int ora_g1_aggregate(const u8* pts48, const uint32_t* seg, size_t nseg, u8* out48, u8* status, int threads) {
    init();
    parallel_for(nseg, threads, [&](size_t s) {
        if (seg[s + 1] == seg[s]) { status[s] = ST_EMPTY; memset(out48 + 48 * s, 0, 48); return; }
        G1J acc = jac_identity<Fp>(); status[s] = ST_OK;
        for (uint32_t i = seg[s]; i < seg[s + 1]; i++) { G1A a; if (deser_g1(a, pts48 + 48 * (size_t)i) != DE_OK) { status[s] = ST_BAD_PK; break; } acc = add(acc, to_jac(a)); }
        if (status[s] == ST_OK) ser_g1(out48 + 48 * s, to_aff(acc)); else memset(out48 + 48 * s, 0, 48); });
    return 0;
}
int ora_g2_aggregate(const u8* pts96, const uint32_t* seg, size_t nseg, u8* out96, u8* status, int threads) {
    init();
    parallel_for(nseg, threads, [&](size_t s) {
        if (seg[s + 1] == seg[s]) { status[s] = ST_EMPTY; memset(out96 + 96 * s, 0, 96); return; }
        G2J acc = jac_identity<Fp2>(); status[s] = ST_OK;
        for (uint32_t i = seg[s]; i < seg[s + 1]; i++) { G2A a; if (deser_g2(a, pts96 + 96 * (size_t)i) != DE_OK) { status[s] = ST_BAD_SIG; break; } acc = add(acc, to_jac(a)); }
        if (status[s] == ST_OK) ser_g2(out96 + 96 * s, to_aff(acc)); else memset(out96 + 96 * s, 0, 96); });
    return 0;
}
// fast_aggregate_verify: k keys per committee, optional participation bitmap (bit j of committee c at bit c*k+j)
int ora_fast_aggregate_verify(const u8* pks48, const uint64_t* bitmap, size_t k, const u8* msg32, const u8* sig96, size_t ncomm,
                              u8* status, u8* agg48, int threads) {
    init();
    parallel_for(ncomm, threads, [&](size_t c) {
        G1J acc = jac_identity<Fp>(); bool bad = false; size_t used = 0;
        for (size_t j = 0; j < k && !bad; j++) {
            size_t bit = c * k + j; if (bitmap && !((bitmap[bit / 64] >> (bit % 64)) & 1)) continue;
            G1A a; if (deser_g1(a, pks48 + 48 * bit) != DE_OK) { bad = true; break; } acc = add(acc, to_jac(a)); used++;
        }
        G1A agg = to_aff(acc);
        if (agg48) { if (bad || used == 0) { memset(agg48 + 48 * c, 0, 48); } else ser_g1(agg48 + 48 * c, agg); }
        if (bad || used == 0) { status[c] = ST_BAD_PK; return; }                    // empty => None => identity pk => false (tests.rs:312-316)
        G2A sig; if (deser_g2(sig, sig96 + 96 * c) != DE_OK) { status[c] = agg.inf ? ST_BAD_PK : ST_BAD_SIG; return; }
        status[c] = (u8)verify_points(agg, msg32 + 32 * c, 32, sig, nullptr); });
    return 0;
}
// GT of a product of pairings of decoded points (parity hook for GT bytes)
int ora_pairing_gt(const u8* g1_48, const u8* g2_96, size_t npairs, u8* gt576) {
    init(); std::vector<G1A> ps(npairs); std::vector<G2A> qs(npairs);
    for (size_t i = 0; i < npairs; i++) { if (deser_g1(ps[i], g1_48 + 48 * i) != DE_OK) return -2; if (deser_g2(qs[i], g2_96 + 96 * i) != DE_OK) return -3; }
    fp12_to_le(gt576, final_exp(miller(ps.data(), qs.data(), (int)npairs))); return 0;
}
int ora_gt_mul(const u8* a576, const u8* b576, u8* out576) { init(); Fp12 a, b; if (!fp12_from_le(a, a576) || !fp12_from_le(b, b576)) return -1; fp12_to_le(out576, a * b); return 0; }
// R1CS: CSR x3 with canonical 48-byte LE coefficients; z = nwit vectors of ncols canonical 48-byte LE values.
// sat_bits: nwit * ceil(nrows/64) words, bit i set <=> row i satisfied; all_sat[w] = 1 iff every row holds.
int ora_r1cs_check(const uint64_t* const rowptr[3], const uint32_t* const col[3], const u8* const coeff48[3], size_t nrows, size_t ncols,
                   const u8* z48, size_t nwit, uint64_t* sat_bits, u8* all_sat, int threads) {
    init(); size_t words = (nrows + 63) / 64;
    std::vector<Fp> cf[3];
    for (int m = 0; m < 3; m++) { size_t nnz = rowptr[m][nrows]; cf[m].resize(nnz); for (size_t k = 0; k < nnz; k++) if (!fp_from_le(cf[m][k], coeff48[m] + 48 * k)) return -1; }
    std::atomic<int> bad(0);
    parallel_for(nwit, threads, [&](size_t w) {
        std::vector<Fp> z(ncols);
        for (size_t j = 0; j < ncols; j++) if (!fp_from_le(z[j], z48 + 48 * (w * ncols + j))) { bad = 1; return; }
        uint64_t* bits = sat_bits + w * words; memset(bits, 0, words * 8); bool all = true;
        for (size_t i = 0; i < nrows; i++) {
            Fp d[3];
            for (int m = 0; m < 3; m++) { Fp acc = FP_ZERO; for (uint64_t k = rowptr[m][i]; k < rowptr[m][i + 1]; k++) acc = fadd(acc, fmul(cf[m][k], z[col[m][k]])); d[m] = acc; }
            bool ok = feq(fmul(d[0], d[1]), d[2]);
            if (ok) bits[i / 64] |= 1ULL << (i % 64); else all = false;
        }
        all_sat[w] = all; });
    return bad ? -1 : 0;
}
// Eth2 AggregateVerify (draft-irtf-cfrg-bls-signature-05 3.1.1 under the POP scheme; upstream bls12-381-tests "aggregate_verify", the
// category reference tests/readme.md:4-7 names): signature s covers pairs [pair_off[s], pair_off[s+1]); every key KeyValidate-d, the
// signature subgroup-checked, then ONE (k+1)-pair product e(-g1, sig) prod e(pk_j, H(m_j)) == 1.  status: 0 / 1, 2 key, 3 signature, 4 no pairs.
int ora_aggregate_verify(const u8* pks48, const u8* msg, const uint32_t* off, const uint32_t* pair_off, const u8* sig96, size_t nsig, u8* status, int threads) {
    init();
    parallel_for(nsig, threads, [&](size_t s) {
        uint32_t lo = pair_off[s], hi = pair_off[s + 1];
        std::vector<G1A> ps; std::vector<G2A> qs; bool bad_pk = false;
        for (uint32_t j = lo; j < hi && !bad_pk; j++) {
            G1A pk; if (deser_g1(pk, pks48 + 48 * (size_t)j) != DE_OK || pk.inf) { bad_pk = true; break; }
            size_t len; const u8* m = msg_ptr(msg, off, j, len);
            ps.push_back(pk); qs.push_back(hash_to_g2(m, len));
        }
        if (bad_pk) { status[s] = ST_BAD_PK; return; }
        G2A sig; if (deser_g2(sig, sig96 + 96 * s) != DE_OK) { status[s] = ST_BAD_SIG; return; }
        if (lo == hi) { status[s] = ST_EMPTY; return; }
        if (!sig.inf) { ps.push_back(G1_GEN_NEG); qs.push_back(sig); }                 // ark-ec drops pairs that contain the identity
        status[s] = final_exp(miller(ps.data(), qs.data(), (int)ps.size())) == F12_ONE ? ST_OK : ST_FALSE; });
    return 0;
}
// compressed <-> uncompressed point encodings; status = DE_* of the input, undecodable input -> all-zero output
int ora_g1_recode(const u8* in, size_t n, u8* out, u8* status, int to_unc) {
    init(); for (size_t i = 0; i < n; i++) { G1A a; int rc = to_unc ? deser_g1(a, in + 48 * i) : deser_g1_unc(a, in + 96 * i); status[i] = (u8)rc;
        u8* o = out + (to_unc ? 96 : 48) * i; if (rc != DE_OK) memset(o, 0, to_unc ? 96 : 48); else if (to_unc) ser_g1_unc(o, a); else ser_g1(o, a); }
    return 0;
}
int ora_g2_recode(const u8* in, size_t n, u8* out, u8* status, int to_unc) {
    init(); for (size_t i = 0; i < n; i++) { G2A a; int rc = to_unc ? deser_g2(a, in + 96 * i) : deser_g2_unc(a, in + 192 * i); status[i] = (u8)rc;
        u8* o = out + (to_unc ? 192 : 96) * i; if (rc != DE_OK) memset(o, 0, to_unc ? 192 : 96); else if (to_unc) ser_g2_unc(o, a); else ser_g2(o, a); }
    return 0;
}
int ora_hw_threads() { return (int)std::thread::hardware_concurrency(); }
}
