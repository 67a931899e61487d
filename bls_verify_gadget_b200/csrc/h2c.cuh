// hash_to_g2: RFC 9380 BLS12381G2_XMD:SHA-256_SSWU_RO_ with the POP domain separation tag.
// Replaces reference src/bls.rs:477-493 (MapToCurveBasedHasher<G2, DefaultFieldHasher<Sha256,128>, WBMap>);
// the in-repo algorithm spec is src/hasher.rs: expand_message_xmd 110-173, hash_to_field 58-107,
// SSWU 352-502 (constants 229-258), 3-isogeny 294-348, cofactor clearing 664-673 (psi constants 600-616).
//
// B200-first choices (results are bit-identical by construction, the maps are functions):
//  * sqrt_ratio in Fp2 is done with two Fp exponentiations via the norm ("complex method") instead of the
//    758-bit Fp2 exponentiation of hasher.rs:532-548, and without any inversion: 1/N(v) falls out of the
//    same exponentiation that decides squareness;
//  * SSWU output stays a fraction, the isogeny is evaluated homogeneously into Jacobian coordinates;
//  * cofactor clearing uses the psi endomorphism (Budroni-Pintore), equal to [h_eff]P of hasher.rs:666.
#pragma once
#include "curve.cuh"

namespace bls {

// ------------------------------------------------------------------------------------------------ SHA-256
BLS_CONST uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

struct sha256_ctx { uint32_t h[8]; uint32_t w[16]; uint32_t fill; uint32_t total; };   // w: big-endian packed block buffer

BLS_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
BLS_NOINLINE void sha256_compress(uint32_t* h, uint32_t* w) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3), s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t t1 = hh + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i & 15];
        uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
BLS_HD void sha256_init(sha256_ctx& c) {
    const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    for (int i = 0; i < 8; i++) c.h[i] = iv[i];
    for (int i = 0; i < 16; i++) c.w[i] = 0;
    c.fill = 0; c.total = 0;
}
BLS_HD void sha256_put(sha256_ctx& c, uint32_t byte) {
    uint32_t idx = c.fill >> 2, sh = 24 - 8 * (c.fill & 3);
    c.w[idx] |= byte << sh;
    c.fill++; c.total++;
    if (c.fill == 64) { sha256_compress(c.h, c.w); for (int i = 0; i < 16; i++) c.w[i] = 0; c.fill = 0; }
}
BLS_HD void sha256_final(sha256_ctx& c, uint32_t* digest) {     // digest as 8 big-endian words
    uint32_t bits = c.total * 8;
    sha256_put(c, 0x80);
    while (c.fill != 56) sha256_put(c, 0);
    c.w[14] = 0; c.w[15] = bits;
    sha256_compress(c.h, c.w);
    for (int i = 0; i < 8; i++) digest[i] = c.h[i];
}

// DST' = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_" || 0x2b  (reference src/bls.rs:482; 44 bytes)
BLS_CONST uint8_t DST_PRIME[44] = {'B','L','S','_','S','I','G','_','B','L','S','1','2','3','8','1','G','2','_','X','M','D',':','S','H','A','-','2','5','6','_',
                                   'S','S','W','U','_','R','O','_','P','O','P','_', 43};
BLS_HD void sha256_put_dst(sha256_ctx& c) { for (int i = 0; i < 44; i++) sha256_put(c, DST_PRIME[i]); }

// expand_message_xmd(msg, DST, 256) -> 64 big-endian words (hasher.rs:110-173)
BLS_HD void expand_xmd_256(uint32_t* uni, const uint8_t* msg, uint32_t mlen) {
    sha256_ctx c; uint32_t b0[8], bi[8];
    sha256_init(c);
    for (int i = 0; i < 64; i++) sha256_put(c, 0);                      // z_pad (hasher.rs:128)
    for (uint32_t i = 0; i < mlen; i++) sha256_put(c, msg[i]);
    sha256_put(c, 0x01); sha256_put(c, 0x00);                           // l_i_b_str = 256, big-endian (hasher.rs:130)
    sha256_put(c, 0x00);
    sha256_put_dst(c);
    sha256_final(c, b0);
    for (int k = 1; k <= 8; k++) {
        sha256_init(c);
        for (int i = 0; i < 8; i++) {
            uint32_t w = k == 1 ? b0[i] : (b0[i] ^ bi[i]);
            sha256_put(c, w >> 24); sha256_put(c, (w >> 16) & 0xff); sha256_put(c, (w >> 8) & 0xff); sha256_put(c, w & 0xff);
        }
        sha256_put(c, (uint32_t)k);
        sha256_put_dst(c);
        sha256_final(c, bi);
        for (int i = 0; i < 8; i++) uni[8 * (k - 1) + i] = bi[i];
    }
}
BLS_CONST uint32_t C_2_256[12] = BLS_C_2_256;
// 64 big-endian bytes (16 BE words) -> integer mod p, Montgomery form (hasher.rs:71-104).  Both halves are < 2^256 < p.
BLS_HD fp fp_from_be64_words(const uint32_t* w) {
    fp hi = fp_zero(), lo = fp_zero();
    for (int i = 0; i < 8; i++) { hi.l[i] = w[7 - i]; lo.l[i] = w[15 - i]; }
    fp r2 = fp_r2();
    return fp_add(fp_mul(fp_mul(hi, r2), fp_const(C_2_256)), fp_mul(lo, r2));
}
BLS_HD void hash_to_field(fp2& u0, fp2& u1, const uint8_t* msg, uint32_t mlen) {
    uint32_t uni[64];
    expand_xmd_256(uni, msg, mlen);
    u0.c0 = fp_from_be64_words(uni); u0.c1 = fp_from_be64_words(uni + 16);
    u1.c0 = fp_from_be64_words(uni + 32); u1.c1 = fp_from_be64_words(uni + 48);
}

// ------------------------------------------------------------------------------------------------ SSWU
BLS_CONST fp2 C_ISO_A = BLS_C_ISO_A;
BLS_CONST fp2 C_ISO_B = BLS_C_ISO_B;
BLS_CONST fp2 C_SSWU_Z = BLS_C_SSWU_Z;
BLS_CONST uint32_t C_SQRT_M5[12] = BLS_C_SQRT_M5;
BLS_CONST fp2 C_ISO_K1[4] = BLS_C_ISO_K1;
BLS_CONST fp2 C_ISO_K2[3] = BLS_C_ISO_K2;
BLS_CONST fp2 C_ISO_K3[4] = BLS_C_ISO_K3;
BLS_CONST fp2 C_ISO_K4[4] = BLS_C_ISO_K4;

// sqrt_ratio(u, v), v != 0 (RFC 9380 F.2.1): returns is_square(u/v) and y with y^2 = u/v, or y^2 = Z u/v otherwise.
//   u/v = b / n^2 with n = N(v) in Fp and b = u conj(v) n, so sqrt(u/v) = sqrt(b)/n.
//   N(b) is a residue in Fp  <=>  u/v is a square in Fp2.  With t = N(b)^((p-3)/4): s = N(b) t is the candidate root
//   of the norm, chi = s t = +-1 the Legendre symbol, and 1/N(b) = chi t^2, whence 1/n = N(u) n^2 / N(b).
//   Non-square: N(Z b) = 5 N(b) and sqrt(5 N(b)) = sqrt(-5) s because s^2 = -N(b).
BLS_HD bool fp2_sqrt_ratio(fp2& y, const fp2& u, const fp2& v) {
    fp n = fp2_norm(v), nu = fp2_norm(u);
    fp2 b = fp2_mul_fp(fp2_mul(u, fp2_conj(v)), n);
    fp n2 = fp_sqr(n);
    fp nb = fp_mul(nu, fp_mul(n2, n));                          // N(b) = N(u) N(conj v) n^2 = N(u) n^3
    fp t = fp_pow_pm3d4(nb);
    fp s = fp_mul(nb, t);
    bool is_sq = fp_eq(fp_sqr(s), nb);
    fp t2 = fp_sqr(t);
    fp inv_nb = fp_csel(is_sq, t2, fp_neg(t2));                 // chi t^2 = 1/N(b)   (0 when N(b) = 0)
    fp2 bb = fp2_csel(is_sq, b, fp2_mul(b, C_SSWU_Z));
    fp ss = fp_csel(is_sq, s, fp_mul(s, fp_const(C_SQRT_M5)));
    fp2 root = fp2_sqrt_with_norm_root(bb, ss);
    // 1/n = N(b)^-1 * N(b)/n  with N(b)/n = N(u) n^2
    fp inv_n = fp_mul(inv_nb, fp_mul(nu, n2));
    y = fp2_mul_fp(root, inv_n);
    return is_sq;
}

// Simplified SWU onto E': y^2 = x^3 + A'x + B' (RFC 9380 F.2 straight line; hasher.rs:361-496).  x = xn/xd.
BLS_NOINLINE void sswu_map(fp2& xn, fp2& xd, fp2& y, const fp2& u) {
    fp2 A = C_ISO_A, B = C_ISO_B, Z = C_SSWU_Z;
    fp2 tv1 = fp2_mul(Z, fp2_sqr(u));
    fp2 tv2 = fp2_add(fp2_sqr(tv1), tv1);
    fp2 tv3 = fp2_mul(B, fp2_add(tv2, fp2_one()));
    fp2 tv4 = fp2_mul(A, fp2_csel(fp2_is_zero(tv2), Z, fp2_neg(tv2)));
    fp2 t2 = fp2_sqr(tv3), tv6 = fp2_sqr(tv4);
    fp2 tv5 = fp2_mul(A, tv6);
    t2 = fp2_mul(fp2_add(t2, tv5), tv3);
    tv6 = fp2_mul(tv6, tv4);
    tv5 = fp2_mul(B, tv6);
    t2 = fp2_add(t2, tv5);                                       // gx1 = t2 / tv6
    fp2 x2n = fp2_mul(tv1, tv3);
    fp2 y1; bool sq = fp2_sqrt_ratio(y1, t2, tv6);
    fp2 y2 = fp2_mul(fp2_mul(tv1, u), y1);
    xn = fp2_csel(sq, tv3, x2n);
    y = fp2_csel(sq, y1, y2);
    if (fp2_sgn0(u) != fp2_sgn0(y)) y = fp2_neg(y);
    xd = tv4;
}

// 3-isogeny E' -> E2 evaluated on x = xn/xd, y exact, into Jacobian coordinates (hasher.rs:294-348; RFC 9380 E.3).
//   XN = sum k1_i xn^i xd^(3-i), XD = xd * sum k2_i xn^i xd^(2-i), YN, YD likewise (degree 3)
//   x' = XN/XD, y' = y YN/YD  ->  Z = XD YD, X = XN XD YD^2, Y = y YN XD^3 YD^2.   XD YD = 0 gives the identity.
BLS_NOINLINE void iso3_map(g2_jac& r, const fp2& xn, const fp2& xd, const fp2& y) {
    fp2 xn2 = fp2_sqr(xn), xd2 = fp2_sqr(xd);
    fp2 xn3 = fp2_mul(xn2, xn), xd3 = fp2_mul(xd2, xd);
    fp2 xn2xd = fp2_mul(xn2, xd), xnxd2 = fp2_mul(xn, xd2), xnxd = fp2_mul(xn, xd);
    fp2 XN = fp2_add(fp2_add(fp2_mul(C_ISO_K1[0], xd3), fp2_mul(C_ISO_K1[1], xnxd2)), fp2_add(fp2_mul(C_ISO_K1[2], xn2xd), fp2_mul(C_ISO_K1[3], xn3)));
    fp2 XD = fp2_mul(xd, fp2_add(fp2_add(fp2_mul(C_ISO_K2[0], xd2), fp2_mul(C_ISO_K2[1], xnxd)), xn2));
    fp2 YN = fp2_add(fp2_add(fp2_mul(C_ISO_K3[0], xd3), fp2_mul(C_ISO_K3[1], xnxd2)), fp2_add(fp2_mul(C_ISO_K3[2], xn2xd), fp2_mul(C_ISO_K3[3], xn3)));
    fp2 YD = fp2_add(fp2_add(fp2_mul(C_ISO_K4[0], xd3), fp2_mul(C_ISO_K4[1], xnxd2)), fp2_add(fp2_mul(C_ISO_K4[2], xn2xd), xn3));
    fp2 Zr = fp2_mul(XD, YD);
    fp2 YD2 = fp2_sqr(YD), XD2 = fp2_sqr(XD);
    r.X = fp2_mul(fp2_mul(XN, XD), YD2);
    r.Y = fp2_mul(fp2_mul(fp2_mul(y, YN), fp2_mul(XD2, XD)), YD2);
    r.Z = Zr;
    if (fp2_is_zero(Zr)) jac_set_identity(r);
}

// clear_cofactor via psi (Budroni-Pintore): [x^2 - x - 1]P + [x - 1]psi(P) + psi^2(2P), x = -|x|; equals [h_eff]P (hasher.rs:664-673)
BLS_NOINLINE void g2_clear_cofactor(g2_jac& out, const g2_jac& P) {
    g2_jac t1, t2, t3, nP;
    jac_mul_x_abs_j(t1, P); jac_neg(t1, t1);            // t1 = [x]P
    g2_psi(t2, P);                                      // t2 = psi(P)
    jac_dbl(t3, P); g2_psi2(t3, t3);                    // t3 = psi^2(2P)
    g2_jac n2; jac_neg(n2, t2); jac_add(t3, t3, n2);    // t3 = psi^2(2P) - psi(P)
    jac_add(t2, t1, t2);                                // t2 = [x]P + psi(P)
    jac_mul_x_abs_j(t2, t2); jac_neg(t2, t2);           // t2 = [x^2]P + [x]psi(P)
    jac_add(t3, t3, t2);
    jac_neg(t1, t1); jac_add(t3, t3, t1);               // - [x]P
    jac_neg(nP, P); jac_add(out, t3, nP);               // - P
}

// message -> H(m) in G2 (Jacobian).  One message per thread.
BLS_NOINLINE void hash_to_g2_jac(g2_jac& out, const uint8_t* msg, uint32_t mlen, bool clear) {
    fp2 u0, u1; hash_to_field(u0, u1, msg, mlen);
    fp2 xn, xd, y; g2_jac q0, q1;
    sswu_map(xn, xd, y, u0); iso3_map(q0, xn, xd, y);
    sswu_map(xn, xd, y, u1); iso3_map(q1, xn, xd, y);
    jac_add(q0, q0, q1);
    if (clear) g2_clear_cofactor(out, q0); else out = q0;
}

}  // namespace bls
