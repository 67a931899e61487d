// Per-item stage functions of the verify path and the limb-SoA HBM layout they exchange data through.
// One item (a (pk, msg, sig) triple, a message, a point) per thread; stages are separate kernels so that each
// gets its own register budget and the intermediate state round-trips through HBM fully coalesced.
//
// HBM layout of intermediates ("limb-SoA"): an array of N records of K field elements each is stored as
//   uint4 [K][3][N]        element k, 128-bit chunk c of item i at ((k*3 + c) * N + i)
// so that a warp reading chunk c of element k of 32 consecutive items touches 512 contiguous bytes.
//
// Control flow mirrors BLS::verify, reference src/bls.rs:427-458, and the test harness' substitution rules
// (tests/tests.rs:244-254); status codes are the C-ABI ones of include/blsgpu.h.
#pragma once
#include "pairing.cuh"
#include "h2c.cuh"

namespace bls {

#if defined(__CUDACC__)
typedef uint4 u32x4;
#else
struct u32x4 { uint32_t x, y, z, w; };
#endif

enum { ST_OK = 0, ST_FALSE = 1, ST_BAD_PK = 2, ST_BAD_SIG = 3, ST_EMPTY = 4, ST_BAD_SK = 5 };
// per-item flags carried beside the decoded signature / hash point
enum { FL_SIG_INF = 1, FL_HM_INF = 2 };

BLS_HD void soa_store_fp(u32x4* base, size_t n, size_t i, int k, const fp& v) {
    u32x4 a, b, c;
    a.x = v.l[0]; a.y = v.l[1]; a.z = v.l[2]; a.w = v.l[3];
    b.x = v.l[4]; b.y = v.l[5]; b.z = v.l[6]; b.w = v.l[7];
    c.x = v.l[8]; c.y = v.l[9]; c.z = v.l[10]; c.w = v.l[11];
    base[(size_t)(3 * k) * n + i] = a; base[(size_t)(3 * k + 1) * n + i] = b; base[(size_t)(3 * k + 2) * n + i] = c;
}
BLS_HD fp soa_load_fp(const u32x4* base, size_t n, size_t i, int k) {
    u32x4 a = base[(size_t)(3 * k) * n + i], b = base[(size_t)(3 * k + 1) * n + i], c = base[(size_t)(3 * k + 2) * n + i];
    fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    v.l[8] = c.x; v.l[9] = c.y; v.l[10] = c.z; v.l[11] = c.w;
    return v;
}
BLS_HD void soa_store_fp2(u32x4* base, size_t n, size_t i, int k, const fp2& v) { soa_store_fp(base, n, i, 2 * k, v.c0); soa_store_fp(base, n, i, 2 * k + 1, v.c1); }
BLS_HD fp2 soa_load_fp2(const u32x4* base, size_t n, size_t i, int k) { fp2 v; v.c0 = soa_load_fp(base, n, i, 2 * k); v.c1 = soa_load_fp(base, n, i, 2 * k + 1); return v; }
BLS_HD void soa_store_g1(u32x4* base, size_t n, size_t i, const g1_aff& p) { soa_store_fp(base, n, i, 0, p.x); soa_store_fp(base, n, i, 1, p.y); }
BLS_HD void soa_load_g1(g1_aff& p, const u32x4* base, size_t n, size_t i) { p.x = soa_load_fp(base, n, i, 0); p.y = soa_load_fp(base, n, i, 1); }
BLS_HD void soa_store_g2(u32x4* base, size_t n, size_t i, const g2_aff& p) { soa_store_fp2(base, n, i, 0, p.x); soa_store_fp2(base, n, i, 1, p.y); }
BLS_HD void soa_load_g2(g2_aff& p, const u32x4* base, size_t n, size_t i) { p.x = soa_load_fp2(base, n, i, 0); p.y = soa_load_fp2(base, n, i, 1); }
BLS_HD void soa_store_fp12(u32x4* base, size_t n, size_t i, const fp12& f) {
    soa_store_fp2(base, n, i, 0, f.c0.c0); soa_store_fp2(base, n, i, 1, f.c0.c1); soa_store_fp2(base, n, i, 2, f.c0.c2);
    soa_store_fp2(base, n, i, 3, f.c1.c0); soa_store_fp2(base, n, i, 4, f.c1.c1); soa_store_fp2(base, n, i, 5, f.c1.c2);
}
BLS_HD void soa_load_fp12(fp12& f, const u32x4* base, size_t n, size_t i) {
    f.c0.c0 = soa_load_fp2(base, n, i, 0); f.c0.c1 = soa_load_fp2(base, n, i, 1); f.c0.c2 = soa_load_fp2(base, n, i, 2);
    f.c1.c0 = soa_load_fp2(base, n, i, 3); f.c1.c1 = soa_load_fp2(base, n, i, 4); f.c1.c2 = soa_load_fp2(base, n, i, 5);
}

// ---- stage 1: public key bytes -> affine G1 (bls.rs:219-223 + the identity test and check() of bls.rs:434-442)
BLS_HD uint8_t stage_decode_pk(g1_aff& pk, const uint8_t* pk48) {
    int rc = g1_decode(pk, pk48);
    return rc == DEC_OK ? ST_OK : ST_BAD_PK;                 // identity, undecodable, off-curve, wrong subgroup
}
// ---- stage 2: signature bytes -> affine G2 (bls.rs:316-320 + check() of bls.rs:443-447); the identity is accepted
BLS_HD uint8_t stage_decode_sig(g2_aff& sig, uint8_t& flags, const uint8_t* sig96) {
    int rc = g2_decode(sig, sig96);
    flags = rc == DEC_INF ? FL_SIG_INF : 0;
    return (rc == DEC_OK || rc == DEC_INF) ? ST_OK : ST_BAD_SIG;
}
// ---- stage 3: message -> H(m) affine (bls.rs:452, 477-493)
BLS_HD void stage_hash(g2_aff& hm, uint8_t& flags, const uint8_t* msg, uint32_t mlen) {
    g2_jac h; hash_to_g2_jac(h, msg, mlen, true);
    flags = jac_to_aff(hm, h) ? 0 : FL_HM_INF;
}
// ---- stage 4: Miller loop of e(-g1, sig) e(pk, H(m))  (bls.rs:449-455, first half of multi_pairing)
BLS_HD void stage_miller(fp12& f, const g1_aff& pk, const g2_aff& hm, const g2_aff& sig, uint8_t flags) {
    g1_aff ng; ng.x = fp_const(C_G1X); ng.y = fp_const(C_G1Y_NEG);
    miller_loop2(f, ng, sig, !(flags & FL_SIG_INF), pk, hm, !(flags & FL_HM_INF));
}
// ---- stage 5: final exponentiation and is_one (bls.rs:455-457)
BLS_HD uint8_t stage_final(fp12& gt, const fp12& f) {
    final_exponentiation(gt, f);
    return fp12_is_one(gt) ? ST_OK : ST_FALSE;
}

}  // namespace bls
