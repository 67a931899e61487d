// Lazy reduction for the tower: unreduced 768-bit sums of Fp products in SPLIT accumulators, one Montgomery reduction
// per output coefficient instead of one per product.
//
// Why split: a 32x32->64 MAC is one IMAD.WIDE.U32(.X) only when its 64-bit addend sits in an ALIGNED register pair.  In a
// row a * b_i the products a_j b_i land on limbs i+j, i+j+1: even and odd i+j need differently aligned pairs, so one
// 24-limb array forces ptxas to shuffle registers (the first lazy attempt, fp_mul_wide in fp.cuh, compiles to 270 MOVs
// per product and was slower than reducing every product).  Here the accumulator is two arrays,
//     e[k] <-> limb k          (pairs (e[2m], e[2m+1]) hold the products with i+j even)
//     o[k] <-> limb k + 1      (pairs (o[2m], o[2m+1]) hold the products with i+j odd)
// plus one small counter per upper limb, c[k] <-> limb 12 + k, that collects the carry out of every 6-MAC chain (a chain
// that ends at limb L carries into limb L+1 >= 12; counting instead of rippling keeps the chains short and independent).
// Value = e + (o << 32) + (c << 384).  Every MAC chain is the same `cmad_n` as the interleaved product in fp.cuh.
//
// Cost model per Fp2 output coefficient that is a sum of n Fp2 products (Karatsuba inside each): 3n wide products
// (144 IMAD.WIDE each) + 2 reductions (156) against n * 900 for separately reduced products; the additions between them
// act on unreduced values (no conditional subtractions).  Used by the Fp6/Fp12 routines of tower.cuh (reference
// src/bls.rs:454-457 through ark-ff's Fp6/Fp12 arithmetic).
#pragma once
#include "fp2.cuh"

namespace bls {

// 64-bit accumulator slots keep every (lo, hi) pair in an aligned register pair across loops and calls (as separate
// 32-bit variables, loop-carried pairs were split by the register allocator and copied back before every IMAD.WIDE).
//   e[m] <-> limbs 2m, 2m+1      o[m] <-> limbs 2m+1, 2m+2      c[k] <-> limb 12+k
struct wacc { uint64_t e[12]; uint64_t o[12]; uint32_t c[14]; };       // o[11], c[12], c[13] stay zero (value < 2^768)

BLS_HD uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
BLS_HD void wacc_zero(wacc& w) {
#pragma unroll
    for (int i = 0; i < 12; i++) { w.e[i] = 0; w.o[i] = 0; }
#pragma unroll
    for (int i = 0; i < 14; i++) w.c[i] = 0;
}
BLS_HD void wacc_set(wacc& w, const uint32_t* v /*24 words*/) {
    wacc_zero(w);
#pragma unroll
    for (int i = 0; i < 12; i++) w.e[i] = pack64(v[2 * i], v[2 * i + 1]);
}
// acc[0..5] (six consecutive 64-bit slots) += (a0, a2, a4, a6, a8, a10) * bi along one carry chain; carry out -> top
#if defined(__CUDA_ARCH__)
BLS_HD void cmad64(uint64_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t a8, uint32_t a10, uint32_t bi) {
    BLS_ASM("{\n\t.reg .u32 l0, h0, l1, h1, l2, h2, l3, h3, l4, h4, l5, h5;\n\t"
        "mov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\tmov.b64 {l2, h2}, %2;\n\tmov.b64 {l3, h3}, %3;\n\tmov.b64 {l4, h4}, %4;\n\tmov.b64 {l5, h5}, %5;\n\t"
        "mad.lo.cc.u32 l0, %7, %13, l0;\n\tmadc.hi.cc.u32 h0, %7, %13, h0;\n\t"
        "madc.lo.cc.u32 l1, %8, %13, l1;\n\tmadc.hi.cc.u32 h1, %8, %13, h1;\n\t"
        "madc.lo.cc.u32 l2, %9, %13, l2;\n\tmadc.hi.cc.u32 h2, %9, %13, h2;\n\t"
        "madc.lo.cc.u32 l3, %10, %13, l3;\n\tmadc.hi.cc.u32 h3, %10, %13, h3;\n\t"
        "madc.lo.cc.u32 l4, %11, %13, l4;\n\tmadc.hi.cc.u32 h4, %11, %13, h4;\n\t"
        "madc.lo.cc.u32 l5, %12, %13, l5;\n\tmadc.hi.cc.u32 h5, %12, %13, h5;\n\t"
        "addc.u32 %6, %6, 0;\n\t"
        "mov.b64 %0, {l0, h0};\n\tmov.b64 %1, {l1, h1};\n\tmov.b64 %2, {l2, h2};\n\tmov.b64 %3, {l3, h3};\n\tmov.b64 %4, {l4, h4};\n\tmov.b64 %5, {l5, h5};\n\t}"
        : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2]), "+l"(acc[3]), "+l"(acc[4]), "+l"(acc[5]), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(a8), "r"(a10), "r"(bi));
}
// x += y, g = carry out
BLS_HD void add64c(uint64_t& x, uint32_t& g, uint64_t y) {
    BLS_ASM("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+l"(x), "=r"(g) : "l"(y));
}
#else
BLS_HD void cmad64(uint64_t* acc, uint32_t& top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t a8, uint32_t a10, uint32_t bi) {
    const uint32_t a[6] = {a0, a2, a4, a6, a8, a10};
    uint64_t c = 0;
    for (int j = 0; j < 6; j++) {
        unsigned __int128 t = (unsigned __int128)a[j] * bi + acc[j] + c;
        acc[j] = (uint64_t)t; c = (uint64_t)(t >> 64);
    }
    top += (uint32_t)c;
}
BLS_HD void add64c(uint64_t& x, uint32_t& g, uint64_t y) { unsigned __int128 s = (unsigned __int128)x + y; x = (uint64_t)s; g = (uint32_t)(s >> 64); }
#endif
// w += a * b   (a, b < 2^384; the caller keeps the running value below 2^768)
BLS_HD void wmac(wacc& w, const fp& a, const fp& b) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        if ((i & 1) == 0) {
            cmad64(&w.e[i / 2], w.c[i], a.l[0], a.l[2], a.l[4], a.l[6], a.l[8], a.l[10], b.l[i]);           // limbs i .. i+11, carry -> limb i+12
            cmad64(&w.o[i / 2], w.c[i + 1], a.l[1], a.l[3], a.l[5], a.l[7], a.l[9], a.l[11], b.l[i]);       // limbs i+1 .. i+12, carry -> limb i+13
        } else {
            cmad64(&w.o[(i - 1) / 2], w.c[i], a.l[0], a.l[2], a.l[4], a.l[6], a.l[8], a.l[10], b.l[i]);
            cmad64(&w.e[(i + 1) / 2], w.c[i + 1], a.l[1], a.l[3], a.l[5], a.l[7], a.l[9], a.l[11], b.l[i]);
        }
    }
}
// T = e + (o << 32) + (c << 384), resolved
BLS_HD void wmerge(fpw& T, const wacc& w) {
    uint32_t el[24], ol[24];
#pragma unroll
    for (int m = 0; m < 12; m++) { el[2 * m] = (uint32_t)w.e[m]; el[2 * m + 1] = (uint32_t)(w.e[m] >> 32); ol[2 * m] = (uint32_t)w.o[m]; ol[2 * m + 1] = (uint32_t)(w.o[m] >> 32); }
    uint32_t bl[12], bh[12];
    bl[0] = 0;
#pragma unroll
    for (int k = 1; k < 12; k++) bl[k] = ol[k - 1];
#pragma unroll
    for (int k = 0; k < 11; k++) bh[k] = ol[11 + k];
    bh[11] = 0;
    uint32_t cy = add12c(T.l, el, bl, 0);
    add12c(T.l + 12, el + 12, bh, cy);
    uint32_t hi[12];
#pragma unroll
    for (int k = 0; k < 12; k++) hi[k] = T.l[12 + k];
    add12c(T.l + 12, hi, w.c, 0);
}
// Montgomery reduction of the accumulated value X (destroys w): returns X / 2^384 mod p in [0, p).
// Requires X < 9.8 p^2 (then (X + m p) / R < 2p and one conditional subtraction is enough).
// Row i: the slot that starts at limb i first absorbs the other array's share of limb i (the high word of the slot that
// starts at limb i-1) and the carry that row i-2 pushed out of its slot; m_i = limb_i * (-p^-1); two MAC chains add m_i p.
BLS_HD fp wredc(wacc& w) {
    uint32_t g0 = 0, g1 = 0;                                 // carries into limb i (from row i-2) for even / odd i
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint64_t* x = (i & 1) ? &w.o[(i - 1) / 2] : &w.e[i / 2];
        uint64_t other = i == 0 ? 0 : ((i & 1) ? w.e[(i - 1) / 2] : w.o[i / 2 - 1]);
        uint32_t& gin = (i & 1) ? g1 : g0;
        uint32_t g;
        add64c(x[0], g, (uint64_t)(uint32_t)(other >> 32) + gin);
        gin = g;                                             // carries into limb i+2
        uint32_t m = (uint32_t)x[0] * BLS_M0;
        cmad64(x, w.c[i], BLS_P0, BLS_P2, BLS_P4, BLS_P6, BLS_P8, BLS_P10, m);                        // limb i becomes 0
        cmad64((i & 1) ? &w.e[(i + 1) / 2] : &w.o[i / 2], w.c[i + 1], BLS_P1, BLS_P3, BLS_P5, BLS_P7, BLS_P9, BLS_P11, m);
    }
    w.c[0] += g0; w.c[1] += g1;                              // limbs 12, 13
    uint32_t eh[12], bh[12];
#pragma unroll
    for (int m = 0; m < 6; m++) { eh[2 * m] = (uint32_t)w.e[6 + m]; eh[2 * m + 1] = (uint32_t)(w.e[6 + m] >> 32); }
    bh[0] = (uint32_t)(w.o[5] >> 32);
#pragma unroll
    for (int m = 0; m < 5; m++) { bh[1 + 2 * m] = (uint32_t)w.o[6 + m]; bh[2 + 2 * m] = (uint32_t)(w.o[6 + m] >> 32); }
    bh[11] = 0;
    fp t, u;
    add12c(t.l, eh, bh, 0);
    add12c(u.l, t.l, w.c, 0);
    return fp_reduce_once(u);
}
BLS_HD fp wredc_merged(const fpw& T) {
    wacc w; wacc_set(w, T.l);
    return wredc(w);
}

BLS_HD void wacc_set_3p2(wacc& w) {
    const uint32_t Q[24] = BLS_C_3P_SQUARED;
    wacc_set(w, Q);
}

// single product through the split accumulator (parity hook: equals fp_mul)
BLS_HD fp fp_mul_lz(const fp& a, const fp& b) { wacc w; wacc_zero(w); wmac(w, a, b); return wredc(w); }

// r = sum_{k < N} x_k * y_k over Fp2, N in 1..3, operands canonical (< p).  Karatsuba inside every product, three
// unreduced accumulations:
//   B = sum x1 y1;   RE = (3p^2 - B) + sum x0 y0 in (0, 6p^2);   IM = (3p^2 - RE - 2B) + sum (x0+x1)(y0+y1) in [0, 6p^2)
// Each accumulator STARTS at the combination of the previous ones (mod 2^768), so RE and IM come out of the MAC chains
// already combined; IM stays in split form and goes straight to the reduction.  r may alias any operand.
// Code size matters as much as instruction count here: the instruction cache behind the 6 KB per-sub-partition L0 holds
// 32 KB, and the pairing loops are far larger, so `no_instruction` stalls were 15-19 % of all samples.  The dot product is
// therefore ONE copy of the MAC stream (2.7 KB) inside a phase loop and one copy of the reduction inside a two-pass loop.
template <int N> BLS_NOINLINE void fp2_dot_t(fp2& r, const fp2* x0, const fp2* y0, const fp2* x1, const fp2* y1, const fp2* x2, const fp2* y2) {
    const fp2* xs[3] = {x0, x1, x2}; const fp2* ys[3] = {y0, y1, y2};
    const uint32_t Q3[24] = BLS_C_3P_SQUARED;
    fpw B, RE;
#pragma unroll
    for (int i = 0; i < 24; i++) { B.l[i] = 0; RE.l[i] = 0; }
    wacc w; wacc_zero(w);
    // one flat loop over (phase, term): phase 0: B = sum x1 y1;  phase 1: RE = (3p^2 - B) + sum x0 y0;
    // phase 2: IM = (3p^2 - RE - 2B) + sum (x0+x1)(y0+y1).  The operands of the NEXT step are fetched before the MAC stream
    // of the current one (local-memory latency was 16 % of this function's stall samples without it).
    int phase = 0, k = 0;
    fp pa = x0->c1, pb = y0->c1, pc = pa, pd = pb;
#pragma unroll 1
    for (int j = 0; j < 3 * N; j++) {
        fp a, b;
        if (phase == 2) { fp_add_raw(a, pa, pc); fp_add_raw(b, pb, pd); }       // < 2p < 2^382
        else { a = pa; b = pb; }
        int nk = k + 1, nphase = phase;
        if (nk == N) { nk = 0; nphase++; }
        if (nphase < 3) {
            const fp2* xp = xs[nk]; const fp2* yp = ys[nk];
            if (nphase == 0) { pa = xp->c1; pb = yp->c1; }
            else { pa = xp->c0; pb = yp->c0; if (nphase == 2) { pc = xp->c1; pd = yp->c1; } }
        }
        if (k == 0 && phase) {
            fpw t;
#pragma unroll
            for (int i = 0; i < 24; i++) t.l[i] = Q3[i];
            fpw_sub(t, t, B);
            if (phase == 2) { fpw_sub(t, t, RE); fpw_sub(t, t, B); }
            wacc_set(w, t.l);
        }
        wmac(w, a, b);
        if (k == N - 1) {
            if (phase == 0) wmerge(B, w);
            else if (phase == 1) wmerge(RE, w);
        }
        k = nk; phase = nphase;
    }
    fp2 t;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        if (pass) wacc_set(w, RE.l);
        fp v = wredc(w);
        if (pass) t.c0 = v; else t.c1 = v;
    }
    r = t;
}
BLS_HD void fp2_dot1(fp2& r, const fp2& x0, const fp2& y0) { fp2_dot_t<1>(r, &x0, &y0, &x0, &y0, &x0, &y0); }
BLS_HD void fp2_dot2(fp2& r, const fp2& x0, const fp2& y0, const fp2& x1, const fp2& y1) { fp2_dot_t<2>(r, &x0, &y0, &x1, &y1, &x1, &y1); }
BLS_HD void fp2_dot3(fp2& r, const fp2& x0, const fp2& y0, const fp2& x1, const fp2& y1, const fp2& x2, const fp2& y2) { fp2_dot_t<3>(r, &x0, &y0, &x1, &y1, &x2, &y2); }

}  // namespace bls
