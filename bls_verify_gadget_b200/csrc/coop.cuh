// Warp-cooperative Fp12 arithmetic: one Fp12 element is spread over SIX lanes, lane k holding the Fp2 coefficient of w^k
// (Fp12 = Fp2[w]/(w^6 - xi)), five elements per warp (lanes 30, 31 idle).  Everything an item needs stays in registers
// (24 limbs per value per lane); operands are exchanged through a small shared-memory window per group.
//
// Why: in the thread-per-item kernels an item's Fp12 state (576 B per value, 3.6-6 KB per thread with temporaries) lives in
// local memory, 0.9 MB per SM against 228 KB of L1, and the pairing kernels reach only ~60 % of the IMAD.WIDE pipe
// (profiles/r01_tuning.md, "throughput ladder").  Here a product is schoolbook over the six coefficients with LAZY reduction:
// lane k accumulates its six Fp2 products (3 unreduced 768-bit products each) and reduces twice -- 2,904 IMAD.WIDE per lane and
// multiplication, the same volume as Karatsuba in one thread, but perfectly parallel, in registers, and with the latency of an
// item divided by six.
//
// w-basis <-> tower (c0.c0 + c0.c1 v + c0.c2 v^2) + (c1.c0 + c1.c1 v + c1.c2 v^2) w, v = w^2:
//   k = 0: c0.c0   1: c1.c0   2: c0.c1   3: c1.c1   4: c0.c2   5: c1.c2
// Replaces the hard part of ark-ec's Bls12::final_exponentiation (reference src/bls.rs:455-457); device-only (uses
// __syncwarp and shared memory), checked against the thread-per-item path and the oracle by the `-m gpu` tests.
#pragma once
#include "pairing.cuh"

namespace bls {

#define COOP_GROUPS 6          // groups addressed per warp (5 active + 1 dummy for lanes 30/31 so that every access is in bounds)
struct coop_smem { fp2 A[COOP_GROUPS][6]; fp2 B[COOP_GROUPS][6]; };          // 6.75 KB per warp

__device__ __constant__ int COOP_TOWER_POS[6] = {0, 3, 1, 4, 2, 5};           // Fp2 slot of w^k in the tower-ordered limb-SoA record

struct coop_lane { int k; fp2* A; fp2* B; };                                   // this lane's coefficient index and its group's windows

__device__ __forceinline__ coop_lane coop_init(coop_smem* sm_warp) {
    int lane = threadIdx.x & 31, g = lane / 6;
    coop_lane c; c.k = lane - 6 * g; c.A = sm_warp->A[g]; c.B = sm_warp->B[g]; return c;
}
__device__ __constant__ uint32_t COOP_OFF_RE[6][24] = BLS_C_COOP_OFF_RE;      // (16 - 2k) p^2
__device__ __constant__ uint32_t COOP_OFF_IM[6][24] = BLS_C_COOP_OFF_IM;      // (5 - k) p^2
// value in [0, 4p) -> [0, p)
__device__ __forceinline__ fp fp_reduce_4p(const fp& a) {
    fp p2; fp_add_raw(p2, fp_modulus(), fp_modulus());
    fp t; uint32_t br = fp_sub_raw(t, a, p2);
    return fp_reduce_once(fp_select(br, a, t));
}
// Montgomery reduction of X < 24 p^2 to [0, p): (X + m p)/R < 24 p (p/R) + p < 3.5 p
__device__ __forceinline__ fp fp_redc_wide_4p(const fpw& Xin) {
    fpw X = Xin; uint32_t C[14];
#pragma unroll
    for (int i = 0; i < 14; i++) C[i] = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t m = X.l[i] * BLS_M0;
        cmad_n(&X.l[i], C[i], BLS_P0, BLS_P2, BLS_P4, BLS_P6, BLS_P8, BLS_P10, m);
        if (i < 11) cmad_n(&X.l[i + 1], C[i + 1], BLS_P1, BLS_P3, BLS_P5, BLS_P7, BLS_P9, BLS_P11, m);
        else { uint32_t drop = 0; cmad_n(&X.l[12], drop, BLS_P1, BLS_P3, BLS_P5, BLS_P7, BLS_P9, BLS_P11, m); }
    }
    fp hi, cc, t;
#pragma unroll
    for (int k = 0; k < 12; k++) { hi.l[k] = X.l[12 + k]; cc.l[k] = C[k]; }
    fp_add_raw(t, hi, cc);
    return fp_reduce_4p(t);
}

// c = a * b.  Every lane passes its own coefficients a_k, b_k and receives c_k.
//   c_k = sum_{i+j=k} a_i b_j + xi * sum_{i+j=k+6} a_i b_j ;  with j = s, i = (k - s) mod 6 the term wraps iff s > k.
//   Per term (Karatsuba on unreduced products T0 = a0 b0, T1 = a1 b1, T2 = (a0+a1)(b0+b1)):
//     plain:  re += T0 - T1        im += T2 - T0 - T1
//     * xi :  re += 2 T0 - T2      im += T2 - 2 T1            ((x + y u)(1 + u) = (x - y) + (x + y) u)
//   Offsets of (16 - 2k) p^2 on re and (5 - k) p^2 on im keep both accumulators in [0, 22 p^2].
__device__ __noinline__ fp2 coop_mul(const coop_lane& c, fp2 a, fp2 b) {
    __syncwarp();
    c.A[c.k] = a; c.B[c.k] = b;
    __syncwarp();
    fpw re, im;
#pragma unroll
    for (int i = 0; i < 24; i++) { re.l[i] = COOP_OFF_RE[c.k][i]; im.l[i] = COOP_OFF_IM[c.k][i]; }     // offsets keep both sums non-negative
#pragma unroll 1
    for (int s = 0; s < 6; s++) {
        int i = c.k - s; bool wrap = i < 0; if (wrap) i += 6;
        fp2 x = c.A[i], y = c.B[s];
        fpw T0, T1, T2;
        fp_mul_wide(T0, x.c0, y.c0); fp_mul_wide(T1, x.c1, y.c1);
        fp sx, sy; fp_add_raw(sx, x.c0, x.c1); fp_add_raw(sy, y.c0, y.c1);
        fp_mul_wide(T2, sx, sy);
        fpw_sub(T2, T2, T0); fpw_sub(T2, T2, T1);             // D = a0 b1 + a1 b0  (imaginary part, >= 0)
        fpw_sub(T0, T0, T1);                                  // E = a0 b0 - a1 b1  (real part, mod 2^768)
        fpw_add(re, re, T0); fpw_add(im, im, T2);             // plain term: (E, D)
        if (wrap) { fpw_sub(re, re, T2); fpw_add(im, im, T0); }   // * xi: (E - D, E + D)
    }
    fp2 r; r.c0 = fp_redc_wide_4p(re); r.c1 = fp_redc_wide_4p(im);
    return r;
}

// Granger-Scott squaring in the cyclotomic subgroup, one Fp2 product per lane.
// Pairs (w^q, w^(q+3)), q = 0,1,2: low lane q computes T = a b, high lane q+3 computes S = (a + b)(a + xi b); then
//   t_even[q] = S - T - xi T,  t_odd[q] = 2 T   and (arkworks' cyclotomic_square_in_place, csrc/tower.cuh fp12_cyclo_sqr)
//   w^0 <- 3 t_even[0] - 2 z   w^3 <- 3 t_odd[0] + 2 z   w^1 <- 3 xi t_odd[2] + 2 z   w^4 <- 3 t_even[2] - 2 z
//   w^2 <- 3 t_even[1] - 2 z   w^5 <- 3 t_odd[1] + 2 z
__device__ __noinline__ fp2 coop_cyclo_sqr(const coop_lane& c, fp2 z) {
    __syncwarp();
    c.A[c.k] = z;
    __syncwarp();
    int q = c.k % 3; bool high = c.k >= 3;
    fp2 a = c.A[q], b = c.A[q + 3];
    fp2 x = high ? fp2_add(a, b) : a;
    fp2 y = high ? fp2_add(a, fp2_mul_xi(b)) : b;
    fp2 prod = fp2_mul(x, y);
    c.B[c.k] = prod;                                          // B[q] = T_q, B[q+3] = S_q
    __syncwarp();
    // which pair feeds this lane: w^0,w^3 <- pair 0 ; w^1,w^4 <- pair 2 ; w^2,w^5 <- pair 1
    int src = (c.k % 3 == 0) ? 0 : (c.k % 3 == 1 ? 2 : 1);
    bool even_type = (c.k == 0) | (c.k == 4) | (c.k == 2);    // lanes that take t_even (and subtract 2 z)
    fp2 T = c.B[src], S = c.B[src + 3];
    fp2 te = fp2_sub(fp2_sub(S, T), fp2_mul_xi(T));
    fp2 to = fp2_dbl(T); if (c.k == 1) to = fp2_mul_xi(to);
    fp2 t = fp2_csel(even_type, te, to);
    fp2 zz = fp2_csel(even_type, fp2_neg(z), z);
    fp2 d = fp2_add(t, zz);                                   // t -+ z
    return fp2_add(fp2_dbl(d), t);                            // 3 t -+ 2 z
}
__device__ __forceinline__ fp2 coop_conj(const coop_lane& c, const fp2& a) { return (c.k & 1) ? fp2_neg(a) : a; }     // a^(p^6): w -> -w
__device__ __forceinline__ fp2 coop_frob(const coop_lane& c, const fp2& a) { fp2 t = fp2_conj(a); return c.k ? fp2_mul(t, FROB1[c.k]) : t; }
__device__ __forceinline__ fp2 coop_frob2(const coop_lane& c, const fp2& a) { return c.k ? fp2_mul_fp(a, FROB2[c.k]) : a; }

// a^x for a in the cyclotomic subgroup (x negative: conjugate at the end)
__device__ __noinline__ fp2 coop_exp_by_x(const coop_lane& c, fp2 a) {
    fp2 acc = a;
    const uint64_t x = BLS_X_ABS;
    for (int i = 62; i >= 0; i--) {
        acc = coop_cyclo_sqr(c, acc);
        if ((x >> i) & 1) acc = coop_mul(c, acc, a);
    }
    return coop_conj(c, acc);
}
// hard part of the final exponentiation, the exact chain of final_exponentiation() in pairing.cuh; r = f^((p^6-1)(p^2+1))
__device__ __forceinline__ fp2 coop_final_exp_hard(const coop_lane& c, fp2 r) {
    fp2 y0 = coop_cyclo_sqr(c, r);
    fp2 y1 = coop_exp_by_x(c, r);
    fp2 y2 = coop_conj(c, r);
    y1 = coop_mul(c, y1, y2);
    y2 = coop_exp_by_x(c, y1);
    y1 = coop_conj(c, y1);
    y1 = coop_mul(c, y1, y2);
    y2 = coop_exp_by_x(c, y1);
    y1 = coop_frob(c, y1);
    y1 = coop_mul(c, y1, y2);
    r = coop_mul(c, r, y0);
    y0 = coop_exp_by_x(c, y1);
    y2 = coop_exp_by_x(c, y0);
    y0 = coop_frob2(c, y1);
    y1 = coop_conj(c, y1);
    y1 = coop_mul(c, y1, y2);
    y1 = coop_mul(c, y1, y0);
    return coop_mul(c, r, y1);
}

}  // namespace bls
