// Warp-cooperative Fp12 arithmetic: one Fp12 element is spread over SIX lanes, lane k holding the Fp2 coefficient of w^k
// (Fp12 = Fp2[w]/(w^6 - xi)), five elements per warp (lanes 30, 31 idle).  Everything an item needs stays in registers
// (24 limbs per value per lane); operands are exchanged through small shared-memory windows per group and read from there
// by the lazy dot products of wide.cuh.
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
// operand windows of one warp: A = left operand coefficients, B = right operand coefficients, XB = xi * B (published by the owner lane,
// so a product term that wraps around w^6 = xi is an ordinary term of the lazy dot product)
struct coop_smem { fp2 A[COOP_GROUPS][6]; fp2 B[COOP_GROUPS][6]; fp2 XB[COOP_GROUPS][6]; };          // 10,368 B per warp

__device__ __constant__ int COOP_TOWER_POS[6] = {0, 3, 1, 4, 2, 5};           // Fp2 slot of w^k in the tower-ordered limb-SoA record

struct coop_lane { int k; fp2* A; fp2* B; fp2* XB; };                          // this lane's coefficient index and its group's windows

__device__ __forceinline__ coop_lane coop_init(coop_smem* sm_warp) {
    int lane = threadIdx.x & 31, g = lane / 6;
    coop_lane c; c.k = lane - 6 * g; c.A = sm_warp->A[g]; c.B = sm_warp->B[g]; c.XB = sm_warp->XB[g]; return c;
}

// c = a * b.  Every lane passes its own coefficients a_k, b_k and receives
//   c_k = sum_{s = 0..5} a_{(k - s) mod 6} * (s > k ? xi b_s : b_s)
// as two 3-term LAZY dot products (wide.cuh: unreduced 768-bit accumulation, Karatsuba inside each Fp2 product, two Montgomery
// reductions per dot) whose operands are read straight from the shared-memory windows: 2 x 1,608 IMAD.WIDE per lane, nothing of
// the item's Fp12 state ever touches local memory.
__device__ __noinline__ fp2 coop_mul(const coop_lane& c, fp2 a, fp2 b) {
    __syncwarp();
    c.A[c.k] = a; c.B[c.k] = b; c.XB[c.k] = fp2_mul_xi(b);
    __syncwarp();
    const fp2* x[6]; const fp2* y[6];
#pragma unroll
    for (int s = 0; s < 6; s++) {
        int i = c.k - s; bool wrap = i < 0; if (wrap) i += 6;
        x[s] = &c.A[i]; y[s] = wrap ? &c.XB[s] : &c.B[s];
    }
    fp2 r0, r1;
    fp2_dot_t<3>(r0, x[0], y[0], x[1], y[1], x[2], y[2]);
    fp2_dot_t<3>(r1, x[3], y[3], x[4], y[4], x[5], y[5]);
    return fp2_add(r0, r1);
}

// Granger-Scott squaring in the cyclotomic subgroup, one (lazy) Fp2 product per lane.
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
    fp2 prod; fp2_dot1(prod, x, y);
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
