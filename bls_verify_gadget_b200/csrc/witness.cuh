// GPU witness generation for the verify circuit: replays the witness program recorded by the host-side
// builder (bls_verify_gadget_b200/gadget: one rule per variable, in allocation order) for 32 assignments per warp and
// writes them straight into the layout the satisfaction kernels gather from -- so (pk, msg, sig) bytes go in and the
// per-constraint bits come out without the 34 MB-per-assignment host synthesis and PCIe transfer.
//
// The reference has no counterpart (its witnesses come from running the gadget code under ark-relations'
// ConstraintSystem in "prove" mode, src/constraints.rs:335-370); the values are checked bit for bit against the host
// builder's assignment, which is pinned by the reference's own expectations (tests/test_gadget_circuit.py).
//
// Mapping: lane <-> assignment, warp <-> group of 32, rules strictly in order (a rule reads variables written by earlier
// rules of the same lane: program order makes them visible).  Every branch is warp-uniform (rule kind, coefficient class).
// Included after r1cs.cuh in blsgpu.cu.
#pragma once
#include <cooperative_groups.h>

#ifndef WIT_PF
#define WIT_PF 1          // gathers issued together per batch of terms in the replay (r1cs_meta_eval<PF>): 4 / 8 / 16 measured 1.3x / 1.8x / 4x SLOWER (spills), profiles/r02_tuning.md
#endif
enum { WR_MULADD = 0, WR_INV = 1, WR_NEQ = 2, WR_NEQMULT = 3, WR_BIT = 4, WR_FP2INV = 5, WR_FP12INV = 6, WR_INPUT = 7 };
struct wit_rule { uint8_t kind, pad; uint16_t aux; uint32_t a, b, d; };
// rule in level order with its term ranges resolved (one dependent load less per combination); for WR_FP12INV a_lo is the
// id of the first of twelve consecutive combinations
struct wit_xrule { uint8_t kind, pad; uint16_t aux; uint32_t var; uint32_t a_lo, a_hi, b_lo, b_hi, d_lo, d_hi; };
struct wit_prog {
    size_t nvars, nout, nlc, nterms, nlevels;          // nvars: columns of the program (circuit variables, then scratch values); nout: circuit variables
    wit_rule* rules; uint64_t* lc_ptr; uint32_t* col; fp* coeff; fp* coeffc; uint8_t* cls;
    wit_xrule* xrules; uint64_t* level_ptr;          // dependency levels: rules of one level are independent
    size_t msg_len, ninputs, nkeys;                  // input slots of the recorded circuit: nkeys = 0: verify (6 + 8 msg_len); nkeys = n: aggregate_verify (3n + 4 + 8 msg_len)
};
#define WIT_POINT_INPUTS 6      // input slots 0..5: pk.x, pk.y, sig x.c0, x.c1, y.c0, y.c1; slot 6 + 8 i + b = bit b of message byte i (any message length)

__device__ __forceinline__ void wit_store(u32x4* zt, size_t col, int lane, const fp& v) {
    u32x4 a, b, c;
    a.x = v.l[0]; a.y = v.l[1]; a.z = v.l[2]; a.w = v.l[3]; b.x = v.l[4]; b.y = v.l[5]; b.z = v.l[6]; b.w = v.l[7]; c.x = v.l[8]; c.y = v.l[9]; c.z = v.l[10]; c.w = v.l[11];
    zt[(col * 3) * 32 + lane] = a; zt[(col * 3 + 1) * 32 + lane] = b; zt[(col * 3 + 2) * 32 + lane] = c;
}
// value of the combination with terms [lo, hi) on this lane's assignment, canonical.  zb (nullable) = the group's packed 0/1 view: a
// column that an earlier level found to be 0/1 in all 32 assignments is read as ONE broadcast 8-byte load instead of a 1.5 KB gather and
// its +-1 / small coefficients go to a 64-bit integer side sum -- the same term rule as the satisfaction kernels (r1cs_term), which is
// what 93 % of the verify circuit's columns and most of the witness program's 4.0 M terms are.
__device__ __forceinline__ fp wit_lc_range(const wit_prog& p, uint64_t lo, uint64_t hi, const u32x4* zt, const uint2* zb, int lane) {
    fp acc = fp_zero();
    if (zb) {
        r1cs_sys view; view.coeff[0] = p.coeff; view.coeffc[0] = p.coeffc; view.col[0] = p.col; view.cls[0] = p.cls;
        int64_t side = 0; bool touched = false;
        r1cs_accum ac; r1cs_accum_zero(ac);                    // metadata of up to 32 terms fetched in one round: the replay is latency-bound
        wacc wg; int wn = 0; wacc_zero(wg);                    // general coefficients on field values: lazy 768-bit sum (r1cs_term)
        for (uint64_t base = lo; base < hi; base += 32) { r1cs_meta t; r1cs_meta_fetch(t, view, 0, base, hi, lane); r1cs_meta_views(t, zb, lane); r1cs_meta_eval<WIT_PF>(t, view, 0, zt, lane, ac, &wg, &wn); }
        if (wn) ac.acc = fp_add(ac.acc, wredc(wg));
        acc = r1cs_accum_close(ac, side, touched);
        return r1cs_finalize(acc, side, touched);
    }
    for (uint64_t k = lo, e = hi; k < e; k++) {
        uint32_t cj = p.col[k]; uint8_t c = p.cls[k];
        fp zv = r1cs_load_z(zt, cj, lane);
        if (c == R1_PLUS_ONE) acc = fp_add(acc, zv);
        else if (c == R1_MINUS_ONE) acc = fp_sub(acc, zv);
        else {
            bool small = zv.l[0] < 2 && !(zv.l[1] | zv.l[2] | zv.l[3] | zv.l[4] | zv.l[5] | zv.l[6] | zv.l[7] | zv.l[8] | zv.l[9] | zv.l[10] | zv.l[11]);
            if (__all_sync(0xffffffffu, small)) acc = fp_add(acc, fp_select(0u - zv.l[0], p.coeffc[k], fp_zero()));       // coefficient times a 0/1 value
            else if (c == R1_SMALL_POS) acc = fp_add(acc, fp_mul_small(zv, p.coeffc[k].l[0]));
            else if (c == R1_SMALL_NEG) acc = fp_sub(acc, fp_mul_small(zv, BLS_P0 - p.coeffc[k].l[0]));
            else acc = fp_add(acc, fp_mul(p.coeff[k], zv));
        }
    }
    return acc;
}
// The three combinations of a product rule (a b + d) with their metadata rounds OVERLAPPED: the replay is bound by the chain of dependent
// loads inside one task (rule -> columns -> packed views -> values), so the column / class / coefficient loads of all three ranges are
// issued together, then the packed views of all three, and only then are the terms evaluated.  Ranges longer than 32 terms finish
// through the chunked loop.
__device__ __forceinline__ void wit_lc3(const wit_prog& p, const wit_xrule& r, const u32x4* zt, const uint2* zb, int lane, fp& a, fp& b, fp& d) {
    r1cs_sys view; view.coeff[0] = p.coeff; view.coeffc[0] = p.coeffc; view.col[0] = p.col; view.cls[0] = p.cls;
    r1cs_meta ta, tb, td;
    r1cs_meta_fetch(ta, view, 0, r.a_lo, r.a_hi, lane); r1cs_meta_fetch(tb, view, 0, r.b_lo, r.b_hi, lane); r1cs_meta_fetch(td, view, 0, r.d_lo, r.d_hi, lane);
    r1cs_meta_views(ta, zb, lane); r1cs_meta_views(tb, zb, lane); r1cs_meta_views(td, zb, lane);
    const r1cs_meta* ts[3] = {&ta, &tb, &td}; const uint32_t his[3] = {r.a_hi, r.b_hi, r.d_hi}; fp* outs[3] = {&a, &b, &d};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r1cs_accum acc; r1cs_accum_zero(acc);
        wacc wg; int wn = 0; wacc_zero(wg);
        r1cs_meta_eval<WIT_PF>(*ts[k], view, 0, zt, lane, acc, &wg, &wn);
        for (uint64_t base = ts[k]->base + 32; base < his[k]; base += 32) { r1cs_meta t; r1cs_meta_fetch(t, view, 0, base, his[k], lane); r1cs_meta_views(t, zb, lane); r1cs_meta_eval<WIT_PF>(t, view, 0, zt, lane, acc, &wg, &wn); }
        if (wn) acc.acc = fp_add(acc.acc, wredc(wg));
        int64_t side; bool touched; fp v = r1cs_accum_close(acc, side, touched);
        *outs[k] = r1cs_finalize(v, side, touched);
    }
}
__device__ __forceinline__ fp wit_lc(const wit_prog& p, uint32_t id, const u32x4* zt, const uint2* zb, int lane) {
    if (!id) return fp_zero();
    return wit_lc_range(p, p.lc_ptr[id - 1], p.lc_ptr[id], zt, zb, lane);
}
__device__ __forceinline__ fp wit_inv_canon(const fp& a) { return fp_inv_raw(a); }                 // canonical in and out: the divsteps inverse needs no Montgomery factor
// a b for canonical a, b: small-integer shortcut as in r1cs_product_ok
__device__ __forceinline__ fp wit_mul_canon(const fp& a, const fp& b) {
    uint32_t ah = a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7] | a.l[8] | a.l[9] | a.l[10] | a.l[11];
    uint32_t bh = b.l[2] | b.l[3] | b.l[4] | b.l[5] | b.l[6] | b.l[7] | b.l[8] | b.l[9] | b.l[10] | b.l[11];
    bool small = (ah | bh) == 0 && (a.l[1] == 0 || b.l[1] == 0);
    if (__all_sync(0xffffffffu, small)) {
        uint64_t x = ((uint64_t)a.l[1] << 32) | a.l[0], y = ((uint64_t)b.l[1] << 32) | b.l[0];
        uint64_t lo = x * y, hi = __umul64hi(x, y);
        fp r = fp_zero(); r.l[0] = (uint32_t)lo; r.l[1] = (uint32_t)(lo >> 32); r.l[2] = (uint32_t)hi; r.l[3] = (uint32_t)(hi >> 32);
        return r;                                                     // < 2^96 < p
    }
    return fp_mul(fp_to_mont(a), b);
}
// inputs: fp [6 + 8 msg_len][nwit_padded] canonical (slot-major, so a warp reads 32 consecutive elements)
__global__ void __launch_bounds__(32) k_witness_gen(wit_prog p, const fp* inputs, size_t nwit_padded, u32x4* zt_all) {
    size_t group = blockIdx.x; int lane = threadIdx.x;
    u32x4* zt = zt_all + group * p.nvars * 96;                        // 3 chunks x 32 lanes per variable
    size_t w = group * 32 + lane;
    fp one = fp_zero(); one.l[0] = 1;
    wit_store(zt, 0, lane, one);
    for (size_t v = 1; v < p.nvars; v++) {
        wit_rule r = p.rules[v];
        fp out;
        switch (r.kind) {
            case WR_MULADD: {
                fp d = wit_lc(p, r.d, zt, nullptr, lane);
                if (r.a) { fp a = wit_lc(p, r.a, zt, nullptr, lane), b = wit_lc(p, r.b, zt, nullptr, lane); out = fp_add(wit_mul_canon(a, b), d); }
                else out = d;
                break;
            }
            case WR_INV: out = wit_inv_canon(wit_lc(p, r.a, zt, nullptr, lane)); break;
            case WR_NEQ: { fp a = wit_lc(p, r.a, zt, nullptr, lane); out = fp_zero(); out.l[0] = fp_is_zero(a) ? 0u : 1u; break; }
            case WR_NEQMULT: { fp a = wit_lc(p, r.a, zt, nullptr, lane); out = fp_is_zero(a) ? one : wit_inv_canon(a); break; }
            case WR_BIT: { fp a = wit_lc(p, r.a, zt, nullptr, lane); out = fp_zero(); out.l[0] = (a.l[r.aux >> 5] >> (r.aux & 31)) & 1u; break; }
            case WR_FP2INV: {
                fp2 x; x.c0 = fp_to_mont(wit_lc(p, r.a, zt, nullptr, lane)); x.c1 = fp_to_mont(wit_lc(p, r.b, zt, nullptr, lane));
                fp2 iv = fp2_inv(x); out = fp_from_mont(r.aux ? iv.c1 : iv.c0); break;
            }
            case WR_FP12INV: {
                fp12 x, iv; fp* xf = &x.c0.c0.c0;
                for (int k = 0; k < 12; k++) xf[k] = fp_to_mont(wit_lc(p, r.a + k, zt, nullptr, lane));
                fp12_inv(iv, x); out = fp_from_mont((&iv.c0.c0.c0)[r.aux]); break;
            }
            default: out = inputs[(size_t)r.aux * nwit_padded + w]; break;                 // WR_INPUT
        }
        wit_store(zt, v, lane, out);
    }
}
// Level-synchronous form: one warp per (rule of the current level, group of 32 assignments); a grid-wide barrier between
// levels (cooperative launch).  The verify circuit has 8,816 levels of ~80 rules: with 16 groups a level is ~1,300 independent
// warp tasks, against ONE warp per group walking 707,809 rules in sequence in k_witness_gen (7.3 s, latency bound).
#ifndef WIT_TPB
#define WIT_TPB 128
#endif
#ifndef WIT_BPS
#define WIT_BPS 2
#endif
#ifdef WIT_TRACE
// tuning builds only (profiles/tools/wit_trace.py): the global timer at the start of every level, written by warp 0
__device__ unsigned long long g_wit_trace[32768];
extern "C" int blsgpu_debug_wit_trace(unsigned long long* out, size_t n) { return cudaMemcpyFromSymbol(out, g_wit_trace, 8 * (n < 32768 ? n : 32768)) == cudaSuccess ? 0 : -1; }
#endif
__global__ void __launch_bounds__(WIT_TPB, WIT_BPS) k_witness_levels(wit_prog p, const fp* inputs, size_t nwit_padded, u32x4* zt_all, size_t groups, uint2* zbool_all) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (size_t)blockDim.x) >> 5; int lane = threadIdx.x & 31;
    fp one = fp_zero(); one.l[0] = 1;
    uint64_t lo = p.level_ptr[0], hi = p.level_ptr[1];
    wit_xrule rn; bool pre = warp < (hi - lo) * groups;              // this warp's first rule of the next level is fetched before the barrier (static data)
    if (pre) rn = p.xrules[lo + warp / groups];
    for (size_t level = 0; level < p.nlevels; level++) {
        uint64_t n = hi - lo, hi_next = level + 2 <= p.nlevels ? p.level_ptr[level + 2] : hi;
#ifdef WIT_TRACE
        if (warp == 0 && lane == 0 && level < 32768) { unsigned long long tm; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm)); g_wit_trace[level] = tm; }
#endif
        for (size_t t = warp; t < n * groups; t += nwarps) {
            wit_xrule r = (t == warp && pre) ? rn : p.xrules[lo + t / groups]; size_t group = t % groups;
            u32x4* zt = zt_all + group * p.nvars * 96; const uint2* zb = zbool_all ? zbool_all + group * p.nvars : nullptr;
            fp out;
            switch (r.kind) {
                case WR_MULADD: {
                    if (zb) { fp a, b, d; wit_lc3(p, r, zt, zb, lane, a, b, d); out = r.a_hi > r.a_lo ? fp_add(wit_mul_canon(a, b), d) : d; break; }
                    fp d = wit_lc_range(p, r.d_lo, r.d_hi, zt, zb, lane);
                    if (r.a_hi > r.a_lo) { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane), b = wit_lc_range(p, r.b_lo, r.b_hi, zt, zb, lane); out = fp_add(wit_mul_canon(a, b), d); }
                    else out = d;
                    break;
                }
                case WR_INV: out = wit_inv_canon(wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane)); break;
                case WR_NEQ: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_zero(); out.l[0] = fp_is_zero(a) ? 0u : 1u; break; }
                case WR_NEQMULT: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_is_zero(a) ? one : wit_inv_canon(a); break; }
                case WR_BIT: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_zero(); out.l[0] = (a.l[r.aux >> 5] >> (r.aux & 31)) & 1u; break; }
                case WR_FP2INV: {
                    fp2 x; x.c0 = fp_to_mont(wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane)); x.c1 = fp_to_mont(wit_lc_range(p, r.b_lo, r.b_hi, zt, zb, lane));
                    fp2 iv = fp2_inv(x); out = fp_from_mont(r.aux ? iv.c1 : iv.c0); break;
                }
                case WR_FP12INV: {
                    fp12 x, iv; fp* xf = &x.c0.c0.c0;
                    for (int k = 0; k < 12; k++) xf[k] = fp_to_mont(wit_lc(p, r.a_lo + k, zt, zb, lane));
                    fp12_inv(iv, x); out = fp_from_mont((&iv.c0.c0.c0)[r.aux]); break;
                }
                default: out = r.aux == 0xffff ? one : inputs[(size_t)r.aux * nwit_padded + group * 32 + lane]; break;       // WR_INPUT; 0xffff: the constant ONE (variable 0)
            }
            if (!zbool_all) wit_store(zt, r.var, lane, out);
            else {                                  // the packed 0/1 view the satisfaction kernels read (r1cs.cuh: k_r1cs_transpose writes the same)
                bool small = out.l[0] < 2 && !(out.l[1] | out.l[2] | out.l[3] | out.l[4] | out.l[5] | out.l[6] | out.l[7] | out.l[8] | out.l[9] | out.l[10] | out.l[11]);
                bool all = __all_sync(0xffffffffu, small); uint32_t pack = __ballot_sync(0xffffffffu, out.l[0] & 1u);
                if (!all) wit_store(zt, r.var, lane, out);           // a 0/1 column lives in its packed word only (every reader tests the flag first)
                if (lane == 0) zbool_all[group * p.nvars + r.var] = make_uint2(all ? pack : 0u, all ? 1u : 0u);
            }
        }
        pre = level + 1 < p.nlevels && warp < (hi_next - hi) * groups;
        if (pre) rn = p.xrules[hi + warp / groups];
        grid.sync();
        lo = hi; hi = hi_next;
    }
}
// Cluster form of the level-synchronous replay: one thread-block CLUSTER (WIT_CL CTAs on WIT_CL SMs, 8 x 256 threads = 64 warps) owns
// one group of 32 assignments, one warp per rule of the current level, and the barrier between levels is the hardware cluster barrier
// (~0.2 us) instead of a grid-wide barrier over every resident CTA (several us, 14.7 k times per launch).  Groups run independently
// side by side (18 clusters fill the 148 SMs; further groups queue behind them), so a launch costs (levels x level latency) per WAVE
// of 18 groups.  The next level's rule record -- static data -- is fetched BEFORE the barrier, so that after it only the loads
// that depend on the previous level's values remain on the critical path.
#ifndef WIT_CL
#define WIT_CL 8
#endif
#define WIT_CL_TPB 256
__device__ __forceinline__ void wit_eval_rule(const wit_prog& p, const wit_xrule& r, const fp* inputs, size_t nwit_padded, size_t group, u32x4* zt, uint2* zbw, int lane) {
    const uint2* zb = zbw;
    fp one = fp_zero(); one.l[0] = 1;
    fp out;
    switch (r.kind) {
        case WR_MULADD: {
            if (zb) { fp a, b, d; wit_lc3(p, r, zt, zb, lane, a, b, d); out = r.a_hi > r.a_lo ? fp_add(wit_mul_canon(a, b), d) : d; break; }
            fp d = wit_lc_range(p, r.d_lo, r.d_hi, zt, zb, lane);
            if (r.a_hi > r.a_lo) { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane), b = wit_lc_range(p, r.b_lo, r.b_hi, zt, zb, lane); out = fp_add(wit_mul_canon(a, b), d); }
            else out = d;
            break;
        }
        case WR_INV: out = wit_inv_canon(wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane)); break;
        case WR_NEQ: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_zero(); out.l[0] = fp_is_zero(a) ? 0u : 1u; break; }
        case WR_NEQMULT: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_is_zero(a) ? one : wit_inv_canon(a); break; }
        case WR_BIT: { fp a = wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane); out = fp_zero(); out.l[0] = (a.l[r.aux >> 5] >> (r.aux & 31)) & 1u; break; }
        case WR_FP2INV: {
            fp2 x; x.c0 = fp_to_mont(wit_lc_range(p, r.a_lo, r.a_hi, zt, zb, lane)); x.c1 = fp_to_mont(wit_lc_range(p, r.b_lo, r.b_hi, zt, zb, lane));
            fp2 iv = fp2_inv(x); out = fp_from_mont(r.aux ? iv.c1 : iv.c0); break;
        }
        case WR_FP12INV: {
            fp12 x, iv; fp* xf = &x.c0.c0.c0;
            for (int k = 0; k < 12; k++) xf[k] = fp_to_mont(wit_lc(p, r.a_lo + k, zt, zb, lane));
            fp12_inv(iv, x); out = fp_from_mont((&iv.c0.c0.c0)[r.aux]); break;
        }
        default: out = r.aux == 0xffff ? one : inputs[(size_t)r.aux * nwit_padded + group * 32 + lane]; break;       // WR_INPUT; 0xffff: the constant ONE (variable 0)
    }
    if (!zbw) { wit_store(zt, r.var, lane, out); return; }
    bool small = out.l[0] < 2 && !(out.l[1] | out.l[2] | out.l[3] | out.l[4] | out.l[5] | out.l[6] | out.l[7] | out.l[8] | out.l[9] | out.l[10] | out.l[11]);
    bool all = __all_sync(0xffffffffu, small); uint32_t pack = __ballot_sync(0xffffffffu, out.l[0] & 1u);
    if (!all) wit_store(zt, r.var, lane, out);                // a 0/1 column lives in its packed word only (every reader tests the flag first)
    if (lane == 0) zbw[r.var] = make_uint2(all ? pack : 0u, all ? 1u : 0u);
}
__global__ void __launch_bounds__(WIT_CL_TPB, 1) k_witness_cluster(wit_prog p, const fp* inputs, size_t nwit_padded, u32x4* zt_all, size_t groups, uint2* zbool_all) {
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
    const unsigned CW = WIT_CL * (WIT_CL_TPB / 32);                                             // warps of the cluster
    size_t group = blockIdx.x / WIT_CL; unsigned cw = cluster.block_rank() * (WIT_CL_TPB / 32) + (threadIdx.x >> 5); int lane = threadIdx.x & 31;
    u32x4* zt = zt_all + group * p.nvars * 96; uint2* zbw = zbool_all ? zbool_all + group * p.nvars : nullptr;
    uint64_t lo = p.level_ptr[0], hi = p.level_ptr[1];
    wit_xrule rn; bool pre = lo + cw < hi;
    if (pre) rn = p.xrules[lo + cw];
    for (size_t level = 0; level < p.nlevels; level++) {
        uint64_t hi_next = level + 2 <= p.nlevels ? p.level_ptr[level + 2] : hi;
        for (uint64_t t = lo + cw; t < hi; t += CW) {
            wit_xrule r = (t == lo + cw && pre) ? rn : p.xrules[t];
            wit_eval_rule(p, r, inputs, nwit_padded, group, zt, zbw, lane);
        }
        pre = level + 1 < p.nlevels && hi + cw < hi_next;
        if (pre) rn = p.xrules[hi + cw];                      // static data: travels while the cluster gathers at the barrier
        cluster.sync();
        lo = hi; hi = hi_next;
    }
}
// input slots from the decoded points and the message bytes; items whose key or signature does not decode get all-zero inputs
__global__ void __launch_bounds__(TPB, BLS_MINB) k_witness_inputs(const u32x4* pk_soa, const uint8_t* code_pk, const u32x4* sig_soa, const uint8_t* code_sig, const uint8_t* msg, size_t msg_len,
                                                                  size_t nwit, size_t nwit_padded, fp* inputs, uint8_t* status) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= nwit_padded) return;
    bool live = i < nwit; uint8_t st = ST_OK;
    if (live) { if (code_pk[i] != DEC_OK) st = ST_BAD_PK; else if (code_sig[i] != DEC_OK) st = ST_BAD_SIG; }      // the circuit inverts z of both points: the identity has no assignment
    bool ok = live && st == ST_OK;
    fp zero = fp_zero();
    for (size_t k = 0; k < 8 * msg_len; k++) { fp b = zero; if (ok) b.l[0] = (msg[msg_len * i + (k >> 3)] >> (k & 7)) & 1u; inputs[(size_t)(WIT_POINT_INPUTS + k) * nwit_padded + i] = b; }
    g1_aff pk; g2_aff sg;
    if (ok) { soa_load_g1(pk, pk_soa, nwit, i); soa_load_g2(sg, sig_soa, nwit, i); }
    inputs[(size_t)0 * nwit_padded + i] = ok ? fp_from_mont(pk.x) : zero; inputs[(size_t)1 * nwit_padded + i] = ok ? fp_from_mont(pk.y) : zero;
    inputs[(size_t)2 * nwit_padded + i] = ok ? fp_from_mont(sg.x.c0) : zero; inputs[(size_t)3 * nwit_padded + i] = ok ? fp_from_mont(sg.x.c1) : zero;
    inputs[(size_t)4 * nwit_padded + i] = ok ? fp_from_mont(sg.y.c0) : zero; inputs[(size_t)5 * nwit_padded + i] = ok ? fp_from_mont(sg.y.c1) : zero;
    if (live && status) status[i] = st;
}
// the same for the aggregate_verify circuit: n keys (decoded into key_soa, item i key k at i * n + k), one bitmap byte per key, message, signature
__global__ void __launch_bounds__(TPB, BLS_MINB) k_witness_inputs_agg(const u32x4* key_soa, const uint8_t* code_key, size_t n, const uint8_t* bitmap, const u32x4* sig_soa, const uint8_t* code_sig,
                                                                      const uint8_t* msg, size_t msg_len, size_t nwit, size_t nwit_padded, fp* inputs, uint8_t* status) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= nwit_padded) return;
    bool live = i < nwit; uint8_t st = ST_OK;
    if (live) { for (size_t k = 0; k < n; k++) if (code_key[i * n + k] != DEC_OK) { st = ST_BAD_PK; break; } if (st == ST_OK && code_sig[i] != DEC_OK) st = ST_BAD_SIG; }
    bool ok = live && st == ST_OK;
    fp zero = fp_zero();
    for (size_t k = 0; k < n; k++) {
        g1_aff pk; if (ok) soa_load_g1(pk, key_soa, nwit * n, i * n + k);
        inputs[(2 * k) * nwit_padded + i] = ok ? fp_from_mont(pk.x) : zero; inputs[(2 * k + 1) * nwit_padded + i] = ok ? fp_from_mont(pk.y) : zero;
        fp b = zero; if (ok) b.l[0] = bitmap[i * n + k] ? 1u : 0u; inputs[(2 * n + k) * nwit_padded + i] = b;
    }
    g2_aff sg; if (ok) soa_load_g2(sg, sig_soa, nwit, i);
    inputs[(3 * n + 0) * nwit_padded + i] = ok ? fp_from_mont(sg.x.c0) : zero; inputs[(3 * n + 1) * nwit_padded + i] = ok ? fp_from_mont(sg.x.c1) : zero;
    inputs[(3 * n + 2) * nwit_padded + i] = ok ? fp_from_mont(sg.y.c0) : zero; inputs[(3 * n + 3) * nwit_padded + i] = ok ? fp_from_mont(sg.y.c1) : zero;
    for (size_t k = 0; k < 8 * msg_len; k++) { fp b = zero; if (ok) b.l[0] = (msg[msg_len * i + (k >> 3)] >> (k & 7)) & 1u; inputs[(3 * n + 4 + k) * nwit_padded + i] = b; }
    if (live && status) status[i] = st;
}
// transposed group -> z[w][col] (48-byte LE canonical), the layout of blsgpu_r1cs_check
__global__ void __launch_bounds__(256) k_witness_untranspose(const u32x4* zt_all, const uint2* zbool_all, size_t ncols, size_t nout, size_t nwit, u32x4* z) {
    size_t col = blockIdx.x * (size_t)8 + (threadIdx.x >> 5); int lane = threadIdx.x & 31; size_t group = blockIdx.y;
    size_t w = group * 32 + lane;
    if (col >= nout || w >= nwit) return;
    const u32x4* zt = zt_all + group * ncols * 96;
    u32x4* dst = z + (w * nout + col) * 3;
    if (zbool_all) {                                          // a 0/1 column exists only as its packed word
        uint2 f = zbool_all[group * ncols + col];
        if (f.y) { u32x4 v; v.x = (f.x >> lane) & 1u; v.y = v.z = v.w = 0; u32x4 zero; zero.x = zero.y = zero.z = zero.w = 0; dst[0] = v; dst[1] = zero; dst[2] = zero; return; }
    }
    dst[0] = zt[(col * 3) * 32 + lane]; dst[1] = zt[(col * 3 + 1) * 32 + lane]; dst[2] = zt[(col * 3 + 2) * 32 + lane];
}

static void wit_release(wit_prog* p) {
    cudaFree(p->rules); cudaFree(p->lc_ptr); cudaFree(p->col); cudaFree(p->coeff); cudaFree(p->coeffc); cudaFree(p->cls);
    cudaFree(p->xrules); cudaFree(p->level_ptr);
    delete p;
}

extern "C" {
int blsgpu_witness_load(blsgpu_ctx* ctx, const uint8_t* rules16, const uint64_t* lc_ptr, const uint32_t* lc_col, const uint8_t* lc_coef48, size_t nvars, size_t nout, size_t nlc, size_t nterms,
                        const uint32_t* order, const uint64_t* level_ptr, size_t nlevels, int* handle) {
    ENTER(); if (!rules16 || !lc_ptr || !lc_col || !lc_coef48 || !handle || !nvars || !nout || nout > nvars) return fail(ctx, BLSGPU_ERR_ARG, "bad argument");
    if (ctx->ptr_mode != BLSGPU_HOST) return fail(ctx, BLSGPU_ERR_ARG, "blsgpu_witness_load takes host pointers (the level-ordered rule table is built on the host)");
    int h = -1; for (int i = 0; i < 4; i++) if (!ctx->wit[i]) { h = i; break; }
    if (h < 0) return fail(ctx, BLSGPU_ERR_ARG, "too many witness programs loaded");
    wit_prog* p = new (std::nothrow) wit_prog(); if (!p) return fail(ctx, BLSGPU_ERR_ALLOC, "out of host memory");
    memset(p, 0, sizeof *p); p->nvars = nvars; p->nout = nout; p->nlc = nlc; p->nterms = nterms;
    struct guard_t { wit_prog* p; ~guard_t() { if (p) wit_release(p); } } undo{p};          // a failed load leaves nothing behind
    {   // the message length of the recorded circuit follows from the highest input slot: slots 0..5 are the points, 6 + 8 i + b the message bits
        const wit_rule* hr = reinterpret_cast<const wit_rule*>(rules16); size_t top = 0;
        for (size_t k = 0; k < nvars; k++) if (hr[k].kind == WR_INPUT && hr[k].aux != 0xffff && (size_t)hr[k].aux + 1 > top) top = (size_t)hr[k].aux + 1;
        p->ninputs = top; p->nkeys = 0;                    // a verify program until blsgpu_witness_set_aggregate says otherwise
        p->msg_len = (top >= WIT_POINT_INPUTS && (top - WIT_POINT_INPUTS) % 8 == 0) ? (top - WIT_POINT_INPUTS) / 8 : (size_t)-1;
    }
    cudaMemcpyKind kind = ctx->ptr_mode == BLSGPU_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    size_t nt = nterms ? nterms : 1;
    CU(cudaMalloc(&p->rules, 16 * nvars)); CU(cudaMalloc(&p->lc_ptr, 8 * (nlc + 1))); CU(cudaMalloc(&p->col, 4 * nt));
    CU(cudaMalloc(&p->coeff, 48 * nt)); CU(cudaMalloc(&p->coeffc, 48 * nt)); CU(cudaMalloc(&p->cls, nt));
    CU(cudaMemcpyAsync(p->rules, rules16, 16 * nvars, kind, ctx->stream)); CU(cudaMemcpyAsync(p->lc_ptr, lc_ptr, 8 * (nlc + 1), kind, ctx->stream));
    if (nterms) {
        CU(cudaMemcpyAsync(p->col, lc_col, 4 * nterms, kind, ctx->stream));
        dev_tmp raw; CU(cudaMalloc(&raw.p, 48 * nterms));
        CU(cudaMemcpyAsync(raw.p, lc_coef48, 48 * nterms, kind, ctx->stream));
        LAUNCH(k_r1cs_prepare, nblk(nterms), TPB, (const uint8_t*)raw.p, nterms, p->coeff, p->coeffc, p->cls);
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (order && level_ptr && nlevels) {                   // level-ordered rules with resolved term ranges
        const wit_rule* hr = reinterpret_cast<const wit_rule*>(rules16);
        std::vector<wit_xrule> xr(nvars);
        auto lo_of = [&](uint32_t id) { return id ? (uint32_t)lc_ptr[id - 1] : 0u; }; auto hi_of = [&](uint32_t id) { return id ? (uint32_t)lc_ptr[id] : 0u; };
        for (size_t k = 0; k < nvars; k++) {
            const wit_rule& r = hr[order[k]]; wit_xrule& x = xr[k];
            x.kind = r.kind; x.pad = 0; x.aux = r.aux; x.var = order[k];
            if (r.kind == WR_FP12INV) { x.a_lo = r.a; x.a_hi = r.a + 12; x.b_lo = x.b_hi = x.d_lo = x.d_hi = 0; }
            else { x.a_lo = lo_of(r.a); x.a_hi = hi_of(r.a); x.b_lo = lo_of(r.b); x.b_hi = hi_of(r.b); x.d_lo = lo_of(r.d); x.d_hi = hi_of(r.d); }
        }
        p->nlevels = nlevels;
        CU(cudaMalloc(&p->xrules, sizeof(wit_xrule) * nvars)); CU(cudaMalloc(&p->level_ptr, 8 * (nlevels + 1)));
        CU(cudaMemcpyAsync(p->xrules, xr.data(), sizeof(wit_xrule) * nvars, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(p->level_ptr, level_ptr, 8 * (nlevels + 1), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    undo.p = nullptr; ctx->wit[h] = p; *handle = h; return 0;
}
// bytes per message of the circuit the loaded program was recorded for (blsgpu_witness_gen / _check take nwit x that many message bytes)
long blsgpu_witness_msg_len(blsgpu_ctx* ctx, int handle) { return (!ctx || handle < 0 || handle >= 4 || !ctx->wit[handle]) ? -1 : (long)ctx->wit[handle]->msg_len; }
// declares a loaded program to be one of the aggregate_verify circuit (src/constraints.rs:153-191) with nkeys public keys: its input slots are
// [0, 2n) key coordinates, [2n, 3n) bitmap bits, [3n, 3n + 4) the signature, then 8 bits per message byte (blsgadget_aggregate_verify_program)
int blsgpu_witness_set_aggregate(blsgpu_ctx* ctx, int handle, size_t nkeys) {
    if (!ctx || handle < 0 || handle >= 4 || !ctx->wit[handle] || !nkeys) return BLSGPU_ERR_ARG;
    wit_prog* p = ctx->wit[handle];
    if (p->ninputs < 3 * nkeys + 4 || (p->ninputs - 3 * nkeys - 4) % 8) return fail(ctx, BLSGPU_ERR_ARG, "witness program: %zu input slots is not 3 x %zu keys + 4 + 8 x message bytes", p->ninputs, nkeys);
    p->nkeys = nkeys; p->msg_len = (p->ninputs - 3 * nkeys - 4) / 8; return 0;
}
int blsgpu_witness_free(blsgpu_ctx* ctx, int handle) {
    if (!ctx || handle < 0 || handle >= 4 || !ctx->wit[handle]) return BLSGPU_ERR_ARG;
    dev_guard guard_; cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
    wit_release(ctx->wit[handle]); ctx->wit[handle] = nullptr; return 0;
}
}
// decode, input slots and the level-synchronous replay for nwit triples: leaves the assignments in the transposed group layout
// (group stride nvars * 96 u32x4) in the workspace; `extra` bytes of workspace are reserved for the caller's own buffers, which it
// takes AFTER this returns.  want_zbool: also the packed 0/1 view per (group, variable).
// bitmap == NULL: verify program (pk48 = nwit keys); else aggregate program (pk48 = nwit x p.nkeys keys, bitmap = nwit x p.nkeys bytes)
static int witness_run(blsgpu_ctx* ctx, const wit_prog& p, const uint8_t* pk48, const uint8_t* bitmap, const uint8_t* msg, const uint8_t* sig96, size_t nwit, uint8_t* status, size_t extra,
                       bool want_zbool, u32x4** zt_out, uint2** zbool_out, uint8_t** dstatus_out) {
    if (p.msg_len == (size_t)-1) return fail(ctx, BLSGPU_ERR_ARG, "witness program: %zu input slots fit no verify circuit (6 + 8 x message bytes); an aggregate program needs blsgpu_witness_set_aggregate", p.ninputs);
    if ((bitmap != nullptr) != (p.nkeys != 0)) return fail(ctx, BLSGPU_ERR_ARG, p.nkeys ? "this is an aggregate_verify program: use the _aggregate entry points" : "this is a verify program: use blsgpu_witness_gen / _check");
    size_t groups = (nwit + 31) / 32, np = groups * 32, ninputs = p.ninputs, nk = p.nkeys ? p.nkeys : 1, nkeys_total = nwit * nk;
    if (int rc = ws_reserve(ctx, al(48 * nkeys_total) + al(nkeys_total) + al(96 * nwit) + al(p.msg_len * nwit + 1) + al(96 * nkeys_total) + al(192 * nwit) + 2 * al(np) + al(nkeys_total + np) + al(48 * ninputs * np) + al(groups * p.nvars * 1536) +
                                 (want_zbool ? al(groups * p.nvars * 8) : 0) + extra + 65536)) return rc;
    const uint8_t *dpk, *dsig, *dmsg, *dbm = nullptr;
    if (int rc = stage_in(ctx, dpk, pk48, 48 * nkeys_total)) return rc;
    if (bitmap) { if (int rc = stage_in(ctx, dbm, bitmap, nkeys_total)) return rc; }
    if (int rc = stage_in(ctx, dsig, sig96, 96 * nwit)) return rc;
    if (int rc = stage_in(ctx, dmsg, msg, p.msg_len * nwit ? p.msg_len * nwit : 1)) return rc;
    u32x4* pk_soa = ws_take<u32x4>(ctx, 6 * nkeys_total); u32x4* sig_soa = ws_take<u32x4>(ctx, 12 * nwit);
    uint8_t* code_pk = ws_take<uint8_t>(ctx, nkeys_total + np); uint8_t* code_sig = ws_take<uint8_t>(ctx, np);
    uint8_t* dstatus = status ? stage_out(ctx, status, nwit) : nullptr;
    fp* inputs = ws_take<fp>(ctx, ninputs * np);
    u32x4* zt_all = ws_take<u32x4>(ctx, groups * p.nvars * 96);
    uint2* zbool_all = want_zbool ? ws_take<uint2>(ctx, groups * p.nvars) : nullptr;
    LAUNCH(k_decode_g1, nblk(nkeys_total), TPB, dpk, nkeys_total, pk_soa, code_pk);
    LAUNCH(k_decode_g2, nblk(nwit), TPB, dsig, nwit, sig_soa, code_sig);
    if (bitmap) LAUNCH(k_witness_inputs_agg, nblk(np), TPB, (const u32x4*)pk_soa, (const uint8_t*)code_pk, p.nkeys, dbm, (const u32x4*)sig_soa, (const uint8_t*)code_sig, dmsg, p.msg_len, nwit, np, inputs, dstatus);
    else LAUNCH(k_witness_inputs, nblk(np), TPB, (const u32x4*)pk_soa, (const uint8_t*)code_pk, (const u32x4*)sig_soa, (const uint8_t*)code_sig, dmsg, p.msg_len, nwit, np, inputs, dstatus);
    if (p.xrules && ctx->wit_cluster) {                    // level-synchronous per group, hardware cluster barrier between levels
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3((unsigned)(groups * WIT_CL)); cfg.blockDim = dim3(WIT_CL_TPB); cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = WIT_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        wit_prog pc = p; const fp* in_c = inputs;
        CU(cudaLaunchKernelEx(&cfg, k_witness_cluster, pc, in_c, np, zt_all, groups, zbool_all)); ctx->launches++;
    } else if (p.xrules) {                                 // level-synchronous over the whole grid, cooperative launch: every block must be resident
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_witness_levels, WIT_TPB, 0)); if (per_sm > WIT_BPS) per_sm = WIT_BPS; CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, BLSGPU_ERR_CUDA, "k_witness_levels does not fit on an SM");
        wit_prog pc = p; const fp* in_c = inputs; void* args[] = {(void*)&pc, (void*)&in_c, (void*)&np, (void*)&zt_all, (void*)&groups, (void*)&zbool_all};
        CU(cudaLaunchCooperativeKernel((void*)k_witness_levels, dim3((unsigned)(per_sm * sms)), dim3(WIT_TPB), args, 0, ctx->stream)); ctx->launches++;
    } else {
        if (p.nvars != p.nout || want_zbool) return fail(ctx, BLSGPU_ERR_ARG, "the sequential replay needs a program without scratch columns (load it with its level order)");
        LAUNCH(k_witness_gen, (unsigned)groups, 32, p, (const fp*)inputs, np, zt_all);
    }
    *zt_out = zt_all; if (zbool_out) *zbool_out = zbool_all; *dstatus_out = dstatus;
    return 0;
}
extern "C" {
// assignments of the verify circuit for nwit (pk48, msg, sig96) triples: z48 = nwit * nout * 48 bytes (the layout of
// blsgpu_r1cs_check), status[i] = 0, or 2 / 3 when the key / signature does not decode to a non-identity point (its assignment
// is then all zeros except z[0] = 1 and the constants).  Pointers follow the context's pointer mode.
static int witness_gen_core(blsgpu_ctx* ctx, int handle, const uint8_t* pk48, const uint8_t* bitmap, const uint8_t* msg, const uint8_t* sig96, size_t nwit, uint8_t* z48, uint8_t* status) {
    if (handle < 0 || handle >= 4 || !ctx->wit[handle] || !pk48 || !msg || !sig96 || !z48) return fail(ctx, BLSGPU_ERR_ARG, "bad argument");
    if (!nwit) return 0;
    wit_prog p = *ctx->wit[handle];
    size_t groups = (nwit + 31) / 32;
    bool host = ctx->ptr_mode == BLSGPU_HOST;
    size_t zbytes = nwit * p.nout * 48;
    u32x4* zt_all; uint2* zbool_all = nullptr; uint8_t* dstatus;
    if (int rc = witness_run(ctx, p, pk48, bitmap, msg, sig96, nwit, status, host ? al(zbytes) : 0, p.xrules != nullptr, &zt_all, &zbool_all, &dstatus)) return rc;
    uint8_t* dz = host ? ws_take<uint8_t>(ctx, zbytes) : z48;
    { dim3 grid(nblk(p.nout, 8), (unsigned)groups); k_witness_untranspose<<<grid, 256, 0, ctx->stream>>>((const u32x4*)zt_all, (const uint2*)zbool_all, p.nvars, p.nout, nwit, (u32x4*)dz); ctx->launches++; CU(cudaGetLastError()); }
    if (host) CU(cudaMemcpyAsync(z48, dz, zbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (status) { if (int rc = finish_out(ctx, status, dstatus, nwit)) return rc; }
    return finish_call(ctx);
}
// Generation and satisfaction check in one call: the assignments never leave the transposed group layout (no 48-byte row-major
// copy, no second transpose; 34 MB per assignment stay out of the caller's memory).  sat_bits / all_sat as blsgpu_r1cs_check,
// status as blsgpu_witness_gen; the R1CS system must be the one the program was recorded with (same column count).
static int witness_check_core(blsgpu_ctx* ctx, int wit_handle, int r1cs_handle, const uint8_t* pk48, const uint8_t* bitmap, const uint8_t* msg, const uint8_t* sig96, size_t nwit,
                              uint64_t* sat_bits, uint8_t* all_sat, uint8_t* status) {
    if (wit_handle < 0 || wit_handle >= 4 || !ctx->wit[wit_handle] || r1cs_handle < 0 || r1cs_handle >= 16 || !ctx->r1cs[r1cs_handle] || !pk48 || !msg || !sig96 || !sat_bits)
        return fail(ctx, BLSGPU_ERR_ARG, "bad argument");
    if (!nwit) return 0;
    wit_prog p = *ctx->wit[wit_handle]; r1cs_sys s = *ctx->r1cs[r1cs_handle];
    if (p.nout != s.ncols) return fail(ctx, BLSGPU_ERR_ARG, "witness program and R1CS system have different column counts");
    if (!p.xrules) return fail(ctx, BLSGPU_ERR_ARG, "blsgpu_witness_check needs a program loaded with its level order");
    size_t groups = (nwit + 31) / 32, words = (s.nrows + 63) / 64, part_bytes = (s.n_seg ? s.n_seg : 1) * 32 * 48;
    bool host = ctx->ptr_mode == BLSGPU_HOST;
    u32x4* zt_all; uint2* zbool_all; uint8_t* dstatus;
    if (int rc = witness_run(ctx, p, pk48, bitmap, msg, sig96, nwit, status, al(part_bytes) + (host ? al(8 * words * nwit) + al(nwit) : 0), true, &zt_all, &zbool_all, &dstatus)) return rc;
    u32x4* part = ws_take<u32x4>(ctx, part_bytes / 16);
    uint64_t* dbits = host ? ws_take<uint64_t>(ctx, words * nwit) : sat_bits;
    uint8_t* dall = all_sat ? (host ? ws_take<uint8_t>(ctx, nwit) : all_sat) : nullptr;
    CU(cudaMemsetAsync(dbits, 0, 8 * words * nwit, ctx->stream));
    for (size_t gi = 0; gi < groups; gi++) {
        size_t w0 = gi * 32, g = nwit - w0 < 32 ? nwit - w0 : 32;
        if (int rc = r1cs_check_group(ctx, s, zt_all + gi * p.nvars * 96, zbool_all + gi * p.nvars, part, w0, g, words, dbits)) return rc;
    }
    if (dall) LAUNCH(k_r1cs_all, nblk(nwit, 8), 256, (const uint64_t*)dbits, nwit, words, s.nrows, dall);
    if (host) {
        CU(cudaMemcpyAsync(sat_bits, dbits, 8 * words * nwit, cudaMemcpyDeviceToHost, ctx->stream));
        if (all_sat) CU(cudaMemcpyAsync(all_sat, dall, nwit, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (status) { if (int rc = finish_out(ctx, status, dstatus, nwit)) return rc; }
    return finish_call(ctx);
}
int blsgpu_witness_gen(blsgpu_ctx* ctx, int handle, const uint8_t* pk48, const uint8_t* msg, const uint8_t* sig96, size_t nwit, uint8_t* z48, uint8_t* status) {
    ENTER(); return witness_gen_core(ctx, handle, pk48, nullptr, msg, sig96, nwit, z48, status);
}
int blsgpu_witness_check(blsgpu_ctx* ctx, int wit_handle, int r1cs_handle, const uint8_t* pk48, const uint8_t* msg, const uint8_t* sig96, size_t nwit,
                         uint64_t* sat_bits, uint8_t* all_sat, uint8_t* status) {
    ENTER(); return witness_check_core(ctx, wit_handle, r1cs_handle, pk48, nullptr, msg, sig96, nwit, sat_bits, all_sat, status);
}
// the aggregate_verify circuit (src/constraints.rs:153-191; program from blsgadget_aggregate_verify_program + blsgpu_witness_set_aggregate):
// pks48 = nwit x nkeys compressed keys, bitmap = nwit x nkeys bytes (0 / non-zero: the participation bits), msg = nwit x L bytes, sig96 = nwit
// aggregate signatures.  status: 2 when any of an item's keys does not decode to a non-identity point (masked-out keys are witnesses too), 3 for the signature.
int blsgpu_witness_gen_aggregate(blsgpu_ctx* ctx, int handle, const uint8_t* pks48, const uint8_t* bitmap, const uint8_t* msg, const uint8_t* sig96, size_t nwit, uint8_t* z48, uint8_t* status) {
    ENTER(); if (!bitmap) return fail(ctx, BLSGPU_ERR_ARG, "null bitmap"); return witness_gen_core(ctx, handle, pks48, bitmap, msg, sig96, nwit, z48, status);
}
int blsgpu_witness_check_aggregate(blsgpu_ctx* ctx, int wit_handle, int r1cs_handle, const uint8_t* pks48, const uint8_t* bitmap, const uint8_t* msg, const uint8_t* sig96, size_t nwit,
                                   uint64_t* sat_bits, uint8_t* all_sat, uint8_t* status) {
    ENTER(); if (!bitmap) return fail(ctx, BLSGPU_ERR_ARG, "null bitmap"); return witness_check_core(ctx, wit_handle, r1cs_handle, pks48, bitmap, msg, sig96, nwit, sat_bits, all_sat, status);
}
}
