// Host emulation of the CUDA path's per-item device functions (TEST TOOL ONLY, never shipped).
// The headers under bls_verify_gadget_b200/csrc compile for the host with plain-C fallbacks of the PTX carry
// chains; this file loops the very same stage functions the kernels call, so the algorithm layer can be checked
// against the oracle in the GPU-less authoring container.  The `-m gpu` tests exercise the real kernels.
#include "../devcheck/ops.h"
#include <cstring>
#include <vector>
using namespace bls;

static inline const uint8_t* msg_at(const uint8_t* msg, const uint32_t* off, size_t i, uint32_t& len) {
    if (off) { len = off[i + 1] - off[i]; return msg + off[i]; } len = 32; return msg + 32 * i;
}
extern "C" {
int emu_run_op(int op, const uint8_t* in, uint8_t* out, size_t n) {
    op_desc d = op_shape(op); if (!d.n_in) return -1;
    for (size_t i = 0; i < n; i++) { fp a[24], r[24]; memcpy(a, in + i * d.n_in * 48, d.n_in * 48); for (int k = 0; k < d.n_out; k++) r[k] = fp_zero(); run_op(op, a, r); memcpy(out + i * d.n_out * 48, r, d.n_out * 48); }
    return 0;
}
void emu_op_shape(int op, int* n_in, int* n_out) { op_desc d = op_shape(op); *n_in = d.n_in; *n_out = d.n_out; }
void emu_fp_mul_raw(const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) { fp x, y; memcpy(&x, a + 48 * i, 48); memcpy(&y, b + 48 * i, 48); fp z = fp_mul(x, y); memcpy(out + 48 * i, &z, 48); }
}
void emu_deser_g1(const uint8_t* in, size_t n, uint8_t* st) { for (size_t i = 0; i < n; i++) { g1_aff p; st[i] = (uint8_t)g1_decode(p, in + 48 * i); } }
void emu_deser_g2(const uint8_t* in, size_t n, uint8_t* st) { for (size_t i = 0; i < n; i++) { g2_aff p; st[i] = (uint8_t)g2_decode(p, in + 96 * i); } }
void emu_recode_g1(const uint8_t* in, size_t n, uint8_t* out) { for (size_t i = 0; i < n; i++) { g1_aff p; int rc = g1_decode(p, in + 48 * i); g1_encode(out + 48 * i, p, rc != DEC_OK); } }
void emu_recode_g2(const uint8_t* in, size_t n, uint8_t* out) { for (size_t i = 0; i < n; i++) { g2_aff p; int rc = g2_decode(p, in + 96 * i); g2_encode(out + 96 * i, p, rc != DEC_OK); } }
void emu_hash_to_g2(const uint8_t* msg, const uint32_t* off, size_t n, uint8_t* out96, int cleared) {
    for (size_t i = 0; i < n; i++) {
        uint32_t len; const uint8_t* m = msg_at(msg, off, i, len);
        g2_jac h; hash_to_g2_jac(h, m, len, cleared != 0);
        g2_aff a; bool ok = jac_to_aff(a, h); g2_encode(out96 + 96 * i, a, !ok);
    }
}
void emu_verify(const uint8_t* pk48, const uint8_t* msg, const uint32_t* off, const uint8_t* sig96, size_t n, uint8_t* status, uint8_t* gt576_each) {
    for (size_t i = 0; i < n; i++) {
        uint32_t len; const uint8_t* m = msg_at(msg, off, i, len);
        g1_aff pk; g2_aff sig, hm; uint8_t fl = 0, fl2 = 0;
        uint8_t st = stage_decode_pk(pk, pk48 + 48 * i);
        if (st == ST_OK) st = stage_decode_sig(sig, fl, sig96 + 96 * i);
        if (gt576_each) memset(gt576_each + 576 * i, 0, 576);
        if (st == ST_OK) {
            stage_hash(hm, fl2, m, len);
            fp12 f, gt; stage_miller(f, pk, hm, sig, fl | fl2);
            st = stage_final(gt, f);
            if (gt576_each) fp12_to_bytes(gt576_each + 576 * i, gt);
        }
        status[i] = st;
    }
}
void emu_sk_to_pk(const uint8_t* sk_le, size_t n, uint8_t* pk48) {
    for (size_t i = 0; i < n; i++) {
        uint32_t k[8]; memcpy(k, sk_le + 32 * i, 32);
        g1_aff g; g.x = fp_const(C_G1X); g.y = fp_const(C_G1Y);
        g1_jac r; jac_mul_scalar(r, g, k); g1_aff a; bool ok = jac_to_aff(a, r); g1_encode(pk48 + 48 * i, a, !ok);
    }
}
void emu_sign(const uint8_t* sk_le, const uint8_t* msg, const uint32_t* off, size_t n, uint8_t* sig96) {
    for (size_t i = 0; i < n; i++) {
        uint32_t len; const uint8_t* m = msg_at(msg, off, i, len);
        uint32_t k[8]; memcpy(k, sk_le + 32 * i, 32);
        g2_aff hm; uint8_t fl; stage_hash(hm, fl, m, len);
        g2_jac r; jac_mul_scalar(r, hm, k); g2_aff a; bool ok = jac_to_aff(a, r); g2_encode(sig96 + 96 * i, a, !ok);
    }
}
void emu_g1_sum(const uint8_t* pts48, size_t n, uint8_t* out48) {
    g1_jac acc; jac_set_identity(acc);
    for (size_t i = 0; i < n; i++) { g1_aff p; int rc = g1_decode(p, pts48 + 48 * i); if (rc == DEC_OK) jac_add_mixed(acc, acc, p); }
    g1_aff a; bool ok = jac_to_aff(a, acc); g1_encode(out48, a, !ok);
}
void emu_g2_sum(const uint8_t* pts96, size_t n, uint8_t* out96) {
    g2_jac acc; jac_set_identity(acc);
    for (size_t i = 0; i < n; i++) { g2_aff p; int rc = g2_decode(p, pts96 + 96 * i); if (rc == DEC_OK) { g2_jac pj; jac_from_aff(pj, p); jac_add(acc, acc, pj); } }
    g2_aff a; bool ok = jac_to_aff(a, acc); g2_encode(out96, a, !ok);
}
void emu_pairing_gt(const uint8_t* g1_48, const uint8_t* g2_96, size_t npairs, uint8_t* gt576) {   // npairs in {1, 2}
    g1_aff p[2]; g2_aff q[2]; bool use[2] = {false, false};
    for (size_t i = 0; i < npairs && i < 2; i++) { int a = g1_decode(p[i], g1_48 + 48 * i), b = g2_decode(q[i], g2_96 + 96 * i); use[i] = a == DEC_OK && b == DEC_OK; }
    if (npairs < 2) { p[1] = p[0]; q[1] = q[0]; }
    fp12 f, gt; miller_loop2(f, p[0], q[0], use[0], p[1], q[1], use[1]); final_exponentiation(gt, f); fp12_to_bytes(gt576, gt);
}
}
