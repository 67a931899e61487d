// C ABI of the host-side circuit builder (libblsgadget.so): synthesises the reference's circuits and exports
// (A, B, C) as CSR with canonical 48-byte little-endian coefficients plus the assignment z, exactly the layout
// blsgpu_r1cs_load / blsgpu_r1cs_check (include/blsgpu.h) consume.  Host code only: no CUDA, no oracle.
#include "r1cs_verify.hpp"
#include <memory>
#include <functional>
#include <mutex>
using namespace gadget;

namespace {
struct Circuit { ConstraintSystem cs; uint8_t out[96]; int out_len = 0; int result = -1; };
std::mutex g_mu; std::vector<std::unique_ptr<Circuit>> g_tab;
int put(std::unique_ptr<Circuit> c) { std::lock_guard<std::mutex> l(g_mu); for (size_t i = 0; i < g_tab.size(); i++) if (!g_tab[i]) { g_tab[i] = std::move(c); return (int)i; } g_tab.push_back(std::move(c)); return (int)g_tab.size() - 1; }
Circuit* get(int h) { std::lock_guard<std::mutex> l(g_mu); return h >= 0 && (size_t)h < g_tab.size() ? g_tab[h].get() : nullptr; }
// Montgomery -> canonical little-endian; 93 % of an assignment is 0 / 1 (boolean variables): those skip the Montgomery product
void le48(uint8_t* o, const fp& m) {
    static const fp one = fp_one();
    if (fp_is_zero(m)) { memset(o, 0, 48); return; }
    if (fp_eq(m, one)) { memset(o, 0, 48); o[0] = 1; return; }
    fp a = fp_from_mont(m); memcpy(o, a.l, 48);
}
}

extern "C" {
// hash_to_g2_with_cons(cs, message) of src/hasher.rs:727-740; the message bytes are witnesses (as in the reference's tests,
// UInt8::new_witness_vec) or instance variables.  out96 = compressed H(m) as computed by the circuit.
int blsgadget_hash_to_g2(const uint8_t* msg, size_t len, int msg_is_instance, uint8_t out96[96]) {
    try {
        auto c = std::make_unique<Circuit>();
        std::vector<UInt8> m(len); for (size_t i = 0; i < len; i++) m[i] = msg_is_instance ? u8_input(c->cs, msg[i]) : u8_witness(c->cs, msg[i]);
        G2Var h = hash_to_g2_with_cons(c->cs, m);
        g2_aff a; bool ok = h.value_affine(a); g2_encode(c->out, a, !ok); c->out_len = 96;
        if (out96) memcpy(out96, c->out, 96);
        return put(std::move(c));
    } catch (...) { return -1; }
}
// BlsSignatureVerifyGadget::verify of src/constraints.rs:90-128 on (pk, msg, sig) given as the reference's byte formats;
// *result = the Boolean the gadget outputs.  Returns -2 when a point does not decode (the native path rejects it first).
int blsgadget_verify(const uint8_t pk48[48], const uint8_t* msg, size_t len, const uint8_t sig96[96], int* result, uint8_t* gt576 /*nullable: GT computed by the circuit*/) {
    try {
        g1_aff pk; g2_aff sig;
        int rp = g1_decode(pk, pk48), rs = g2_decode(sig, sig96);
        if (rp != DEC_OK || rs != DEC_OK) return -2;
        auto c = std::make_unique<Circuit>();
        fp12 gt; c->result = synthesize_verify(c->cs, pk, msg, len, sig, &gt) ? 1 : 0;
        if (gt576) fp12_to_bytes(gt576, gt);
        if (result) *result = c->result;
        return put(std::move(c));
    } catch (...) { return -1; }
}
// aggregate_verify of src/constraints.rs:153-167: n compressed keys, one bitmap byte per key (0 / non-zero)
int blsgadget_aggregate_verify(const uint8_t* pks48, size_t n, const uint8_t* bitmap, const uint8_t* msg, size_t len, const uint8_t sig96[96], int* result, uint32_t* count) {
    try {
        std::vector<g1_aff> pks(n); g2_aff sig;
        for (size_t i = 0; i < n; i++) {
            if (i && !memcmp(pks48 + 48 * i, pks48 + 48 * (i - 1), 48)) { pks[i] = pks[i - 1]; continue; }                // repeated keys (the reference's test uses 511 copies) decode once
            if (g1_decode(pks[i], pks48 + 48 * i) != DEC_OK) return -2;
        }
        if (g2_decode(sig, sig96) != DEC_OK) return -2;
        auto c = std::make_unique<Circuit>();
        uint32_t cnt = 0;
        c->result = synthesize_aggregate_verify(c->cs, pks, std::vector<uint8_t>(bitmap, bitmap + n), msg, len, sig, &cnt) ? 1 : 0;
        if (result) *result = c->result;
        if (count) *count = cnt;
        return put(std::move(c));
    } catch (...) { return -1; }
}
// The witness program of the verify circuit for messages of `len` bytes (the circuit's shape depends on the length through the
// number of SHA-256 blocks; len <= 8000): a full synthesis with rule recording on a sample input.  The program is otherwise
// input-independent; it is exported with blsgadget_program_export and replayed on the GPU by blsgpu_witness_gen.
int blsgadget_verify_program(const uint8_t pk48[48], const uint8_t* msg, size_t len, const uint8_t sig96[96]) {
    try {
        g1_aff pk; g2_aff sig;
        if (g1_decode(pk, pk48) != DEC_OK || g2_decode(sig, sig96) != DEC_OK) return -2;
        auto c = std::make_unique<Circuit>();
        c->cs.record_rules = true;
        if (len > 8000) return -4;
        c->result = synthesize_verify(c->cs, pk, msg, len, sig) ? 1 : 0;
        if (c->cs.rules.size() != c->cs.z.size()) return -3;
        return put(std::move(c));
    } catch (...) { return -1; }
}
// The witness program of the aggregate_verify circuit (constraints.rs:153-191) for n keys and messages of `len` bytes, recorded on a sample
// input; input slots: [0, 2n) key coordinates, [2n, 3n) bitmap bits, [3n, 3n + 4) signature, then 8 len message bits.
int blsgadget_aggregate_verify_program(const uint8_t* pks48, size_t n, const uint8_t* bitmap, const uint8_t* msg, size_t len, const uint8_t sig96[96]) {
    try {
        std::vector<g1_aff> pks(n); g2_aff sig;
        for (size_t i = 0; i < n; i++) {
            if (i && !memcmp(pks48 + 48 * i, pks48 + 48 * (i - 1), 48)) { pks[i] = pks[i - 1]; continue; }
            if (g1_decode(pks[i], pks48 + 48 * i) != DEC_OK) return -2;
        }
        if (g2_decode(sig, sig96) != DEC_OK) return -2;
        auto c = std::make_unique<Circuit>();
        c->cs.record_rules = true;
        uint32_t cnt = 0;
        c->result = synthesize_aggregate_verify(c->cs, pks, std::vector<uint8_t>(bitmap, bitmap + n), msg, len, sig, &cnt) ? 1 : 0;
        if (c->cs.rules.size() != c->cs.z.size()) return -3;
        return put(std::move(c));
    } catch (...) { return -1; }
}
// nvars: circuit variables (the assignment's length); ncols = nvars + scratch columns (the rule count of the program)
int blsgadget_program_shape(int h, uint64_t* nvars, uint64_t* ncols, uint64_t* nlc, uint64_t* nterms) {
    Circuit* c = get(h); if (!c || c->cs.rules.size() != c->cs.z.size()) return -1;
    *nvars = c->cs.rules.size(); *ncols = c->cs.rules.size() + c->cs.scratch_rules.size(); *nlc = c->cs.lc_ptr.size() - 1; *nterms = c->cs.lc_col.size(); return 0;
}
// rules16: ncols records (variables, then scratch columns) {u8 kind, u8 0, u16 aux, u32 a, u32 b, u32 d}; lc_ptr: nlc + 1; lc_col / lc_coef48: nterms
int blsgadget_program_export(int h, uint8_t* rules16, uint64_t* lc_ptr, uint32_t* lc_col, uint8_t* lc_coef48) {
    Circuit* c = get(h); if (!c || c->cs.rules.size() != c->cs.z.size()) return -1;
    static_assert(sizeof(Rule) == 16, "rule record layout");
    size_t nv = c->cs.rules.size();
    if (rules16) { memcpy(rules16, c->cs.rules.data(), 16 * nv); memcpy(rules16 + 16 * nv, c->cs.scratch_rules.data(), 16 * c->cs.scratch_rules.size()); }
    if (lc_ptr) memcpy(lc_ptr, c->cs.lc_ptr.data(), 8 * c->cs.lc_ptr.size());
    if (lc_col) for (size_t k = 0; k < c->cs.lc_col.size(); k++) { uint32_t v = c->cs.lc_col[k]; lc_col[k] = v >= ConstraintSystem::SCRATCH_BASE ? (uint32_t)(nv + (v - ConstraintSystem::SCRATCH_BASE)) : v; }
    if (lc_coef48) for (size_t k = 0; k < c->cs.lc_val.size(); k++) le48(lc_coef48 + 48 * k, c->cs.lc_val[k]);
    return 0;
}
// Dependency levels of the witness program: level(v) = 1 + max level of the variables its rule reads (inputs and z[0]: level 0).
// Rules of one level are independent, so the GPU evaluates a level with many warps and synchronises between levels.
// order[nvars]: variable indices sorted by level (stable); level_ptr[nlevels + 1]; returns nlevels (or < 0).
long blsgadget_program_levels(int h, uint32_t* order, uint64_t* level_ptr, uint64_t level_cap) {
    Circuit* c = get(h); if (!c || c->cs.rules.size() != c->cs.z.size()) return -1;
    const ConstraintSystem& cs = c->cs; size_t nv = cs.rules.size(), n = nv + cs.scratch_rules.size();
    std::vector<uint32_t> lvl(n, 0); uint32_t maxl = 0;
    auto col_of = [&](uint32_t v) { return v >= ConstraintSystem::SCRATCH_BASE ? (uint32_t)(nv + (v - ConstraintSystem::SCRATCH_BASE)) : v; };
    auto lc_level = [&](uint32_t id) { uint32_t m = 0; if (!id) return m; for (uint64_t k = cs.lc_ptr[id - 1]; k < cs.lc_ptr[id]; k++) m = std::max(m, lvl[col_of(cs.lc_col[k])]); return m; };
    // a scratch column only reads variables allocated before the rules that read it, and is read only by later variables: a
    // single pass in allocation order works if scratch levels are resolved on first use
    std::vector<uint8_t> done(cs.scratch_rules.size(), 0);
    std::function<void(uint32_t)> resolve_scratch = [&](uint32_t id) {
        if (!id) return;
        for (uint64_t k = cs.lc_ptr[id - 1]; k < cs.lc_ptr[id]; k++) {
            uint32_t v = cs.lc_col[k]; if (v < ConstraintSystem::SCRATCH_BASE) continue;
            size_t si = v - ConstraintSystem::SCRATCH_BASE; if (done[si]) continue;
            done[si] = 1; resolve_scratch(cs.scratch_rules[si].d);                   // partial sums first
            lvl[nv + si] = lc_level(cs.scratch_rules[si].d) + 1; maxl = std::max(maxl, lvl[nv + si]);
        }
    };
    for (size_t v = 1; v < nv; v++) {
        const Rule& r = cs.rules[v]; uint32_t m = 0;
        if (r.kind != RULE_INPUT && r.kind != RULE_FP12INV) { resolve_scratch(r.a); resolve_scratch(r.b); resolve_scratch(r.d); }
        if (r.kind == RULE_INPUT) { lvl[v] = 0; continue; }
        if (r.kind == RULE_FP12INV) { for (uint32_t k = 0; k < 12; k++) m = std::max(m, lc_level(r.a + k)); }
        else m = std::max(lc_level(r.a), std::max(lc_level(r.b), lc_level(r.d)));
        lvl[v] = m + 1; maxl = std::max(maxl, lvl[v]);
    }
    size_t nlev = (size_t)maxl + 1;
    if (!order || !level_ptr) return (long)nlev;
    if (level_cap < nlev + 1) return -2;
    std::vector<uint64_t> cnt(nlev + 1, 0);
    for (size_t v = 0; v < n; v++) cnt[lvl[v] + 1]++;
    for (size_t l = 0; l < nlev; l++) cnt[l + 1] += cnt[l];
    memcpy(level_ptr, cnt.data(), 8 * (nlev + 1));
    std::vector<uint64_t> pos(cnt.begin(), cnt.end() - 1);
    for (size_t v = 0; v < n; v++) order[pos[lvl[v]]++] = (uint32_t)v;
    return (long)nlev;
}
// assignment only (witness-only synthesis: no matrices are built): z48 must hold ncols * 48 bytes, ncols from a full synthesis
int blsgadget_verify_assignment(const uint8_t pk48[48], const uint8_t* msg, size_t len, const uint8_t sig96[96], uint8_t* z48, size_t ncols, int* result) {
    struct Guard { Guard() { witness_only_mode() = true; } ~Guard() { witness_only_mode() = false; } } guard;
    try {
        g1_aff pk; g2_aff sig;
        if (g1_decode(pk, pk48) != DEC_OK || g2_decode(sig, sig96) != DEC_OK) return -2;
        ConstraintSystem cs;
        bool r = synthesize_verify(cs, pk, msg, len, sig);
        if (cs.z.size() != ncols) return -3;
        for (size_t i = 0; i < cs.z.size(); i++) le48(z48 + 48 * i, cs.z[i]);
        if (result) *result = r ? 1 : 0;
        return 0;
    } catch (...) { return -1; }
}
int blsgadget_shape(int h, uint64_t* nrows, uint64_t* ncols, uint64_t* ninstance, uint64_t nnz[3]) {
    Circuit* c = get(h); if (!c) return -1;
    *nrows = c->cs.num_constraints(); *ncols = c->cs.num_variables(); *ninstance = c->cs.num_instance;
    for (int m = 0; m < 3; m++) nnz[m] = c->cs.col[m].size();
    return 0;
}
// any pointer may be NULL; rowptr[m] has nrows + 1 entries, col[m] / coeff48[m] nnz[m], z48 ncols * 48 bytes
int blsgadget_export(int h, uint64_t* const rowptr[3], uint32_t* const col[3], uint8_t* const coeff48[3], uint8_t* z48) {
    Circuit* c = get(h); if (!c) return -1;
    for (int m = 0; m < 3; m++) {
        if (rowptr && rowptr[m]) memcpy(rowptr[m], c->cs.rowptr[m].data(), 8 * c->cs.rowptr[m].size());
        if (col && col[m]) memcpy(col[m], c->cs.col[m].data(), 4 * c->cs.col[m].size());
        if (coeff48 && coeff48[m]) for (size_t k = 0; k < c->cs.val[m].size(); k++) le48(coeff48[m] + 48 * k, c->cs.val[m][k]);
    }
    if (z48) for (size_t i = 0; i < c->cs.z.size(); i++) le48(z48 + 48 * i, c->cs.z[i]);
    return 0;
}
// builder self-check on the host: index of the first unsatisfied row of the stored assignment, -1 if none
long blsgadget_first_unsatisfied(int h) { Circuit* c = get(h); return c ? c->cs.first_unsatisfied() : -2; }
int blsgadget_free(int h) { std::lock_guard<std::mutex> l(g_mu); if (h < 0 || (size_t)h >= g_tab.size() || !g_tab[h]) return -1; g_tab[h].reset(); return 0; }
}
