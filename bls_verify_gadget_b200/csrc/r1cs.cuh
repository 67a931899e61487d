// K8: R1CS satisfaction check  (A z) o (B z) == C z  over Fq (the BLS12-381 base field, the constraint field of
// the reference's circuits: src/constraints.rs:18, src/hasher.rs:32), as three CSR sparse mat-vecs.
// Replaces ark-relations' ConstraintSystem::is_satisfied applied to the system synthesised by
// BlsSignatureVerifyGadget::verify (src/constraints.rs:90-128); unlike arkworks' serial early-exit loop every row
// is reported.  Included at the end of blsgpu.cu (same translation unit: one copy of the Fp core).
//
// Mapping: lane <-> witness, warp <-> 64 consecutive rows.  All lanes of a warp walk the same CSR row, so column
// indices and coefficients are warp-uniform broadcast loads and the only divergent traffic is the gather of z,
// which is coalesced by transposing each group of 32 witnesses to  uint4 [col][3][32]  first.
// Coefficients +1 / -1 (the bulk of boolean/uint gadget rows) skip the Montgomery multiply; the branch is
// warp-uniform.  z stays canonical: coeff(Montgomery) x z(canonical) -> canonical, no conversion of z needed.
#pragma once

struct r1cs_sys {
    size_t nrows, ncols, nnz[3];
    uint64_t* rowptr[3]; uint32_t* col[3]; fp* coeff[3]; uint8_t* cls[3];
};
enum { R1_GENERAL = 0, R1_PLUS_ONE = 1, R1_MINUS_ONE = 2 };
#define R1_GROUP 32

__global__ void __launch_bounds__(TPB, BLS_MINB) k_r1cs_prepare(const uint8_t* coeff48, size_t nnz, fp* out, uint8_t* cls) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= nnz) return;
    fp v; const uint8_t* b = coeff48 + 48 * i;
    for (int w = 0; w < 12; w++) v.l[w] = b[4 * w] | ((uint32_t)b[4 * w + 1] << 8) | ((uint32_t)b[4 * w + 2] << 16) | ((uint32_t)b[4 * w + 3] << 24);
    fp one = fp_zero(); one.l[0] = 1;
    fp m1; fp_sub_raw(m1, fp_modulus(), one);
    cls[i] = fp_eq(v, one) ? R1_PLUS_ONE : (fp_eq(v, m1) ? R1_MINUS_ONE : R1_GENERAL);
    out[i] = fp_to_mont(v);
}
// z[w][col] (48-byte LE canonical) for witnesses w0 .. w0+g-1  ->  zt[(col*3 + c)*32 + lane]
__global__ void __launch_bounds__(256) k_r1cs_transpose(const u32x4* z, size_t ncols, size_t w0, size_t g, u32x4* zt) {
    size_t col = blockIdx.x * (size_t)8 + (threadIdx.x >> 5); int lane = threadIdx.x & 31;
    if (col >= ncols) return;
    u32x4 zero; zero.x = zero.y = zero.z = zero.w = 0;
    const u32x4* src = z + ((w0 + lane) * ncols + col) * 3;
    bool live = (size_t)lane < g;
    for (int c = 0; c < 3; c++) zt[(col * 3 + c) * 32 + lane] = live ? src[c] : zero;
}
__device__ __forceinline__ fp r1cs_load_z(const u32x4* zt, uint32_t col, int lane) {
    u32x4 a = zt[((size_t)col * 3) * 32 + lane], b = zt[((size_t)col * 3 + 1) * 32 + lane], c = zt[((size_t)col * 3 + 2) * 32 + lane];
    fp v; v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w; v.l[8] = c.x; v.l[9] = c.y; v.l[10] = c.z; v.l[11] = c.w;
    return v;
}
__device__ __forceinline__ fp r1cs_row_dot(const uint64_t* rowptr, const uint32_t* col, const fp* coeff, const uint8_t* cls, size_t row, const u32x4* zt, int lane) {
    fp acc = fp_zero();
    for (uint64_t k = rowptr[row], e = rowptr[row + 1]; k < e; k++) {
        fp zv = r1cs_load_z(zt, col[k], lane);
        uint8_t c = cls[k];
        if (c == R1_PLUS_ONE) acc = fp_add(acc, zv);
        else if (c == R1_MINUS_ONE) acc = fp_sub(acc, zv);
        else acc = fp_add(acc, fp_mul(coeff[k], zv));
    }
    return acc;
}
// one warp per block of 64 rows; lane = witness in the group.  sat word for (witness, row block) written by its lane.
__global__ void __launch_bounds__(TPB, BLS_MINB) k_r1cs_rows(r1cs_sys s, const u32x4* zt, size_t w0, size_t g, size_t words, uint64_t* sat_bits) {
    size_t rb = blockIdx.x * (size_t)(TPB / 32) + (threadIdx.x >> 5); int lane = threadIdx.x & 31;
    if (rb >= words) return;
    uint64_t bits = 0;
    size_t r_end = rb * 64 + 64 < s.nrows ? rb * 64 + 64 : s.nrows;
    for (size_t row = rb * 64; row < r_end; row++) {
        fp a = r1cs_row_dot(s.rowptr[0], s.col[0], s.coeff[0], s.cls[0], row, zt, lane);
        fp b = r1cs_row_dot(s.rowptr[1], s.col[1], s.coeff[1], s.cls[1], row, zt, lane);
        fp c = r1cs_row_dot(s.rowptr[2], s.col[2], s.coeff[2], s.cls[2], row, zt, lane);
        fp ab = fp_mul(fp_to_mont(a), b);                       // (aR)(b)/R = ab, canonical
        if (fp_eq(ab, c)) bits |= 1ull << (row & 63);
    }
    if ((size_t)lane < g) sat_bits[(w0 + lane) * words + rb] = bits;
}
__global__ void k_r1cs_all(const uint64_t* sat_bits, size_t nwit, size_t words, size_t nrows, uint8_t* all_sat) {
    size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (w >= nwit) return;
    bool all = true;
    for (size_t i = 0; i < words; i++) {
        uint64_t want = (i + 1 == words && (nrows & 63)) ? ((1ull << (nrows & 63)) - 1) : ~0ull;
        all &= sat_bits[w * words + i] == want;
    }
    all_sat[w] = all ? 1 : 0;
}

extern "C" {
int blsgpu_r1cs_load(blsgpu_ctx* ctx, const uint64_t* const rowptr[3], const uint32_t* const col[3], const uint8_t* const coeff48[3], size_t nrows, size_t ncols, int* handle) {
    ENTER(); if (!rowptr || !col || !coeff48 || !handle || !nrows || !ncols) return fail(ctx, BLSGPU_ERR_ARG, "bad argument");
    int h = -1; for (int i = 0; i < 16; i++) if (!ctx->r1cs[i]) { h = i; break; }
    if (h < 0) return fail(ctx, BLSGPU_ERR_ARG, "too many R1CS systems loaded");
    r1cs_sys* s = new (std::nothrow) r1cs_sys(); if (!s) return fail(ctx, BLSGPU_ERR_ALLOC, "out of host memory");
    memset(s, 0, sizeof *s); s->nrows = nrows; s->ncols = ncols;
    cudaMemcpyKind kind = ctx->ptr_mode == BLSGPU_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    ctx->r1cs[h] = s;
    for (int m = 0; m < 3; m++) {
        uint64_t last = 0;
        if (ctx->ptr_mode == BLSGPU_DEVICE) { CU(cudaMemcpyAsync(&last, rowptr[m] + nrows, 8, cudaMemcpyDeviceToHost, ctx->stream)); CU(cudaStreamSynchronize(ctx->stream)); }
        else last = rowptr[m][nrows];
        size_t nnz = s->nnz[m] = last, na = nnz ? nnz : 1;
        CU(cudaMalloc(&s->rowptr[m], 8 * (nrows + 1))); CU(cudaMalloc(&s->col[m], 4 * na)); CU(cudaMalloc(&s->coeff[m], 48 * na)); CU(cudaMalloc(&s->cls[m], na));
        CU(cudaMemcpyAsync(s->rowptr[m], rowptr[m], 8 * (nrows + 1), kind, ctx->stream));
        if (nnz) {
            CU(cudaMemcpyAsync(s->col[m], col[m], 4 * nnz, kind, ctx->stream));
            uint8_t* raw; CU(cudaMalloc(&raw, 48 * nnz));
            CU(cudaMemcpyAsync(raw, coeff48[m], 48 * nnz, kind, ctx->stream));
            LAUNCH(k_r1cs_prepare, nblk(nnz), TPB, (const uint8_t*)raw, nnz, s->coeff[m], s->cls[m]);
            CU(cudaStreamSynchronize(ctx->stream)); cudaFree(raw);
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    *handle = h; return 0;
}
int blsgpu_r1cs_free(blsgpu_ctx* ctx, int handle) {
    if (!ctx || handle < 0 || handle >= 16 || !ctx->r1cs[handle]) return BLSGPU_ERR_ARG;
    cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
    r1cs_sys* s = ctx->r1cs[handle];
    for (int m = 0; m < 3; m++) { cudaFree(s->rowptr[m]); cudaFree(s->col[m]); cudaFree(s->coeff[m]); cudaFree(s->cls[m]); }
    delete s; ctx->r1cs[handle] = nullptr; return 0;
}
int blsgpu_r1cs_check(blsgpu_ctx* ctx, int handle, const uint8_t* z48, size_t nwit, uint64_t* sat_bits, uint8_t* all_sat) {
    ENTER(); if (handle < 0 || handle >= 16 || !ctx->r1cs[handle] || !z48 || !sat_bits) return fail(ctx, BLSGPU_ERR_ARG, "bad argument");
    if (!nwit) return 0;
    r1cs_sys s = *ctx->r1cs[handle];
    size_t words = (s.nrows + 63) / 64;
    bool host = ctx->ptr_mode == BLSGPU_HOST;
    // host mode stages one group of 32 witnesses at a time (32 * ncols * 48 bytes) so the workspace stays bounded
    size_t zgroup = (size_t)R1_GROUP * s.ncols * 48;
    if (int rc = ws_reserve(ctx, (host ? al(zgroup) : 0) + al(zgroup) + (host ? al(8 * words * nwit) + al(nwit) : 0) + 8192)) return rc;
    u32x4* zstage = host ? ws_take<u32x4>(ctx, zgroup / 16) : nullptr;
    u32x4* zt = ws_take<u32x4>(ctx, zgroup / 16);
    uint64_t* dbits = host ? ws_take<uint64_t>(ctx, words * nwit) : sat_bits;
    uint8_t* dall = all_sat ? (host ? ws_take<uint8_t>(ctx, nwit) : all_sat) : nullptr;
    for (size_t w0 = 0; w0 < nwit; w0 += R1_GROUP) {
        size_t g = nwit - w0 < R1_GROUP ? nwit - w0 : R1_GROUP;
        const u32x4* zsrc; size_t wbase;
        if (host) { CU(cudaMemcpyAsync(zstage, z48 + w0 * s.ncols * 48, g * s.ncols * 48, cudaMemcpyHostToDevice, ctx->stream)); zsrc = zstage; wbase = 0; }
        else { zsrc = (const u32x4*)z48; wbase = w0; }
        LAUNCH(k_r1cs_transpose, nblk(s.ncols, 8), 256, zsrc, s.ncols, wbase, g, zt);
        LAUNCH(k_r1cs_rows, nblk(words, TPB / 32), TPB, s, (const u32x4*)zt, w0, g, words, dbits);
    }
    if (dall) LAUNCH(k_r1cs_all, nblk(nwit), TPB, (const uint64_t*)dbits, nwit, words, s.nrows, dall);
    if (host) {
        CU(cudaMemcpyAsync(sat_bits, dbits, 8 * words * nwit, cudaMemcpyDeviceToHost, ctx->stream));
        if (all_sat) CU(cudaMemcpyAsync(all_sat, dall, nwit, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return finish_call(ctx);
}
}
