// Multi-GPU verify behind the C ABI (SURVEY 8(b) draft "blsgpu_create(ctx**, const int* devices, int ndev) owns streams + NCCL comm",
// 8(e)): one process drives every GPU of the box, so a Rust / C caller of the reference's BLS::verify
// (src/bls.rs:427-458) gets the sharded path without torch.distributed.
//
// The batch is cut into contiguous index ranges (multiples of 64, so ok-bitmap words never straddle GPUs), one per device; every
// device runs the whole single-GPU path on its range from its own host thread -- no data-path collective.  The one exchange step is
// an NCCL all-gather of the packed ok-bitmap shards and of the 576-byte GT partial products (one grouped call over all
// communicators), after which every device folds the partials in rank order (an Fp12 product is not an NCCL reduction operator) and
// holds the whole-batch bitmap and GT accumulator; device 0's copies go back to the caller.
//
// NCCL is bound lazily with dlopen("libnccl.so.2") inside blsgpu_create_multi: single-GPU users never load it, and a process that
// already carries an NCCL (PyTorch's) shares that copy instead of mapping a second one.  Included at the end of blsgpu.cu.
#pragma once
#include <dlfcn.h>
#include <thread>
#include <vector>
#include <string>

typedef struct ncclComm* blsgpu_nccl_comm;
struct nccl_api {
    void* so;
    int (*GetVersion)(int*);
    int (*CommInitAll)(blsgpu_nccl_comm*, int, const int*);
    int (*CommDestroy)(blsgpu_nccl_comm);
    int (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, blsgpu_nccl_comm, cudaStream_t);
    int (*GroupStart)(); int (*GroupEnd)();
    const char* (*GetErrorString)(int);
};
#define BLS_NCCL_UINT8 1          // ncclUint8 in every NCCL 2.x header

struct multi_dev {
    blsgpu_ctx* ctx; blsgpu_nccl_comm comm;
    uint8_t* arena; size_t arena_bytes;
    // carved per call
    uint8_t *pk, *sig, *msg, *status, *gt, *all_gt, *gt_folded; uint32_t* off; uint64_t *bitmap, *all_bitmap;
    int rc; size_t lo, hi;
};
struct blsgpu_multi {
    int ndev; std::vector<multi_dev> dev; nccl_api nccl; char err[512]; int nccl_version;
};
static int mfail(blsgpu_multi* m, int code, const char* fmt, ...) {
    if (m) { va_list ap; va_start(ap, fmt); vsnprintf(m->err, sizeof m->err, fmt, ap); va_end(ap); }
    return code;
}
static bool nccl_bind(nccl_api& a, std::string& why) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    a.so = nullptr;
    for (const char* n : names) { a.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.so) break; }
    if (!a.so) { why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?"); return false; }
#define BIND(field, sym) do { *(void**)(&a.field) = dlsym(a.so, sym); if (!a.field) { why = std::string("NCCL symbol missing: ") + sym; return false; } } while (0)
    BIND(GetVersion, "ncclGetVersion"); BIND(CommInitAll, "ncclCommInitAll"); BIND(CommDestroy, "ncclCommDestroy"); BIND(AllGather, "ncclAllGather");
    BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    return true;
}
// contiguous shard [lo, hi) of device r: sizes are multiples of 64 except the last (the layout of bls_verify_gadget_b200/dist.py)
static void multi_shard(size_t n, int ndev, int r, size_t& lo, size_t& hi, size_t& words) {
    size_t per = (n + ndev - 1) / ndev; per = (per + 63) / 64 * 64;
    lo = (size_t)r * per < n ? (size_t)r * per : n; hi = lo + per < n ? lo + per : n; words = per / 64;
}
static int multi_arena(blsgpu_multi* m, multi_dev& d, size_t bytes) {
    if (bytes <= d.arena_bytes) return 0;
    if (d.arena) { cudaStreamSynchronize(d.ctx->stream); cudaFree(d.arena); d.arena = nullptr; d.arena_bytes = 0; }
    size_t want = bytes + (bytes >> 3);
    if (cudaMalloc(&d.arena, want) != cudaSuccess) { cudaGetLastError(); return mfail(m, BLSGPU_ERR_ALLOC, "cudaMalloc of %zu staging bytes failed on device %d", want, d.ctx->device); }
    d.arena_bytes = want; return 0;
}

extern "C" {
int blsgpu_create_multi(blsgpu_multi** out, const int* devices, int ndev) {
    if (!out) return BLSGPU_ERR_ARG;
    *out = nullptr;
    int have = 0; if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { cudaGetLastError(); return BLSGPU_ERR_CUDA; }
    if (ndev <= 0) ndev = have;                                   // devices == NULL or ndev <= 0: every visible device, in order
    if (ndev > 64) return BLSGPU_ERR_ARG;
    std::vector<int> list(ndev);
    for (int i = 0; i < ndev; i++) { list[i] = (devices && i < ndev) ? devices[i] : i; if (list[i] < 0 || list[i] >= have) return BLSGPU_ERR_ARG; for (int j = 0; j < i; j++) if (list[j] == list[i]) return BLSGPU_ERR_ARG; }
    blsgpu_multi* m = new (std::nothrow) blsgpu_multi(); if (!m) return BLSGPU_ERR_ALLOC;
    m->ndev = ndev; m->err[0] = 0; m->dev.resize(ndev); m->nccl_version = 0;
    for (auto& d : m->dev) memset(&d, 0, sizeof d);
    dev_guard guard_;
    std::string why;
    if (!nccl_bind(m->nccl, why)) { delete m; return BLSGPU_ERR_CUDA; }
    m->nccl.GetVersion(&m->nccl_version);
    for (int i = 0; i < ndev; i++) if (int rc = blsgpu_create(&m->dev[i].ctx, list[i])) { for (int j = 0; j < i; j++) blsgpu_destroy(m->dev[j].ctx); delete m; return rc; }
    std::vector<blsgpu_nccl_comm> comms(ndev);
    int nrc = m->nccl.CommInitAll(comms.data(), ndev, list.data());
    if (nrc != 0) { for (auto& d : m->dev) blsgpu_destroy(d.ctx); delete m; return BLSGPU_ERR_CUDA; }
    for (int i = 0; i < ndev; i++) { m->dev[i].comm = comms[i]; blsgpu_set_pointer_mode(m->dev[i].ctx, BLSGPU_DEVICE); }
    *out = m; return 0;
}
void blsgpu_destroy_multi(blsgpu_multi* m) {
    if (!m) return;
    dev_guard guard_;
    for (auto& d : m->dev) {
        cudaSetDevice(d.ctx->device); cudaStreamSynchronize(d.ctx->stream);
        if (d.comm) m->nccl.CommDestroy(d.comm);
        if (d.arena) cudaFree(d.arena);
        blsgpu_destroy(d.ctx);
    }
    delete m;
}
const char* blsgpu_multi_last_error(blsgpu_multi* m) { return m ? m->err : "no multi-GPU context (no usable sm_100 device, bad device list, or NCCL not loadable)"; }
int blsgpu_multi_ndev(blsgpu_multi* m) { return m ? m->ndev : 0; }
int blsgpu_multi_nccl_version(blsgpu_multi* m) { return m ? m->nccl_version : 0; }
/* the single-GPU context of device slot i (0 <= i < ndev): for per-device calls of the rest of the ABI; it is in DEVICE pointer mode */
blsgpu_ctx* blsgpu_multi_ctx(blsgpu_multi* m, int i) { return (m && i >= 0 && i < m->ndev) ? m->dev[i].ctx : nullptr; }

// BLS::verify over a batch sharded across the context's GPUs.  HOST pointers (pinned for full PCIe rate); semantics and outputs of
// blsgpu_verify_batch: status[n], ok_bitmap (nullable) ceil(n/64) words, gt_acc_le576 (nullable) = product over the whole batch.
int blsgpu_multi_verify_batch(blsgpu_multi* m, const uint8_t* pk48, const uint8_t* msg, const uint32_t* msg_off, const uint8_t* sig96, size_t n,
                              uint8_t* status, uint64_t* ok_bitmap, uint8_t* gt_acc_le576) {
    if (!m) return BLSGPU_ERR_ARG;
    if (!pk48 || !msg || !sig96 || !status) return mfail(m, BLSGPU_ERR_ARG, "null pointer");
    if (!n) return 0;
    dev_guard guard_;
    int G = m->ndev; size_t words = 0;
    for (int r = 0; r < G; r++) multi_shard(n, G, r, m->dev[r].lo, m->dev[r].hi, words);
    // phase 1 (one host thread per device): stage the shard, run the single-GPU path on it in device-pointer mode
    auto work = [&](int r) {
        multi_dev& d = m->dev[r]; d.rc = 0;
        size_t cnt = d.hi - d.lo, mb0 = msg_off ? msg_off[d.lo] : 32 * d.lo, mb = msg_off ? msg_off[d.hi] - msg_off[d.lo] : 32 * cnt;
        if (cudaSetDevice(d.ctx->device) != cudaSuccess) { d.rc = BLSGPU_ERR_CUDA; return; }
        size_t need = al(48 * cnt + 16) + al(96 * cnt + 16) + al(mb + 16) + al(4 * (cnt + 1)) + al(cnt + 16) + al(8 * words) + al(8 * words * G) + 3 * al(576) + al(576 * (size_t)G) + 4096;
        if ((d.rc = multi_arena(m, d, need))) return;
        uint8_t* p = d.arena; auto take = [&](size_t b) { uint8_t* q = p; p += al(b); return q; };
        d.pk = take(48 * cnt + 16); d.sig = take(96 * cnt + 16); d.msg = take(mb + 16); d.off = (uint32_t*)take(4 * (cnt + 1)); d.status = take(cnt + 16);
        d.bitmap = (uint64_t*)take(8 * words); d.all_bitmap = (uint64_t*)take(8 * words * G); d.gt = take(576); d.gt_folded = take(576); d.all_gt = take(576 * (size_t)G);
        cudaStream_t st = d.ctx->stream;
        cudaError_t e = cudaMemsetAsync(d.bitmap, 0, 8 * words, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d.gt, 0, 576, st);
        const uint8_t one = 1; if (e == cudaSuccess) e = cudaMemcpyAsync(d.gt, &one, 1, cudaMemcpyHostToDevice, st);          // GT one = 01 00 .. 00: the partial of an empty shard
        if (cnt && e == cudaSuccess) e = cudaMemcpyAsync(d.pk, pk48 + 48 * d.lo, 48 * cnt, cudaMemcpyHostToDevice, st);
        if (cnt && e == cudaSuccess) e = cudaMemcpyAsync(d.sig, sig96 + 96 * d.lo, 96 * cnt, cudaMemcpyHostToDevice, st);
        if (cnt && mb && e == cudaSuccess) e = cudaMemcpyAsync(d.msg, msg + mb0, mb, cudaMemcpyHostToDevice, st);
        std::vector<uint32_t> rebased;
        if (cnt && msg_off && e == cudaSuccess) {
            rebased.resize(cnt + 1); for (size_t i = 0; i <= cnt; i++) rebased[i] = msg_off[d.lo + i] - (uint32_t)mb0;
            e = cudaMemcpyAsync(d.off, rebased.data(), 4 * (cnt + 1), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);                  // `rebased` is a stack-lifetime buffer
        }
        if (e != cudaSuccess) { d.rc = mfail(m, BLSGPU_ERR_CUDA, "staging on device %d failed: %s", d.ctx->device, cudaGetErrorString(e)); return; }
        if (cnt) {
            d.rc = blsgpu_verify_batch(d.ctx, d.pk, d.msg, msg_off ? d.off : nullptr, d.sig, cnt, d.status, ok_bitmap ? d.bitmap : nullptr, gt_acc_le576 ? d.gt : nullptr);
            if (d.rc) { mfail(m, d.rc, "device %d: %s", d.ctx->device, blsgpu_last_error(d.ctx)); return; }
            if (cudaMemcpyAsync(status + d.lo, d.status, cnt, cudaMemcpyDeviceToHost, st) != cudaSuccess) d.rc = mfail(m, BLSGPU_ERR_CUDA, "status copy failed on device %d", d.ctx->device);
        }
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < G; r++) th.emplace_back(work, r);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < G; r++) if (m->dev[r].rc) { for (auto& d : m->dev) { cudaSetDevice(d.ctx->device); cudaStreamSynchronize(d.ctx->stream); } return m->dev[r].rc; }
    // phase 2: the one exchange step -- grouped all-gathers of the bitmap shards and the GT partials over every communicator
    if (ok_bitmap || gt_acc_le576) {
        int nrc = m->nccl.GroupStart();
        for (int r = 0; r < G && nrc == 0; r++) {
            multi_dev& d = m->dev[r];
            if (ok_bitmap) nrc = m->nccl.AllGather(d.bitmap, d.all_bitmap, 8 * words, BLS_NCCL_UINT8, d.comm, d.ctx->stream);
            if (gt_acc_le576 && nrc == 0) nrc = m->nccl.AllGather(d.gt, d.all_gt, 576, BLS_NCCL_UINT8, d.comm, d.ctx->stream);
        }
        int erc = m->nccl.GroupEnd(); if (nrc == 0) nrc = erc;
        if (nrc != 0) return mfail(m, BLSGPU_ERR_CUDA, "NCCL all-gather failed: %s", m->nccl.GetErrorString(nrc));
        // phase 3: every device folds the G partials in rank order; device 0 reports
        if (gt_acc_le576) for (int r = 0; r < G; r++) {
            multi_dev& d = m->dev[r];
            if (int rc = blsgpu_gt_fold(d.ctx, d.all_gt, (size_t)G, d.gt_folded)) return mfail(m, rc, "device %d: %s", d.ctx->device, blsgpu_last_error(d.ctx));
        }
        multi_dev& d0 = m->dev[0];
        if (cudaSetDevice(d0.ctx->device) != cudaSuccess) return mfail(m, BLSGPU_ERR_CUDA, "cudaSetDevice failed");
        if (ok_bitmap && cudaMemcpyAsync(ok_bitmap, d0.all_bitmap, 8 * ((n + 63) / 64), cudaMemcpyDeviceToHost, d0.ctx->stream) != cudaSuccess) return mfail(m, BLSGPU_ERR_CUDA, "bitmap copy failed");
        if (gt_acc_le576 && cudaMemcpyAsync(gt_acc_le576, d0.gt_folded, 576, cudaMemcpyDeviceToHost, d0.ctx->stream) != cudaSuccess) return mfail(m, BLSGPU_ERR_CUDA, "GT copy failed");
    }
    for (auto& d : m->dev) {
        if (cudaSetDevice(d.ctx->device) != cudaSuccess || cudaStreamSynchronize(d.ctx->stream) != cudaSuccess) return mfail(m, BLSGPU_ERR_CUDA, "synchronising device %d failed: %s", d.ctx->device, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
// test / diagnostics hook: the whole-batch bitmap and GT accumulator as device slot i holds them after the last blsgpu_multi_verify_batch
// (every device must hold the same bytes as device 0)
int blsgpu_multi_peek(blsgpu_multi* m, int i, size_t n, uint64_t* ok_bitmap, uint8_t* gt_le576) {
    if (!m || i < 0 || i >= m->ndev || !m->dev[i].arena) return BLSGPU_ERR_ARG;
    dev_guard guard_; multi_dev& d = m->dev[i];
    if (cudaSetDevice(d.ctx->device) != cudaSuccess) return BLSGPU_ERR_CUDA;
    if (ok_bitmap && cudaMemcpy(ok_bitmap, d.all_bitmap, 8 * ((n + 63) / 64), cudaMemcpyDeviceToHost) != cudaSuccess) return BLSGPU_ERR_CUDA;
    if (gt_le576 && cudaMemcpy(gt_le576, d.gt_folded, 576, cudaMemcpyDeviceToHost) != cudaSuccess) return BLSGPU_ERR_CUDA;
    return 0;
}
}
