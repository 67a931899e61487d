// How many resident warps per SM sub-partition does the IMAD.WIDE pipe need?  (sm_100a)
// Each kernel is launched as sms * w blocks of 128 threads (w warps per sub-partition, one block wave), w = 1, 2, 3, 4.
//   B   two independent carry chains of 8 IMAD.WIDE.U32.X each per step (16 mad.lo/madc.hi pairs fuse to 8 IMAD.WIDE; no ALU work)
//   M   dependent chain of out-of-line Montgomery products (csrc/fp.cuh fp_mul)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bls_verify_gadget_b200/csrc -o build/micro/micro_warps profiles/micro/micro_warps.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace bls;
#define CHAIN16(v, x, y) asm volatile( \
    "mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\tmadc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t" \
    "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\tmadc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.cc.u32 %7, %19, %24, %7;\n\t" \
    "madc.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\tmadc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t" \
    "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\tmadc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;" \
    : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) \
    : "r"(x), "r"(x + 1), "r"(x + 2), "r"(x + 3), "r"(x + 4), "r"(x + 5), "r"(x + 6), "r"(x + 7), "r"(y))
__global__ void __launch_bounds__(128) kB(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t e[16], o[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { e[j] = a + j; o[j] = b + j; }
    for (int it = 0; it < iters; it++) { CHAIN16(e, a, o[3]); CHAIN16(o, b, e[5]); }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= e[j] ^ o[j];
    if (s == 0x12345u) sink[0] = s;
}
// four independent chains in flight
__global__ void __launch_bounds__(128) kB4(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t e[16], o[16], f[16], g[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { e[j] = a + j; o[j] = b + j; f[j] = a ^ j; g[j] = b ^ j; }
    for (int it = 0; it < iters; it++) { CHAIN16(e, a, b); CHAIN16(o, b, a); CHAIN16(f, a, b); CHAIN16(g, b, a); a += e[3]; b += o[2]; }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= e[j] ^ o[j] ^ f[j] ^ g[j];
    if (s == 0x12345u) sink[0] = s;
}
__global__ void __launch_bounds__(128) kM(uint32_t* sink, int iters) {
    fp x, y;
#pragma unroll
    for (int j = 0; j < 12; j++) { x.l[j] = threadIdx.x * 77u + j; y.l[j] = blockIdx.x + 3u * j; }
    x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) x = fp_mul(x, y);
    if (x.l[0] == 0x12345u) sink[0] = x.l[3];
}
// two independent products per step (ILP across calls is impossible for out-of-line calls: this is the inlined form)
__global__ void __launch_bounds__(128) kM2(uint32_t* sink, int iters) {
    fp x, y, z;
#pragma unroll
    for (int j = 0; j < 12; j++) { x.l[j] = threadIdx.x * 77u + j; y.l[j] = blockIdx.x + 3u * j; z.l[j] = threadIdx.x + 5u * j; }
    x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu; z.l[11] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) { x = fp_mul_inl(x, y); z = fp_mul_inl(z, y); }
    if ((x.l[0] ^ z.l[0]) == 0x12345u) sink[0] = x.l[3];
}
template <class K> void run(K k, uint32_t* sink, int sms, int w, int iters, double imads_per_iter, const char* name, double ghz) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
    for (int r = 0; r < 4; r++) { cudaEventRecord(e0); k<<<sms * w, 128>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double cyc = best * 1e-3 * ghz * 1e9;             // cycles for the whole (single-wave) launch
    double per_imad = cyc / (iters * imads_per_iter * w);   // SMSP cycles per warp-level IMAD.WIDE
    printf("%-34s w=%d  %8.3f ms  %7.2f SMSP-cycles per IMAD.WIDE (4.00 = pipe rate)  %8.0f cycles/iter/warp\n", name, w, best, per_imad, cyc / iters);
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* sink; cudaMalloc(&sink, 64); double g = khz / 1e6;
    for (int w = 1; w <= 4; w++) run(kB, sink, sms, w, 4096, 16, "B  2 chains x 8 IMAD.WIDE.X", g);
    for (int w = 1; w <= 4; w++) run(kB4, sink, sms, w, 4096, 32, "B4 4 chains x 8 IMAD.WIDE.X (iterations serialised by a, b updates)", g);
    for (int w = 1; w <= 4; w++) run(kM, sink, sms, w, 2048, 300, "M  fp_mul chain (300 IMAD)", g);
    for (int w = 1; w <= 4; w++) run(kM2, sink, sms, w, 1024, 600, "M2 two inlined fp_mul (600 IMAD)", g);
    return 0;
}
