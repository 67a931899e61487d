// Integer-pipe microbenchmarks that size the roofline denominator and drive the Fp multiplier design (sm_100a).
//   A  independent IMAD.WIDE.U32 (mad.wide.u32), 16 accumulators per thread
//   B  mad.lo.cc/madc.hi.cc carry chains -> IMAD.WIDE.U32.X with predicate carry in/out
//   C  "row" form: 12 independent mad.wide.u32 (a_j*b + t_j) followed by one 12-long add.cc/addc.cc chain
//   D  add.cc/addc.cc chains only (IADD3.X)
//   E  mad.lo.u32 + mad.hi.u32 pairs (32-bit IMAD / IMAD.HI), no carries
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_imad micro_imad.cu ; run: ./micro_imad
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void __launch_bounds__(256) kA(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    unsigned long long acc[16];
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = (unsigned long long)j * 0x9e3779b97f4a7c15ull + a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 5) & 15]), "r"(b));   // loop-varying multiplicand: no strength reduction
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= acc[j];
    if (s == 0x1234567ull) sink[0] = (uint32_t)s;
}
#define CHAIN16(v, x, y) asm volatile( \
    "mad.lo.cc.u32 %0, %16, %24, %0;\n\tmadc.hi.cc.u32 %1, %16, %24, %1;\n\tmadc.lo.cc.u32 %2, %17, %24, %2;\n\tmadc.hi.cc.u32 %3, %17, %24, %3;\n\t" \
    "madc.lo.cc.u32 %4, %18, %24, %4;\n\tmadc.hi.cc.u32 %5, %18, %24, %5;\n\tmadc.lo.cc.u32 %6, %19, %24, %6;\n\tmadc.hi.cc.u32 %7, %19, %24, %7;\n\t" \
    "madc.lo.cc.u32 %8, %20, %24, %8;\n\tmadc.hi.cc.u32 %9, %20, %24, %9;\n\tmadc.lo.cc.u32 %10, %21, %24, %10;\n\tmadc.hi.cc.u32 %11, %21, %24, %11;\n\t" \
    "madc.lo.cc.u32 %12, %22, %24, %12;\n\tmadc.hi.cc.u32 %13, %22, %24, %13;\n\tmadc.lo.cc.u32 %14, %23, %24, %14;\n\tmadc.hi.u32 %15, %23, %24, %15;" \
    : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) \
    : "r"(x), "r"(x + 1), "r"(x + 2), "r"(x + 3), "r"(x + 4), "r"(x + 5), "r"(x + 6), "r"(x + 7), "r"(y))
__global__ void __launch_bounds__(256) kB(uint32_t* sink, int iters) {
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
// C: 12 independent wide MACs then one carry chain folding hi_{j-1} into lo_j  (13 ALU adds per 12 MACs)
__global__ void __launch_bounds__(256) kC(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t t[13];
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = a ^ (j * 77u);
    for (int it = 0; it < iters; it++) {
        unsigned long long p[12];
#pragma unroll
        for (int j = 0; j < 12; j++) asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p[j]) : "r"(a + j), "r"(t[12 - j] ^ b), "l"((unsigned long long)t[j]));
        uint32_t lo[12], hi[12];
#pragma unroll
        for (int j = 0; j < 12; j++) { lo[j] = (uint32_t)p[j]; hi[j] = (uint32_t)(p[j] >> 32); }
        asm volatile("add.cc.u32 %0, %13, %24;\n\taddc.cc.u32 %1, %14, %25;\n\taddc.cc.u32 %2, %15, %26;\n\taddc.cc.u32 %3, %16, %27;\n\t"
                     "addc.cc.u32 %4, %17, %28;\n\taddc.cc.u32 %5, %18, %29;\n\taddc.cc.u32 %6, %19, %30;\n\taddc.cc.u32 %7, %20, %31;\n\t"
                     "addc.cc.u32 %8, %21, %32;\n\taddc.cc.u32 %9, %22, %33;\n\taddc.cc.u32 %10, %23, %34;\n\taddc.u32 %11, %35, 0;"
                     : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[0])
                     : "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(lo[8]), "r"(lo[9]), "r"(lo[10]), "r"(lo[11]),
                       "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]), "r"(hi[8]), "r"(hi[9]), "r"(hi[10]), "r"(hi[11]));
        t[0] = lo[0];
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 13; j++) s ^= t[j];
    if (s == 0x12345u) sink[0] = s;
}
__global__ void __launch_bounds__(256) kD(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t t[16], u[16];
#pragma unroll
    for (int j = 0; j < 16; j++) { t[j] = a ^ (j * 77u); u[j] = b + j; }
    for (int it = 0; it < iters; it++) {
        asm volatile("add.cc.u32 %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %18;\n\taddc.cc.u32 %3, %3, %19;\n\t"
                     "addc.cc.u32 %4, %4, %20;\n\taddc.cc.u32 %5, %5, %21;\n\taddc.cc.u32 %6, %6, %22;\n\taddc.cc.u32 %7, %7, %23;\n\t"
                     "addc.cc.u32 %8, %8, %24;\n\taddc.cc.u32 %9, %9, %25;\n\taddc.cc.u32 %10, %10, %26;\n\taddc.cc.u32 %11, %11, %27;\n\t"
                     "addc.cc.u32 %12, %12, %28;\n\taddc.cc.u32 %13, %13, %29;\n\taddc.cc.u32 %14, %14, %30;\n\taddc.u32 %15, %15, %31;"
                     : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
                     : "r"(u[0] ^ t[7]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]));
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s ^= t[j];
    if (s == 0x12345u) sink[0] = s;
}
__global__ void __launch_bounds__(256) kE(uint32_t* sink, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 977u;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = a + j; hi[j] = b + j; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[j]) : "r"(hi[(j + 3) & 7]), "r"(b)); asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[j]) : "r"(lo[(j + 5) & 7]), "r"(b)); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
    if (s == 0x12345u) sink[0] = s;
}
template <class K> double run(K k, uint32_t* sink, int blocks, double per_iter, const char* name, double clk_ghz, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<<<blocks, 256>>>(sink, ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double ops = (double)blocks * 256 * ITERS * per_iter; double rate = ops / (best * 1e-3);
    printf("%-44s %8.3f ms  %8.3f Tops/s  %6.2f ops/clk/SM @%.3f GHz\n", name, best, rate / 1e12, rate / (clk_ghz * 1e9 * sms), clk_ghz);
    return rate;
}
int main() {
    int sms, khz; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* sink; cudaMalloc(&sink, 64);
    double g = khz / 1e6; int blocks = sms * 8;
    printf("SMs %d, clock %.3f GHz (attr)\n", sms, g);
    run(kA, sink, blocks, 16, "A independent IMAD.WIDE.U32 (MAC32)", g, sms);
    run(kB, sink, blocks, 16, "B IMAD.WIDE.U32.X carry chains (MAC32)", g, sms);
    run(kC, sink, blocks, 12, "C 12 IMAD.WIDE + 12 IADD3.X chain (MAC32)", g, sms);
    run(kD, sink, blocks, 16, "D IADD3.X chains (adds)", g, sms);
    run(kE, sink, blocks, 8, "E IMAD.LO + IMAD.HI pairs (MAC32)", g, sms);
    return 0;
}
