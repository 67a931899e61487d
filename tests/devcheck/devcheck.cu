// Device unit-test library (TEST TOOL ONLY): applies one primitive of the CUDA device library per thread.
#include <cuda_runtime.h>
#include "ops.h"
using namespace bls;
__global__ void __launch_bounds__(128, 2) k_run_op(int op, const fp* in, fp* out, size_t n, int n_in, int n_out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i >= n) return;
    fp a[24], r[12];
    for (int k = 0; k < n_in; k++) a[k] = in[i * n_in + k];
    for (int k = 0; k < n_out; k++) r[k] = fp_zero();
    run_op(op, a, r);
    for (int k = 0; k < n_out; k++) out[i * n_out + k] = r[k];
}
extern "C" int dev_run_op(int op, const uint8_t* in, uint8_t* out, size_t n) {
    op_desc d = op_shape(op); if (!d.n_in) return -1;
    fp *din, *dout; size_t bi = n * d.n_in * 48, bo = n * d.n_out * 48;
    if (cudaMalloc(&din, bi) != cudaSuccess || cudaMalloc(&dout, bo) != cudaSuccess) return -2;
    cudaMemcpy(din, in, bi, cudaMemcpyHostToDevice);
    k_run_op<<<(unsigned)((n + 127) / 128), 128>>>(op, din, dout, n, d.n_in, d.n_out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out, dout, bo, cudaMemcpyDeviceToHost); cudaFree(din); cudaFree(dout);
    return e == cudaSuccess ? 0 : -3;
}
extern "C" void dev_op_shape(int op, int* n_in, int* n_out) { op_desc d = op_shape(op); *n_in = d.n_in; *n_out = d.n_out; }
