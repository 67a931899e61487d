import ctypes, sys
import os
D = ctypes.CDLL(os.environ.get("DEVCHECK_SO", "tests/_hostemu/libdevcheck.so")); D.dev_bench_op.restype = ctypes.c_float
# (op, name, Fp products per application)
ops = [(21, "fp_mul", 1), (29, "wide+redc", 1), (30, "fp2_mul lazy inline", 3), (31, "fp_mul_lz (split wide)", 1), (32, "fp2_dot n=1", 3), (33, "fp2_dot n=3", 9), (1, "fp2_mul", 3), (2, "fp2_sqr", 2), (3, "fp2_add", 0), (8, "fp6_mul", 18), (9, "fp12_mul", 54), (10, "fp12_sqr", 36), (11, "fp12_mul_by_014", 39), (12, "fp12_cyclo_sqr", 18), (5, "jac_dbl<fp2>", 16), (6, "jac_add_mixed<fp2>", 29)]
for wps in (2,):
    n = 148 * 128 * wps
    for op, name, m in ops:
        reps = max(8, 3000 // max(m, 1))
        ms = D.dev_bench_op(op, ctypes.c_size_t(n), reps)
        cyc = ms * 1e-3 * 1.965e9 / reps
        if m: print(f"{name:22s} warps/SMSP={wps} {cyc:9.0f} cycles/op  {cyc / m:7.0f} cycles per Fp product per warp  -> pipe efficiency {wps * 1200 * m / cyc * 100 / 1:5.1f} % of 4 cyc/IMAD.WIDE")
        else: print(f"{name:22s} warps/SMSP={wps} {cyc:9.0f} cycles/op")

D.dev_coop_bench.restype = ctypes.c_float
for which, name, m, ref in ((0, "coop_mul (6 lanes/item)", 54, "fp12_mul"), (1, "coop_cyclo_sqr (6 lanes/item)", 18, "fp12_cyclo_sqr")):
    reps = 400
    ms = D.dev_coop_bench(which, 2, reps)
    cyc = ms * 1e-3 * 1.965e9 / reps
    print(f"{name:30s} warps/SMSP=2 {cyc:9.0f} cycles per op per warp (5 items) = {cyc / 5:8.0f} warp-cycles per item   [thread-per-item {ref}: cycles/op above / 32 items]")
