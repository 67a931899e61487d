import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from bls_verify_gadget_b200 import Context
ctx = Context(0)
rng = np.random.default_rng(0)
for wps in (1, 2, 3, 4, 8):
    n = 148 * 128 * wps
    a = rng.integers(0, 256, size=(n, 48), dtype=np.uint8); a[:, 47] &= 0x0f
    b = rng.integers(0, 256, size=(n, 48), dtype=np.uint8); b[:, 47] &= 0x0f
    da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda(); do = torch.empty_like(da)
    ctx.set_pointer_mode(True); s = torch.cuda.current_stream(); ctx.set_stream(s.cuda_stream)
    import ctypes
    from bls_verify_gadget_b200._lib import lib
    def run(reps):
        rc = lib().blsgpu_fp_mul_raw(ctx._h, ctypes.c_void_p(da.data_ptr()), ctypes.c_void_p(db.data_ptr()), ctypes.c_size_t(n), ctypes.c_void_p(do.data_ptr()), reps); assert rc == 0
    run(10); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 4000
    e0.record(); run(reps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cyc = ms * 1e-3 * 1.965e9
    print(f"warps/SMSP={wps}: {ms:.2f} ms, cycles per fp_mul per warp = {cyc / reps:.0f}, per SMSP-mul = {cyc / reps / wps:.0f}, MAC32/s = {n * reps * 300 / (ms * 1e-3) / 1e12:.2f} T")
