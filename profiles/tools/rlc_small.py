import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from bls_verify_gadget_b200 import Context, synth
ctx = Context(0)
for n in (64, 4096):
    pk, msg, sig, exp = synth.verify_batch_inputs(ctx, n, every=10**9, fast=True)
    msgs = msg
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); ok, st = ctx.verify_rlc(pk, msg, sig, bytes(16), fixed32=True); dt = time.perf_counter() - t0
    print("n", n, "rlc call ms", round(dt * 1e3, 1), ok)
    for rep in range(2):
        t0 = time.perf_counter(); s2 = ctx.verify(pk, msg, sig, fixed32=True); dt = time.perf_counter() - t0
    print("n", n, "per-item verify ms", round(dt * 1e3, 1))
