"""Latency of a small blsgpu_verify_batch call with the library defaults (python profiles/tools/latency_default.py)"""
import sys, os, time, json, numpy as np
sys.path.insert(0, os.getcwd())
from bls_verify_gadget_b200 import Context, synth
ctx = Context(0)
pk, msg, sig, exp = synth.verify_batch_inputs(ctx, 4096, every=64, fast=True)
out = []
for n in (1, 32, 1024, 4096):
    best = 1e9
    for rep in range(6):
        t0 = time.perf_counter(); st = ctx.verify(pk[:48 * n], msg[:32 * n], sig[:96 * n], fixed32=True); dt = time.perf_counter() - t0
        assert list(st) == list(exp[:n])
        if rep: best = min(best, dt)
    out.append({"n": n, "defaults": True, "ms_per_call": round(1e3 * best, 2), "verifies_per_sec": round(n / best)}); print(out[-1], flush=True)
json.dump(out, open("gpurun_out/latency_default_r02.json", "w"), indent=1)
