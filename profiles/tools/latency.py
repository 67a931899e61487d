"""Latency of blsgpu_verify_batch against the batch size (host pointers, one call at a time): python profiles/tools/latency.py"""
import sys, os, time, json, numpy as np
sys.path.insert(0, os.getcwd())
from bls_verify_gadget_b200 import Context, synth
ctx = Context(0)
N = 1 << 17
pk, msg, sig, exp = synth.verify_batch_inputs(ctx, N, every=64, fast=True)
out = []
for n in (1, 32, 1024, 8192, 37888, 65536, 131072):
    for split, coop in ((1, 0), (1, 1), (0, 0)):
        ctx.set_split(split); ctx.set_coop(coop); best = 1e9
        for rep in range(4):
            t0 = time.perf_counter(); st = ctx.verify(pk[:48 * n], msg[:32 * n], sig[:96 * n], fixed32=True); dt = time.perf_counter() - t0
            assert list(st) == list(exp[:n])
            if rep: best = min(best, dt)
        out.append({"n": n, "split_stage_kernels": bool(split), "six_lane_small_pass_kernels": bool(coop), "ms_per_call": round(1e3 * best, 2), "verifies_per_sec": round(n / best)})
        print(out[-1], flush=True)
json.dump(out, open("gpurun_out/latency_r02.json", "w"), indent=1)
