"""Stage times of a small verify pass (serial kernels, CUDA events at the stage boundaries): python profiles/tools/latency_stages.py"""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from bls_verify_gadget_b200 import Context, synth
ctx = Context(0); ctx.set_profiling(True); ctx.set_lanes(1)
pk, msg, sig, exp = synth.verify_batch_inputs(ctx, 4096, every=64, fast=True)
for n in (1, 1024, 4096):
    for rep in range(3): st = ctx.verify(pk[:48 * n], msg[:32 * n], sig[:96 * n], fixed32=True)
    print(n, {k: round(v, 2) for k, v in ctx.stage_times().items()}, flush=True)
