"""Per-level timing of the witness replay (tuning build with -DWIT_TRACE: BLSGPU_SO=build_var/wit_trace.so): which dependency levels
cost what, correlated with the rules they hold.  python profiles/tools/wit_trace.py [n]"""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.getcwd())
from bls_verify_gadget_b200 import Context, synth, gadget as G
from bls_verify_gadget_b200._lib import lib
dev = torch.device("cuda", 0); ctx = Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pk, msg, sig, exp = synth.verify_batch_inputs(ctx, n, every=10**9, fast=True)
prog = G.verify_program(pk[:48].tobytes(), msg[:32].tobytes(), sig[:96].tobytes()); nvars = prog["nvars"]; nlev = prog["level_ptr"].size - 1
h = ctx.witness_load(prog)
c = G.verify_circuit(pk[:48].tobytes(), msg[:32].tobytes(), sig[:96].tobytes()); mats = c.matrices()
rh = ctx.r1cs_load([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], c.nrows, c.ncols)
for rep in range(2): bits, al, st = ctx.witness_check(h, rh, pk, msg, sig, c.nrows)
tr = np.zeros(nlev, np.uint64); assert lib().blsgpu_debug_wit_trace(tr.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nlev)) == 0
d = np.diff(tr.astype(np.int64)) / 1e3          # us per level (all but the last)
print("split", ctx.witness_shape(h))
nl = ctx.witness_shape(h)["field_levels"]; d = d[:nl - 1]
print(f"field levels {nl}, traced total {d.sum() / 1e3:.1f} ms, mean {d.mean():.2f} us, median {np.median(d):.2f}, p90 {np.percentile(d, 90):.2f}, p99 {np.percentile(d, 99):.2f}, max {d.max():.1f}")
hist = np.histogram(d, bins=[0, 2, 4, 6, 8, 10, 15, 20, 30, 50, 100, 1000, 1e9])
for lo_, hi_, cnt in zip(hist[1][:-1], hist[1][1:], hist[0]): print(f"  {lo_:>6.0f}-{hi_:<8.0f} us: {cnt:6d} levels, {d[(d >= lo_) & (d < hi_)].sum() / 1e3:7.1f} ms")
