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
rules = prog["rules16"].reshape(-1, 16); kind = rules[:, 0]; order = prog["order"]; lp = prog["level_ptr"].astype(np.int64)
lc_ptr = prog["lc_ptr"].astype(np.int64)
a = rules[:, 4:8].copy().view(np.uint32).reshape(-1); b = rules[:, 8:12].copy().view(np.uint32).reshape(-1); dd = rules[:, 12:16].copy().view(np.uint32).reshape(-1)
def lclen(ids): return np.where(ids > 0, lc_ptr[np.maximum(ids, 1)] - lc_ptr[np.maximum(ids, 1) - 1], 0)
print(f"levels {nlev}, total {d.sum() / 1e3:.1f} ms, mean {d.mean():.2f} us, median {np.median(d):.2f}, p90 {np.percentile(d, 90):.2f}, p99 {np.percentile(d, 99):.2f}, max {d.max():.1f}")
hist = np.histogram(d, bins=[0, 2, 4, 6, 8, 10, 15, 20, 30, 50, 100, 1000, 1e9])
for lo_, hi_, cnt in zip(hist[1][:-1], hist[1][1:], hist[0]): print(f"  {lo_:>6.0f}-{hi_:<8.0f} us: {cnt:6d} levels, {d[(d >= lo_) & (d < hi_)].sum() / 1e3:7.1f} ms")
# per level: number of rules, max terms of a rule, kinds
nr = np.diff(lp)[:-1]; terms = np.zeros(nlev - 1); kinds = np.zeros((nlev - 1, 8), int)
for L in range(nlev - 1):
    vs = order[lp[L]:lp[L + 1]]; k = kind[vs]
    isf12 = k == 6
    t = lclen(np.where(isf12, 0, a[vs])) + lclen(np.where(isf12, 0, b[vs])) + lclen(np.where(isf12, 0, dd[vs]))
    terms[L] = t.max() if len(t) else 0
    for q in range(8): kinds[L, q] = (k == q).sum()
print("corr(time, rules in level) %.2f   corr(time, longest rule's terms) %.2f" % (np.corrcoef(d, nr)[0, 1], np.corrcoef(d, terms)[0, 1]))
for name, sel in (("levels with an inversion rule", kinds[:, [1, 3, 5, 6]].sum(1) > 0), ("levels with only MULADD/BIT/NEQ/INPUT", kinds[:, [1, 3, 5, 6]].sum(1) == 0)):
    print(f"  {name}: {sel.sum()} levels, {d[sel].sum() / 1e3:.1f} ms, mean {d[sel].mean():.2f} us")
for lo_, hi_ in ((0, 2), (2, 4), (4, 8), (8, 16), (16, 64), (64, 10**6)):
    sel = (terms >= lo_) & (terms < hi_); 
    if sel.sum(): print(f"  longest rule has {lo_}..{hi_} terms: {sel.sum()} levels, {d[sel].sum() / 1e3:.1f} ms, mean {d[sel].mean():.2f} us, mean rules/level {nr[sel].mean():.0f}")
for lo_, hi_ in ((0, 8), (8, 32), (32, 74), (74, 148), (148, 600), (600, 10**7)):
    sel = (nr >= lo_) & (nr < hi_)
    if sel.sum(): print(f"  {lo_}..{hi_} rules in the level: {sel.sum()} levels, {d[sel].sum() / 1e3:.1f} ms, mean {d[sel].mean():.2f} us")
