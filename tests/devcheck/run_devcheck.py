"""Host-emulation vs device comparison of every primitive in tests/devcheck/ops.h on random field elements."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hostemu import emu
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
def dev_lib():
    so = os.environ.get("DEVCHECK_SO") or os.path.join(ROOT, "tests", "_hostemu", "libdevcheck.so")
    return ctypes.CDLL(so)
def rand_fp(rng, n):
    a = rng.integers(0, 256, size=(n, 48), dtype=np.uint8); a[:, 47] &= 0x0f      # < 2^380 < p: a valid Montgomery image
    return a
def check(ops=range(1, 34), n=256, seed=1):
    D = dev_lib(); rng = np.random.default_rng(seed); bad = []
    for op in ops:
        n_in, n_out = emu.op_shape(op)
        if not n_in: continue
        x = rand_fp(rng, n * n_in).reshape(-1)
        if op in (16,):          # make half of the inputs squares
            pass
        want = emu.run_op(op, x)
        got = np.zeros(48 * n_out * n, np.uint8)
        rc = D.dev_run_op(int(op), x.ctypes.data_as(ctypes.c_void_p), got.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
        ok = rc == 0 and np.array_equal(got, want)
        nbad = int((got.reshape(n, -1) != want.reshape(n, -1)).any(axis=1).sum())
        print(f"op {op:2d} in={n_in:2d} out={n_out:2d} rc={rc} {'ok' if ok else 'MISMATCH items=' + str(nbad)}")
        if not ok:
            bad.append(op); g = got.reshape(n, n_out, 48); w = want.reshape(n, n_out, 48)
            print('   mismatching output fps (item 0):', [k for k in range(n_out) if not np.array_equal(g[0, k], w[0, k])], ' items:', np.nonzero((g != w).any(axis=(1, 2)))[0][:8])
    return bad
if __name__ == "__main__":
    bad = check(); print("bad ops:", bad); sys.exit(1 if bad else 0)
