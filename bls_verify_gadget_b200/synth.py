"""Deterministic synthetic workloads: keys, messages, signatures and committees derived from
seed = 0x424c53 with SHA-256 in counter mode.  Point generation (pk = sk*g1, sig = sk*H(m)) runs on the GPU through
the library's own sk_to_pk / sign kernels and is excluded from every timed region."""
import hashlib
import numpy as np

SEED = bytes.fromhex("424c53")
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

def _h(tag, i): return hashlib.sha256(SEED + tag + int(i).to_bytes(8, "little")).digest()

def secret_keys(n, start=0):
    """sk_i = (SHA256(seed || "sk" || i) mod (r-1)) + 1, 32-byte little-endian"""
    out = bytearray(32 * n)
    for k in range(n):
        v = int.from_bytes(_h(b"sk", start + k), "big") % (R_ORDER - 1) + 1
        out[32 * k:32 * k + 32] = v.to_bytes(32, "little")
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()

def messages(n, start=0, tag=b"msg"):
    return np.frombuffer(b"".join(_h(tag, start + k) for k in range(n)), dtype=np.uint8).copy()

def fast_random_bytes(n, seed):
    """bulk pseudo-random bytes for the large configs (numpy PCG64; hashing 2^20 items in Python is too slow)"""
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)

def fast_secret_keys(n, seed=0x424c53):
    """n scalars in [1, 2^248): canonical Fr (top byte cleared), never zero"""
    sk = fast_random_bytes(32 * n, seed).reshape(n, 32)
    sk[:, 31] = 0; sk[:, 0] |= 1
    return sk.reshape(-1).copy()

CORRUPTIONS = ("wrong_msg", "pk_swap", "sig_swap", "sig_tamper", "pk_inf")

def verify_batch_inputs(ctx, n, every=64, fast=True):
    """cfg 2: n (pk, msg, sig) triples as compressed bytes; item i with i % every == every-1 is corrupted, cycling through
    CORRUPTIONS.  Returns (pk48, msg32, sig96, expected_status) -- expected by construction."""
    if fast: sk = fast_secret_keys(n); msg = fast_random_bytes(32 * n, 0x6d7367)
    else: sk = secret_keys(n); msg = messages(n)
    pk, st = ctx.sk_to_pk(sk); assert not st.any()
    sig, st = ctx.sign(sk, msg, fixed32=True); assert not st.any()
    pk = pk.reshape(n, 48).copy(); sig = sig.reshape(n, 96).copy(); msg = msg.reshape(n, 32).copy()
    exp = np.zeros(n, dtype=np.uint8)
    pk0, sig0 = pk.copy(), sig.copy()
    for j, i in enumerate(range(every - 1, n, every)):
        kind = CORRUPTIONS[j % len(CORRUPTIONS)]; nxt = (i + 1) % n
        if kind == "wrong_msg": msg[i, 0] ^= 1; exp[i] = 1
        elif kind == "pk_swap": pk[i] = pk0[nxt]; exp[i] = 1
        elif kind == "sig_swap": sig[i] = sig0[nxt]; exp[i] = 1          # a valid subgroup point, wrong signature
        elif kind == "sig_tamper": sig[i, 92:] = 0xff; exp[i] = 3         # last 4 bytes ffffffff => not on curve (w.h.p.)
        elif kind == "pk_inf": pk[i] = 0; pk[i, 0] = 0xc0; exp[i] = 2
    return pk.reshape(-1), msg.reshape(-1), sig.reshape(-1), exp

def committees(ctx, ncomm, k=512, pool=1 << 16, seed=0x636d):
    """cfg 3: validator pool of `pool` keys, ncomm committees of k pool indices; all members sign msg_c, so
    sig_c = (sum of member sks) * H(msg_c).  Returns (pks48 [ncomm*k*48], msg32, sig96, pool_pk48, idx)."""
    sk = fast_secret_keys(pool); pool_pk, st = ctx.sk_to_pk(sk); assert not st.any()
    pool_pk = pool_pk.reshape(pool, 48)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, pool, size=(ncomm, k))
    sk_int = [int.from_bytes(sk[32 * i:32 * i + 32].tobytes(), "little") for i in range(pool)]
    agg = bytearray(32 * ncomm)
    for c in range(ncomm):
        s = sum(sk_int[i] for i in idx[c]) % R_ORDER
        agg[32 * c:32 * c + 32] = s.to_bytes(32, "little")
    msg = fast_random_bytes(32 * ncomm, seed + 1)
    sig, st = ctx.sign(np.frombuffer(bytes(agg), dtype=np.uint8), msg, fixed32=True); assert not st.any()
    pks = pool_pk[idx.reshape(-1)].reshape(-1).copy()
    return pks, msg, sig, pool_pk, idx

def r1cs_system(nrows, ncols, seed=0x7231, frac_general=0.1, p=None):
    """Synthetic verify-shaped R1CS: per matrix 1 + Geom(0.5) non-zeros per row, 90% coefficients +-1 / small,
    10% uniform in Fq; C has one extra entry on a fresh 'product' column so that a satisfying z can be planted."""
    from .bls import R_ORDER as _r  # noqa
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    rng = np.random.default_rng(seed)
    nfree = ncols - nrows                    # columns [0, nfree) are free variables (col 0 = constant 1); column nfree+i is row i's product slot
    assert nfree >= 2
    mats = []
    for m in range(3):
        cnt = np.minimum(rng.geometric(0.5, size=nrows), 6).astype(np.int64)
        if m == 2: cnt[:] = 1
        rowptr = np.zeros(nrows + 1, dtype=np.uint64); rowptr[1:] = np.cumsum(cnt)
        nnz = int(rowptr[-1])
        col = rng.integers(0, nfree, size=nnz).astype(np.uint32)
        coeff = np.zeros((nnz, 48), dtype=np.uint8)
        kind = rng.random(nnz)
        gen = kind < frac_general
        coeff[gen] = rng.integers(0, 256, size=(int(gen.sum()), 48), dtype=np.uint8); coeff[gen, 47] &= 0x0f     # < 2^380 < p
        plus = (~gen) & (kind < frac_general + 0.5); coeff[plus, 0] = 1
        minus = (~gen) & ~plus
        coeff[minus] = np.frombuffer((P - 1).to_bytes(48, "little"), dtype=np.uint8)
        if m == 2:                            # C row i = 1 * z[nfree + i]
            col = (nfree + np.arange(nrows)).astype(np.uint32); coeff[:] = 0; coeff[:, 0] = 1
        mats.append((rowptr, col, coeff.reshape(-1)))
    return mats, nfree
