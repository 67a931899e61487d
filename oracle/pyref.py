"""Big-integer CPU restatement of the reference's BLS12-381 verify path (TEST INFRASTRUCTURE ONLY).

This file is the *slow, obviously-right* oracle: Python ints, affine formulas, naive
exponentiations.  It pins the faster C oracle (oracle/bls_oracle.cpp) and generates the golden
vectors under tests/golden/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import it; the product (bls_verify_gadget_b200/) never does.

The arithmetic of the reference lives in un-vendored arkworks 0.4.x crates (ark-ff, ark-ec,
ark-bls12-381, ark-serialize; /root/reference/Cargo.toml:17-29), so each function restates the
published algorithm and cites the reference call site it stands behind:

  verify            src/bls.rs:427-458      sign/keygen        src/bls.rs:411-425, 210-216
  hash_to_g2        src/bls.rs:477-493      aggregate          src/bls.rs:183-195, 288-300
  expand / h2f      src/hasher.rs:58-173    sswu               src/hasher.rs:352-502 (RFC 9380 6.6.2 form)
  iso               src/hasher.rs:294-348   clear cofactor     src/hasher.rs:664-673 (h_eff)
  codecs            src/bls.rs:219-260, 316-357 (ZCash compressed format, big-endian)

Parity status: pinned by all 78 JSON fixtures of tests/test_cases and the inline KATs
(bls.rs:572, 622, 645; hasher.rs:822-862) -- see tests/test_oracle_golden.py.
GT bytes and R1CS vectors are parity-UNPINNED (no fixture in the reference observes them).
"""
import hashlib

p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X = 0xd201000000010000          # |x|; the BLS parameter x is negative
DST = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"   # bls.rs:482
H = lambda s: int(s, 16)

# ---------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1)
ZERO = (0, 0); ONE = (1, 0)
def add(a, b): return ((a[0] + b[0]) % p, (a[1] + b[1]) % p)
def sub(a, b): return ((a[0] - b[0]) % p, (a[1] - b[1]) % p)
def neg(a): return ((-a[0]) % p, (-a[1]) % p)
def mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
def sqr(a): return mul(a, a)
def conj(a): return (a[0], (-a[1]) % p)
def fpmul(a, s): return ((a[0] * s) % p, (a[1] * s) % p)
def inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], p - 2, p)
    return ((a[0] * n) % p, (-a[1] * n) % p)
def fpow(a, e):
    out = ONE
    for bit in bin(e)[2:]:
        out = sqr(out)
        if bit == '1': out = mul(out, a)
    return out
def is_square(a): return pow(a[0] * a[0] + a[1] * a[1], (p - 1) // 2, p) in (0, 1)
def sqrt(a):
    """Fp2 square root for p = 3 mod 4 (any root; callers fix the sign). None if a is a non-residue."""
    if a == ZERO: return ZERO
    a1 = fpow(a, (p - 3) // 4); alpha = mul(sqr(a1), a); x0 = mul(a1, a)
    if alpha == (p - 1, 0): x = mul((0, 1), x0)
    else: x = mul(fpow(add(ONE, alpha), (p - 1) // 2), x0)
    return x if sqr(x) == a else None
def sgn0(a):                                   # hasher.rs:520-530
    return (a[0] & 1) | ((a[0] == 0) & (a[1] & 1))

# ---------------------------------------------------------------- hash_to_field (hasher.rs:58-173)
def expand(msg, dst, n):
    ell = (n + 31) // 32; dstp = dst + bytes([len(dst)])
    b0 = hashlib.sha256(b"\0" * 64 + msg + n.to_bytes(2, 'big') + b"\0" + dstp).digest()
    bi = hashlib.sha256(b0 + b"\1" + dstp).digest(); out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dstp).digest(); out += bi
    return out[:n]
def h2f(msg, dst=DST):
    u = expand(msg, dst, 256); e = [int.from_bytes(u[64 * i:64 * i + 64], 'big') % p for i in range(4)]
    return (e[0], e[1]), (e[2], e[3])

# ---------------------------------------------------------------- SSWU on E' (hasher.rs:229-240, 352-502)
A_ISO = (0, 240); B_ISO = (1012, 1012); Z_SSWU = ((-2) % p, (-1) % p)
def g_iso(x): return add(add(mul(sqr(x), x), mul(A_ISO, x)), B_ISO)
def sswu(u):
    zu2 = mul(Z_SSWU, sqr(u)); ta = add(sqr(zu2), zu2)
    if ta == ZERO: x1 = mul(B_ISO, inv(mul(Z_SSWU, A_ISO)))
    else: x1 = mul(mul(neg(B_ISO), inv(A_ISO)), add(ONE, inv(ta)))
    gx1 = g_iso(x1)
    if is_square(gx1): x, y = x1, sqrt(gx1)
    else:
        x = mul(zu2, x1); y = sqrt(g_iso(x))
    if sgn0(u) != sgn0(y): y = neg(y)
    return x, y

# 3-isogeny E' -> E2 (RFC 9380 E.3; hasher.rs:294-348), coefficients in ascending degree
K1 = [(H("5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6"),) * 2,
      (0, H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a")),
      (H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e"),
       H("8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d")),
      (H("171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1"), 0)]
K2 = [(0, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa63")),
      (0xc, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa9f")), ONE]
K3 = [(H("1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706"),) * 2,
      (0, H("5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be")),
      (H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c"),
       H("8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f")),
      (H("124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10"), 0)]
K4 = [(H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa8fb"),) * 2,
      (0, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa9d3")),
      (0x12, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa99")), ONE]
def horner(k, x):
    acc = ZERO
    for c in reversed(k): acc = add(mul(acc, x), c)
    return acc
def iso(P):
    x, y = P
    xd = horner(K2, x); yd = horner(K4, x)
    if xd == ZERO or yd == ZERO: return None       # kernel of the isogeny -> identity (hasher.rs:340-345)
    return mul(horner(K1, x), inv(xd)), mul(y, mul(horner(K3, x), inv(yd)))

# ---------------------------------------------------------------- E2: y^2 = x^3 + 4(1+u), affine, None = identity
B2 = (4, 4)
def on_g2(P): return P is None or sqr(P[1]) == add(mul(sqr(P[0]), P[0]), B2)
def padd(P, Q):
    if P is None: return Q
    if Q is None: return P
    if P[0] == Q[0]:
        if P[1] != Q[1] or P[1] == ZERO: return None
        l = mul(fpmul(sqr(P[0]), 3), inv(fpmul(P[1], 2)))
    else: l = mul(sub(Q[1], P[1]), inv(sub(Q[0], P[0])))
    x = sub(sub(sqr(l), P[0]), Q[0]); return x, sub(mul(l, sub(P[0], x)), P[1])
def pneg(P): return None if P is None else (P[0], neg(P[1]))
def smul(k, P):
    R = None
    for bit in bin(k)[2:]:
        R = padd(R, R)
        if bit == '1': R = padd(R, P)
    return R
HEFF = H("0bc69f08f2ee75b3584c6a0ea91b352888e2a8e9145ad7689986ff031508ffe1329c2f178731db956d82bf015d1212b0"
         "2ec0ec69d7477c1ae954cbc06689f6a359894c0adebbf6b4e8020005aaa95551")        # hasher.rs:666
def map_to_g2_uncleared(msg, dst=DST):
    u0, u1 = h2f(msg, dst); return padd(iso(sswu(u0)), iso(sswu(u1)))
def hash_to_g2(msg, dst=DST):                 # bls.rs:477-493
    return smul(HEFF, map_to_g2_uncleared(msg, dst))
# psi endomorphism constants (hasher.rs:600-616), used only to cross-check fast paths
PSI_X = inv(fpow((1, 1), (p - 1) // 3)); PSI_Y = inv(fpow((1, 1), (p - 1) // 2))
def psi(P): return None if P is None else (mul(PSI_X, conj(P[0])), mul(PSI_Y, conj(P[1])))

# ---------------------------------------------------------------- E1: y^2 = x^3 + 4
def g1add(P, Q):
    if P is None: return Q
    if Q is None: return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0: return None
        l = 3 * P[0] * P[0] * pow(2 * P[1], p - 2, p) % p
    else: l = (Q[1] - P[1]) * pow(Q[0] - P[0], p - 2, p) % p
    x_ = (l * l - P[0] - Q[0]) % p; return x_, (l * (P[0] - x_) - P[1]) % p
def g1mul(k, P):
    R = None
    for bit in bin(k)[2:]:
        R = g1add(R, R)
        if bit == '1': R = g1add(R, P)
    return R
def g1neg(P): return None if P is None else (P[0], (-P[1]) % p)
G1 = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
      0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2 = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
      (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))

# ---------------------------------------------------------------- ZCash compressed codecs (bls.rs:219-260, 316-357)
class DeserErr(Exception): pass
def lex_largest(y):                        # arkworks Fp2 order: c1 first, then c0
    ny = neg(y); return (y[1], y[0]) > (ny[1], ny[0])
def deser_g1(b, subgroup=True):
    if len(b) < 48: raise DeserErr("short")
    b = b[:48]                             # the ark reader ignores trailing bytes
    c, i, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c: raise DeserErr("flags")
    if i: return None
    x_ = int.from_bytes(bytes([b[0] & 0x1f]) + b[1:], 'big')
    if x_ >= p: raise DeserErr("x>=p")
    y2 = (x_ ** 3 + 4) % p; y = pow(y2, (p + 1) // 4, p)
    if y * y % p != y2: raise DeserErr("not on curve")
    if (y > (p - y) % p) != bool(s): y = (p - y) % p
    P = (x_, y)
    if subgroup and g1mul(r, P) is not None: raise DeserErr("subgroup")
    return P
def ser_g1(P):
    if P is None: return bytes([0xc0]) + bytes(47)
    b = bytearray(P[0].to_bytes(48, 'big')); b[0] |= 0x80
    if P[1] > (p - P[1]) % p: b[0] |= 0x20
    return bytes(b)
def deser_g2(b, subgroup=True):
    if len(b) < 96: raise DeserErr("short")
    b = b[:96]; c, i, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c: raise DeserErr("flags")
    if i: return None
    x1 = int.from_bytes(bytes([b[0] & 0x1f]) + b[1:48], 'big'); x0 = int.from_bytes(b[48:], 'big')
    if x1 >= p or x0 >= p: raise DeserErr("x>=p")
    x_ = (x0, x1); y = sqrt(add(mul(sqr(x_), x_), B2))
    if y is None: raise DeserErr("not on curve")
    if lex_largest(y) != bool(s): y = neg(y)
    P = (x_, y)
    if subgroup and smul(r, P) is not None: raise DeserErr("subgroup")
    return P
def ser_g2(P):
    if P is None: return bytes([0xc0]) + bytes(95)
    b = bytearray(P[0][1].to_bytes(48, 'big') + P[0][0].to_bytes(48, 'big')); b[0] |= 0x80
    if lex_largest(P[1]): b[0] |= 0x20
    return bytes(b)

# ---------------------------------------------------------------- tower Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v)
XI = (1, 1)
def m6(a, b):
    a0, a1, a2 = a; b0, b1, b2 = b
    return (add(mul(a0, b0), mul(XI, add(mul(a1, b2), mul(a2, b1)))),
            add(add(mul(a0, b1), mul(a1, b0)), mul(XI, mul(a2, b2))),
            add(add(mul(a0, b2), mul(a1, b1)), mul(a2, b0)))
def a6(a, b): return tuple(add(x_, y_) for x_, y_ in zip(a, b))
def v6(a): return (mul(XI, a[2]), a[0], a[1])
Z6 = (ZERO, ZERO, ZERO); O6 = (ONE, ZERO, ZERO); O12 = (O6, Z6)
def m12(a, b):
    return (a6(m6(a[0], b[0]), v6(m6(a[1], b[1]))), a6(m6(a[0], b[1]), m6(a[1], b[0])))
def mul_by_014(f, c0, c1, c4): return m12(f, ((c0, c1, ZERO), (ZERO, c4, ZERO)))
def p12(a, e):
    R = O12
    for bit in bin(e)[2:]:
        R = m12(R, R)
        if bit == '1': R = m12(R, a)
    return R

# ---------------------------------------------------------------- optimal-ate Miller loop, ark-ec 0.4 style
TWO_INV = pow(2, p - 2, p)
def prepare_g2(Q):
    qx, qy = Q; rx, ry, rz = qx, qy, ONE; co = []
    for bit in bin(X)[3:]:
        a = fpmul(mul(rx, ry), TWO_INV); b = sqr(ry); c = sqr(rz)
        e = mul(B2, add(add(c, c), c)); f = add(add(e, e), e); g = fpmul(add(b, f), TWO_INV)
        h = sub(sqr(add(ry, rz)), add(b, c)); i = sub(e, b); j = sqr(rx); es = sqr(e)
        rx = mul(a, sub(b, f)); ry = sub(sqr(g), add(add(es, es), es)); rz = mul(b, h)
        co.append((i, add(add(j, j), j), neg(h)))
        if bit == '1':
            th = sub(ry, mul(qy, rz)); la = sub(rx, mul(qx, rz)); c = sqr(th); d = sqr(la); e = mul(la, d)
            f = mul(rz, c); g = mul(rx, d); h = sub(add(e, f), add(g, g))
            rx = mul(la, h); ry = sub(mul(th, sub(g, h)), mul(e, ry)); rz = mul(rz, e)
            j = sub(mul(th, qx), mul(la, qy)); co.append((j, neg(th), la))
    return co
def miller(pairs):
    """pairs: [(G1 affine, G2 affine)]; pairs with an identity member are dropped (ark-ec multi_miller_loop)."""
    pr = [(P, iter(prepare_g2(Q))) for P, Q in pairs if P is not None and Q is not None]
    f = O12
    def ell(f, c, P): return mul_by_014(f, c[0], fpmul(c[1], P[0]), fpmul(c[2], P[1]))
    for bit in bin(X)[3:]:
        f = m12(f, f)
        for P, it in pr: f = ell(f, next(it), P)
        if bit == '1':
            for P, it in pr: f = ell(f, next(it), P)
    return (f[0], tuple(neg(t) for t in f[1]))      # x < 0  =>  conjugate
EXPF = (p ** 12 - 1) // r
def final_exp(f):
    """arkworks' final exponentiation raises to 3*(p^12-1)/r (hard-part chain)."""
    return p12(f, 3 * EXPF)
def pairing_gt(pairs): return final_exp(miller(pairs))
def ser12(f):                               # 12 x 48-byte LE canonical, tower order
    return b"".join(c.to_bytes(48, 'little') for h in f for f2 in h for c in f2)

# ---------------------------------------------------------------- scheme (bls.rs:379-475)
ST_OK, ST_FALSE, ST_BAD_PK, ST_BAD_SIG, ST_EMPTY, ST_BAD_SK = 0, 1, 2, 3, 4, 5
def verify_points(pk, msg, sig):
    """bls.rs:427-458 on already-validated points.  Returns (status, GT or None)."""
    if pk is None: return ST_BAD_PK, None                    # identity test, bls.rs:434
    gt = pairing_gt([(g1neg(G1), sig), (pk, hash_to_g2(msg))])
    return (ST_OK if gt == O12 else ST_FALSE), gt
def verify_bytes(pk48, msg, sig96):
    """Batch-ABI semantics: undecodable pk -> 2, undecodable sig -> 3."""
    try: pk = deser_g1(pk48)
    except DeserErr: return ST_BAD_PK, None
    if pk is None: return ST_BAD_PK, None
    try: sig = deser_g2(sig96)
    except DeserErr: return ST_BAD_SIG, None
    return verify_points(pk, msg, sig)
def sk_to_pk(sk): return g1mul(sk, G1)                        # bls.rs:210-216
def sign(sk, msg):                                            # bls.rs:411-425
    if sk % r == 0: raise ValueError("InvalidSecretKey")
    return smul(sk, hash_to_g2(msg))
def aggregate_g1(pts):                                        # bls.rs:183-195 (None on empty)
    if not pts: return "empty"
    a = None
    for P in pts: a = g1add(a, P)
    return a
def aggregate_g2(pts):                                        # bls.rs:288-300
    if not pts: return "empty"
    a = None
    for P in pts: a = padd(a, P)
    return a

# ---------------------------------------------------------------- R1CS check (ark-relations is_satisfied)
def r1cs_check(A, B, C, z):
    """A,B,C: list of rows, each row a list of (coeff, col).  Returns the per-row satisfaction list."""
    dot = lambda row: sum(c * z[j] for c, j in row) % p
    return [(dot(a) * dot(b) - dot(c)) % p == 0 for a, b, c in zip(A, B, C)]
