#!/usr/bin/env python3
"""Generates consts.cuh: every BLS12-381 constant the CUDA path needs, in Montgomery form (R = 2^384) as
12 x u32 little-endian limbs.  Pure integer arithmetic, no dependency on oracle/ (build tooling of the product).
Sources of the values: SURVEY Appendix A (curve parameters), reference src/hasher.rs:229-258 (SSWU constants),
hasher.rs:600-616 (psi constants), RFC 9380 E.3 (3-isogeny coefficients)."""
import os

p = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X = 0xd201000000010000
R = (1 << 384) % p
H = lambda s: int(s, 16)

def limbs(v, n=12): return [(v >> (32 * i)) & 0xffffffff for i in range(n)]
def fmt(ws): return "{" + ", ".join("0x%08xu" % w for w in ws) + "}"
def mont(v): return limbs(v % p * R % p)

# Fp2 helpers
def f2mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)
def f2pow(a, e):
    out = (1, 0)
    for bit in bin(e)[2:]:
        out = f2mul(out, out)
        if bit == '1': out = f2mul(out, a)
    return out
def f2inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], p - 2, p); return (a[0] * n % p, -a[1] * n % p)
def f2conj(a): return (a[0], -a[1] % p)

out = []
def emit_fp(name, v): out.append(f"#define {name} {fmt(mont(v))}")
def emit_fp2(name, v): out.append(f"#define {name} {{{fmt(mont(v[0]))}, {fmt(mont(v[1]))}}}")
def emit_words(name, v, n): out.append(f"#define {name} {fmt(limbs(v, n))}")

assert limbs(p)[0] == 0xffffaaab and (-pow(p, -1, 1 << 32)) % (1 << 32) == 0xfffcfffd
emit_words("BLS_C_P", p, 12)
emit_fp("BLS_C_ONE", 1)
emit_words("BLS_C_R2", R * R % p, 12)
emit_words("BLS_C_EXP_PM2", p - 2, 12)
emit_words("BLS_C_R3", R * R * R % p, 12)           # fix-up factor of the divsteps inversion (fp.cuh: fp_inv)
out.append("#define BLS_C_MU414 0x%xULL" % ((1 << 414) // p))      # Barrett factor of fp_mul_small
out.append("#define BLS_C_K400 {" + ", ".join("0x%08xu" % ((pow(2, 400, p) >> (32 * i)) & 0xffffffff) for i in range(12)) + "}")      # 2^400 mod p: folds the top of a lazy 14-limb sum (fp_lacc_reduce)
out.append("#define BLS_C_P_INV62 0x%016xULL" % pow(p, -1, 1 << 62))
emit_words("BLS_C_EXP_PM3D4", (p - 3) // 4, 12)
emit_fp("BLS_C_TWO_INV", pow(2, p - 2, p))
emit_fp("BLS_C_FOUR", 4)
emit_fp("BLS_C_2_256", 1 << 256)
emit_words("BLS_C_P_SQUARED", p * p, 24)
emit_words("BLS_C_3P_SQUARED", 3 * p * p, 24)      # offset of the lazily reduced Fp2 dot products (wide.cuh)
# offsets of the lazily reduced cooperative Fp12 product (coop.cuh): (16 - 2k) p^2 on the real and (5 - k) p^2 on the imaginary accumulator of lane k
out.append("#define BLS_C_COOP_OFF_RE {" + ", ".join(fmt(limbs((16 - 2 * k) * p * p, 24)) for k in range(6)) + "}")
out.append("#define BLS_C_COOP_OFF_IM {" + ", ".join(fmt(limbs((5 - k) * p * p, 24)) for k in range(6)) + "}")

# --- G1 generator and its negation
G1X = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
G1Y = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
assert (G1Y * G1Y - G1X ** 3 - 4) % p == 0
emit_fp("BLS_C_G1X", G1X); emit_fp("BLS_C_G1Y", G1Y); emit_fp("BLS_C_G1Y_NEG", p - G1Y)

# --- G1 endomorphism sigma(x,y) = (beta x, y): pick the cube root of unity with sigma(P) = -[x^2]P on G1
def g1add(P, Q):
    if P is None: return Q
    if Q is None: return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0: return None
        l = 3 * P[0] * P[0] * pow(2 * P[1], p - 2, p) % p
    else: l = (Q[1] - P[1]) * pow(Q[0] - P[0], p - 2, p) % p
    x_ = (l * l - P[0] - Q[0]) % p; return x_, (l * (P[0] - x_) - P[1]) % p
def g1mul(k, P):
    Rr = None
    for bit in bin(k)[2:]:
        Rr = g1add(Rr, Rr)
        if bit == '1': Rr = g1add(Rr, P)
    return Rr
G = (G1X, G1Y)
x2G = g1mul(X * X % r, G); target = (x2G[0], (-x2G[1]) % p)
beta = None
for g in range(2, 50):
    b = pow(g, (p - 1) // 3, p)
    if b != 1:
        for cand in (b, b * b % p):
            if (cand * G1X % p, G1Y) == target: beta = cand
        break
assert beta is not None and pow(beta, 3, p) == 1
emit_fp("BLS_C_BETA", beta)

# --- SSWU / isogeny constants (hasher.rs:229-240)
emit_fp2("BLS_C_ISO_A", (0, 240)); emit_fp2("BLS_C_ISO_B", (1012, 1012)); emit_fp2("BLS_C_SSWU_Z", (p - 2, p - 1))
# sqrt(-5): N(Z) = 5 is a non-residue, -1 too, so -5 is a residue (used to turn sqrt(-N) into sqrt(5N))
s5 = pow(p - 5, (p + 1) // 4, p); assert s5 * s5 % p == p - 5
emit_fp("BLS_C_SQRT_M5", s5)
K1 = [(H("5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6"),) * 2,
      (0, H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a")),
      (H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e"),
       H("8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d")),
      (H("171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1"), 0)]
K2 = [(0, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa63")),
      (0xc, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa9f")), (1, 0)]
K3 = [(H("1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706"),) * 2,
      (0, H("5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be")),
      (H("11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c"),
       H("8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f")),
      (H("124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10"), 0)]
K4 = [(H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa8fb"),) * 2,
      (0, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffa9d3")),
      (0x12, H("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaa99")), (1, 0)]
def emit_fp2_arr(name, ks):
    out.append(f"#define {name} {{" + ", ".join("{%s, %s}" % (fmt(mont(k[0])), fmt(mont(k[1]))) for k in ks) + "}")
emit_fp2_arr("BLS_C_ISO_K1", K1); emit_fp2_arr("BLS_C_ISO_K2", K2); emit_fp2_arr("BLS_C_ISO_K3", K3); emit_fp2_arr("BLS_C_ISO_K4", K4)

# --- psi endomorphism (hasher.rs:600-616): psi(x,y) = (PSI_X conj(x), PSI_Y conj(y)); psi^2(x,y) = (PSI2_X x, -y)
xi = (1, 1)
PSI_X = f2inv(f2pow(xi, (p - 1) // 3)); PSI_Y = f2inv(f2pow(xi, (p - 1) // 2))
assert PSI_X == (0, 4002409555221667392624310435006688643935503118305586438271171395842971157480381377015405980053539358417135540939437)
assert PSI_Y == (2973677408986561043442465346520108879172042883009249989176415018091420807192182638567116318576472649347015917690530,
                 1028732146235106349975324479215795277384839936929757896155643118032610843298655225875571310552543014690878354869257)
PSI2_X = f2mul(PSI_X, f2conj(PSI_X))
assert PSI2_X == (4002409555221667392624310435006688643935503118305586438271171395842971157480381377015405980053539358417135540939436, 0)
assert f2mul(PSI_Y, f2conj(PSI_Y)) == (p - 1, 0)
emit_fp2("BLS_C_PSI_X", PSI_X); emit_fp2("BLS_C_PSI_Y", PSI_Y); emit_fp("BLS_C_PSI2_X", PSI2_X[0])

# --- Frobenius coefficients: (w^i)^(p^k) = w^i * xi^(i(p^k-1)/6)
g1 = f2pow(xi, (p - 1) // 6)
F1 = [f2pow(g1, i) for i in range(6)]
emit_fp2_arr("BLS_C_FROB1", F1)
g2 = f2pow(xi, (p * p - 1) // 6)
F2 = [f2pow(g2, i) for i in range(6)]
assert all(c[1] == 0 for c in F2)
out.append("#define BLS_C_FROB2 {" + ", ".join(fmt(mont(c[0])) for c in F2) + "}")

emit_words("BLS_C_R_ORDER", r, 8)
out.append("#define BLS_X_ABS 0xd201000000010000ULL")

hdr = "// GENERATED by gen_consts.py -- do not edit.  Montgomery-form (R = 2^384) 12 x u32 LE limbs unless noted.\n#pragma once\n"
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "consts.cuh")
open(path, "w").write(hdr + "\n".join(out) + "\n")
print("wrote", path, len(out), "constants")
