"""Pins the CPU oracles (oracle/bls_oracle.cpp via ctypes, oracle/pyref.py) against every fixture and
known-answer test the reference holds for the path: tests/test_cases/** (78 JSON) and the
inline KATs of src/bls.rs:572,622,645 and src/hasher.rs:822-862.  Test harness mirrors
/root/reference/tests/tests.rs (same substitution rules: undecodable input => identity => false)."""
import hashlib
import numpy as np
import pytest
from conftest import hx
from oracle import cwrap as C
from oracle import pyref as R

def test_xmd_kats(eth):                              # hasher.rs:820-886
    k = eth["inline_kats"]; dst = hx(k["xmd_dst_hex"])
    assert C.expand_xmd(b"abc", dst, 32).hex() == k["xmd_abc_32"]
    assert C.expand_xmd(b"abc", dst, 128).hex().startswith(k["xmd_abc_128_prefix"])
    assert R.expand(b"abc", dst, 32).hex() == k["xmd_abc_32"]
    for n in (0, 1, 55, 56, 64, 119, 120, 250):
        m = bytes(range(256))[:n]
        assert C.expand_xmd(m, R.DST, 256) == R.expand(m, R.DST, 256)

def test_hash_to_g2_kat(eth, pyv):                   # bls.rs:643-652
    assert C.hash_to_g2([bytes(32)]).tobytes().hex() == eth["inline_kats"]["hash_to_g2_zero32"]
    msgs = [hx(v["msg"]) for v in pyv["hash_to_g2"]]
    out = C.hash_to_g2(msgs).reshape(-1, 96); unc = C.hash_to_g2(msgs, cleared=False).reshape(-1, 96)
    for i, v in enumerate(pyv["hash_to_g2"]):
        assert out[i].tobytes().hex() == v["out"] and unc[i].tobytes().hex() == v["uncleared"]

def test_sk_and_aggregate_kats(eth):                 # bls.rs:572-573, 620-641
    k = eth["inline_kats"]
    sk = hx(k["sk_le_hex"])
    assert int.from_bytes(sk, "little") == sum(l << (64 * i) for i, l in enumerate(k["sk_limbs64"]))
    sks = b"".join(hx(s) for s in k["aggregate_sks_le_hex"])
    pks = C.sk_to_pk(sks)
    agg, st = C.g1_aggregate(pks, [0, 4])
    assert st[0] == 0 and agg.tobytes().hex() == k["aggregate_pk"]
    assert C.deser_g1(hx(k["pubkey_roundtrip"]))[0] == 0 and C.deser_g2(hx(k["signature_roundtrip"]))[0] == 0
    # round trip through aggregate-of-one = decode + encode
    assert C.g1_aggregate(hx(k["pubkey_roundtrip"]), [0, 1])[0].tobytes().hex() == k["pubkey_roundtrip"]
    assert C.g2_aggregate(hx(k["signature_roundtrip"]), [0, 1])[0].tobytes().hex() == k["signature_roundtrip"]

def test_sign_fixtures(eth):                         # tests.rs:203-237
    for c in eth["sign"]:
        sk = hx(c["input"]["privkey"])[::-1]          # BE fixture -> LE (tests.rs:207-210)
        sig, st = C.sign(sk, [hx(c["input"]["message"])])
        if c["output"] is None: assert st[0] == 5, c["name"]
        else: assert st[0] == 0 and sig.tobytes() == hx(c["output"]), c["name"]

def test_verify_fixtures(eth):                       # tests.rs:240-268
    for c in eth["verify"]:
        i = c["input"]; st = C.verify(hx(i["pubkey"]), [hx(i["message"])], hx(i["signature"]))[0]
        assert (st == 0) == c["output"], c["name"]

def test_aggregate_fixtures(eth):                    # tests.rs:271-294
    for c in eth["aggregate"]:
        sigs = b"".join(hx(s) for s in c["input"])
        out, st = C.g2_aggregate(sigs, [0, len(c["input"])])
        if c["output"] is None: assert st[0] == 4, c["name"]
        else: assert st[0] == 0 and out.tobytes() == hx(c["output"]), c["name"]

def test_fast_aggregate_verify_fixtures(eth):        # tests.rs:297-334
    for c in eth["fast_aggregate_verify"]:
        i = c["input"]; pks = b"".join(hx(s) for s in i["pubkeys"])
        st = C.fast_aggregate_verify(pks, len(i["pubkeys"]), hx(i["message"]), hx(i["signature"]))[0]
        assert (st == 0) == c["output"], c["name"]

def test_aggregate_verify_pinned_by_the_fast_aggregate_fixtures(eth):
    """Eth2 AggregateVerify (oracle/bls_oracle.cpp ora_aggregate_verify) has no fixture of its own in the reference
    (tests/readme.md:4-7 names the category, tests/test_cases/ does not vendor it).  By bilinearity AggregateVerify(pks, [m] * k, sig) ==
    FastAggregateVerify(pks, m, sig): every fast_aggregate_verify fixture (tests.rs:297-334) must give the same verdict through it; and a
    single pair is BLS::verify (the 29 verify fixtures)."""
    import numpy as np
    for c in eth["fast_aggregate_verify"]:
        i = c["input"]; keys = [hx(s) for s in i["pubkeys"]]; sgn = hx(i["signature"])
        if any(len(k) != 48 for k in keys) or len(sgn) != 96: continue
        st = C.aggregate_verify(b"".join(keys), [hx(i["message"])] * len(keys), np.array([0, len(keys)], np.uint32), sgn)
        assert (st[0] == 0) == c["output"], c["name"]
    n = 0
    for c in eth["verify"]:
        i = c["input"]; pk, sgn = hx(i["pubkey"]), hx(i["signature"])
        if len(pk) != 48 or len(sgn) != 96: continue
        st = C.aggregate_verify(pk, [hx(i["message"])], np.array([0, 1], np.uint32), sgn); n += 1
        assert (st[0] == 0) == c["output"], c["name"]
    assert n >= 20

def test_uncompressed_codec_round_trip_and_fixture_verdicts(eth):
    """the uncompressed ZCash encodings of the oracle: every decodable deserialisation fixture survives compressed -> uncompressed ->
    compressed byte for byte, undecodable ones are rejected in the first step with the same verdict as tests.rs:337-364"""
    import numpy as np
    for kind, key, size, rec in (("deserialization_G1", "pubkey", 48, C.g1_recode), ("deserialization_G2", "signature", 96, C.g2_recode)):
        seen = 0
        for c in eth[kind]:
            s = c["input"][key]
            if len(s) != 2 * size: continue
            raw = np.frombuffer(bytes.fromhex(s), np.uint8); unc, st = rec(raw, True)
            assert (st[0] == 0) == c["output"], c["name"]
            if st[0] == 0:
                back, st2 = rec(unc, False); assert st2[0] == 0 and (np.array_equal(back, raw) or (raw[0] & 0x40)), c["name"]; seen += 1      # (a lenient identity encoding normalises)
        assert seen >= 2

@pytest.mark.parametrize("kind,key,size,fn", [("deserialization_G1", "pubkey", 48, C.deser_g1), ("deserialization_G2", "signature", 96, C.deser_g2)])
def test_deser_fixtures(eth, kind, key, size, fn):   # tests.rs:337-364
    for c in eth[kind]:
        s = c["input"][key]
        ok = len(s) % 2 == 0 and len(s) >= 2 * size and fn(bytes.fromhex(s)[:size])[0] == 0
        assert ok == c["output"], c["name"]

def test_gt_anchor_and_bilinearity(pyv):             # SURVEY A.9 (parity-unpinned vs arkworks; pinned C oracle == big-int oracle)
    g1 = R.ser_g1(R.G1); g2 = R.ser_g2(R.G2)
    gt = C.pairing_gt(g1, g2)
    assert hashlib.sha256(gt.tobytes()).hexdigest() == pyv["gt_anchor"]["sha256"] == "ff9912603bb02b77bc6ec1deaeddf9d1fee40ac17a781fb13c9c6e7a9f74d22b"
    tp = pyv["gt_two_pair"]
    gt2 = C.pairing_gt(b"".join(hx(s) for s in tp["g1"]), b"".join(hx(s) for s in tp["g2"]))
    assert gt2.tobytes().hex() == tp["bytes"]
    # e(2 g1, g2) == e(g1,g2)^2
    assert C.pairing_gt(R.ser_g1(R.g1mul(2, R.G1)), g2).tobytes() == C.gt_mul(gt, gt).tobytes()

def test_pyref_items_and_subgroup(pyv):
    it = pyv["sign_items"]
    sks = b"".join(hx(v["sk_le"]) for v in it); msgs = [hx(v["msg"]) for v in it]
    assert C.sk_to_pk(sks).tobytes() == b"".join(hx(v["pk"]) for v in it)
    sig, st = C.sign(sks, msgs)
    assert sig.tobytes() == b"".join(hx(v["sig"]) for v in it) and not st.any()
    st, gt = C.verify(C.sk_to_pk(sks), msgs, sig, want_gt=True)
    assert not st.any() and gt.tobytes() == R.ser12(R.O12)
    assert C.deser_g1(hx(pyv["g1_not_in_subgroup"]))[0] == 4 and C.deser_g2(hx(pyv["g2_not_in_subgroup"]))[0] == 4

def test_pyref_against_fixture_subset(eth):          # the big-int oracle itself, on a subset (it is slow)
    k = eth["inline_kats"]
    assert R.ser_g2(R.hash_to_g2(bytes(32))).hex() == k["hash_to_g2_zero32"]
    c = next(c for c in eth["sign"] if c["output"])
    assert R.ser_g2(R.sign(int(c["input"]["privkey"][2:], 16), hx(c["input"]["message"]))) == hx(c["output"])
    for name in ("verify_valid_case_195246ee3bd3b6ec", "verify_wrong_pubkey_case_195246ee3bd3b6ec"):
        c = next(c for c in eth["verify"] if c["name"] == name); i = c["input"]
        assert (R.verify_bytes(hx(i["pubkey"]), hx(i["message"]), hx(i["signature"]))[0] == 0) == c["output"]

def test_r1cs_oracles_agree():
    rng = np.random.default_rng(7); nrows, ncols = 40, 17
    def mat():
        rows = [[(int(rng.integers(1, 1 << 62)) * int(rng.integers(1, 1 << 62)) % R.p, int(rng.integers(0, ncols))) for _ in range(int(rng.integers(0, 4)))] for _ in range(nrows)]
        return rows
    A, B = mat(), mat()
    z = [1] + [int.from_bytes(rng.bytes(47), "little") for _ in range(ncols - 1)]
    dot = lambda row: sum(c * z[j] for c, j in row) % R.p
    # C row i: one fresh entry on column 0 (=1) carrying the product, so the system is satisfied; then break some rows
    Cm = [[(dot(a) * dot(b) % R.p, 0)] for a, b in zip(A, B)]
    for i in (3, 17, 39): Cm[i] = [((Cm[i][0][0] + 1) % R.p, 0)]
    def csr(M):
        rp = [0]; cl = []; cf = b""
        for row in M:
            for c, j in row: cl.append(j); cf += c.to_bytes(48, "little")
            rp.append(len(cl))
        return np.array(rp, dtype=np.uint64), np.array(cl, dtype=np.uint32), np.frombuffer(cf + b"\0", dtype=np.uint8)[:-1]
    ms = [csr(M) for M in (A, B, Cm)]
    z48 = b"".join(v.to_bytes(48, "little") for v in z)
    bits, allsat = C.r1cs_check([m[0] for m in ms], [m[1] for m in ms], [m[2] for m in ms], nrows, ncols, z48, 1)
    want = R.r1cs_check(A, B, Cm, z)
    got = [bool((int(bits[0, i // 64]) >> (i % 64)) & 1) for i in range(nrows)]
    assert got == want and want.count(False) == 3 and allsat[0] == 0

def test_fast_set_matches_simple_set(eth, pyv):
    """the arkworks-style algorithm set used for the timed CPU baseline gives the same bytes as the simple set used by the parity tests"""
    from conftest import hx
    rng = np.random.default_rng(44)
    pk = b"".join(hx(c["input"]["pubkey"]) for c in eth["verify"]); sig = b"".join(hx(c["input"]["signature"]) for c in eth["verify"])
    msgs = [hx(c["input"]["message"]) for c in eth["verify"]]
    mm = [rng.bytes(int(l)) for l in rng.integers(0, 120, size=12)]
    pts1 = b"".join(hx(c["input"]["pubkey"])[:48].ljust(48, b"\0") for c in eth["deserialization_G1"]) + hx(pyv["g1_not_in_subgroup"])
    pts2 = b"".join(hx(c["input"]["signature"])[:96].ljust(96, b"\0") for c in eth["deserialization_G2"] if len(c["input"]["signature"]) % 2 == 0) + hx(pyv["g2_not_in_subgroup"])
    def run():
        st, gt = C.verify(pk, msgs, sig, want_gt=True, threads=8)
        return st.tobytes(), gt.tobytes(), C.hash_to_g2(mm, threads=8).tobytes(), C.deser_g1(pts1).tobytes(), C.deser_g2(pts2).tobytes()
    slow = run()
    C.set_fast(True)
    try: fast = run()
    finally: C.set_fast(False)
    assert slow == fast
