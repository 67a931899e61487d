#!/usr/bin/env python3
"""Builds tests/golden/*.json.  Run in the authoring container only (needs /root/reference).

 * eth_vectors.json  -- the reference's own fixtures (tests/test_cases/**, the public Ethereum
   bls12-381-tests v0.1.2 vectors, tests/readme.md:4-7) consolidated into one file, plus the inline
   KAT constants of src/bls.rs:572-573, 622-641, 645-652 and src/hasher.rs:822-862.  DATA only.
 * pyref_vectors.json -- extra vectors computed by oracle/pyref.py (big-int restatement), used to
   pin the C oracle and the CUDA path on cases the reference has no fixture for (GT bytes, psi,
   uncleared map-to-curve output, edge-case encodings).  These are NOT reference outputs.
"""
import glob, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference/tests/test_cases"

def eth():
    out = {}
    for d in ("sign", "verify", "aggregate", "fast_aggregate_verify", "deserialization_G1", "deserialization_G2"):
        cases = []
        for f in sorted(glob.glob(f"{REF}/{d}/*.json")):
            c = json.load(open(f)); c["name"] = os.path.basename(f)[:-5]; cases.append(c)
        out[d] = cases
    out["inline_kats"] = {
        "sk_le_hex": "88c522e40e4d57abd3386ff6cb2c5496d767606488f3c9f9494cd363741d4e67",       # bls.rs:572
        "sk_limbs64": [12346421629811869064, 10832332258257352915, 17999185152888039383, 7443919619818212425],
        "pubkey_roundtrip": "a491d1b0ecd9bb917989f0e74f0dea0422eac4a873e5e2644f368dffb9a6e20fd6e10c1b77654d067c0618f6e5a7f79a",  # bls.rs:590
        "aggregate_sks_le_hex": ["88c522e40e4d57abd3386ff6cb2c5496d767606488f3c9f9494cd363741d4e" + s for s in ("67", "68", "69", "6a")],
        "aggregate_pk": "88843ab5f8471de849950c06674238f68899e242cbc72f81bda95647caea52513139792c6511b18eaf2942d04fc54cae",     # bls.rs:622
        "hash_to_g2_zero32": "97502412bcfc3f1d88b71f1ad9b60fa37c332d19466fba1dc991d42bcd09bcd9f1c22a562646ffce0922793b6c69938b"
                             "076e5cd6cfb3c361fc767e5f40ce05486e1668825ffeecab89d7daa455a179736a387ae93b9b15d283d45ffa14cd4af7",  # bls.rs:645
        "signature_roundtrip": "b2cc74bc9f089ed9764bbceac5edba416bef5e73701288977b9cac1ccb6964269d4ebf78b4e8aa7792ba09d3e49c8e6a"
                               "1351bdf582971f796bbaf6320e81251c9d28f674d720cca07ed14596b96697cf18238e0e03ebd7fc1353d885a39407e0",  # bls.rs:560
        "xmd_dst_hex": "412717974da474d0f8c420f320ff81e8432adb7c927d9bd082b4fb4d16c0a236",       # hasher.rs:822
        "xmd_abc_32": "52dbf4f36cf560fca57dedec2ad924ee9c266341d8f3d6afe5171733b16bbb12",         # hasher.rs:848
        "xmd_abc_128_prefix": "1a30a5e36fbdb87077552b9d18b9f0ae",                                  # hasher.rs:883
    }
    return out

def pyref_vectors():
    from oracle import pyref as R
    out = {}
    msgs = [b"", b"abc", bytes(32), bytes(range(32)), b"\xff" * 55, b"\x01" * 56, b"\x02" * 64,
            "h2c long message: 你好 BLS12-381 ".encode() * 8]
    out["hash_to_g2"] = [{"msg": m.hex(), "uncleared": R.ser_g2(R.map_to_g2_uncleared(m)).hex(),
                          "out": R.ser_g2(R.hash_to_g2(m)).hex()} for m in msgs]
    gt0 = R.pairing_gt([(R.G1, R.G2)])
    out["gt_anchor"] = {"sha256": hashlib.sha256(R.ser12(gt0)).hexdigest(), "bytes": R.ser12(gt0).hex()}
    # a two-pair GT value that is not one: e(-g1, [5]g2) * e([3]g1, g2)  = e(g1,g2)^-2
    gt1 = R.pairing_gt([(R.g1neg(R.G1), R.smul(5, R.G2)), (R.g1mul(3, R.G1), R.G2)])
    out["gt_two_pair"] = {"g1": [R.ser_g1(R.g1neg(R.G1)).hex(), R.ser_g1(R.g1mul(3, R.G1)).hex()],
                          "g2": [R.ser_g2(R.smul(5, R.G2)).hex(), R.ser_g2(R.G2).hex()], "bytes": R.ser12(gt1).hex()}
    # seeded sign/verify items incl. corrupted ones (SURVEY 8(d) cfg 2 recipe, first 8 items)
    seed = b"BLS"
    items = []
    for i in range(8):
        sk = int.from_bytes(hashlib.sha256(seed + b"sk" + i.to_bytes(8, "little")).digest(), "big") % (R.r - 1) + 1
        msg = hashlib.sha256(seed + b"msg" + i.to_bytes(8, "little")).digest()
        pk = R.ser_g1(R.sk_to_pk(sk)); sig = R.ser_g2(R.sign(sk, msg))
        items.append({"sk_le": sk.to_bytes(32, "little").hex(), "msg": msg.hex(), "pk": pk.hex(), "sig": sig.hex()})
    out["sign_items"] = items
    # psi / subgroup edge points: a point on E2 outside G2, and on E1 outside G1 (x found by search)
    x = 1
    while True:
        try:
            P = R.deser_g1(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big"), subgroup=False)
            if R.g1mul(R.r, P) is not None: break
        except R.DeserErr: pass
        x += 1
    out["g1_not_in_subgroup"] = R.ser_g1(P).hex()
    x = 1
    while True:
        try:
            Q = R.deser_g2(bytes([0x80]) + bytes(47) + x.to_bytes(48, "big"), subgroup=False)
            if R.smul(R.r, Q) is not None: break
        except R.DeserErr: pass
        x += 1
    out["g2_not_in_subgroup"] = R.ser_g2(Q).hex()
    # G2 x with zero imaginary rhs etc. are covered by random differential tests against the C oracle.
    return out

if __name__ == "__main__":
    json.dump(eth(), open(os.path.join(HERE, "eth_vectors.json"), "w"), indent=0, sort_keys=True)
    json.dump(pyref_vectors(), open(os.path.join(HERE, "pyref_vectors.json"), "w"), indent=0, sort_keys=True)
    print("wrote golden vectors")
