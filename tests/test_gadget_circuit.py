"""The host-side circuit builder (bls_verify_gadget_b200/gadget/): C++ mirror of the reference's gadget code.

Pins, all from the reference's own tests: the in-circuit hash-to-G2 must equal the native one (src/hasher.rs:1005-1026,
KAT src/bls.rs:645), and the verify circuit must return the Booleans of src/constraints.rs:326-332 (true, false, false)
for its (pk, msg, sig) cases.  The GT element the circuit computes is compared with the oracle's, the exported system with
its assignment is checked by the oracle's R1CS check (every row), and a perturbed assignment must be reported unsatisfied.
"""
import hashlib
import numpy as np
import pytest
from oracle import cwrap as C
from bls_verify_gadget_b200 import gadget as G

PK = bytes.fromhex("a491d1b0ecd9bb917989f0e74f0dea0422eac4a873e5e2644f368dffb9a6e20fd6e10c1b77654d067c0618f6e5a7f79a")
SIG = bytes.fromhex("882730e5d03f6b42c3abc26d3372625034e1d871b65a8a6b900a56dae22da98abbe1b68f85e49fe7652a55ec3d0591c2"
                    "0767677e33e5cbb1207315c41a9ac03be39c2e7668edc043d6cb1d9fd93033caa8a1c5b0e84bedaeb6c64972503a43eb")
CASES = [(bytes.fromhex("56" * 32), True), (bytes.fromhex("56" * 31 + "57"), False), (bytes.fromhex("78" * 32), False)]      # constraints.rs:326-332

def _oracle_check(c, z, nwit=1):
    mats = c.matrices()
    return C.r1cs_check([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], c.nrows, c.ncols, z, nwit, threads=C.hw_threads())

def test_hash_to_g2_circuit_matches_native(eth):
    c = G.hash_to_g2_circuit(bytes(32))
    assert c.output.hex() == eth["inline_kats"]["hash_to_g2_zero32"]                    # src/bls.rs:645
    assert c.ninstance == 1 and c.first_unsatisfied() == -1
    assert 600_000 < c.nrows < 900_000 and c.nrows < c.ncols + 20_000                   # SHA-256 dominated: ~39 k rows per compression x 18
    bits, allsat = _oracle_check(c, c.assignment())
    assert allsat[0] == 1
    # ragged message, message bytes as instance variables: same function as the native path, instance block first
    msg = b"abcdefghijklmnopqrstuvwxyz0123456789-ragged"
    c2 = G.hash_to_g2_circuit(msg, message_is_instance=True)
    assert c2.output == C.hash_to_g2([msg]).tobytes() and c2.ninstance == 1 + 8 * len(msg) and c2.first_unsatisfied() == -1
    z = c2.assignment().reshape(-1, 48)
    want_bits = np.unpackbits(np.frombuffer(msg, np.uint8), bitorder="little")
    assert np.array_equal(z[1:1 + 8 * len(msg), 0], want_bits) and not z[1:1 + 8 * len(msg), 1:].any()

def test_verify_circuit_reference_cases():
    shapes = set(); digests = set()
    for msg, expect in CASES:
        c = G.verify_circuit(PK, msg, SIG)
        assert c.result == expect
        st, gt = C.verify(np.frombuffer(PK, np.uint8), [msg], np.frombuffer(SIG, np.uint8), want_gt=True, threads=1)
        assert (st[0] == 0) == expect and c.gt == gt.tobytes()                          # the circuit computes the same GT element as the native path
        assert c.first_unsatisfied() == -1                                              # the gadget RETURNS a Boolean, it does not enforce validity
        shapes.add((c.nrows, c.ncols, tuple(c.nnz)))
        digests.add(hashlib.sha256(b"".join(x.tobytes() for m in c.matrices() for x in m)).hexdigest())
        c.free()
    assert len(shapes) == 1 and len(digests) == 1                                       # the matrices do not depend on the inputs

def test_verify_circuit_oracle_check_and_perturbation():
    c = G.verify_circuit(PK, CASES[0][0], SIG)
    z = c.assignment().reshape(c.ncols, 48).copy()
    zw, rw = G.verify_witnesses([(PK, CASES[0][0], SIG), (PK, CASES[2][0], SIG)], threads=2, ncols=c.ncols)      # witness-only synthesis
    assert np.array_equal(zw[0], z.reshape(-1)) and list(rw) == [True, False]
    rng = np.random.default_rng(3)
    bad = z.copy(); victims = sorted(int(v) for v in rng.integers(1, c.ncols, size=3))
    for v in victims: bad[v, 0] ^= 1
    zz = np.concatenate([z.reshape(-1), bad.reshape(-1)])
    bits, allsat = _oracle_check(c, zz, nwit=2)
    assert list(allsat) == [1, 0]
    nun = c.nrows - int(np.unpackbits(bits[1].view(np.uint8), bitorder="little")[:c.nrows].sum())
    assert 1 <= nun <= 64                                                               # a flipped variable breaks only the rows that read it
    with pytest.raises(RuntimeError): G.verify_circuit(b"\xc0" + bytes(47), CASES[0][0], SIG)    # identity public key: rejected before synthesis (bls.rs:434)

def test_aggregate_verify_circuit_reference_cases():
    """constraints.rs:378-520: 512 keys (pk1, then 511 copies of pk2); bitmap {0, 1} set -> true with count 2; all set -> false"""
    pk1 = PK
    pk2 = bytes.fromhex("b301803f8b5ac4a1133581fc676dfedc60d891dd5fa99028805e5ea5b08d3491af75d0707adab3b70c6a6a580217bf81")
    sig = bytes.fromhex("912c3615f69575407db9392eb21fee18fff797eeb2fbe1816366ca2a08ae574d8824dbfafb4c9eaa1cf61b63c6f9b69911f269b664c42947dd1b53ef1081926c"
                        "1e82bb2a465f927124b08391a5249036146d6f3f1e17ff5f162f779746d830d1")
    msg = bytes.fromhex("56" * 32); shapes = set()
    for bitmap, expect, count in (([1, 1] + [0] * 510, True, 2), ([1] * 512, False, 512)):
        c = G.aggregate_verify_circuit(pk1 + pk2 * 511, bitmap, msg, sig)
        assert c.result == expect and c.count == count and c.first_unsatisfied() == -1
        shapes.add((c.nrows, c.ncols, tuple(c.nnz))); c.free()
    assert len(shapes) == 1
    # native cross-check of the first case: fast_aggregate_verify over the two selected keys (tests/test_cases/fast_aggregate_verify)
    st = C.fast_aggregate_verify(np.frombuffer(pk1 + pk2, np.uint8), 2, np.frombuffer(msg, np.uint8), np.frombuffer(sig, np.uint8), threads=1)
    assert st[0] == 0
