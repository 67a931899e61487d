"""Host-side mirror of the reference's public surface (src/bls.rs) on top of the C ABI -- same names, argument
meaning and error behaviour, so the parity tests read like the reference's own tests (tests/tests.rs).

  Parameters / PrivateKey / PublicKey / Signature      src/bls.rs:25-357  (value types; bytes are the reference's
                                                       serialisations: pk 48 B, sig 96 B ZCash compressed, sk 32 B LE)
  PublicKey.aggregate / Signature.aggregate            src/bls.rs:183-195, 288-300  (None on empty input)
  BLS.setup / keygen / sign / verify                   src/bls.rs:391-458
  hash_to_g2                                           src/bls.rs:477-493
  BLSError {InvalidSecretKey, InvalidPublicKey, InvalidSignature}   src/bls.rs:359-377

Every operation runs on the GPU through libblsgpu (no CPU arithmetic here).  The reference is one-item-at-a-time;
the *_batch methods are the batch-first form of the same calls and are what the benchmark drives.
The reference's host language (Rust) is not available in this image; INTEGRATION.md holds the Rust shim."""
import os
import numpy as np
from ._lib import Context, BlsGpuError

R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
G1_GENERATOR_COMPRESSED = bytes.fromhex("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")

class BLSError(Exception):
    """src/bls.rs:359-377"""
    InvalidSecretKey = "InvalidSecretKey"; InvalidPublicKey = "InvalidPublicKey"; InvalidSignature = "InvalidSignature"
    def __init__(self, kind): super().__init__(kind); self.kind = kind
class SerializationError(ValueError):
    """ark_serialize::SerializationError returned by TryFrom<&[u8]> (src/bls.rs:219-223, 316-320)"""

_ctx = None
def default_context():
    global _ctx
    if _ctx is None: _ctx = Context(int(os.environ.get("LOCAL_RANK", "-1")) if "LOCAL_RANK" in os.environ else -1)
    return _ctx

def _from_hex(s):                       # malformed hex panics in the reference (.unwrap(), src/bls.rs:83,230,327): raise ValueError
    return bytes.fromhex(s[2:] if s.startswith("0x") else s)

class Parameters:
    """src/bls.rs:26-36: holds the G1 generator."""
    def __init__(self): self.g1_generator = G1_GENERATOR_COMPRESSED

class PrivateKey:
    """src/bls.rs:53-121: Fr element, 32-byte little-endian canonical."""
    def __init__(self, sk_le=bytes(32)): self.private_key = bytes(sk_le)
    @classmethod
    def try_from(cls, v):
        b = _from_hex(v) if isinstance(v, str) else bytes(v)
        if len(b) < 32: raise SerializationError("short")
        b = b[:32]
        if int.from_bytes(b, "little") >= R_ORDER: raise SerializationError("not a canonical Fr element")
        return cls(b)
    def to_bytes(self): return self.private_key
    def to_hex(self): return self.private_key.hex()
    def __eq__(self, o): return isinstance(o, PrivateKey) and self.private_key == o.private_key

class PublicKey:
    """src/bls.rs:136-260: a G1 point; kept as its 48-byte compressed encoding."""
    def __init__(self, pk48=bytes([0xc0]) + bytes(47)): self.public_key = bytes(pk48)      # default = identity (bls.rs:140-146)
    @classmethod
    def try_from(cls, v, ctx=None):
        b = _from_hex(v) if isinstance(v, str) else bytes(v)
        if len(b) < 48: raise SerializationError("short")
        b = b[:48]                                  # the ark reader consumes exactly 48 bytes
        code = (ctx or default_context()).deserialize_g1(b)[0]
        if code > 1: raise SerializationError(f"invalid G1 encoding (code {code})")
        return cls(bytes([0xc0]) + bytes(47) if code == 1 else b)
    @classmethod
    def from_private(cls, sk, ctx=None):            # From<&PrivateKey>, src/bls.rs:210-216
        pk, _ = (ctx or default_context()).sk_to_pk(sk.private_key); return cls(pk.tobytes())
    @staticmethod
    def aggregate(public_keys, ctx=None):           # src/bls.rs:183-195
        if len(public_keys) == 0: return None
        out, st = (ctx or default_context()).g1_aggregate(b"".join(p.public_key for p in public_keys), [0, len(public_keys)])
        if st[0] != 0: raise SerializationError("aggregate member failed to decode")
        return PublicKey(out.tobytes())
    def to_bytes(self): return self.public_key
    def to_hex(self): return self.public_key.hex()
    def to_uncompressed(self, ctx=None):            # ark-serialize's serialize_uncompressed: 96 bytes x || y
        out, st = (ctx or default_context()).g1_uncompress(self.public_key); return out.tobytes()
    @classmethod
    def try_from_uncompressed(cls, b96, ctx=None):  # deserialize_uncompressed with Validate::Yes
        if len(b96) < 96: raise SerializationError("short")
        out, st = (ctx or default_context()).g1_compress(bytes(b96[:96]))
        if st[0] > 1: raise SerializationError(f"invalid uncompressed G1 encoding (code {st[0]})")
        return cls(out.tobytes())
    def __eq__(self, o): return isinstance(o, PublicKey) and self.public_key == o.public_key
    def __hash__(self): raise NotImplementedError("unimplemented!() in the reference, src/bls.rs:176-180")

class Signature:
    """src/bls.rs:263-357: a G2 point; kept as its 96-byte compressed encoding."""
    def __init__(self, sig96=bytes([0xc0]) + bytes(95)): self.sig = bytes(sig96)
    @classmethod
    def try_from(cls, v, ctx=None):
        b = _from_hex(v) if isinstance(v, str) else bytes(v)
        if len(b) < 96: raise SerializationError("short")
        b = b[:96]
        code = (ctx or default_context()).deserialize_g2(b)[0]
        if code > 1: raise SerializationError(f"invalid G2 encoding (code {code})")
        return cls(bytes([0xc0]) + bytes(95) if code == 1 else b)
    @staticmethod
    def aggregate(signatures, ctx=None):            # src/bls.rs:288-300
        if len(signatures) == 0: return None
        out, st = (ctx or default_context()).g2_aggregate(b"".join(s.sig for s in signatures), [0, len(signatures)])
        if st[0] != 0: raise SerializationError("aggregate member failed to decode")
        return Signature(out.tobytes())
    def to_bytes(self): return self.sig
    def to_hex(self): return self.sig.hex()
    def to_uncompressed(self, ctx=None):            # 192 bytes x.c1 || x.c0 || y.c1 || y.c0
        out, st = (ctx or default_context()).g2_uncompress(self.sig); return out.tobytes()
    @classmethod
    def try_from_uncompressed(cls, b192, ctx=None):
        if len(b192) < 192: raise SerializationError("short")
        out, st = (ctx or default_context()).g2_compress(bytes(b192[:192]))
        if st[0] > 1: raise SerializationError(f"invalid uncompressed G2 encoding (code {st[0]})")
        return cls(out.tobytes())
    def __eq__(self, o): return isinstance(o, Signature) and self.sig == o.sig

def hash_to_g2(message, ctx=None):
    """src/bls.rs:477-493 -> the G2 point, as a Signature-typed value like the reference's test does (bls.rs:646-650)."""
    return Signature((ctx or default_context()).hash_to_g2([bytes(message)]).tobytes())

class BLS:
    """impl SignatureScheme for BLS<P>, src/bls.rs:379-475."""
    @staticmethod
    def setup(rng=None): return Parameters()                                        # bls.rs:391-393
    @staticmethod
    def keygen(parameters, rng, ctx=None):                                          # bls.rs:395-409: Fr::rand then pk = g1 * sk
        sk = PrivateKey((int.from_bytes(rng.bytes(48), "little") % R_ORDER).to_bytes(32, "little"))
        return PublicKey.from_private(sk, ctx), sk
    @staticmethod
    def sign(parameters, sk, message, rng=None, ctx=None):                          # bls.rs:411-425
        sig, st = (ctx or default_context()).sign(sk.private_key, [bytes(message)])
        if st[0] == 5: raise BLSError(BLSError.InvalidSecretKey)
        return Signature(sig.tobytes())
    @staticmethod
    def verify(parameters, pk, message, signature, ctx=None):                       # bls.rs:427-458 -> Ok(bool) | Err(BLSError)
        st = (ctx or default_context()).verify(pk.public_key, [bytes(message)], signature.sig)[0]
        if st == 2: raise BLSError(BLSError.InvalidPublicKey)
        if st == 3: raise BLSError(BLSError.InvalidSignature)
        return st == 0
    @staticmethod
    def randomize_public_key(*a): raise NotImplementedError("unimplemented!() in the reference, src/bls.rs:460-466")
    @staticmethod
    def randomize_signature(*a): raise NotImplementedError("unimplemented!() in the reference, src/bls.rs:468-474")
    # ---- the Eth2 calls the reference's harness composes from aggregate + verify (tests/tests.rs:297-334) or does not vendor (tests/readme.md:4-7)
    @staticmethod
    def fast_aggregate_verify(parameters, public_keys, message, signature, ctx=None):           # PublicKey::aggregate then verify; None / Err collapse to False (tests.rs:312-316, 328)
        agg = PublicKey.aggregate(public_keys, ctx)
        if agg is None: return False
        try: return BLS.verify(parameters, agg, message, signature, ctx)
        except BLSError: return False
    @staticmethod
    def aggregate_verify(parameters, public_keys, messages, signature, ctx=None):               # distinct messages: one signature over the (pk_i, msg_i) pairs
        if len(public_keys) != len(messages): raise ValueError("one message per public key")
        st = (ctx or default_context()).aggregate_verify(b"".join(p.public_key for p in public_keys), [bytes(m) for m in messages], [0, len(public_keys)], signature.sig)[0]
        return st == 0
    # ---- batch-first forms (no reference counterpart: the reference verifies one triple per call)
    @staticmethod
    def verify_batch(pk48, msgs, sig96, ctx=None, **kw): return (ctx or default_context()).verify(pk48, msgs, sig96, **kw)
    @staticmethod
    def fast_aggregate_verify_batch(pks48, k, msg32, sig96, bitmap=None, ctx=None, **kw):
        return (ctx or default_context()).fast_aggregate_verify(pks48, k, msg32, sig96, bitmap=bitmap, **kw)
