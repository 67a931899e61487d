"""bls_verify_gadget_b200 -- B200-native (sm_100a) BLS12-381 batch verification / hash-to-G2 / aggregation /
R1CS-check engine: a drop-in for the hot path of lightec-xyz/bls-verify-gadget (src/bls.rs, src/hasher.rs,
the satisfaction check of the src/constraints.rs circuit).  The product is csrc/ (CUDA kernels + C ABI,
include/blsgpu.h); this package is the thin host-side mirror of the reference's API used by tests and bench."""
from ._lib import Context, BlsGpuError, build, lib, SO_PATH, EXPORTS
from .bls import BLS, Parameters, PrivateKey, PublicKey, Signature, BLSError, hash_to_g2, default_context
