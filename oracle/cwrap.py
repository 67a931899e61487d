"""ctypes binding of the C++ CPU oracle (oracle/bls_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

def build(force=False):
    src = os.path.join(_HERE, "bls_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO

_lib = None
def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build()); _lib.ora_init()
    return _lib

def _u8(a):
    if isinstance(a, np.ndarray): return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    b = bytes(a)
    return np.frombuffer(b if b else b"\0", dtype=np.uint8)[:len(b)].copy() if b else np.zeros(0, dtype=np.uint8)
def _p(a):
    if a is None: return None
    if a.size == 0: a = np.zeros(1, dtype=a.dtype)
    return a.ctypes.data_as(ctypes.c_void_p)
_sz = ctypes.c_size_t
def hw_threads(): return lib().ora_hw_threads()
def set_fast(on):
    """algorithm set of the oracle: False (default) = the simple forms every parity test uses; True = arkworks-style fast forms,
    for the timed CPU-baseline legs only (both produce identical bytes: tests/test_oracle_golden.py::test_fast_set_matches_simple_set)"""
    lib().ora_set_fast(1 if on else 0)

def msgs_pack(msgs):
    """list of bytes -> (flat uint8 array, uint32 offsets[n+1])"""
    off = np.zeros(len(msgs) + 1, dtype=np.uint32)
    if msgs: off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    flat = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    return flat, off

def fp_mul_raw(a, b):
    a = _u8(a); b = _u8(b); n = a.size // 48; out = np.empty(n * 48, dtype=np.uint8)
    lib().ora_fp_mul_raw(_p(a), _p(b), _p(out), _sz(n)); return out
def expand_xmd(msg, dst, n):
    out = np.empty(n, dtype=np.uint8); m = _u8(msg); d = _u8(dst)
    lib().ora_expand_xmd(_p(m), _sz(len(msg)), _p(d), _sz(len(dst)), _p(out), _sz(n)); return out.tobytes()
def deser_g1(in48):
    a = _u8(in48); n = a.size // 48; st = np.empty(n, dtype=np.uint8); lib().ora_deser_g1(_p(a), _sz(n), _p(st)); return st
def deser_g2(in96):
    a = _u8(in96); n = a.size // 96; st = np.empty(n, dtype=np.uint8); lib().ora_deser_g2(_p(a), _sz(n), _p(st)); return st
def hash_to_g2(msgs, cleared=True, threads=1):
    flat, off = msgs_pack(msgs); out = np.empty(96 * len(msgs), dtype=np.uint8)
    lib().ora_hash_to_g2(_p(flat), _p(off), _sz(len(msgs)), _p(out), int(cleared), threads); return out
def verify(pk48, msgs, sig96, want_gt=False, threads=1):
    pk = _u8(pk48); sg = _u8(sig96); flat, off = msgs_pack(msgs); n = len(msgs)
    st = np.empty(n, dtype=np.uint8); gt = np.empty(576, dtype=np.uint8) if want_gt else None
    lib().ora_verify(_p(pk), _p(flat), _p(off), _p(sg), _sz(n), _p(st), _p(gt), threads)
    return (st, gt) if want_gt else st
def sk_to_pk(sk_le, threads=1):
    sk = _u8(sk_le); n = sk.size // 32; out = np.empty(48 * n, dtype=np.uint8)
    lib().ora_sk_to_pk(_p(sk), _sz(n), _p(out), threads); return out
def sign(sk_le, msgs, threads=1):
    sk = _u8(sk_le); flat, off = msgs_pack(msgs); n = len(msgs); out = np.empty(96 * n, dtype=np.uint8); st = np.empty(n, dtype=np.uint8)
    lib().ora_sign(_p(sk), _p(flat), _p(off), _sz(n), _p(out), _p(st), threads); return out, st
def g1_aggregate(pts48, seg_off, threads=1):
    a = _u8(pts48); seg = np.ascontiguousarray(seg_off, dtype=np.uint32); ns = seg.size - 1
    out = np.empty(48 * ns, dtype=np.uint8); st = np.empty(ns, dtype=np.uint8)
    lib().ora_g1_aggregate(_p(a), _p(seg), _sz(ns), _p(out), _p(st), threads); return out, st
def g2_aggregate(pts96, seg_off, threads=1):
    a = _u8(pts96); seg = np.ascontiguousarray(seg_off, dtype=np.uint32); ns = seg.size - 1
    out = np.empty(96 * ns, dtype=np.uint8); st = np.empty(ns, dtype=np.uint8)
    lib().ora_g2_aggregate(_p(a), _p(seg), _sz(ns), _p(out), _p(st), threads); return out, st
def fast_aggregate_verify(pks48, k, msg32, sig96, bitmap=None, want_agg=False, threads=1):
    pk = _u8(pks48); m = _u8(msg32); sg = _u8(sig96); nc = sg.size // 96
    st = np.empty(nc, dtype=np.uint8); agg = np.empty(48 * nc, dtype=np.uint8) if want_agg else None
    bm = np.ascontiguousarray(bitmap, dtype=np.uint64) if bitmap is not None else None
    lib().ora_fast_aggregate_verify(_p(pk), _p(bm), _sz(k), _p(m), _p(sg), _sz(nc), _p(st), _p(agg), threads)
    return (st, agg) if want_agg else st
def aggregate_verify(pks48, msgs, pair_off, sig96, threads=1):
    """Eth2 AggregateVerify: signature s covers pairs [pair_off[s], pair_off[s+1]) of (pks48[j], msgs[j]); status per signature"""
    pk = _u8(pks48); sg = _u8(sig96); ns = sg.size // 96; flat, off = msgs_pack(msgs); po = np.ascontiguousarray(pair_off, dtype=np.uint32)
    st = np.empty(ns, dtype=np.uint8)
    lib().ora_aggregate_verify(_p(pk), _p(flat), _p(off), _p(po), _p(sg), _sz(ns), _p(st), threads)
    return st
def g1_recode(data, to_uncompressed):
    a = _u8(data); n = a.size // (48 if to_uncompressed else 96); out = np.empty((96 if to_uncompressed else 48) * n, dtype=np.uint8); st = np.empty(n, dtype=np.uint8)
    lib().ora_g1_recode(_p(a), _sz(n), _p(out), _p(st), 1 if to_uncompressed else 0); return out, st
def g2_recode(data, to_uncompressed):
    a = _u8(data); n = a.size // (96 if to_uncompressed else 192); out = np.empty((192 if to_uncompressed else 96) * n, dtype=np.uint8); st = np.empty(n, dtype=np.uint8)
    lib().ora_g2_recode(_p(a), _sz(n), _p(out), _p(st), 1 if to_uncompressed else 0); return out, st
def pairing_gt(g1_48, g2_96):
    a = _u8(g1_48); b = _u8(g2_96); out = np.empty(576, dtype=np.uint8)
    rc = lib().ora_pairing_gt(_p(a), _p(b), _sz(a.size // 48), _p(out))
    if rc: raise ValueError(f"pairing_gt rc={rc}")
    return out
def gt_mul(a, b):
    a = _u8(a); b = _u8(b); out = np.empty(576, dtype=np.uint8)
    if lib().ora_gt_mul(_p(a), _p(b), _p(out)): raise ValueError("bad GT")
    return out
def r1cs_check(rowptr, col, coeff48, nrows, ncols, z48, nwit, threads=1):
    """rowptr/col/coeff48: 3-lists of numpy arrays (uint64, uint32, uint8)."""
    rp = [np.ascontiguousarray(x, dtype=np.uint64) for x in rowptr]; cl = [np.ascontiguousarray(x, dtype=np.uint32) for x in col]
    cf = [np.ascontiguousarray(x, dtype=np.uint8) for x in coeff48]; z = _u8(z48)
    P3 = ctypes.c_void_p * 3
    words = (nrows + 63) // 64; bits = np.zeros(nwit * words, dtype=np.uint64); allsat = np.zeros(nwit, dtype=np.uint8)
    rc = lib().ora_r1cs_check(P3(*[_p(x) for x in rp]), P3(*[_p(x) for x in cl]), P3(*[_p(x) for x in cf]), _sz(nrows), _sz(ncols),
                              _p(z), _sz(nwit), _p(bits), _p(allsat), threads)
    if rc: raise ValueError("non-canonical field element")
    return bits.reshape(nwit, words), allsat
