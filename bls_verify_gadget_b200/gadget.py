"""Host-side circuit builder: ctypes binding of libblsgadget.so (gadget/*.hpp), the C++ mirror of the reference's gadget
code -- hash_to_g2_with_cons (src/hasher.rs:727-740) and BlsSignatureVerifyGadget::verify (src/constraints.rs:90-128).

It produces what the K8 kernel consumes: the R1CS matrices (A, B, C) as CSR with canonical 48-byte little-endian
coefficients and the assignment z = [1, instance.., witness..], in exactly the layout of `Context.r1cs_load` /
`Context.r1cs_check`.  The satisfaction check itself runs on the GPU (csrc/r1cs.cuh); nothing here computes it for the
product path.
"""
import ctypes, os, subprocess
import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "libblsgadget.so")
_SRC = [os.path.join(_PKG, "gadget", f) for f in ("gadget_api.cpp", "r1cs_core.hpp", "r1cs_hasher.hpp", "r1cs_verify.hpp")] + \
       [os.path.join(_PKG, "csrc", f) for f in ("fp.cuh", "fp2.cuh", "wide.cuh", "tower.cuh", "curve.cuh", "h2c.cuh", "pairing.cuh", "stages.cuh", "consts.cuh")]

def build(force=False):
    stale = not os.path.exists(SO_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in _SRC)
    if force or stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(_PKG, "csrc"), "-o", SO_PATH, _SRC[0]])
    return SO_PATH

_lib = None
def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.blsgadget_first_unsatisfied.restype = ctypes.c_long
        _lib = L
    return _lib

def _u8(b): return np.frombuffer(bytes(b) + b"\0", dtype=np.uint8)[:len(b)].copy()
def _p(a): return a.ctypes.data_as(ctypes.c_void_p)

class Circuit:
    """A synthesised constraint system with its assignment.  `result` is the value of the Boolean the verify gadget returns
    (None for the hash circuit), `output` the compressed H(m) of the hash circuit, `gt` the 576-byte GT element of the verify
    circuit."""
    def __init__(self, handle, result=None, output=None, gt=None):
        if handle < 0: raise RuntimeError({-1: "synthesis failed", -2: "public key or signature does not decode"}.get(handle, f"error {handle}"))
        self.h = handle; self.result = result; self.output = output; self.gt = gt
        nr = ctypes.c_uint64(); nc = ctypes.c_uint64(); ni = ctypes.c_uint64(); nnz = (ctypes.c_uint64 * 3)()
        lib().blsgadget_shape(handle, ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(ni), nnz)
        self.nrows, self.ncols, self.ninstance, self.nnz = nr.value, nc.value, ni.value, list(nnz)
    def matrices(self):
        """[(rowptr u64[nrows+1], col u32[nnz], coeff48 u8[nnz*48])] for A, B, C"""
        rp = [np.empty(self.nrows + 1, np.uint64) for _ in range(3)]; cl = [np.empty(max(n, 1), np.uint32) for n in self.nnz]; cf = [np.empty(max(n, 1) * 48, np.uint8) for n in self.nnz]
        P3 = ctypes.c_void_p * 3
        lib().blsgadget_export(self.h, P3(*[_p(x) for x in rp]), P3(*[_p(x) for x in cl]), P3(*[_p(x) for x in cf]), None)
        return [(rp[m], cl[m][:self.nnz[m]], cf[m][:48 * self.nnz[m]]) for m in range(3)]
    def assignment(self):
        """z as ncols * 48 bytes, canonical little-endian"""
        z = np.empty(self.ncols * 48, np.uint8); lib().blsgadget_export(self.h, None, None, None, _p(z)); return z
    def first_unsatisfied(self): return int(lib().blsgadget_first_unsatisfied(self.h))
    def free(self):
        if self.h >= 0: lib().blsgadget_free(self.h); self.h = -1
    def __del__(self):
        try: self.free()
        except Exception: pass

def hash_to_g2_circuit(message, message_is_instance=False):
    m = _u8(message); out = np.zeros(96, np.uint8)
    h = lib().blsgadget_hash_to_g2(_p(m), ctypes.c_size_t(len(message)), int(message_is_instance), _p(out))
    return Circuit(h, output=out.tobytes())

def verify_circuit(pk48, message, sig96):
    pk = _u8(pk48); m = _u8(message); sg = _u8(sig96); res = ctypes.c_int(-1); gt = np.zeros(576, np.uint8)
    h = lib().blsgadget_verify(_p(pk), _p(m), ctypes.c_size_t(len(message)), _p(sg), ctypes.byref(res), _p(gt))
    return Circuit(h, result=bool(res.value) if h >= 0 else None, gt=gt.tobytes())

def aggregate_verify_circuit(pks48, bitmap, message, sig96):
    """BlsSignatureVerifyGadget::aggregate_verify (constraints.rs:153-167): `pks48` n*48 bytes, `bitmap` n truthy/falsy values.
    The returned circuit carries `result` and `count` (the UInt32 participant count the gadget outputs)."""
    pk = _u8(pks48); n = len(pk) // 48; bm = np.array([1 if b else 0 for b in bitmap], dtype=np.uint8); assert len(bm) == n
    m = _u8(message); sg = _u8(sig96); res = ctypes.c_int(-1); cnt = ctypes.c_uint32(0)
    h = lib().blsgadget_aggregate_verify(_p(pk), ctypes.c_size_t(n), _p(bm), _p(m), ctypes.c_size_t(len(message)), _p(sg), ctypes.byref(res), ctypes.byref(cnt))
    c = Circuit(h, result=bool(res.value) if h >= 0 else None); c.count = cnt.value
    return c

def _export_program(h, extra):
    nv = ctypes.c_uint64(); ncol = ctypes.c_uint64(); nl = ctypes.c_uint64(); nt = ctypes.c_uint64(); lib().blsgadget_program_shape(h, ctypes.byref(nv), ctypes.byref(ncol), ctypes.byref(nl), ctypes.byref(nt))
    rules = np.empty(16 * ncol.value, np.uint8); lp = np.empty(nl.value + 1, np.uint64); lc = np.empty(max(nt.value, 1), np.uint32); cf = np.empty(48 * max(nt.value, 1), np.uint8)
    lib().blsgadget_program_export(h, _p(rules), _p(lp), _p(lc), _p(cf))
    lib().blsgadget_program_levels.restype = ctypes.c_long
    nlev = lib().blsgadget_program_levels(h, None, None, ctypes.c_uint64(0))
    order = np.empty(ncol.value, np.uint32); level_ptr = np.empty(nlev + 1, np.uint64)
    assert lib().blsgadget_program_levels(h, _p(order), _p(level_ptr), ctypes.c_uint64(nlev + 1)) == nlev
    lib().blsgadget_free(h)
    d = {"rules16": rules, "lc_ptr": lp, "lc_col": lc[:nt.value], "lc_coef48": cf[:48 * nt.value], "nvars": nv.value, "ncols": ncol.value, "order": order, "level_ptr": level_ptr}
    d.update(extra); return d

def aggregate_verify_program(pks48, bitmap, msg, sig96):
    """the witness program of the aggregate_verify circuit (constraints.rs:153-191) for len(pks48) / 48 keys and messages of len(msg) bytes,
    recorded on the given sample input; for Context.witness_load(program) and Context.witness_gen_aggregate / witness_check_aggregate"""
    pk = _u8(pks48); n = len(pk) // 48; bm = np.array([1 if b else 0 for b in bitmap], dtype=np.uint8); assert len(bm) == n
    m = _u8(msg); sg = _u8(sig96)
    h = lib().blsgadget_aggregate_verify_program(_p(pk), ctypes.c_size_t(n), _p(bm), _p(m) if len(msg) else None, ctypes.c_size_t(len(msg)), _p(sg))
    if h < 0: raise RuntimeError(f"blsgadget_aggregate_verify_program failed ({h})")
    return _export_program(h, {"msg_len": len(msg), "nkeys": n})

def verify_program(pk48, msg, sig96):
    """the witness program of the verify circuit for messages of len(msg) bytes (otherwise input-independent; recorded on the given sample triple):
    dict(rules16 u8[ncols*16] (ncols = nvars + scratch columns), lc_ptr u64[nlc+1], lc_col u32[nterms], lc_coef48 u8[nterms*48], nvars, order u32[nvars] = variables
    sorted by dependency level, level_ptr u64[nlevels+1]) for Context.witness_load"""
    pk = _u8(pk48); m = _u8(msg); sg = _u8(sig96)
    h = lib().blsgadget_verify_program(_p(pk), _p(m) if len(msg) else None, ctypes.c_size_t(len(msg)), _p(sg))
    if h < 0: raise RuntimeError(f"blsgadget_verify_program failed ({h})")
    return _export_program(h, {"msg_len": len(msg), "nkeys": 0})

def verify_witnesses(triples, threads=None, ncols=None):
    """assignments of the verify circuit for a list of (pk48, msg, sig96) (all with the same message length), synthesised on
    `threads` host threads in witness-only mode (no matrices: they do not depend on the inputs); returns
    (z [n, ncols*48] u8, results [n] bool).  `ncols` comes from one full synthesis (done here when not given)."""
    from concurrent.futures import ThreadPoolExecutor
    if ncols is None:
        c = verify_circuit(*triples[0]); ncols = c.ncols; c.free()
    z = np.empty((len(triples), ncols * 48), np.uint8); res = np.zeros(len(triples), bool)
    def one(i):
        pk, m, sg = (_u8(x) for x in triples[i]); r = ctypes.c_int(-1)
        rc = lib().blsgadget_verify_assignment(_p(pk), _p(m), ctypes.c_size_t(len(triples[i][1])), _p(sg), _p(z[i]), ctypes.c_size_t(ncols), ctypes.byref(r))
        if rc != 0: raise RuntimeError(f"blsgadget_verify_assignment failed ({rc})")
        res[i] = bool(r.value)
    with ThreadPoolExecutor(max_workers=threads or os.cpu_count() or 1) as ex: list(ex.map(one, range(len(triples))))
    return z, res


def write_r1cs_file(path, nrows, ncols, ninstance, mats, assignments):
    """the BLSR1CS1 exchange format of blsgpu_r1cs_load_file / blsgpu_r1cs_check_file (layout: csrc/r1cs.cuh); the Rust exporter
    rust/examples/export_r1cs.rs writes the same bytes from arkworks' cs.to_matrices().  mats = [(rowptr u64, col u32, coeff48 u8)] x 3,
    assignments = iterable of ncols * 48-byte vectors."""
    import struct
    zs = [np.ascontiguousarray(z, dtype=np.uint8).reshape(-1) for z in assignments]
    with open(path, "wb") as f:
        f.write(b"BLSR1CS1" + struct.pack("<II", 1, 48) + struct.pack("<QQQ", nrows, ncols, ninstance) + struct.pack("<QQQ", *[int(m[0][-1]) for m in mats]) + struct.pack("<Q", len(zs)))
        for rp, cl, cf in mats:
            nnz = int(rp[-1]); assert len(rp) == nrows + 1 and len(cl) >= nnz and len(cf) >= 48 * nnz
            f.write(np.ascontiguousarray(rp, dtype="<u8").tobytes()); f.write(np.ascontiguousarray(cl[:nnz], dtype="<u4").tobytes())
            if nnz % 2: f.write(b"\0" * 4)
            f.write(np.ascontiguousarray(cf[:48 * nnz], dtype=np.uint8).tobytes())
        for z in zs:
            assert z.size == ncols * 48; f.write(z.tobytes())
