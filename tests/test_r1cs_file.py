"""The BLSR1CS1 exchange format (blsgpu_r1cs_load_file / blsgpu_r1cs_check_file): how matrices and assignments produced by arkworks
(rust/examples/export_r1cs.rs, after /root/reference/src/constraints.rs:335-367) reach the GPU kernel.  No Rust toolchain exists here, so
the file is written by the in-repo builder's exporter (gadget.write_r1cs_file, same bytes) and must give, through the file, exactly the
bits of the direct blsgpu_r1cs_load / blsgpu_r1cs_check path and of the CPU oracle."""
import os, struct, tempfile
import numpy as np
import pytest
from bls_verify_gadget_b200 import gadget as G

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab

def _small_system():
    """x * y = t, (t + 3) * 1 = out, b * (1 - b) = 0 over columns [1, x, y, t, out, b]"""
    def f(v): return (v % P).to_bytes(48, "little")
    A = [[(1, 1)], [(3, 1), (0, 3)], [(5, 1)]]; B = [[(2, 1)], [(0, 1)], [(0, 1), (5, P - 1)]]; Cm = [[(3, 1)], [(4, 1)], []]
    def csr(M):
        rp = [0]; cl = []; cf = b""
        for row in M:
            for c, v in row: cl.append(c); cf += f(v)
            rp.append(len(cl))
        return np.array(rp, np.uint64), np.array(cl, np.uint32), np.frombuffer(cf + b"\0", np.uint8)[:len(cf)]
    z_ok = b"".join(f(v) for v in [1, 6, 7, 42, 45, 1]); z_bad = b"".join(f(v) for v in [1, 6, 7, 42, 46, 2])
    return [csr(A), csr(B), csr(Cm)], [np.frombuffer(z_ok, np.uint8), np.frombuffer(z_bad, np.uint8)]

def test_file_layout_matches_the_documented_header():
    mats, zs = _small_system()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.r1cs"); G.write_r1cs_file(path, 3, 6, 1, mats, zs)
        raw = open(path, "rb").read()
    assert raw[:8] == b"BLSR1CS1" and struct.unpack_from("<II", raw, 8) == (1, 48)
    nrows, ncols, ninst, na, nb, nc, nwit = struct.unpack_from("<7Q", raw, 16)
    assert (nrows, ncols, ninst, na, nb, nc, nwit) == (3, 6, 1, 4, 4, 2, 2)
    off = 72
    for nnz in (na, nb, nc): off += 8 * (nrows + 1) + ((4 * nnz + 7) & ~7) + 48 * nnz
    assert len(raw) == off + nwit * ncols * 48 and raw[off:off + 48] == (1).to_bytes(48, "little")

@pytest.mark.gpu
def test_system_through_the_file_equals_direct_load():
    from bls_verify_gadget_b200 import Context
    from bls_verify_gadget_b200._lib import BlsGpuError
    from oracle import cwrap as C
    ctx = Context(0)
    mats, zs = _small_system()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s.r1cs"); G.write_r1cs_file(path, 3, 6, 1, mats, zs)
        h, shape = ctx.r1cs_load_file(path); assert shape == {"nrows": 3, "ncols": 6, "ninstance": 1, "nwit": 2}
        bits, allsat = ctx.r1cs_check_file(h, path, 0, 2, 3); ctx.r1cs_free(h)
        assert [int(b) for b in bits[:, 0]] == [0b111, 0b001] and list(allsat) == [1, 0]
        open(os.path.join(d, "bad.r1cs"), "wb").write(b"NOTR1CS1" + bytes(100))
        with pytest.raises(BlsGpuError): ctx.r1cs_load_file(os.path.join(d, "bad.r1cs"))
        with pytest.raises(BlsGpuError): ctx.r1cs_load_file(os.path.join(d, "missing.r1cs"))
        # the hash-to-G2 circuit of the builder (hasher.rs:727-740; ~640 k rows) with a satisfying and a perturbed assignment
        c = G.hash_to_g2_circuit(b"file round trip"); m = c.matrices(); z = c.assignment()
        bad = z.reshape(c.ncols, 48).copy(); bad[c.ncols // 3, 0] ^= 1
        big = os.path.join(d, "h2c.r1cs"); G.write_r1cs_file(big, c.nrows, c.ncols, c.ninstance, m, [z, bad.reshape(-1), z])
        h, shape = ctx.r1cs_load_file(big); assert (shape["nrows"], shape["ncols"], shape["nwit"]) == (c.nrows, c.ncols, 3)
        fbits, fall = ctx.r1cs_check_file(h, big, 0, 3, c.nrows); tail_bits, tail_all = ctx.r1cs_check_file(h, big, 2, 1, c.nrows); ctx.r1cs_free(h)
        h2 = ctx.r1cs_load([x[0] for x in m], [x[1] for x in m], [x[2] for x in m], c.nrows, c.ncols)
        dbits, dall = ctx.r1cs_check(h2, np.concatenate([z, bad.reshape(-1), z]), 3, c.nrows); ctx.r1cs_free(h2)
        obits, oall = C.r1cs_check([x[0] for x in m], [x[1] for x in m], [x[2] for x in m], c.nrows, c.ncols, np.concatenate([z, bad.reshape(-1), z]), 3, threads=C.hw_threads())
        assert np.array_equal(fbits, dbits) and np.array_equal(fbits, obits) and list(fall) == list(dall) == list(oall) == [1, 0, 1]
        assert np.array_equal(tail_bits[0], fbits[2]) and list(tail_all) == [1]
    ctx.close()
