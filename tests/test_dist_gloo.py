"""world_size-2 gloo test (CPU) of the N>1 host logic: contiguous sharding, the all-gather exchange of bitmap shards and
GT partials, and the rank-order fold.  Kernels are replaced by the CPU oracle here (no GPU in this container); the
exchange code is the one bench.py runs over NCCL."""
import os, socket, sys
import numpy as np
import torch, torch.distributed as dist, torch.multiprocessing as mp
from conftest import ROOT

N = 200                      # not a multiple of 64 * world: ragged last shard

def _inputs():
    from oracle import cwrap as C
    rng = np.random.default_rng(12)
    sk = rng.integers(0, 256, size=(N, 32), dtype=np.uint8); sk[:, 31] &= 0x3f; sk[:, 0] |= 1
    msgs = [rng.bytes(32) for _ in range(N)]
    pk = C.sk_to_pk(sk, threads=8).reshape(N, 48); sig = C.sign(sk, msgs, threads=8)[0].reshape(N, 96).copy()
    for i in (5, 70, 130, 199): sig[i] = sig[(i + 1) % N]                  # valid points, wrong signatures
    return pk, msgs, sig

def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cwrap as C
    from bls_verify_gadget_b200.dist import shard_range, shard_words, exchange
    pk, msgs, sig = _inputs()
    lo, hi = shard_range(N, world, rank); words = shard_words(N, world)
    st, gt = C.verify(pk[lo:hi].reshape(-1), msgs[lo:hi], sig[lo:hi].reshape(-1), want_gt=True, threads=2) if hi > lo else (np.zeros(0, np.uint8), C.pairing_gt(b"", b""))
    bits = np.zeros(words * 64, dtype=np.uint8); bits[:hi - lo] = st == 0
    bm = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int64).copy())
    def fold(parts):
        acc = parts[:576].numpy()
        for r in range(1, world): acc = C.gt_mul(acc, parts[576 * r:576 * (r + 1)].numpy())
        return torch.from_numpy(np.asarray(acc).copy())
    full_bm, gt_all = exchange(bm, torch.from_numpy(gt.copy()), fold)
    q.put((rank, lo, hi, full_bm.numpy().tobytes(), gt_all.numpy().tobytes()))
    dist.destroy_process_group()

def test_two_rank_exchange_matches_single_process():
    from oracle import cwrap as C
    from bls_verify_gadget_b200.dist import shard_range, shard_words
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in ps: p.join(60)
    pk, msgs, sig = _inputs()
    st, gt = C.verify(pk.reshape(-1), msgs, sig.reshape(-1), want_gt=True, threads=8)
    assert (st != 0).sum() == 4
    words = shard_words(N, world)
    assert [r[1:3] for r in res] == [shard_range(N, world, r) for r in range(world)] == [(0, 128), (128, 200)]
    for rank, lo, hi, bm, gt_all in res:
        bits = np.unpackbits(np.frombuffer(bm, dtype=np.uint8), bitorder="little")
        got = np.concatenate([bits[r * words * 64: r * words * 64 + (shard_range(N, world, r)[1] - shard_range(N, world, r)[0])] for r in range(world)])
        assert np.array_equal(got, (st == 0).astype(np.uint8))             # every rank holds the whole bitmap
        assert gt_all == gt.tobytes()                                      # and the same folded GT accumulator


# ---- R1CS check: assignments sharded across ranks, one all-gather of the per-assignment flags
def _r1cs_case():
    from bls_verify_gadget_b200 import synth
    nrows, ncols, nwit = 96, 130, 11
    mats, nfree = synth.r1cs_system(nrows, ncols, seed=0x99)
    rng = np.random.default_rng(4); P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    zs = []
    for w in range(nwit):
        z = [1] + [int.from_bytes(rng.bytes(47), "little") for _ in range(nfree - 1)] + [0] * nrows
        for i in range(nrows):
            d = []
            for m in range(2):
                rp, col, cf = mats[m]
                d.append(sum(int.from_bytes(cf[48 * k:48 * k + 48].tobytes(), "little") * z[col[k]] for k in range(int(rp[i]), int(rp[i + 1]))) % P)
            z[nfree + i] = d[0] * d[1] % P
        if w in (2, 9): z[nfree + 7 * w] = (z[nfree + 7 * w] + 1) % P
        zs.append(b"".join(v.to_bytes(48, "little") for v in z))
    return mats, nrows, ncols, nwit, np.frombuffer(b"".join(zs), np.uint8)

def _r1cs_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cwrap as C
    from bls_verify_gadget_b200.dist import witness_shard, gather_flags
    mats, nrows, ncols, nwit, z = _r1cs_case()
    lo, hi = witness_shard(nwit, world, rank)
    bits, allsat = C.r1cs_check([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols, z[48 * ncols * lo:48 * ncols * hi], hi - lo, threads=2)
    flags = gather_flags(torch.from_numpy(np.asarray(allsat, dtype=np.uint8).copy()), nwit)
    q.put((rank, lo, hi, flags.numpy().tobytes()))
    dist.destroy_process_group()

def test_two_rank_r1cs_flags_match_single_process():
    from oracle import cwrap as C
    from bls_verify_gadget_b200.dist import witness_shard
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    ps = [ctx.Process(target=_r1cs_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in ps: p.join(60)
    mats, nrows, ncols, nwit, z = _r1cs_case()
    bits, allsat = C.r1cs_check([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols, z, nwit, threads=4)
    assert [r[1:3] for r in res] == [witness_shard(nwit, world, r) for r in range(world)] == [(0, 6), (6, 11)]
    want = np.asarray(allsat, dtype=np.uint8).tobytes()
    assert list(allsat) == [1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1]
    for rank, lo, hi, flags in res: assert flags == want
