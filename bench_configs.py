#!/usr/bin/env python3
"""Secondary measurements for the other BASELINE.json configs (bench.py is the driver's contract for configs[1]).
One JSON line per config, single GPU, device-resident inputs, CUDA events on the launching stream; each line carries the
algorithmic work figure of SURVEY 8(d) and a CPU-oracle rate on a bounded sample.

  cfg2r random-linear-combination batch check of 2^20 valid triples (the additional fast path of SURVEY 8(f)-3, not BLS::verify's
        per-item semantics: one Boolean per batch)
  cfg4  hash-to-G2 of 2^22 synthetic 32-byte messages
  cfg3a sync-committee fast_aggregate_verify over a resident pool of pre-decoded keys (committees = index lists)
  cfg3b sync-committee fast_aggregate_verify: 512 compressed keys per committee (decode + subgroup check included)
  cfg5  R1CS satisfaction check of a synthetic verify-shaped system over 512 witnesses (one GPU's share of 4096)
  cfg5r R1CS satisfaction check of the REAL verify circuit (constraints.rs:90-128, built by bls_verify_gadget_b200/gadget):
        714 k rows, assignments synthesised on the host for distinct (pk, msg, sig) triples
"""
import argparse, json, sys, time, os
ROOT = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, ROOT)
import numpy as np

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab

def timed(fn, steps, stream):
    import torch
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

def plant_witness(mats, nfree, nrows, rng):
    """satisfying z for synth.r1cs_system: free variables random, product slot i = (A_i z)(B_i z) (python ints)"""
    z = np.empty(nfree + nrows, dtype=object); z[0] = 1
    z[1:nfree] = [int.from_bytes(rng.bytes(47), "little") for _ in range(nfree - 1)]
    dots = []
    for m in range(2):
        rp, col, cf = mats[m]; c = cf.reshape(-1, 48)
        coeff = np.array([int.from_bytes(c[k].tobytes(), "little") for k in range(len(col))], dtype=object)
        prod = coeff * z[col]
        d = np.add.reduceat(prod, rp[:-1].astype(np.int64)) % P
        dots.append(d)
    z[nfree:] = (dots[0] * dots[1]) % P
    return b"".join(int(v).to_bytes(48, "little") for v in z)

def run_cfg(cfg, ctx, dev, stream, steps=2, scale=1.0, peak=None, thr=None):
    """one secondary configuration on one GPU -> its JSON-able result line (None for an unknown name); used by main() below and by bench.py"""
    import torch
    from bls_verify_gadget_b200 import synth
    from oracle import cwrap as C
    if thr is None: thr = C.hw_threads()
    if peak is None: peak = NOMINAL_IMAD_PEAK
    if cfg == "2r":
        n = int((1 << 20) * scale)
        ctx.set_pointer_mode(False)
        pk, msg, sig, exp = synth.verify_batch_inputs(ctx, n, every=10 ** 12)                  # all valid
        d_pk, d_msg, d_sig = (torch.from_numpy(x).to(dev) for x in (pk, msg, sig))
        seed = torch.arange(16, dtype=torch.uint8, device=dev); ok = torch.zeros(1, dtype=torch.uint8, device=dev); st = torch.empty(n, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.verify_rlc_ptr(d_pk.data_ptr(), d_msg.data_ptr(), None, d_sig.data_ptr(), n, seed.data_ptr(), st.data_ptr(), ok.data_ptr()), steps, stream)
        assert int(ok.item()) == 1 and int(st.sum().item()) == 0
        d_sig2 = d_sig.clone(); d_sig2[96 * 5:96 * 6] = d_sig[96 * 6:96 * 7]                    # one wrong (but valid-point) signature must flip the batch
        ctx.verify_rlc_ptr(d_pk.data_ptr(), d_msg.data_ptr(), None, d_sig2.data_ptr(), n, seed.data_ptr(), st.data_ptr(), ok.data_ptr()); torch.cuda.synchronize()
        assert int(ok.item()) == 0
        work = (1510 + 2650 + 7472 + 1320 + 1950 + 6500) * 300
        line = {"config": "random-linear-combination batch check of %d valid (pk,msg,sig) triples from compressed bytes: one pairing-product equation, one final exponentiation (additional fast path, SURVEY 8(f)-3)" % n,
                "metric": "verifies_per_sec_batch_boolean", "value": n / (ms * 1e-3), "ms": ms,
                "roofline": {"bound": "imad", "algorithmic_mac32_per_unit": work, "achieved_TMAC32s": n * work / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12, "frac": n * work / (ms * 1e-3) / peak,
                             "note": "work figure: SURVEY decode + hash counts, 64-bit G1/G2 scalar products (1,320 + 1,950 products), one-pair Miller loop (6,500)"},
                "cpu_baseline": None}
    elif cfg == "4":
        n = int((1 << 22) * scale)
        msg = torch.from_numpy(synth.fast_random_bytes(32 * n, 0x683263)).to(dev); out = torch.empty(96 * n, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.hash_to_g2_ptr(msg.data_ptr(), None, n, out.data_ptr()), steps, stream)
        ns = 2048; sample = [bytes(msg[32 * i:32 * i + 32].cpu().numpy()) for i in range(ns)]
        t0 = time.perf_counter(); ref = C.hash_to_g2(sample, threads=thr); dt = time.perf_counter() - t0
        assert np.array_equal(out[:96 * ns].cpu().numpy(), ref), "hash-to-G2 differs from the oracle on the sample"
        line = {"config": "hash-to-G2 of %d 32-byte messages (BASELINE configs[3])" % n, "metric": "hash_to_g2_per_sec", "value": n / (ms * 1e-3), "ms": ms,
                "roofline": {"bound": "imad", "algorithmic_mac32_per_unit": 7472 * 300, "achieved_TMAC32s": n * 7472 * 300 / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12,
                             "frac": n * 7472 * 300 / (ms * 1e-3) / peak, "note": "SURVEY work figure (758-bit Fp2 sqrt_ratio); this implementation executes ~5.4k products per hash"},
                "cpu_baseline": {"value": ns / dt, "unit": "hashes/s", "cores": thr, "kind": "port", "sample": f"{ns} messages"}}
    elif cfg == "3b":
        nc = int((1 << 14) * scale); k = 512
        ctx.set_pointer_mode(False)
        pks, msg, sig, _, _ = synth.committees(ctx, nc, k=k, pool=1 << 16)
        d_pks = torch.from_numpy(pks).to(dev); d_msg = torch.from_numpy(msg).to(dev); d_sig = torch.from_numpy(sig).to(dev)
        st = torch.empty(nc, dtype=torch.uint8, device=dev); agg = torch.empty(48 * nc, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.fast_aggregate_verify_ptr(d_pks.data_ptr(), None, k, d_msg.data_ptr(), d_sig.data_ptr(), nc, st.data_ptr(), agg.data_ptr()), steps, stream)
        assert int(st.sum().item()) == 0, "a committee failed to verify"
        ns = 8; t0 = time.perf_counter(); ost, oagg = C.fast_aggregate_verify(pks[:48 * k * ns], k, msg[:32 * ns], sig[:96 * ns], want_agg=True, threads=thr); dt = time.perf_counter() - t0
        assert not ost.any() and np.array_equal(agg[:48 * ns].cpu().numpy(), oagg)
        work = 810400 * 300
        line = {"config": "%d committees x %d compressed keys: decode + subgroup check + aggregate + pairing check (BASELINE configs[2], variant 3b)" % (nc, k),
                "metric": "committees_per_sec", "value": nc / (ms * 1e-3), "keys_per_sec": nc * k / (ms * 1e-3), "ms": ms,
                "roofline": {"bound": "imad", "algorithmic_mac32_per_unit": work, "achieved_TMAC32s": nc * work / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12, "frac": nc * work / (ms * 1e-3) / peak},
                "cpu_baseline": {"value": ns / dt, "unit": "committees/s", "cores": thr, "kind": "port", "sample": f"{ns} committees"}}
    elif cfg == "3a":
        nc = int((1 << 14) * scale); k = 512; pool = 1 << 16
        ctx.set_pointer_mode(False)
        pks, msg, sig, pool_pk, idx = synth.committees(ctx, nc, k=k, pool=pool)
        t0 = time.perf_counter(); h, codes = ctx.pool_create(pool_pk.reshape(-1)); t_pool = time.perf_counter() - t0
        assert not codes.any()
        d_idx = torch.from_numpy(idx.astype(np.uint32).reshape(-1)).to(dev); d_msg = torch.from_numpy(msg).to(dev); d_sig = torch.from_numpy(sig).to(dev)
        st = torch.empty(nc, dtype=torch.uint8, device=dev); agg = torch.empty(48 * nc, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.pool_fast_aggregate_verify_ptr(h, d_idx.data_ptr(), None, k, d_msg.data_ptr(), d_sig.data_ptr(), nc, st.data_ptr(), agg.data_ptr()), steps, stream)
        assert int(st.sum().item()) == 0, "a committee failed to verify"
        ns = 8; C.set_fast(True); t0 = time.perf_counter(); ost, oagg = C.fast_aggregate_verify(pks[:48 * k * ns], k, msg[:32 * ns], sig[:96 * ns], want_agg=True, threads=thr); dt = time.perf_counter() - t0; C.set_fast(False)
        assert not ost.any() and np.array_equal(agg[:48 * ns].cpu().numpy(), oagg)
        work = 37300 * 300
        line = {"config": "%d committees x %d members as indices into a resident pool of %d pre-decoded keys: aggregate + pairing check (BASELINE configs[2], variant 3a)" % (nc, k, pool),
                "metric": "committees_per_sec", "value": nc / (ms * 1e-3), "keys_per_sec": nc * k / (ms * 1e-3), "ms": ms, "pool_decode_s": t_pool,
                "roofline": {"bound": "imad", "algorithmic_mac32_per_unit": work, "achieved_TMAC32s": nc * work / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12, "frac": nc * work / (ms * 1e-3) / peak},
                "cpu_baseline": {"value": ns / dt, "unit": "committees/s", "cores": thr, "kind": "port", "sample": f"{ns} committees from compressed keys (the CPU port has no resident pool)"}}
        ctx.pool_free(h)
    elif cfg == "5":
        nrows = int((1 << 18) * scale); ncols = nrows + 4096; nwit = 512; nbase = 4
        mats, nfree = synth.r1cs_system(nrows, ncols)
        rng = np.random.default_rng(5)
        base = [np.frombuffer(plant_witness(mats, nfree, nrows, rng), dtype=np.uint8) for _ in range(nbase)]
        z = np.concatenate([base[w % nbase] for w in range(nwit)]).reshape(nwit, ncols, 48).copy()
        bad = list(range(7, nwit, 64))
        for w in bad: z[w, nfree + (w * 977) % nrows, 0] ^= 1                       # one perturbed product slot
        ctx.set_pointer_mode(False)
        h = ctx.r1cs_load([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols)
        dz = torch.from_numpy(z.reshape(-1)).to(dev); words = (nrows + 63) // 64
        bits = torch.zeros(nwit * words, dtype=torch.int64, device=dev); allsat = torch.zeros(nwit, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.r1cs_check_ptr(h, dz.data_ptr(), nwit, bits.data_ptr(), allsat.data_ptr()), steps, stream)
        a = allsat.cpu().numpy(); assert sorted(np.nonzero(a == 0)[0].tolist()) == bad, "all_sat flags differ from the planted pattern"
        ns = 4; t0 = time.perf_counter(); obits, oall = C.r1cs_check([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols, z[:ns].reshape(-1), ns, threads=thr); dt = time.perf_counter() - t0
        assert np.array_equal(bits.cpu().numpy().view(np.uint64).reshape(nwit, words)[:ns], obits)
        nnz = sum(int(m[0][-1]) for m in mats); gen = sum(int(((m[2].reshape(-1, 48)[:, 1:] != 0).any(axis=1) & ~(m[2].reshape(-1, 48) == np.frombuffer((P - 1).to_bytes(48, "little"), dtype=np.uint8)).all(axis=1)).sum()) for m in mats)
        work_survey = (nnz + nrows) * 300; work_exec = (gen + 2 * nrows) * 300
        line = {"config": "R1CS check, synthetic verify-shaped system: %d rows, %d cols, nnz %d (%d general coefficients), %d witnesses (BASELINE configs[4], one GPU's share)" % (nrows, ncols, nnz, gen, nwit),
                "metric": "constraints_checked_per_sec", "value": nrows * nwit / (ms * 1e-3), "witnesses_per_sec": nwit / (ms * 1e-3), "ms": ms,
                "roofline": {"bound": "imad or L2 gather", "algorithmic_mac32_per_unit": work_survey, "achieved_TMAC32s_survey_count": nwit * work_survey / (ms * 1e-3) / 1e12,
                             "executed_TMAC32s": nwit * work_exec / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12,
                             "gather_GBs": nwit * nnz * 48 / (ms * 1e-3) / 1e9, "z_bytes": int(nwit) * ncols * 48},
                "cpu_baseline": {"value": nrows * ns / dt, "unit": "constraints/s", "cores": thr, "kind": "port", "sample": f"{ns} witnesses"},
                "parity": "unpinned vs arkworks (the reference never calls is_satisfied); bit-equal to the CPU oracle on the sample"}
        ctx.r1cs_free(h)
    elif cfg == "5r":
        from bls_verify_gadget_b200 import gadget as G
        nbase = max(2, int(32 * scale)); nwit = max(nbase, int(256 * scale))
        ctx.set_pointer_mode(False)
        pk, msg, sig, exp = synth.verify_batch_inputs(ctx, nbase, every=4, fast=False)
        triples = []
        for i in range(nbase):
            good = exp[i] in (0, 1)                                                    # decodable points; status 1 = pairing false (still a satisfying assignment)
            j = i if good else 0
            triples.append((pk[48 * j:48 * j + 48].tobytes(), msg[32 * i:32 * i + 32].tobytes(), sig[96 * j:96 * j + 96].tobytes()))
        t0 = time.perf_counter(); zb, res = G.verify_witnesses(triples, threads=thr); t_syn = time.perf_counter() - t0
        ost = C.verify(b"".join(t[0] for t in triples), [t[1] for t in triples], b"".join(t[2] for t in triples), threads=thr)
        assert list(res) == [s == 0 for s in ost], "the circuit's output Boolean differs from the native verify"
        c = G.verify_circuit(*triples[0]); mats = c.matrices(); nrows, ncols = c.nrows, c.ncols
        h = ctx.r1cs_load([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols)
        dzb = torch.from_numpy(zb.reshape(-1)).to(dev).reshape(nbase, ncols * 48)
        dz = dzb.repeat((nwit + nbase - 1) // nbase, 1)[:nwit].contiguous()                # nwit assignments resident in HBM (nbase distinct ones)
        bad = list(range(5, nwit, 32))
        for w in bad: dz[w, 48 * ((w * 7919) % ncols)] ^= 1                             # one perturbed variable
        dz = dz.reshape(-1); words = (nrows + 63) // 64
        bits = torch.zeros(nwit * words, dtype=torch.int64, device=dev); allsat = torch.zeros(nwit, dtype=torch.uint8, device=dev)
        ctx.set_pointer_mode(True)
        ms = timed(lambda: ctx.r1cs_check_ptr(h, dz.data_ptr(), nwit, bits.data_ptr(), allsat.data_ptr()), steps, stream)
        a = allsat.cpu().numpy(); assert sorted(np.nonzero(a == 0)[0].tolist()) == bad, "all_sat flags differ from the planted pattern"
        ns = 2; zs = dz[:ns * ncols * 48].cpu().numpy()
        t0 = time.perf_counter(); obits, oall = C.r1cs_check([m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], nrows, ncols, zs, ns, threads=thr); dt = time.perf_counter() - t0
        assert np.array_equal(bits.cpu().numpy().view(np.uint64).reshape(nwit, words)[:ns], obits)
        nnz = sum(c.nnz); one = np.frombuffer((1).to_bytes(48, "little"), np.uint8); m1 = np.frombuffer((P - 1).to_bytes(48, "little"), np.uint8)
        gen = sum(int((~((m[2].reshape(-1, 48) == one).all(axis=1) | (m[2].reshape(-1, 48) == m1).all(axis=1))).sum()) for m in mats)
        work_survey = (nnz + nrows) * 300; work_exec = (gen + 2 * nrows) * 300
        line = {"config": "R1CS check of the verify circuit (constraints.rs:90-128 via the host builder): %d rows, %d cols, nnz %d (%d general coefficients), %d assignments, %d distinct (BASELINE configs[4], part of one GPU's share)" % (nrows, ncols, nnz, gen, nwit, nbase),
                "metric": "constraints_checked_per_sec", "value": nrows * nwit / (ms * 1e-3), "witnesses_per_sec": nwit / (ms * 1e-3), "ms": ms, "host_synthesis_s_per_assignment": t_syn / nbase * min(thr, nbase),
                "roofline": {"bound": "imad or L2 gather", "algorithmic_mac32_per_unit": work_survey, "achieved_TMAC32s_survey_count": nwit * work_survey / (ms * 1e-3) / 1e12,
                             "executed_TMAC32s": nwit * work_exec / (ms * 1e-3) / 1e12, "peak_TMAC32s": peak / 1e12,
                             "gather_GBs": nwit * nnz * 48 / (ms * 1e-3) / 1e9, "z_bytes": int(nwit) * ncols * 48},
                "cpu_baseline": {"value": nrows * ns / dt, "unit": "constraints/s", "cores": thr, "kind": "port", "sample": f"{ns} assignments"},
                "parity": "circuit values pinned (hash-to-G2 KAT bls.rs:645, verify Booleans constraints.rs:326-332, GT = native); matrices unpinned vs arkworks (variable numbering differs); bit-equal to the CPU oracle on the sample"}
        ctx.r1cs_free(h); c.free()
    else: return None
    return line

NOMINAL_IMAD_PEAK = 148 * 32 * 1.965e9      # IMAD.WIDE.U32: one warp instruction per 4 cycles per SM sub-partition = 32 MAC32/clk/SM (profiles/micro)

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--cfg", default="2r,4,3a,3b,5,5r"); ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the configs down for quick runs")
    args = ap.parse_args()
    import torch
    from bls_verify_gadget_b200 import Context
    dev = torch.device("cuda", 0); torch.cuda.set_device(0)
    ctx = Context(0); stream = torch.cuda.current_stream(dev); ctx.set_stream(stream.cuda_stream)
    for cfg in args.cfg.split(","):
        line = run_cfg(cfg, ctx, dev, stream, args.steps, args.scale)
        if line is not None: print(json.dumps(line), flush=True)

if __name__ == "__main__":
    main()
