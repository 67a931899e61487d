#!/usr/bin/env python3
"""Executed-work figures from `ncu --set full --import-source on` reports: per kernel, the thread-level count of executed IMAD.WIDE
instructions (summed over the SASS source page: "Thread Instructions Executed" of every line whose opcode is IMAD.WIDE*), the DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) and both divided by the number of items the launch processed.

  python profiles/tools/ncu_executed.py ITEMS report1.ncu-rep [report2.ncu-rep ...] > profiles/ncu_r02_executed.json
  python profiles/tools/ncu_executed.py --stages ITEMS report.ncu-rep     # one verify step captured whole: sums over ALL launches of each stage's kernels
"""
import csv, io, json, subprocess, sys, collections

def source_tables(rep, kernel=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", kernel] if kernel else []), capture_output=True, text=True).stdout
    kernels = []; cur = None
    for row in csv.reader(io.StringIO(out)):
        if not row: continue
        if row[0] == "Kernel Name": cur = {"name": row[1].split("(")[0], "hdr": None, "rows": []}; kernels.append(cur)
        elif cur is not None and row[0] == "Address": cur["hdr"] = row
        elif cur is not None and cur["hdr"]: cur["rows"].append(row)
    return kernels

def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); H, U = rows[0], rows[1]; res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(H):
            if h in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"):
                d[h] = (r[i], U[i])
        res.append(d)
    return res

def to_bytes(v, u):
    f = float(v.replace(",", "")); return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)

STAGES = {"k_decode_g1": ("k_decode_g1",), "k_decode_g2": ("k_decode_g2",), "k_hash_to_g2": ("k_hash_to_g2", "k_hash_field", "k_hash_map", "k_hash_clear"),
          "k_miller": ("k_miller", "k_miller_lines", "k_miller_accum"), "k_final_exp": ("k_final_exp", "k_final_squarings", "k_final_step")}

def count_ops(table):
    H = table["hdr"]; si = H.index("Source"); ti = H.index("Thread Instructions Executed"); ops = collections.Counter()
    for row in table["rows"]:
        txt = row[si].strip()
        if txt.startswith("@"): txt = txt.split(None, 1)[1] if " " in txt else txt
        op = txt.split()[0] if txt else ""
        ops[op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += int(row[ti].replace(",", "") or 0)
    return ops

def base_name(n): return n.split("(")[0].replace("void ", "").split("<")[0].strip()

def main_stages():
    items = int(sys.argv[2]); rep = sys.argv[3]; result = {}
    raw = raw_metrics(rep)
    unit_ms = {"us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "ns": 1e-6, "nsecond": 1e-6}
    bases = sorted({base_name(r["Kernel Name"][0]) for r in raw})
    for stage, kernels in STAGES.items():
        ops = collections.Counter(); dram = 0.0; ms = 0.0; per_kernel = {}
        for b in bases:
            if b not in kernels: continue
            k_ops = collections.Counter(); ntab = 0
            for t in source_tables(rep, b):
                if t["hdr"] and base_name(t["name"]) == b: k_ops += count_ops(t); ntab += 1
            rs = [r for r in raw if base_name(r["Kernel Name"][0]) == b]
            if ntab and ntab != len(rs):                       # the source page repeats a launch once per view (SASS, PTX+SASS): count every launch once
                assert ntab % len(rs) == 0, (b, ntab, len(rs))
                for o in k_ops: k_ops[o] //= ntab // len(rs)
            k_ms = sum(float(r["gpu__time_duration.sum"][0].replace(",", "")) * unit_ms[r["gpu__time_duration.sum"][1]] for r in rs)
            dram += sum(to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"]) for r in rs); ms += k_ms; ops += k_ops
            per_kernel[b] = {"launches": len(rs), "source_tables": ntab, "imad_wide_per_item": k_ops.get("IMAD.WIDE", 0) / items, "thread_instructions_per_item": sum(k_ops.values()) / items, "ms_under_ncu": k_ms,
                             "registers": sorted({int(r["launch__registers_per_thread"][0]) for r in rs}), "fma_pipe_active_pct": [float(r["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]) for r in rs][:3],
                             "issue_active_pct": [float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]) for r in rs][:3]}
        if not per_kernel: continue
        result[stage] = {"items": items, "imad_wide_per_item": ops.get("IMAD.WIDE", 0) / items, "thread_instructions_per_item": sum(ops.values()) / items, "dram_bytes_per_item": dram / items,
                         "other_fma_pipe_per_item": (ops.get("IMAD", 0) + ops.get("IMAD.HI", 0)) / items, "ms_under_ncu": ms, "top_ops_per_item": {o: c / items for o, c in ops.most_common(8)},
                         "kernels": per_kernel, "report": rep.split("/")[-1], "note": "sum over all launches of the stage's kernels in one verify step"}
    json.dump(result, sys.stdout, indent=1); print()

def main():
    if sys.argv[1] == "--stages": return main_stages()
    items = int(sys.argv[1]); result = {}
    for rep in sys.argv[2:]:
        raw = raw_metrics(rep); last = {}
        for r in raw: last[r["Kernel Name"][0].split("(")[0]] = r                 # the last captured launch of every kernel
        for name, r in last.items():
            tabs = [t for t in source_tables(rep, name) if t["name"] == name and t["hdr"]]
            if not tabs: continue
            k = tabs[-1]
            H = k["hdr"]; si = H.index("Source"); ti = H.index("Thread Instructions Executed")
            ops = collections.Counter()
            for row in k["rows"]:
                txt = row[si].strip()
                if txt.startswith("@"): txt = txt.split(None, 1)[1] if " " in txt else txt
                op = txt.split()[0] if txt else ""
                ops[op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += int(row[ti].replace(",", "") or 0)
            wide = ops.get("IMAD.WIDE", 0); total = sum(ops.values())
            dram = to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"])
            name = k["name"]
            result[name] = {"items": items, "imad_wide_per_item": wide / items, "thread_instructions_per_item": total / items, "dram_bytes_per_item": dram / items,
                            "other_fma_pipe_per_item": (ops.get("IMAD", 0) + ops.get("IMAD.HI", 0)) / items,
                            "fma_pipe_active_pct": float(r["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]), "issue_active_pct": float(r["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                            "warps_active_pct": float(r["sm__warps_active.avg.pct_of_peak_sustained_active"][0]), "registers": int(r["launch__registers_per_thread"][0]),
                            "time_under_ncu": " ".join(r["gpu__time_duration.sum"]), "top_ops_per_item": {o: c / items for o, c in ops.most_common(8)}, "report": rep.split("/")[-1]}
    json.dump(result, sys.stdout, indent=1); print()

if __name__ == "__main__":
    main()
