#!/usr/bin/env python
"""Turn `ncu --set full` captures (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into the per-kernel numbers bench.py and
the docs quote:  profiles/r02_ncu_traffic.json  (kernel name -> dram bytes per launch, duration, shared-memory pipe, conflicts).

    python profiles/parse_ncu.py gpurun_out/r02_prof_*.ncu-rep [--out profiles/r02_ncu_traffic.json]

bench.py reads `dram_bytes` from that file for `roofline.traffic` (never a constant in the source)."""
import csv, io, json, os, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WANT = {
    "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "duration": "gpu__time_duration.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smem_bank_conflicts_ld": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "inst_executed": "smsp__inst_executed.sum", "registers": "launch__registers_per_thread",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active",
}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        ent = res.setdefault(name, {"launches": 0})
        ent["launches"] += 1
        for key, metric in WANT.items():
            if metric not in col or r[col[metric]] in ("", "n/a"):
                continue
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]]
            if key.startswith("dram_r") or key.startswith("dram_w"):
                v *= UNIT.get(u, 1.0)
            if key == "duration":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            ent.setdefault("_" + key, []).append(v)
    final = {}
    for name, ent in res.items():
        row = {"launches_captured": ent["launches"], "source": os.path.basename(rep)}
        for key in WANT:
            vals = ent.get("_" + key)
            if vals:
                row[key if key != "duration" else "duration_us"] = sum(vals) / len(vals)
        if "dram_read" in row and "dram_write" in row:
            row["dram_bytes"] = int(row["dram_read"] + row["dram_write"])
        final[name] = row
    return final


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    outp = "profiles/r02_ncu_traffic.json"
    if "--out" in sys.argv:
        outp = sys.argv[sys.argv.index("--out") + 1]
        args = [a for a in args if a != outp]
    table = {}
    if os.path.exists(outp):
        table = json.load(open(outp))
    for rep in args:
        table.update(parse(rep))
    json.dump(table, open(outp, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1, sort_keys=True))
