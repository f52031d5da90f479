#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` capture of one bench step.

  ncu --set full --clock-control none --import-source on -k regex:"mixed_gemm|reorder_quantize" -s 28 -c 8 \
      -o gpurun_out/X/prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-prefill
  python tools/ncu_traffic.py gpurun_out/X/prof.ncu-rep profiles/r02_traffic.json [profiles/r02_ncu_step_summary.txt]

The LAST eight captured launches are one step in bench.py's order: quantize / GEMM of qkv, o, gate_up, down.
Per launch: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), duration, the pipe / memory utilisation figures the
rooflines cite.  bench.py copies the byte counts next to its live per-linear numbers.
"""
import csv
import json
import subprocess
import sys

NAMES = ["qkv", "o", "gate_up", "down"]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
           "sm__inst_executed.sum"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in data:
        name = r[ix["Kernel Name"]]
        kind = "gemm" if "mixed_gemm" in name else ("quantize" if "reorder_quantize" in name else None)
        if kind is None:
            continue
        rec = {"kind": kind, "kernel": name[:90]}
        for m in METRICS:
            if m in ix:
                rec[m] = r[ix[m]]
                rec[m + "__unit"] = units[ix[m]]
        rec["dram_bytes"] = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
            to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        launches.append(rec)
    step = launches[-8:]
    assert len(step) == 8 and [l["kind"] for l in step] == ["quantize", "gemm"] * 4, [l["kind"] for l in step]
    res = {"source": rep, "how": "ncu --set full --clock-control none, one bench step (M = 8192), per launch", "gemm": {}, "quantize": {}}
    lines = []
    for i, l in enumerate(step):
        nm = NAMES[i // 2]
        res[l["kind"]][nm] = {"dram_bytes": l["dram_bytes"], "duration_us": float(l["gpu__time_duration.sum"].replace(",", ""))}
        lines.append(f"{l['kind']:9s} {nm:8s} " + "  ".join(
            f"{m.split('.')[0].replace('__', ':')}={l.get(m, '?')}{l.get(m + '__unit', '')}" for m in METRICS))
    json.dump(res, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("# " + res["how"] + " (" + rep + ")\n" + "\n".join(lines) + "\n")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
