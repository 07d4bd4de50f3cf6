#!/usr/bin/env python3
"""Summarise an `ncu --set full` report of tools/kernel_zoo.py into JSON (one record per launch, labelled with the role
printed by kernel_zoo.py):

  ncu -i gpurun_out/zoo_r2.ncu-rep --page raw --csv > /tmp/zoo.csv
  python tools/ncu_zoo_summary.py /tmp/zoo.csv gpurun_out/zoo_r2.log > profiles/r2/ncu_kernel_zoo_summary.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    names = []
    if len(sys.argv) > 2:
        for line in open(sys.argv[2]):
            m = re.match(r"^(\d+) (.+)$", line.strip())
            if m:
                names.append(m.group(2))
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    launches = [r for r in rows[2:] if len(r) == len(hdr)]
    # the capture is restricted to the library's kernels with ncu's -k regex, in launch order
    ours = launches
    for i, r in enumerate(ours):
        rec = {"launch": i, "role": names[i] if i < len(names) else None,
               "kernel": re.sub(r"\(CUtensorMap.*|\((const |float|unsigned|__nv|long).*", "",
                                r[col["Kernel Name"]].replace("dprb::<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")),
               "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k, nm in WANT.items():
            if k in col:
                v = num(r[col[k]])
                u = units[col[k]]
                if v is not None and nm == "duration_us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if v is not None and nm.endswith("_bytes") and u in ("Kbyte", "Mbyte", "Gbyte", "byte"):
                    v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                rec[nm] = v
        if rec.get("duration_us") and rec.get("dram_read_bytes") is not None:
            rec["dram_GBs"] = (rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0)) / rec["duration_us"] / 1e3
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
