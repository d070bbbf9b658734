"""Turn an `ncu --set full` report into the two small JSON files committed under profiles/.

    python tools/ncu_summary.py gpurun_out/r01_kernels.ncu-rep profiles/r01

writes <prefix>_ncu_kernels_summary.json (one row per captured launch) and <prefix>_traffic.json (per pipeline
stage: DRAM bytes per launch, instruction counts, IPC -- what bench.py reads for `roofline.traffic`). Uses
`ncu -i <rep> --page raw --csv`, so it runs anywhere the report and the ncu CLI are (no GPU needed).
"""
import csv
import io
import json
import subprocess
import sys

COLS = {
    "time_us": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
    "block": "launch__block_size", "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_inst": "smsp__inst_executed.sum", "ipc_per_sm": "sm__inst_executed.avg.per_cycle_elapsed",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
}
STAGES = {"render_bwd": "render_bwd", "render_fwd": "render_fwd", "preprocess_bwd": "preprocess_bwd",
          "preprocess_fwd": "preprocess_fwd", "emit_instances": "emit_instances", "tile_ranges": "tile_prefix",
          "tile_sort_pass1": "tile_sort_pass_kernel<1", "tile_sort_pass2": "tile_sort_pass_kernel<0",
          "tile_count": "tile_count"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(row, metric):
        i = col.get(metric)
        if i is None or row[i] in ("", "n/a"):
            return None
        return float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0), units[i]

    out, traffic = [], {}
    for r in body:
        name = r[col["Kernel Name"]]
        d = {"kernel": name}
        for k, m in COLS.items():
            v = val(r, m)
            d[k] = None if v is None else round(v[0], 6)
        out.append(d)
        for stage, pat in STAGES.items():
            if pat in name and stage not in traffic:
                traffic[stage] = {"dram_bytes": (d["dram_read"] or 0) + (d["dram_write"] or 0),
                                  "time_us_under_ncu": d["time_us"], "warp_inst": d["warp_inst"],
                                  "ipc_per_sm": d["ipc_per_sm"], "issue_active_pct": d["issue_active_pct"],
                                  "regs": d["regs"], "kernel": name}
    json.dump(out, open(prefix + "_ncu_kernels_summary.json", "w"), indent=1)
    json.dump({"source": f"ncu --set full --clock-control none --import-source on ({rep}); units: bytes, us, warp instructions",
               "kernels": traffic}, open(prefix + "_traffic.json", "w"), indent=1)
    for d in out:
        print(f'{d["kernel"][:70]:70s} {d["time_us"]:9.1f} us  inst {d["warp_inst"]:.3g}  ipc {d["ipc_per_sm"]}  dram {((d["dram_read"] or 0) + (d["dram_write"] or 0)) / 1e6:.1f} MB')


if __name__ == "__main__":
    main()
