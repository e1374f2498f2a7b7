"""Summarise .ncu-rep files (ncu --set full captures) or their `--page raw --csv` exports into one JSON:
    python tools/ncu_summary.py out.json rep1.ncu-rep|rep1_raw.csv [...]
Per kernel launch: duration, registers, occupancy, pipe utilisation, executed instructions per pipe, stall reasons per issue,
DRAM bytes.  Needs `ncu` on PATH (reads the reports with --page raw --csv)."""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
STALLS = "smsp__average_warps_issue_stalled_"


def summarise(path):
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = {"Kernel Name": d.get("Kernel Name", "")}
        for name, val, unit in zip(hdr, r, units):
            if name in KEEP or (name.startswith(STALLS) and name.endswith("_per_issue_active.ratio")):
                if val not in ("", "n/a"):
                    k[name] = (val + " " + unit).strip()
        # ncu sometimes returns "-nan" for every derived counter of a launch it could not replay consistently: keep the
        # duration, drop the counters, and say so
        if any(v.startswith("-nan") or v.startswith("nan") for v in k.values()):
            k = {n: v for n, v in k.items() if not (v.startswith("-nan") or v.startswith("nan"))}
            k["note"] = "ncu returned nan for the derived counters of this launch"
        res.append(k)
    return res


if __name__ == "__main__":
    out, reps = sys.argv[1], sys.argv[2:]
    doc = {"reports": {}}
    for rep in reps:
        doc["reports"][rep.split("/")[-1]] = summarise(rep)
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote", out)
