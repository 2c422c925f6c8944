"""Summarise one kernel of an .ncu-rep (ncu --set full capture) as the JSON kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep kernel_regex attempts_per_launch > profiles/<round>_ncu_full_<kernel>.json
"""
import csv, io, json, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, pat, attempts = sys.argv[1], re.compile(sys.argv[2]), float(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if not pat.search(d.get("Kernel Name", "")):
            continue
        o = {"kernel": d["Kernel Name"][:120]}
        for k in KEYS:
            if k in d:
                o[k] = "%s %s" % (d[k], units[hdr.index(k)])
        for k in hdr:
            if "issue_stalled" in k and k.endswith(".ratio") and "not_issued" not in k:
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.3:
                    o[k] = v
        o["instructions_per_attempt"] = float(d["smsp__inst_executed.sum"].replace(",", "")) * 32 / attempts
        res.append(o)
    print(json.dumps(res, indent=1))


main()
