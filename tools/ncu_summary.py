"""Summarise one kernel of an .ncu-rep into JSON: python tools/ncu_summary.py <report.ncu-rep> <kernel regex> <out.json>
(reads `ncu -i <report> --page raw --csv`; keeps the metrics DESIGN.md / bench.py cite)."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum$|dram__cycles_active|gpu__dram_throughput|gpu__time_duration\.sum|"
                  r"lts__t_sector_hit_rate|l1tex__t_sector_hit_rate|launch__(block_size|grid_size|registers_per_thread|"
                  r"shared_mem_per_block_dynamic)|sm__warps_active\.avg\.pct|sm__pipe_tensor_cycles_active|"
                  r"sm__inst_executed_pipe_tc|sm__throughput\.avg\.pct|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active)")


def main():
    rep, pat, out = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    picked = [r for r in rows[2:] if pat.search(r[name_col])]
    if not picked:
        sys.exit("no kernel matches")
    r = picked[0]
    res = {"kernel": r[name_col], "report": rep, "launches_in_report": len(picked)}
    for i, h in enumerate(hdr):
        if KEEP.match(h):
            res[h] = [r[i], units[i]]
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k.startswith(("dram__bytes", "gpu__time", "launch__reg"))}))


if __name__ == "__main__":
    main()
