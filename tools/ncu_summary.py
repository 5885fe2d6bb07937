"""Turns an `ncu --set full` report into the text summary kept under profiles/
and merges per-kernel DRAM traffic into profiles/traffic.json (read by bench.py).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xyz.txt "free text header"
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def short(name):
    name = name.replace("void ", "").replace("mgb::", "").replace("(int)", "").replace("(bool)", "")
    if "<" in name:
        name = name[: name.index(">") + 1]
    else:
        name = re.sub(r"\(.*", "", name)
    return name.replace(" ", "").strip()


def main():
    rep, out, header = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    lines = [header, "ncu --set full --clock-control none; per launch; dram bytes = "
             "dram__bytes_read.sum + dram__bytes_write.sum", ""]
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    seen = {}
    for r in rows[2:]:
        name = short(r[h.index("Kernel Name")])
        vals = {}
        for col, key in KEYS:
            if col in h:
                vals[key] = (r[h.index(col)], units[h.index(col)])
        rd = float(vals["dram_read"][0]) * UNIT.get(vals["dram_read"][1], 1)
        wr = float(vals["dram_write"][0]) * UNIT.get(vals["dram_write"][1], 1)
        lines.append(f"{name} | grid {vals['grid'][0]} block {vals['block'][0]} | "
                     f"{float(vals['time_us'][0]):.1f} us | dram read {rd/1e9:.3f} GB write {wr/1e9:.3f} GB "
                     f"| dram {float(vals['dram_pct'][0]):.1f}% | warps active {float(vals['warps_active_pct'][0]):.1f}% "
                     f"| issue active {float(vals['issue_active_pct'][0]):.1f}% | regs {vals['regs'][0]} "
                     f"| L2 hit {float(vals['l2_hit_pct'][0]):.1f}% L1 hit {float(vals['l1_hit_pct'][0]):.1f}% "
                     f"| warp insts {float(vals['warp_insts'][0])/1e6:.1f} M")
        key = re.sub(r"<.*", "", name) if name.startswith("k_half_sweep") else name
        seen.setdefault(key, []).append((rd + wr, float(vals["time_us"][0]), vals["grid"][0]))
    def blocks(grid):
        n = 1
        for x in re.findall(r"\d+", grid):
            n *= int(x)
        return n

    for key, lst in seen.items():
        # one template instance serves several levels: keep the launches of the largest
        # grid (the finest level), so that the figures are per FINEST-level launch
        most = max(blocks(x[2]) for x in lst)
        lst = [x for x in lst if blocks(x[2]) == most]
        traffic[key] = {"dram_bytes_per_launch": sum(x[0] for x in lst) / len(lst),
                        "ncu_time_us": sum(x[1] for x in lst) / len(lst), "launches": len(lst),
                        "grid": lst[0][2], "source": os.path.basename(out)}
    open(out, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
