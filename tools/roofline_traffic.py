"""profiles/roofline_traffic.json from a tools/ncu_summary.py `full` text summary:
    python tools/roofline_traffic.py profiles/rNN_kernels_ncu.txt "<capture note>" > profiles/roofline_traffic.json
Per kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch and sm__inst_executed.avg.per_cycle_active
(read by bench.py for roofline.traffic and stats.hbm_bound_stages)."""
import json
import re
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
STAGE = [("k_blend_fwd", "blend_fwd"), ("k_blend_bwd", "blend_bwd"), ("k_project_bwd", "project_bwd"),
         ("k_project_fwd", "project_fwd"), ("k_depth_sort", "bin_sort_depth"),
         ("k_level_", "bin_tiles_count_and_fill_kernels")]


def main(path, note):
    traffic, ipc, cur = {}, {}, None
    for line in open(path):
        if line.startswith("void k_") or line.startswith("k_"):
            name = line.split("  grid")[0].replace("void ", "")
            cur = next((s for pat, s in STAGE if name.startswith(pat)), None)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+(dram__bytes_(?:read|write)\.sum)\s+([\d.,]+)\s+(\w+)", line)
        if m:
            traffic[cur] = traffic.get(cur, 0) + float(m.group(2).replace(",", "")) * UNIT[m.group(3)]
        m = re.match(r"\s+sm__inst_executed\.avg\.per_cycle_active\s+([\d.]+)", line)
        if m and not cur.startswith("bin_tiles"):
            ipc[cur] = round(float(m.group(1)), 3)
    out = {"ipc": ipc}
    out.update({k: int(v) for k, v in traffic.items()})
    out["_capture"] = note
    out["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (the tile-list entry sums its nine "
                       "kernels) and sm__inst_executed.avg.per_cycle_active (max 4); read by bench.py for roofline.traffic")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none")
