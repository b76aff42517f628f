"""profiles/traffic.json from the round's ncu captures (run HERE, after the gpurun call that produced them).

    python tools/capture_traffic.py gpurun_out/r2_selfplay.ncu-rep gpurun_out/r2_cap_plain.json \
                                    gpurun_out/r2_td.ncu-rep gpurun_out/r2_td_plain.log

Per kernel: DRAM bytes and warp instructions of the captured launch (ncu --set full), divided by the units that launch
processed (plies: 65,536 games x 16; TD steps: printed by tools/td_bench.py), and the sha256 of the kernel's SASS in the
libbgx.so of this tree - bench.py refuses to use the numbers when the loaded kernel hashes differently.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def num(v):
    return float(v.replace(",", ""))


def dram_bytes(vals, units):
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[k]]
        tot += num(vals[k]) * scale
    return tot


def main():
    sp_rep, sp_plain, td_rep, td_plain = sys.argv[1:5]
    import bench                                            # kernel_sass_sha
    out = {}
    v, u = raw(sp_rep)
    line = json.loads(open(sp_plain).read().strip().splitlines()[-1])
    plies = 65536 * 16
    out["k_selfplay"] = {
        "sass_function": "k_selfplayILi20ELi86ELb0", "sass_sha256": bench.kernel_sass_sha("k_selfplayILi20ELi86ELb0"),
        "dram_bytes_per_launch": dram_bytes(v, u), "warp_inst_per_ply": num(v["smsp__inst_executed.sum"]) / plies,
        "tree_edges_per_ply": line["tree_edges_per_ply_rank0"],
        "issue_active_pct": num(v["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "pipe_pct": {k: num(v[f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active"]) for k in ("alu", "fma", "xu", "lsu")},
        "capture": "profiles/r2f_selfplay_ncu_summary.md",
        "note": "one k_selfplay launch of bench.py (65,536 games x 16 plies), ncu --set full --clock-control none; tree edges per ply from the "
                "same command run without ncu"}
    v, u = raw(td_rep)
    m = re.search(r"([\d.]+) M steps/s.*games (\d+)", open(td_plain).read())
    steps_line = re.search(r"play (\d+) plies", open(td_plain).read())
    steps = int(steps_line.group(1))
    out["k_td_replay"] = {
        "sass_function": "k_td_replayILb0", "sass_sha256": bench.kernel_sass_sha("k_td_replayILb0"),
        "dram_bytes_per_launch": dram_bytes(v, u), "warp_inst_per_step": num(v["smsp__inst_executed.sum"]) / steps,
        "td_steps_of_the_launch": steps, "issue_active_pct": num(v["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "l1_hit_pct": num(v["l1tex__t_sector_hit_rate.pct"]), "l2_hit_pct": num(v["lts__t_sector_hit_rate.pct"]),
        "capture": "profiles/r2_td_replay.md",
        "note": "one k_td_replay launch of tools/td_bench.py (greedy self-play round, random-init weights), ncu --set full --clock-control none"}
    os.write(bench._REAL_STDOUT, (json.dumps(out, indent=1) + "\n").encode())
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
