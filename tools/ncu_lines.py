#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel.

Joins `ncu -i REP --page source --csv` (SASS with 'Instructions Executed' and '# Samples')
with `nvdisasm -g -c` line markers of the matching cubin, instruction by instruction.

    python tools/ncu_lines.py REP.ncu-rep path/to/lib.so KERNEL_SUBSTRING [top_n]
"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    ie, ns, src = H.index("Instructions Executed"), H.index("# Samples"), H.index("Source")
    return [(r[src].strip(), float(r[ie] or 0), float(r[ns] or 0)) for r in rows[hdr + 1:] if len(r) == len(H)]


def line_map(lib, kernel, want=None):
    """SASS instructions of every .text section whose name contains `kernel`, with the source line
    each one came from; among several template instantiations the one with `want` instructions wins."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    found = []
    for f in os.listdir(tmp):
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in dis:
            continue
        lines = dis.splitlines()
        for start in (i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l):
            cur, out = ("?", 0), []
            for l in lines[start + 1:]:
                if l.startswith(".text.") or l.startswith(".section"):
                    break
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
                if m:
                    out.append((cur, m.group(1).strip()))
            found.append(out)
    if not found:
        raise SystemExit("kernel not found in " + lib)
    if want is not None:
        for out in found:
            if len(out) == want:
                return out
    return found[0]


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass = sass_rows(rep)
    lm = line_map(lib, kernel, len(sass))
    if len(sass) != len(lm):
        print(f"# warning: {len(sass)} profiled instructions vs {len(lm)} disassembled (library rebuilt since the capture?)")
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
    ops = collections.defaultdict(float)
    for (src, ie, ns), (loc, op) in zip(sass, lm):
        a = agg[loc]
        a[0] += ie; a[1] += ns; a[2] += 1
        ops[re.sub(r"^@!?U?P\d+\s+", "", op).split()[0].split(".")[0]] += ie
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    print(f"# {kernel}: {tot_i:.3e} warp instructions, {tot_s:.0f} samples, {len(sass)} SASS instructions")
    print("# inst%  samp%  sass  file:line")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * a[0] / tot_i:6.2f} {100 * a[1] / tot_s:6.2f} {a[2]:5d}  {loc[0]}:{loc[1]}")
    print("# opcode mix (share of warp instructions)")
    print("  " + "  ".join(f"{k} {100 * v / tot_i:.1f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]))


if __name__ == "__main__":
    main()
