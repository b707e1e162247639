"""Attribute ncu SASS-level metrics to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel-substring> <cubin.sass from `nvdisasm -g -c`> [launch idx]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def sass_line_map(sass_path, kern):
    """offset -> (file line) for the function whose name contains `kern`."""
    m = {}
    cur = None
    infn = False
    for line in open(sass_path, errors="replace"):
        if line.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", line):
            infn = kern in line
            continue
        if re.match(r"^\s*\.section", line):
            infn = False
        if not infn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if mm:
            cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
            continue
        mo = re.match(r"^\s*/\*([0-9a-f]{4,})\*/", line)
        if mo and cur:
            m[int(mo.group(1), 16)] = cur
    return m


def main():
    rep, kern, sass = sys.argv[1], sys.argv[2], sys.argv[3]
    skern = sys.argv[4] if len(sys.argv) > 4 else kern
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name"')
    target = None
    for b in blocks[1:]:
        if kern in b.split("\n", 1)[0]:
            target = b
            break
    rows = list(csv.reader(target.split("\n")[1:]))
    h = rows[0]
    ia, ii, it, ism = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    isrc = h.index("Source")
    lm = sass_line_map(sass, skern)
    base = None
    per = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]
    stalls = defaultdict(lambda: defaultdict(int))
    for r in rows[1:]:
        if len(r) <= ism or not r[ia]:
            continue
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = a
        off = a - base
        key = lm.get(off, ("?", 0))
        try:
            v = (int(r[ii]), int(r[it]), int(r[ism]))
        except ValueError:
            continue
        for j in range(3):
            per[key][j] += v[j]
            tot[j] += v[j]
        for c in stall_cols:
            try:
                stalls[key][h[c]] += int(r[c])
            except ValueError:
                pass
    print("total warp-inst %d thread-inst %d samples %d" % tuple(tot))
    src = {}
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:40]:
        top = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
        line = ""
        try:
            if k[0] not in src:
                src[k[0]] = open("/root/repo/dampr_b200/csrc/" + k[0]).read().split("\n")
            line = src[k[0]][k[1] - 1].strip()[:70]
        except Exception:
            pass
        print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s   [%s]" % (
            100.0 * v[2] / max(1, tot[2]), 100.0 * v[0] / max(1, tot[0]), v[1] / max(1, v[0]), k[0], k[1], line,
            " ".join("%s=%d" % (a.replace("stall_", ""), b) for a, b in top)))


if __name__ == "__main__":
    main()
