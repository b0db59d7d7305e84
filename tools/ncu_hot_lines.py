#!/usr/bin/env python3
"""Attribute ncu per-instruction counters to source lines.

ncu's CSV export of the source page carries SASS only; this joins it with `nvdisasm -g` line info of the same
kernel from the built library, and prints the source lines with the most executed warp instructions and stall samples.

  python tools/ncu_hot_lines.py gpurun_out/prof2.ncu-rep k_sense [top_n]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "puckswarm_b200", "csrc", "libpuckswarm_b200.so")


def sass_lines(kernel):
    """offset -> (file, line) from nvdisasm -g of the cubin embedded in the library."""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    out = subprocess.run(["nvdisasm", "-g", "-c", cubins[0]], capture_output=True, text=True).stdout
    res, cur, infn = {}, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            infn = kernel in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            res[int(m.group(1), 16)] = cur
    return res


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    # the export holds one block per profiled launch: take the first
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {h: i for i, h in enumerate(hdr)}
    lines = sass_lines(kernel)
    opre = re.compile(os.environ["OP"]) if os.environ.get("OP") else None   # e.g. OP='STL|LDL': local-memory traffic only
    base = None
    inst = collections.Counter()
    stall = collections.Counter()
    total_i = total_s = 0
    for r in rows[hdr_i + 1:]:
        if not r or r[0] == "Address" or r[0] == "Kernel Name":
            break
        addr = int(r[ci["Address"]], 16)
        if base is None:
            base = addr
        if opre and not opre.search(r[ci["Source"]]):
            continue
        key = lines.get(addr - base, ("?", 0))
        n = int(float(r[ci["Instructions Executed"]] or 0))
        s = int(float(r[ci["# Samples"]] or 0))
        inst[key] += n
        stall[key] += s
        total_i += n
        total_s += s
    print("kernel %s: %d warp instructions, %d samples" % (kernel, total_i, total_s))
    print("%-22s %6s %14s %6s %10s %6s" % ("file", "line", "warp-inst", "%", "samples", "%"))
    for key, n in sorted(inst.items(), key=lambda kv: -(stall[kv[0]] if os.environ.get("BY","inst") == "stall" else kv[1]))[:top]:
        print("%-22s %6d %14d %5.1f%% %10d %5.1f%%" % (key[0], key[1], n, 100.0 * n / max(1, total_i), stall[key], 100.0 * stall[key] / max(1, total_s)))


if __name__ == "__main__":
    main()
