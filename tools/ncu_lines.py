#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction counters to CUDA source lines (run in the build container, no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'evaluate_kernelILb1' [top_n]

ncu's `--page source --csv` lists SASS instructions with their counters but without line numbers; `nvdisasm -g` on the
cubin extracted from the in-tree library lists the same instructions with `//## File ..., line N` markers.  The two are
joined on the instruction offset.  The library must be the one the capture ran (rebuild -> recapture)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dlii_jssp_b200", "libjssp_b200.so")


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    lines, cur, on = {}, None, False
    for ln in dis:
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines[int(m.group(1), 16)] = cur
    src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src_csv.split("\n")))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows if len(r) > ie and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    agg, samp = collections.Counter(), collections.Counter()
    for r in data:
        key = lines.get(int(r[0], 16) - base)
        agg[key] += float(r[ie]); samp[key] += float(r[isamp])
    tot, ts = sum(agg.values()), sum(samp.values()) or 1.0
    text = {}
    print(f"kernel {kernel}: {tot:.0f} warp instructions, {ts:.0f} samples")
    for key, v in agg.most_common(top):
        if key is None:
            print(f"{v / tot * 100:5.1f}%  (no line info)")
            continue
        f, l = key
        if f not in text:
            p = os.path.join(ROOT, "dlii_jssp_b200", "csrc", f)
            text[f] = open(p).read().split("\n") if os.path.exists(p) else []
        t = text[f][l - 1].strip()[:100] if l - 1 < len(text[f]) else ""
        print(f"{v / tot * 100:5.1f}% inst {samp[key] / ts * 100:5.1f}% samples  {f}:{l}  {t}")


if __name__ == "__main__":
    main()
