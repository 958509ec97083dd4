#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture, without a GPU.

ncu's CSV source page only exports the SASS view; this joins it with the line table of the cubin
(`nvdisasm -g`) so that stall samples and executed instructions can be read per line of the .cuh file.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'k_merge_buckets<(int)256' [launch_skip] [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(rep, kernel_regex, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    ci = {n: k for k, n in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ci["Address"]].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break  # the next kernel of the report
            continue
        data.append((int(r[ci["Address"]], 16), r[ci["Source"]].strip(), int(r[ci["# Samples"]]),
                     int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]])))
    base = data[0][0]
    return name, [(a - base, s, smp, ins, tins) for a, s, smp, ins, tins in data]


def line_table(kernel_mangled_substr):
    """offset -> (file, line) for the first function of engine's cubin whose name contains the substring."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "impg_b200", "libimpgx.so")], cwd=tmp,
                   capture_output=True)
    cubin = os.path.join(tmp, "engine.sm_100a.cubin")
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    table, cur, inside = {}, None, False
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            inside = kernel_mangled_substr in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, rows = sass_rows(rep, rx, skip)
    m = re.search(r"(k_\w+)<([^>]*)>", name)
    base = m.group(1)
    # mangled: template ints appear as Li<N>E
    args = re.findall(r"\)(\d+)", m.group(2))
    sub = base + "I" + "".join(f"Li{a}E" if base != "k_liftover_ends" else f"Lb{a}E" for a in args)
    table = line_table(sub)
    agg = {}
    tot_s = tot_i = 0
    src = {}
    for off, sass, smp, ins, tins in rows:
        key = table.get(off, ("?", 0))
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += smp; a[1] += ins; a[2] += tins
        tot_s += smp; tot_i += ins
    print(name)
    print(f"samples {tot_s}, warp instructions {tot_i}; mangled match '{sub}', {len(table)} SASS lines mapped")
    files = {}
    for (f, l), (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in files:
            p = os.path.join(ROOT, "impg_b200", "csrc", f)
            files[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = files[f][l - 1].strip()[:100] if 0 < l <= len(files[f]) else ""
        print(f"{100 * s / max(tot_s, 1):5.1f}% samples {100 * i / max(tot_i, 1):5.1f}% inst  {f}:{l}  {text}")


if __name__ == "__main__":
    main()
