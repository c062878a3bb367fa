#!/usr/bin/env python
"""Per-source-line view of an ncu report: joins `ncu --page source --csv` (SASS, per-instruction samples and
shared-memory wavefronts) with `nvdisasm -gi` line info of the same cubin (offset -> outermost kernel-file line).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <object .o> <mangled-name substring> [file suffix]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_map(obj, mangled, want_file):
    d = tempfile.mkdtemp()
    import os
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    import glob
    cub = glob.glob(d + "/*.cubin")[0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
    m = {}
    infn = False
    cur = None
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            infn = mangled in ln
            continue
        if not infn:
            continue
        mm = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if mm:
            f, l, rest = mm.group(1), int(mm.group(2)), mm.group(3)
            # with -gi the chain is printed innermost first; keep the entry that lies in the wanted file
            if f.endswith(want_file):
                cur = l
            else:
                m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
                for f2, l2 in m2:
                    if f2.endswith(want_file):
                        cur = int(l2)
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if mm:
            m[int(mm.group(1), 16)] = (cur, mm.group(2).strip())
    return m


def main():
    rep, kre, obj, mangled = sys.argv[1:5]
    want_file = sys.argv[5] if len(sys.argv) > 5 else "lqr_fwd.cuh"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    lm = line_map(obj, mangled, want_file)

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    agg = defaultdict(lambda: defaultdict(float))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in data:
        off = int(r[0], 16) - base
        line, _ = lm.get(off, (None, ""))
        a = agg[line]
        a["samples"] += f(r, "# Samples")
        a["inst"] += f(r, "Instructions Executed")
        a["wf"] += f(r, "L1 Wavefronts Shared")
        a["wfx"] += f(r, "L1 Wavefronts Shared Excessive")
        for s in stalls:
            a[s] += f(r, s)
    ts = sum(a["samples"] for a in agg.values())
    tw = sum(a["wf"] for a in agg.values()) or 1.0
    ti = sum(a["inst"] for a in agg.values())
    src = {}
    try:
        import os
        path = [p for p in (os.path.join(os.path.dirname(obj), "..", want_file), want_file) if os.path.exists(p)][0]
        for i, l in enumerate(open(path), 1):
            src[i] = l.rstrip()
    except Exception:
        pass
    print(f"samples {ts:.0f}  smem wavefronts {tw:.0f}  instructions {ti:.0f}")
    print("line  samples%  inst%  wavefronts% (excess%)  top stalls | source")
    key = sys.argv[6] if len(sys.argv) > 6 else "samples"
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:45]:
        st = sorted(((s, a[s]) for s in stalls), key=lambda x: -x[1])[:3]
        tot_st = sum(a[s] for s in stalls) or 1.0
        sts = " ".join("%s=%.0f%%" % (s.replace("stall_", ""), 100 * v / tot_st) for s, v in st)
        print("%5s %6.1f%% %6.1f%% %6.1f%% (%5.1f%%)  %-44s | %s" % (line, 100 * a["samples"] / ts, 100 * a["inst"] / ti,
                                                                  100 * a["wf"] / tw, 100 * a["wfx"] / tw, sts,
                                                                  src.get(line, "")[:90].strip()))


if __name__ == "__main__":
    main()
