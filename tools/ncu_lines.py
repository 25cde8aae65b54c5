#!/usr/bin/env python3
"""Per-source-line view of an ncu capture without the GUI: joins the SASS page of a .ncu-rep
(instructions executed + warp-stall samples per instruction) with nvdisasm's line table of the
object file the kernel was built from.
    python tools/ncu_lines.py gpurun_out/em.ncu-rep ldsr_b200/build/kernels_pq3.o em_split_kernelILi3ELi4ELi2E [top]
The kernel pattern must select ONE instantiation (give enough of the mangled name): offsets of several
functions would overwrite each other in the line table.
Prints the hottest source lines (by stall samples) and totals per source function region."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, kernel_pat):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
    # with -gi every instruction is preceded by its inline chain: innermost line first, the line of
    # the kernel body (outermost call site) last
    table, chain, fresh, inside = {}, [], True, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = re.search(kernel_pat, ln) is not None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and chain:
            table[int(m.group(1), 16)] = (chain[0], chain[-1])
            fresh = True
    return table


def main():
    rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tab = line_table(obj, pat)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ins, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    per_site = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        inner, outer = tab.get(a - base, (("?", 0), ("?", 0)))
        for e in (per_line[inner], per_site[outer]):
            e[0] += int(r[ins])
            e[1] += int(r[ismp])
            for i in stall_cols:
                if r[i] not in ("", "0"):
                    e[2][hdr[i][6:]] += int(r[i])
        tot_i += int(r[ins])
        tot_s += int(r[ismp])
    print("total warp-instructions %d, stall samples %d" % (tot_i, tot_s))
    for title, d, n in (("by line of the kernel body (inlined callees folded into their call site)", per_site, top),
                        ("by innermost source line", per_line, top)):
        print("\n== %s" % title)
        print("%-28s %12s %6s %8s %6s  top stall reasons" % ("file:line", "inst", "%", "samples", "%"))
        for k, (ni, nsmp, st) in sorted(d.items(), key=lambda kv: -kv[1][1])[:n]:
            why = " ".join("%s=%d" % kv for kv in st.most_common(4))
            print("%-28s %12d %5.1f%% %8d %5.1f%%  %s" % ("%s:%d" % k, ni, 100.0 * ni / tot_i, nsmp,
                                                        100.0 * nsmp / max(tot_s, 1), why))


if __name__ == "__main__":
    main()
