"""Stall samples and executed instructions per CUDA source line of one kernel (ncu report captured with
--import-source on, code compiled with -lineinfo):  python tools/ncu_lines.py REP KERNEL_REGEX [N]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr = "?", None
    agg = defaultdict(lambda: [0, 0, "", defaultdict(int)])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            col = {}
            for i, c in enumerate(hdr):
                col.setdefault(c, i)
            stall_cols = [(c, i) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if hdr is None or len(r) < 8 or r[2] != "-":           # only the per-source-line summary rows
            continue
        try:
            s = int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
            e = int(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        a = agg[(fname, int(r[0]))]
        a[0] += s
        a[1] += e
        a[2] = r[1].strip()[:100]
        for c, i in stall_cols:
            try:
                a[3][c[6:]] += int(r[i] or 0)
            except (ValueError, IndexError):
                pass
    tot = sum(a[0] for a in agg.values())
    texe = sum(a[1] for a in agg.values())
    print(f"{kernel}: {tot} stall samples, {texe} warp instructions")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{k} {v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{a[0]:6d} {100.0 * a[0] / max(tot, 1):5.1f}%  exec {a[1]:9d} {100.0 * a[1] / max(texe, 1):5.1f}%  {f}:{ln}: {a[2]}   [{why}]")


if __name__ == "__main__":
    main()
