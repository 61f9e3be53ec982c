"""Aggregate an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass) by CUDA source line.
usage: python tools/ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":  # line-level rows have '-' as address
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        out.append((int(d["Warp Stall Sampling (All Samples)"]), int(d["Instructions Executed"]), cur_file, r[0], r[1].strip()[:120], d))
    except ValueError:
        pass
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, warp instructions {toti}")
stall_cols = [c for c in out[0][5] if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(o[5][c] or 0) for o in out) for c in stall_cols}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for v, ins, f, ln, src, d in sorted(out, key=lambda o: -o[0])[:top]:
    st = sorted(((int(d[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{100 * v / tot:5.1f}% smp {100 * ins / toti:5.1f}% inst  {f}:{ln:>4s}  [{st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}]  {src}")
