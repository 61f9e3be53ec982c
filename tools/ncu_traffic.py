"""DRAM traffic per launch of the captured kernels (ncu --set full reports) -> profiles/<round>_ncu_traffic.json,
which bench.py reads for roofline.traffic.  usage: python tools/ncu_traffic.py out.json label=report.ncu-rep ..."""
import csv, json, subprocess, sys


def val(x, unit):
    v = float(x.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


out = {}
for arg in sys.argv[2:]:
    label, rep = arg.split("=", 1)
    div = 1.0
    if "/" in rep.split(".ncu-rep")[-1]:
        rep, d = rep.rsplit("/", 1)
        div = float(d)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in txt.splitlines() if not l.startswith("==")]))
    hdr, units, r = rows[0], rows[1], rows[2]
    rd, wr, du = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    out[label] = {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_per_launch": (val(r[rd], units[rd]) + val(r[wr], units[wr])) / div, "divided_by": div,
                  "duration_under_ncu": r[du] + " " + units[du], "report": rep.split("/")[-1], "how": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full --clock-control none capture"}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
