"""ncu --set full report(s) -> profiles/r02_ncu_traffic.json: measured DRAM bytes per launch of our kernels.

usage: python scripts/ncu_traffic.py out.json key=report.ncu-rep[:"note"] ...

For every report: dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum averaged over the captured
launches of the report's kernel.  bench.py reads the json (`roofline.traffic`): the number is a measurement of the same
command's kernel at the same shapes, taken once per round under ncu (never during a timed run)."""
import csv, io, json, subprocess, sys

out, specs = sys.argv[1], sys.argv[2:]
res = {}
try:
    res = json.load(open(out))
except Exception:
    pass
for spec in specs:
    key, rest = spec.split("=", 1)
    rep, _, note = rest.partition(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("no data in", rep)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = float(r[col[name]].replace(",", ""))
        u = units[col[name]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "ns": 1e-6,
                 "us": 1e-3, "ms": 1.0}.get(u, 1)
        return v * scale
    launches = rows[2:]
    rd = [val(r, "dram__bytes_read.sum") for r in launches]
    wr = [val(r, "dram__bytes_write.sum") for r in launches]
    ms = [val(r, "gpu__time_duration.sum") for r in launches]
    n = len(launches)
    res[key] = {"kernel": launches[0][col["Kernel Name"]][:120], "launches_captured": n,
                "dram_bytes_read_per_launch": sum(rd) / n, "dram_bytes_write_per_launch": sum(wr) / n,
                "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "ncu_ms_per_launch": sum(ms) / n,
                "source": f"ncu --set full --clock-control none, {n} launch(es) of {rep.split('/')[-1]}" + (f" ({note})" if note else "")}
    print(key, json.dumps(res[key]))
json.dump(res, open(out, "w"), indent=1)
