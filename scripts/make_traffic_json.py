"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the proof's hot kernels, from `ncu --set full` reports -> the JSON
bench.py reads for `roofline.traffic`.  usage: python scripts/make_traffic_json.py out.json name=rep.ncu-rep [name=rep ...]"""
import csv, json, subprocess, sys


def dram_per_launch(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, ms = [], []
    for r in data:
        if "nan" in (r[ir].lower(), r[iw].lower()) or "nan" in r[ir].lower() or "nan" in r[iw].lower():
            continue     # a replay pass that lost its counters
        tot.append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
        ms.append(float(r[it]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[it], 1.0))
    return {"dram_bytes_per_launch": sum(tot) / len(tot), "launches_captured": len(tot), "ms_per_launch_under_ncu": sum(ms) / len(ms)}


if __name__ == "__main__":
    res = {}
    for a in sys.argv[2:]:
        name, path = a.split("=", 1)
        res[name] = dram_per_launch(path)
        res[name]["source"] = path
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(res))
