"""profiles/traffic.json from `ncu --set full` reports: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
kernel of a workload, stamped with the SHA-256 of the source file the kernel lives in. bench.py refuses an entry whose stamp no longer
matches the tree (roofline.traffic = null) instead of reporting a stale figure.

    python tools/make_traffic.py <workload> "<table_layout_chosen>" <kernel regex> <report.ncu-rep> <source file> [...more 5-tuples]"""
import csv
import hashlib
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles" / "traffic.json"


def dram_bytes(report: str, pattern: str):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        if re.search(pattern, r[idx["Kernel Name"]]):
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
            vals.append((tot, float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]], r[idx["Kernel Name"]].split("(")[0]))
    if not vals:
        raise SystemExit(f"no kernel matching {pattern} in {report}")
    return vals[-1]


def main():
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    args = sys.argv[1:]
    for i in range(0, len(args), 5):
        workload, layout, pattern, report, source = args[i:i + 5]
        b, t, tu, name = dram_bytes(report, pattern)
        data.setdefault(workload, {})[layout] = {
            "kernel": name.replace("void ", ""), "dram_bytes": b, "ncu_duration": f"{t} {tu}", "source": f"{Path(report).name}: ncu --set full --clock-control none, one launch",
            "source_file": source, "source_sha256": hashlib.sha256((ROOT / source).read_bytes()).hexdigest()}
    OUT.write_text(json.dumps(data, indent=1) + "\n")
    print(json.dumps(data, indent=1))


if __name__ == "__main__":
    main()
