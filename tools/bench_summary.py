"""Print the headline and the extras of a bench.py JSON line in a few rows. Usage: python tools/bench_summary.py <file.json>"""
import json
import sys

d = json.load(open(sys.argv[1]))
rf = d.get("roofline") or {}
print("HEAD", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in (rf.get("phases_ms") or d.get("c5_phases") or {}).items() if isinstance(v, float)},
      "job_frac", round((rf.get("job") or {}).get("frac", 0), 3), d["parity"], "e2e_ms", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), "launches", d["gpu_launches"])
for k in ("c2_sparse", "hash_layout", "match_cache", "fused", "c3", "c4", "c5", "ref10m", "ref100m"):
    r = d.get(k)
    if not r:
        continue
    rf = r.get("roofline") or {}
    print(f"{k:12s} {r['ms_per_step']:8.3f} ms", {a: round(b, 3) for a, b in r["phases_ms"].items()}, "job_frac", round(rf.get("job", {}).get("frac", 0), 3), "|", r["table_layout_chosen"], "|",
          "parity", all(r["parity"].values()), r.get("reference_published", {}).get("speedup_vs_published", ""), (r.get("c5_phases") or {}).get("nvlink_frac", ""))
