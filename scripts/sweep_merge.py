"""Merge per-GPU-count sweep files (scripts/sweep.py) into one JSONL with scaling efficiency per point.
    python scripts/sweep_merge.py out.jsonl sweep_1gpu.jsonl sweep_2gpu.jsonl ..."""
import json
import sys

rows = {}
extra = []
for path in sys.argv[2:]:
    for ln in open(path):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        if "N" in d and "n_gpus" in d:
            rows[(d["N"], d["M"], d["Q"], d["n_gpus"])] = d
        else:
            extra.append(d)
with open(sys.argv[1], "w") as f:
    for d in extra:
        f.write(json.dumps(d) + "\n")
    for key in sorted(rows):
        d = dict(rows[key])
        base = rows.get(key[:3] + (1,))
        if base:
            d["scaling_efficiency"] = d["rows_per_s"] / (key[3] * base["rows_per_s"])
        f.write(json.dumps(d) + "\n")
print("merged", len(rows), "points")
