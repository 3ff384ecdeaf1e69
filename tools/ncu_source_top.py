"""Per-instruction stall summary of an `ncu --set full --import-source on` capture, read without a GPU.

    python tools/ncu_source_top.py gpurun_out/prof.ncu-rep [top_n] [kernel-substring]

For every kernel in the report: total samples, the stall reasons summed over all instructions, and the top_n SASS
instructions by sample count with their dominant stall reasons and a few neighbouring opcodes for orientation.
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for k in kernels:
    if want not in k["name"]:
        continue
    hdr = k["hdr"]
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    samp = ci["# Samples"]
    total = sum(int(r[samp] or 0) for r in k["rows"])
    print("=" * 120)
    print(k["name"][:160])
    print(f"instructions {len(k['rows'])}, samples {total}")
    agg = {s: sum(int(r[ci[s]] or 0) for r in k["rows"]) for s in stall_cols}
    print("stall totals: " + ", ".join(f"{s[6:]} {v} ({100.0 * v / max(total, 1):.1f}%)"
                                        for s, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0))
    order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][samp] or 0))[:top_n]
    for i in sorted(order):
        r = k["rows"][i]
        st = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
        st = " ".join(f"{n}:{v}" for v, n in st if v > 0)
        print(f"{i:6d} {int(r[samp]):7d} {100.0 * int(r[samp]) / max(total, 1):5.1f}%  exec {r[ci['Instructions Executed']]:>10}  "
              f"{r[1].strip()[:70]:70s} | {st}")
