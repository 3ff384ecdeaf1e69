"""Reads an `ncu --set full` capture (.ncu-rep) and prints / stores the DRAM traffic and headline counters of the
matching kernel launches.  Used to fill profiles/r02_traffic.json, which bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/joint_fwd.ncu-rep "joint_fwd_kernel<3, 0" KEY [profiles/r02_traffic.json]
"""
import csv
import json
import subprocess
import sys

rep, pattern, key = sys.argv[1], sys.argv[2], sys.argv[3]
out_path = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]
units = rows[1]
res = []
for r in rows[2:]:
    if len(r) != len(hdr) or pattern not in r[ci["Kernel Name"]]:
        continue
    d = {"kernel": r[ci["Kernel Name"]][:120]}
    for w in want:
        if w in ci:
            try:
                d[w] = float(r[ci[w]].replace(",", ""))
                d[w + "|unit"] = units[ci[w]]
            except ValueError:
                pass
    res.append(d)
for d in res:
    print(json.dumps(d))
if out_path and res:
    def to_bytes(d, k):
        v, u = d.get(k), d.get(k + "|unit", "byte")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        return v * mult
    last = res[-1]
    try:
        with open(out_path) as fh:
            table = json.load(fh)
    except FileNotFoundError:
        table = {}
    table[key] = {"dram_bytes_read": to_bytes(last, "dram__bytes_read.sum"),
                  "dram_bytes_write": to_bytes(last, "dram__bytes_write.sum"),
                  "gpu_time": last.get("gpu__time_duration.sum"), "gpu_time_unit": last.get("gpu__time_duration.sum|unit"),
                  "tensor_pipe_active_pct": last.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "source": f"ncu --set full --clock-control none, {rep.split('/')[-1]}, launch {len(res)} of {len(res)} matching"}
    with open(out_path, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    print("wrote", out_path, key)
