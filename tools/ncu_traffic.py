#!/usr/bin/env python3
"""Measured DRAM traffic per frame of the hot kernels from `ncu --set full` captures -> profiles/r1_dram_traffic.json
usage: python tools/ncu_traffic.py <frames_in_capture> name=rep.ncu-rep [name=rep ...]"""
import csv, json, subprocess, sys

frames = int(sys.argv[1])
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(metric):
        i = hdr.index(metric)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                 "msecond": 1e3}.get(u, 1)
        return v * scale

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out[name] = {"frames_in_capture": frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "dram_bytes_per_frame": round((rd + wr) / frames), "duration_us": get("gpu__time_duration.sum"),
                 "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
json.dump(out, open("profiles/r1_dram_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
