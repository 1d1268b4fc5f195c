#!/usr/bin/env python
"""dram traffic per frame of the detector conv kernels from an `ncu --set full` capture of
`scripts/gpu_probe_det.py --frames F --once` -> profiles/ncu_traffic.json (read by bench.py: roofline.traffic).
usage: python scripts/ncu_traffic.py gpurun_out/x.ncu-rep F"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyannote_video_b200.pyrgeom import pyramid_geometry  # noqa: E402


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m[unit]


def main():
    rep, frames = sys.argv[1], int(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {"detconv": 0.0, "conv1_fused": 0.0}
    per = []
    for r in data:
        name = r[ix["Kernel Name"]]
        b = sum(to_bytes(r[ix[k]], units[ix[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        key = "detconv" if ("detconv" in name or "rsconv" in name) else ("conv1_fused" if "conv1_fused" in name else None)
        if key:
            out[key] += b
            per.append({"kernel": name.split("(")[0].strip()[-60:], "dram_bytes": b})
    g = pyramid_geometry(1080, 1920, 1)
    impl = "rsconv" if any("rsconv" in q["kernel"] for q in per) else "detconv"
    j = {"source": os.path.basename(rep), "frames": frames, "impl": impl, "plane": [g.plane_h, g.plane_w],
         "detconv_dram_bytes_per_frame": out["detconv"] / frames,
         "conv1_fused_dram_bytes_per_frame": out["conv1_fused"] / frames, "launches": per}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps({k: v for k, v in j.items() if k != "launches"}))


if __name__ == "__main__":
    main()
