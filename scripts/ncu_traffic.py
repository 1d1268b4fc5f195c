#!/usr/bin/env python
"""dram traffic per frame of the detector conv kernels from an `ncu --set full` capture of one bench step
(`bench.py --ncu-step`, exported with `ncu -i x.ncu-rep --page raw --csv`) -> profiles/ncu_traffic.json (read by bench.py:
roofline.traffic, labelled static).  usage: python scripts/ncu_traffic.py gpurun_out/x_raw.csv FRAMES"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyannote_video_b200.pyrgeom import pyramid_geometry  # noqa: E402

DET = {"rsconv_kernel<32, 32, 5, 5, 2, 0>": "conv3", "rsconv_kernel<32, 48, 5, 5, 1, 0>": "conv4",
       "rsconv_kernel<48, 48, 5, 5, 1, 0>": "conv5/6", "rsconv_kernel<48, 16, 9, 1, 1, 1>": "conv7",
       "rsconv_kernel<16, 32, 5, 5, 2, 0>": "conv2"}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m[unit]


def main():
    rep, frames = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(open(rep).read().splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    rs, c12, c1 = 0.0, 0.0, 0.0
    per, layers = [], []
    for r in data:
        name = r[ix["Kernel Name"]]
        b = sum(to_bytes(r[ix[k]], units[ix[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        hit = next((v for k, v in DET.items() if k in name), None)
        if hit:
            rs += b
            layers.append(hit)
            per.append({"kernel": hit, "dram_bytes": b})
        elif "c12_kernel" in name:
            c12 += b
            per.append({"kernel": "conv1+2 (c12)", "dram_bytes": b})
        elif "conv1_fused" in name:
            c1 += b
    g = pyramid_geometry(1080, 1920, 1)
    names = []
    for l in layers:
        names += ["conv5", "conv6"][len([n for n in names if n in ("conv5", "conv6")]):][:1] if l == "conv5/6" else [l]
    j = {"source": os.path.basename(rep), "frames": frames, "impl": "rsconv", "plane": [g.plane_h, g.plane_w], "layers": names,
         "detconv_dram_bytes_per_frame": rs / frames, "c12_dram_bytes_per_frame": c12 / frames,
         "conv1_fused_dram_bytes_per_frame": c1 / frames, "launches": per}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps({k: v for k, v in j.items() if k != "launches"}))


if __name__ == "__main__":
    main()
