"""GPU probe: srgemm kernel vs the CPU emulator over layouts / grouping / descriptor modes.

Each case runs in its own subprocess (a device-side trap poisons the CUDA context), results are
appended to gpurun_out/probe_srgemm.jsonl.  Usage: python scripts/gpu_probe_srgemm.py [--quick]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = []
# (name, layout kind, B,H,W, cin, cout, k, stride, pad, group, desc_mode, timing)
for dm in (0,):
    for grp in ("tap", "row", "all"):
        CASES.append(("c32_3x3", "padded", 2, 35, 35, 32, 32, 3, 1, 1, grp, dm, False))
for grp in ("tap", "row"):
    CASES.append(("c64_3x3", "padded", 2, 17, 17, 64, 64, 3, 1, 1, grp, 0, False))
    CASES.append(("c128_3x3", "padded", 4, 8, 8, 128, 128, 3, 1, 1, grp, 0, False))
    CASES.append(("c256_3x3", "padded", 8, 4, 4, 256, 256, 3, 1, 1, grp, 0, False))
    CASES.append(("c48_5x5", "padded", 1, 40, 60, 48, 45, 5, 1, 2, grp, 0, False))
    CASES.append(("c16_5x5s2", "parity", 1, 61, 83, 16, 32, 5, 2, 0, grp, 0, False))
    CASES.append(("c32_3x3s2", "parity", 2, 35, 35, 32, 64, 3, 2, 0, grp, 0, False))
CASES.append(("first5", "gathered", 1, 63, 90, 3, 16, 5, 2, 0, "tap", 0, False))
CASES.append(("first7", "gathered", 2, 150, 150, 3, 32, 7, 2, 0, "tap", 0, False))
CASES.append(("c48_9x9_n1", "padded", 1, 30, 40, 48, 1, 9, 1, 4, "tap", 0, False))
# timing cases (large)
for grp in ("tap", "row", "all"):
    CASES.append(("T_embed_L4", "padded", 512, 35, 35, 32, 32, 3, 1, 1, grp, 0, True))
    CASES.append(("T_embed_L3", "padded", 512, 17, 17, 64, 64, 3, 1, 1, grp, 0, True))
CASES.append(("T_det_c1", "gathered", 2, 2182, 14805, 3, 16, 5, 2, 0, "tap", 0, True))
CASES.append(("T_det_c3", "parity", 2, 543, 3699, 32, 32, 5, 2, 0, "row", 0, True))
for grp in ("tap", "row"):
    CASES.append(("T_det_c5", "padded", 1, 270, 1700, 48, 45, 5, 1, 2, grp, 0, True))
    CASES.append(("T_det_c2", "parity", 1, 1080, 3400, 16, 32, 5, 2, 0, grp, 0, True))
    CASES.append(("T_embed_L2", "padded", 512, 8, 8, 128, 128, 3, 1, 1, grp, 0, True))
    CASES.append(("T_embed_L1", "padded", 512, 4, 4, 256, 256, 3, 1, 1, grp, 0, True))


def run_case(idx):
    import torch
    from pyannote_video_b200.plan import ConvPlan, RowLayout, Srgemm
    from srgemm_emu import emulate
    name, kind, B, H, W, cin, cout, k, stride, pad, grp, dm, timing = CASES[idx]
    torch.manual_seed(idx)
    dev = torch.device("cuda:0")
    res = dict(case=name, kind=kind, B=B, H=H, W=W, cin=cin, cout=cout, k=k, stride=stride, group=grp, desc_mode=dm)
    x = torch.randn(B, H, W, cin).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, k, k) * (1.0 / (cin * k * k) ** 0.5)).to(torch.bfloat16).float()
    if kind == "gathered":
        lin = RowLayout("gathered", B, H, W, 3, kw=k)
    else:
        lin = RowLayout(kind, B, H, W, cin, pad=pad)
    cp = ConvPlan(lin, w, stride, pad, group=grp, ctas_per_sm=int(os.environ.get('PV_CTAS', '0')))
    res.update(ctas=cp.ctas_per_sm, stages=len(cp.stages), tail=cp.tail_rows, mma_per_tile=cp.mma_per_tile, N=cp.N)
    f32 = cout == 1
    lout = RowLayout("padded", B, cp.OH, cp.OW, cp.N, pad=1)
    scale = torch.rand(cout) + 0.5
    shift = torch.randn(cout) * 0.1
    if timing:
        xr = torch.randn(lin.rows, lin.cols, device=dev).to(torch.bfloat16)
    else:
        xr_cpu = lin.to_rows(x)
        xr = xr_cpu.to(dev)
    if f32:
        out = torch.zeros(B, cp.OH, cp.OW, dtype=torch.float32, device=dev)
    else:
        out = lout.alloc(dev)
    op = Srgemm(cp, xr, out, lout, scale, shift, relu=not f32, out_f32=f32, acc_split=int(os.environ.get('PV_ACC_SPLIT', '0')),
                mma_warps=(int(os.environ['PV_MMAW']) if 'PV_MMAW' in os.environ else None))
    res.update(op.info())
    res.update(n_slots=cp.n_slots)
    op.run()
    op.check()
    if timing:
        import time as _t
        t0 = _t.time()
        while _t.time() - t0 < 0.4:      # let clocks ramp: >= 0.4 s of back-to-back launches
            for _ in range(20):
                op.run()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            op.run()
        e1.record()
        for _ in range(200):
            op.run()
        smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res["smi"] = smi
        q = lin.plane_rows
        flops = 2.0 * q * cp.N * 16 * cp.mma_per_tile
        useful = 2.0 * B * cp.OH * cp.OW * cout * cin * k * k
        res.update(ms=ms, tflops_issued=flops / ms / 1e9, tflops_useful=useful / ms / 1e9,
                   x_mb=lin.rows * lin.cols * 2 / 1e6)
        op.check()
    else:
        if f32:
            ref = torch.zeros(B, cp.OH, cp.OW)
        else:
            ref = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
        emulate(cp, xr_cpu, ref, lout, scale, shift, relu=not f32, out_f32=f32)
        got = out.cpu().float()
        d = (got - ref.float()).abs()
        tol = 0.02 + 0.02 * ref.float().abs()
        res.update(max_abs=float(d.max()), n_bad=int((d > tol).sum()), n=int(d.numel()),
                   ref_absmax=float(ref.float().abs().max()))
        res["ok"] = res["n_bad"] == 0
    print("RESULT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(int(sys.argv[2]))
        return
    quick = "--quick" in sys.argv
    only_timing = "--timing" in sys.argv
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    outp = os.path.join(ROOT, "gpurun_out", "probe_srgemm.jsonl")
    with open(outp, "a") as f:
        splits = [s_ for s_ in os.environ.get("PV_SPLITS", "0").split(",")]
        for sp_, i, c in [(sp_, i_, c_) for sp_ in splits for (i_, c_) in enumerate(CASES)]:
            if quick and c[-1]:
                continue
            if only_timing and not c[-1]:
                continue
            t0 = time.time()
            try:
                env = dict(os.environ)
                env["PV_ACC_SPLIT"] = sp_
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)],
                                   capture_output=True, text=True, timeout=300, env=env)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                if line:
                    rec = json.loads(line[-1][7:])
                else:
                    rec = dict(case=c[0], group=c[10], desc_mode=c[11], error=(p.stderr or p.stdout)[-600:], rc=p.returncode)
            except subprocess.TimeoutExpired:
                rec = dict(case=c[0], group=c[10], desc_mode=c[11], error="timeout")
            rec["wall_s"] = round(time.time() - t0, 1)
            rec["req_split"] = sp_
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
