#!/usr/bin/env python
"""Kernel sweep (BASELINE config 5): frames x (visual ratio, KV ratio) through bench.py on N GPUs of one box; writes
gpurun_out/sweep_n<N>.jsonl (one record per point).

  python tools_sweep.py                       # 1 GPU, the full grid (20 Qwen2-VL points + 2 LLaVA-Video points)
  python tools_sweep.py --gpus 8 --corners    # 8 GPUs (one video per GPU, torchrun), the four corner points of the grid
  python tools_sweep.py --gpus 8 --llava      # BASELINE config 4: LLaVA-Video shape, 2048 frames, one video per GPU
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def arg(name, default=None):
    if name in sys.argv:
        i = sys.argv.index(name)
        return sys.argv[i + 1] if default is not None else True
    return default


gpus = int(arg("--gpus", "1"))
points = []
if arg("--llava"):
    points = [("llava", 2048, 1.0, -1.0)]
elif arg("--corners"):
    points = [("qwen2vl", 128, 0.5, 0.5), ("qwen2vl", 128, 0.25, 1.0), ("qwen2vl", 2048, 1.0, -1.0), ("qwen2vl", 2048, 0.5, 0.5)]
else:
    for frames in (128, 256, 512, 1024, 2048):
        for rv, rkv in ((1.0, -1.0), (1.0, 0.25), (0.5, 0.5), (0.25, 1.0)):
            points.append(("qwen2vl", frames, rv, rkv))
    points.append(("llava", 1024, 1.0, -1.0))
    points.append(("llava", 2048, 1.0, -1.0))
if arg("--quick"):
    points = points[:2]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
name = "sweep_n%d%s.jsonl" % (gpus, "_llava" if arg("--llava") else "")
with open(os.path.join(ROOT, "gpurun_out", name), "w") as out:
    for i, (shape, frames, rv, rkv) in enumerate(points):
        # enough steps for ~1 s of timed work: the first step after the synchronisation pays the clock ramp of an idle GPU
        # and the host's launch latency (several ms), which a two-step run of a short video would not amortise
        est_ms = frames * rv * (0.30 if shape == "qwen2vl" else 0.65)
        steps = max(3, min(40, int(1000.0 / est_ms) + 1))
        bench = [os.path.join(ROOT, "bench.py"), "--gpus", str(gpus), "--shape", shape, "--frames", str(frames), "--visual-ratio", str(rv),
                 "--kv-ratio", str(rkv), "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline", "--no-immediate-ab", "--no-sharded",
                 "--no-parity"]
        if gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus), "--master-addr",
                   "127.0.0.1", "--master-port", str(29600 + i)] + bench
        else:
            cmd = [sys.executable] + bench + ["--no-e2e"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            out.write(json.dumps({"shape": shape, "frames": frames, "rv": rv, "rkv": rkv, "n_gpus": gpus, "error": r.stderr[-400:]}) + "\n")
            continue
        d = json.loads(line[-1])
        rec = {"shape": shape, "frames": frames, "n_gpus": gpus, "visual_ratio": rv, "kv_ratio": d["config"]["kv_ratio"],
               "frames_per_s": d["value"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "score_ms": d["roofline"]["ms_per_call"],
               "score_frac": d["roofline"]["frac"], "dpselect_ms": d["roofline_dpselect"]["ms_per_call"],
               "dpselect_frac": d["roofline_dpselect"]["frac"], "e2e_frames_per_s": (d.get("e2e") or {}).get("value"),
               "clocks": d.get("clocks"), "build_id": d.get("build_id")}
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(rec, flush=True)
