#!/usr/bin/env python
"""Kernel sweep (BASELINE config 5): frames x (visual ratio, KV ratio) through bench.py on one GPU; writes
gpurun_out/sweep.jsonl (one bench line per point).  `python tools_sweep.py [--quick]`"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
points = []
for frames in (128, 256, 512, 1024, 2048):
    for rv, rkv in ((1.0, -1.0), (1.0, 0.25), (0.5, 0.5), (0.25, 1.0)):
        points.append(("qwen2vl", frames, rv, rkv))
points.append(("llava", 1024, 1.0, -1.0))
points.append(("llava", 2048, 1.0, -1.0))
if "--quick" in sys.argv:
    points = points[:2]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "w") as out:
    for shape, frames, rv, rkv in points:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--shape", shape, "--frames", str(frames), "--visual-ratio", str(rv),
               "--kv-ratio", str(rkv), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-immediate-ab"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            out.write(json.dumps({"shape": shape, "frames": frames, "rv": rv, "rkv": rkv, "error": r.stderr[-400:]}) + "\n")
            continue
        d = json.loads(line[-1])
        rec = {"shape": shape, "frames": frames, "visual_ratio": rv, "kv_ratio": d["config"]["kv_ratio"], "frames_per_s": d["value"],
               "ms_per_step": d["ms_per_step"], "score_ms": d["roofline"]["ms_per_call"], "score_frac": d["roofline"]["frac"],
               "dpselect_ms": d["roofline_dpselect"]["ms_per_call"], "dpselect_frac": d["roofline_dpselect"]["frac"]}
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(rec, flush=True)
