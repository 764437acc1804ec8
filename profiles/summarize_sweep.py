"""gpurun_out/sweep.jsonl (tools_sweep.py) -> markdown table.  python profiles/summarize_sweep.py in.jsonl > out.md"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
print("# Round 1 - kernel sweep on one B200 (BASELINE config 5 + LLaVA-Video shape), `python tools_sweep.py`\n")
print("Each point is `bench.py --steps 2 --warmup 3` on that configuration (inputs resident in HBM, deferred compression: one batched")
print("`rtk_pivot_update_batch` per chunk). `score` = the scoring of one layer (CUDA events around the batched scoring of a chunk inside the")
print("timed region / 28 layers), frac = algorithmic flops / measured sustained bf16 peak; `dpselect` = the whole operator,")
print("frac = algorithmic bytes / measured HBM peak (small videos are launch/latency dominated: 3 kernels for <= 0.5 GB).\n")
print("| shape | frames | r_v | r_kv | frames/s | ms/step | score ms (frac) | dpselect ms (frac) |")
print("|---|---|---|---|---|---|---|---|")
for r in rows:
    if "error" in r:
        print(f"| {r['shape']} | {r['frames']} | {r['rv']} | {r['rkv']} | error | | | |")
        continue
    print(f"| {r['shape']} | {r['frames']} | {r['visual_ratio']} | {r['kv_ratio']:.3f} | {r['frames_per_s']:.0f} | {r['ms_per_step']:.1f} | "
          f"{r['score_ms']:.3f} ({r['score_frac']:.3f}) | {r['dpselect_ms']:.3f} ({r['dpselect_frac']:.2f}) |")
