"""gpurun_out/sweep_n*.jsonl (tools_sweep.py) -> markdown table.  python profiles/summarize_sweep.py title in1.jsonl [in2.jsonl ...] > out.md"""
import json
import sys

title = sys.argv[1]
rows = []
for f in sys.argv[2:]:
    rows += [json.loads(l) for l in open(f) if l.strip()]
print(f"# {title}\n")
print("Each point is `bench.py --steps K --warmup 3` on that configuration (K in the `steps` column; records without it were taken")
print("with K = 2 by an earlier build, which is noted per table; inputs resident in HBM, deferred compression: one batched")
print("`rtk_pivot_update_batch` per chunk; N > 1: one video per GPU under torchrun, frames/s is the aggregate). `score` = the scoring of")
print("one layer (CUDA events around the batched scoring of a chunk inside the timed region / 28 layers), frac = algorithmic flops /")
print("measured sustained bf16 peak; `dpselect` = the whole operator, frac = algorithmic bytes / measured HBM peak.\n")
print("| shape | frames | GPUs | r_v | r_kv | steps | frames/s | per GPU | ms/step | score ms (frac) | dpselect ms (frac) | e2e frames/s | build |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    if "error" in r:
        print(f"| {r['shape']} | {r['frames']} | {r.get('n_gpus', 1)} | {r['rv']} | {r['rkv']} | | error | | | | | | |")
        continue
    n = r.get("n_gpus", 1)
    e2e = r.get("e2e_frames_per_s")
    print(f"| {r['shape']} | {r['frames']} | {n} | {r['visual_ratio']} | {r['kv_ratio']:.3f} | {r.get('steps', 2)} | {r['frames_per_s']:.0f} | {r['frames_per_s'] / n:.0f} | "
          f"{r['ms_per_step']:.1f} | {r['score_ms']:.3f} ({r['score_frac']:.3f}) | {r['dpselect_ms']:.3f} ({r['dpselect_frac']:.2f}) | "
          f"{'%.0f' % e2e if e2e else '-'} | {(r.get('build_id') or '')[:8]} |")
