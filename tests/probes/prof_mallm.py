"""GPU probe: fused MA-LLM compression (rtk_mallm_compress) vs the reference's torch op loop on the same GPU."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200")):
    sys.path.insert(0, p)
import torch
from retake import visual_compression as vc
from oracle import reference_ops as ro

def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = {}
g = torch.Generator(device="cuda").manual_seed(0)
for name, T, N, C, t in (("qwen_256f", 128, 256, 3584, 64), ("qwen_1024f", 512, 256, 3584, 256), ("llava_64f", 64, 729, 1152, 32)):
    x = torch.randn(1, T, N, C, generator=g, device="cuda").to(torch.bfloat16)
    for sync in (False, True):
        for hard in (False, True):
            ours = ev_time(lambda: vc.mallm_compress(x, t, sync=sync, hard=hard), 3)
            ref = ev_time(lambda: ro.mallm_compress(x, t, sync, hard), 1)
            got, gs = vc.mallm_compress(x, t, sync=sync, hard=hard)
            want, ws = ro.mallm_compress(x, t, sync, hard)
            out[f"{name}_sync{int(sync)}_hard{int(hard)}"] = {
                "T": T, "N": N, "C": C, "t": t, "rounds": T - t, "ours_ms": ours, "torch_cuda_ops_ms": ref, "speedup": ref / ours,
                "bit_identical": bool(torch.equal(got, want)) and (hard or bool(torch.equal(gs, ws))),
                "bank_bytes": 2 * T * N * C, "bank_reads_equiv": ours * 1e-3 * 6.5e12 / (2 * T * N * C)}
            print(name, sync, hard, out[f"{name}_sync{int(sync)}_hard{int(hard)}"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "mallm_timing.json"), "w"), indent=1)
