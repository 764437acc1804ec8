"""GPU probe: the three DPSelect kernels at small / medium / headline sizes, run a few times each so that
`ncu --metrics gpu__time_duration.sum` lists their device times (BASELINE config 2 = 256 frames = T 128)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "video-retake_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from retake import visual_compression as vc
for T in (64, 128, 256, 1024):
    x = torch.randn(T, 256, 3584, device="cuda").to(torch.bfloat16)
    for t in (T, T // 2):
        for _ in range(3):
            vc.memory_bank_compress_keyframe(x[None], t, 3, sync=False)
    torch.cuda.synchronize()
    if os.environ.get("EVENTS") == "1":
        for t in (T, T // 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # keep the GPU queue ahead of the host: a long dummy kernel first
            big = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
            big.zero_()
            a.record()
            for _ in range(20):
                vc.memory_bank_compress_keyframe(x[None], t, 3, sync=False)
            b.record()
            torch.cuda.synchronize()
            by = 2.0 * T * 256 * 3584 + 4.0 * T * 256 + 4.0 * t * 256 * 3584
            ms = a.elapsed_time(b) / 20
            print(f"T={T} t={t}: {ms * 1e3:.1f} us per operator call, {by / ms / 1e6:.0f} GB/s ({by / ms / 1e6 / 6531.9:.2f} of measured peak)")
    del x
