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

# cold-cache mode (COLD=1): L2 flushed before every timed call, the GPU queue kept ahead of the host, per-kernel events
if os.environ.get("COLD") == "1":
    from retake import _native as N
    lib = N.lib()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    for T in (64, 128, 256, 512):
        x = torch.randn(T, 256, 3584, device="cuda").to(torch.bfloat16)
        for t in (T, T // 2):
            dis = torch.empty(T, 256, dtype=torch.float32, device="cuda")
            idx = torch.empty(t, 256, dtype=torch.int32, device="cuda")
            mask = torch.empty(t * 256, dtype=torch.bool, device="cuda")
            out = torch.empty(t, 256, 3584, dtype=torch.bfloat16, device="cuda")
            acc = [0.0, 0.0, 0.0, 0.0]
            n = 8
            for it in range(n + 2):
                flush.zero_(); flush.zero_()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                ev[0].record()
                lib.rtk_dpselect_dis(x.data_ptr(), T, 256, 3584, 0, dis.data_ptr(), st); ev[1].record()
                lib.rtk_dpselect_select(dis.data_ptr(), T, 256, t, 0, idx.data_ptr(), mask.data_ptr(), st); ev[2].record()
                lib.rtk_dpselect_gather(x.data_ptr(), T, 256, 3584, idx.data_ptr(), t, 0, out.data_ptr(), st); ev[3].record()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.rtk_dpselect_keyframe(x.data_ptr(), T, 256, 3584, t, 0, dis.data_ptr(), idx.data_ptr(), mask.data_ptr(), out.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    for j in range(3):
                        acc[j] += ev[j].elapsed_time(ev[j + 1]) * 1e3 / n
                    acc[3] += e0.elapsed_time(e1) * 1e3 / n
            by = 2.0 * T * 256 * 3584 + 4.0 * T * 256 + 4.0 * t * 256 * 3584
            print(f"COLD T={T} t={t}: dis {acc[0]:.1f} us, select {acc[1]:.1f} us, gather {acc[2]:.1f} us, one-call operator {acc[3]:.1f} us "
                  f"= {by / acc[3] / 1e3 / 6531.9:.2f} of peak")
        del x
