"""GPU probe: A/B timing of rtk_pivot_score across library variants (build/ab/*/librtk_*.so) in one process."""
import ctypes as C, glob, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H, KVH, D = 28, 4, 128
L = int(os.environ.get("L", "4096"))
q = torch.randn(1, L, H, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
k = torch.randn(1, L, KVH, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
hs = torch.empty(KVH, L, dtype=torch.bfloat16, device="cuda")
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
libs = sorted(glob.glob(os.path.join(ROOT, "build", "ab", "librtk_*.so")))
res = {}
outs = {}
p, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
NCALLS = int(os.environ.get("NCALLS", "50"))
for rnd in range(int(os.environ.get("ROUNDS", "3"))):
    for path in libs:
        lib = C.CDLL(path)
        fn = lib.rtk_pivot_score
        fn.argtypes = [p, i64, i64, i64, p, i64, i64, i64, i64, i64, p, p, sz, p]
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: fn(q.data_ptr(), H, q.stride(1), q.stride(2), k.data_ptr(), KVH, k.stride(1), k.stride(2), L, D,
                          hs.data_ptr(), ws.data_ptr(), ws.numel(), st)
        for _ in range(5):
            assert call() == 0
        torch.cuda.synchronize()
        outs[os.path.basename(path)] = hs.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(NCALLS):
            call()
        b.record()
        torch.cuda.synchronize()
        res.setdefault(os.path.basename(path), []).append(a.elapsed_time(b) / NCALLS)
print(json.dumps(res, indent=1))
ref = outs[sorted(outs)[0]]
print({n: int((o.view(torch.int16) != ref.view(torch.int16)).sum()) for n, o in outs.items()}, 'mismatches vs first variant')
