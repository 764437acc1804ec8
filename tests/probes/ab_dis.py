"""GPU probe: A/B of rtk_dpselect_dis across library variants (build/ab/librtk_*.so) in one process."""
import ctypes as C, glob, json, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
T, N, Cc = 1024, 256, 3584
x = torch.randn(T, N, Cc, device="cuda").to(torch.bfloat16)
dis = torch.empty(T, N, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for rnd in range(3):
    for path in sorted(glob.glob(os.path.join(ROOT, "build", "ab", "librtk_*.so"))):
        lib = C.CDLL(path)
        fn = lib.rtk_dpselect_dis
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        ts = []
        for i in range(13):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            assert fn(x.data_ptr(), T, N, Cc, 0, dis.data_ptr(), st) == 0
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ts.sort()
        res.setdefault(os.path.basename(path), []).append(ts[len(ts) // 2])
print(json.dumps(res, indent=1))
