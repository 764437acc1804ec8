"""GPU probe: board power and SM clock while rtk_pivot_score runs back to back, for library variants
(build/ab/librtk_*.so: default, NOMATH = TMA + MMA + TMEM loads only, NOLOAD = no TMEM reads, NOEXP = no MUFU).
Answers: which part of the scoring kernel drives the GPU into its power cap (bench.py: clocks.reasons sw_power_cap)."""
import ctypes as C, glob, json, os, statistics, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H, KVH, D, L = 28, 4, 128, 4096
SECONDS = float(os.environ.get("SECONDS_PER_VARIANT", "4"))
q = torch.randn(1, L, H, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
k = torch.randn(1, L, KVH, D, device="cuda").to(torch.bfloat16).transpose(1, 2)
hs = torch.empty(KVH, L, dtype=torch.bfloat16, device="cuda")
ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
p, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
res = {}
for path in sorted(glob.glob(os.path.join(ROOT, "build", "ab", "librtk_*.so"))):
    lib = C.CDLL(path)
    fn = lib.rtk_pivot_score
    fn.argtypes = [p, i64, i64, i64, p, i64, i64, i64, i64, i64, p, p, sz, p]
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: fn(q.data_ptr(), H, q.stride(1), q.stride(2), k.data_ptr(), KVH, k.stride(1), k.stride(2), L, D,
                      hs.data_ptr(), ws.data_ptr(), ws.numel(), st)
    for _ in range(20):
        assert call() == 0
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks_throttle_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                            "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    a.record()
    while time.time() - t0 < SECONDS:
        for _ in range(200):
            call()
        n += 200
        torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    smi.terminate()
    rows = [l.split(",") for l in smi.communicate()[0].strip().splitlines() if l.count(",") == 2]
    rows = rows[len(rows) // 3:]                      # drop the ramp
    pw = [float(r[0]) for r in rows]
    ck = [float(r[1]) for r in rows]
    cap = sum("Active" in r[2] and "Not" not in r[2] for r in rows)
    res[os.path.basename(path)] = {"ms_per_call": a.elapsed_time(b) / n, "power_w_median": statistics.median(pw) if pw else None,
                                   "power_w_max": max(pw) if pw else None, "sm_mhz_median": statistics.median(ck) if ck else None,
                                   "samples": len(rows), "power_cap_samples": cap}
    time.sleep(1.0)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_power.json"), "w"), indent=1)
