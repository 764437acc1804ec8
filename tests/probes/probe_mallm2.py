"""GPU probe: vector width of ATen's bf16 mean(-1) (many rows so that the final bf16 rounding does not hide the order)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from oracle import mallm
from oracle.dpselect import _r
g = torch.Generator().manual_seed(5)
out = {}
for n in (64, 128, 136, 256, 729, 1000):
    R = 150000
    v = torch.randn(R, n, generator=g).to(torch.bfloat16)
    gpu16 = v.cuda().mean(-1).float().cpu()
    gpu32 = v.cuda().mean(-1, dtype=torch.float32).cpu()
    row = {}
    rcp = torch.tensor(1.0) / torch.tensor(float(n))
    sums = {vec: mallm.aten_cuda_rowsum_general(v.float(), vec, 2) for vec in (4, 8)}
    for vec in (4, 8):
        m32 = sums[vec] * rcp
        row[f"vec{vec}_bf16out_mismatch"] = int((_r(m32) != gpu16).sum())
        row[f"vec{vec}_f32out_mismatch"] = int((m32 != gpu32).sum())
    row["models_differ_bf16"] = int((_r(sums[4] * rcp) != _r(sums[8] * rcp)).sum())
    out[n] = row
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_mallm2.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
