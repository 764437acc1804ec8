"""GPU probe: characterise the ATen-CUDA semantics the kernels must replay.

Run on the B200 box (``gpurun -- python tests/probes/probe_aten_cuda.py``).  Writes a JSON
summary to ``gpurun_out/probe_aten_cuda.json``.  Nothing in the product or in
the test-suite imports this file; it documents how the rules written down in
DESIGN.md ("ATen-CUDA semantics") were established.

Questions:
  Q1  torch.topk tie rule on CUDA (set semantics): lowest index first among
      ties?  (float32 [N,T] rows, bf16 1-D rows, sorted=True/False)
  Q2  fp32 accumulation order of ``sum(dim=-1)`` / ``linalg_vector_norm`` on a
      contiguous bf16 row of C elements: which input_vec_size (2/4/8)?
  Q3  ``bf16_tensor / python_float``: a * fp32(1/fp32(b)) or true division or
      bf16-rounded divisor?
  Q4  ``bf16.mean(dim)``: bf16(fp32_sum * fp32(1/n)) or true division?
  Q5  softmax(dtype=float32) of bf16 logits: max/expf/div chain vs emulation.
"""
import json
import math
import os

import torch

dev = os.environ.get("PROBE_DEV", "cuda")
out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0) if dev == "cuda" else "cpu"}
torch.manual_seed(0)


# ----------------------------------------------------------------- Q1 topk ties
def lowest_index_topk(x, k):
    # stable descending sort == ties broken by ascending index
    idx = torch.sort(x.float(), dim=-1, descending=True, stable=True).indices[..., :k]
    return idx.sort(dim=-1).values


q1 = {}
for name, rows, n, k, dt, levels in [
    ("f32_rows_T64", 256, 64, 16, torch.float32, 5),
    ("f32_rows_T1024", 256, 1024, 256, torch.float32, 7),
    ("f32_rows_T2048", 729, 2048, 1024, torch.float32, 9),
    ("bf16_1d_L4096", 1, 4096, 1024, torch.bfloat16, 33),
    ("bf16_1d_L6272", 1, 6272, 1568, torch.bfloat16, 20),
    ("bf16_1d_L1024", 1, 1024, 512, torch.bfloat16, 4),
    ("f32_1d_T512", 1, 512, 128, torch.float32, 6),
]:
    ok_sorted = ok_unsorted = True
    for trial in range(8):
        x = torch.randint(0, levels, (rows, n), device=dev).to(dt) * 0.25
        if rows == 1:
            x = x[0]
        want = lowest_index_topk(x, k)
        got_s = torch.topk(x, k, dim=-1, sorted=True).indices.sort(dim=-1).values
        got_u = torch.topk(x, k, dim=-1, sorted=False).indices.sort(dim=-1).values
        ok_sorted &= bool((got_s == want).all())
        ok_unsorted &= bool((got_u == want).all())
    q1[name] = {"sorted_true_lowest_index": ok_sorted, "sorted_false_lowest_index": ok_unsorted}
# signed zeros / NaN ordering
z = torch.tensor([0.0, -0.0, 0.0, -0.0, 1.0, float("nan"), -1.0], device=dev)
q1["zeros_nan_top4"] = torch.topk(z, 4).indices.tolist()
out["Q1_topk"] = q1


# ------------------------------------------------------------ Q2 reduce order
def emulate_rowsum(xf, vec, square=False):
    """ATen Reduce.cuh order for one warp per row: lane l owns vectors l, l+32, ...;
    vec accumulators per lane; accumulators combined 0+1+..; shfl_down tree 16..1."""
    R, C = xf.shape
    nv = C // vec
    assert C % vec == 0
    acc = torch.zeros(R, 32, vec, device=xf.device, dtype=torch.float32)
    v = xf.view(R, nv, vec)
    k = 0
    while k * 32 < nv:
        chunk = v[:, k * 32:(k + 1) * 32]
        lanes = chunk.shape[1]
        val = chunk * chunk if square else chunk
        acc[:, :lanes] = acc[:, :lanes] + val
        k += 1
    lane = acc[:, :, 0].clone()
    for i in range(1, vec):
        lane = lane + acc[:, :, i]
    off = 16
    while off > 0:
        sh = torch.cat([lane[:, off:], lane[:, 32 - off:]], dim=1)  # shfl_down (upper lanes keep own)
        lane = lane + sh
        off >>= 1
    return lane[:, 0]


q2 = {}
for C in (256, 1152, 3584):
    x = torch.randn(4096, C, device=dev).to(torch.bfloat16)
    xf = x.float()
    s_ref = torch.sum(x, dim=-1, dtype=torch.float32)
    s_bf = torch.sum(x, dim=-1)
    n_bf = torch.linalg.vector_norm(x, 2, dim=-1)
    n_f32in = torch.linalg.vector_norm(xf, 2, dim=-1)
    ent = {}
    for vec in (2, 4, 8):
        e = emulate_rowsum(xf, vec)
        ent[f"sum_f32out_vec{vec}_mismatch"] = int((e != s_ref).sum())
        ent[f"sum_bf16out_vec{vec}_mismatch"] = int((e.to(torch.bfloat16) != s_bf).sum())
        e2 = torch.sqrt(emulate_rowsum(xf, vec, square=True))
        ent[f"norm_bf16_vec{vec}_mismatch"] = int((e2.to(torch.bfloat16) != n_bf).sum())
        ent[f"norm_f32in_vec{vec}_mismatch"] = int((e2 != n_f32in).sum())
    ent["rows"] = 4096
    q2[f"C{C}"] = ent
out["Q2_reduce_order"] = q2

# full cosine chain vs F.cosine_similarity for each vec candidate
q2b = {}
for (T, N, C) in ((33, 64, 256), (17, 256, 3584), (9, 729, 1152)):
    X = torch.randn(1, T, N, C, device=dev).to(torch.bfloat16)
    sim_ref = torch.nn.functional.cosine_similarity(X[:, :-1], X[:, 1:], dim=-1)[0]
    xf = X[0].float().reshape(T * N, C)
    ent = {}
    for vec in (4, 8):
        n = torch.sqrt(emulate_rowsum(xf, vec, square=True)).to(torch.bfloat16).float()
        n = torch.clamp_min(n, torch.tensor(1e-8, device=dev).to(torch.bfloat16).float())
        u = (xf / n[:, None]).to(torch.bfloat16).float().view(T, N, C)
        p = (u[:-1] * u[1:]).to(torch.bfloat16).float().reshape((T - 1) * N, C)
        for vec2 in (4, 8):
            sim = emulate_rowsum(p, vec2).to(torch.bfloat16).view(T - 1, N)
            ent[f"norm_vec{vec}_sum_vec{vec2}_mismatch"] = int((sim != sim_ref).sum())
    ent["outputs"] = (T - 1) * N
    q2b[f"T{T}_N{N}_C{C}"] = ent
out["Q2b_cosine_chain"] = q2b

# ------------------------------------------------------- Q3 div by python float
x = torch.randn(1 << 20, device=dev).to(torch.bfloat16) * 40
for D in (64, 128):
    b = math.sqrt(D)
    ref = x / b
    xf = x.float()
    bf = torch.tensor(b, device=dev, dtype=torch.float32)
    cand = {
        "mul_rcp_f32": (xf * (1.0 / bf)).to(torch.bfloat16),
        "true_div_f32": (xf / bf).to(torch.bfloat16),
        "div_bf16_divisor": (xf / bf.to(torch.bfloat16).float()).to(torch.bfloat16),
        "mul_rcp_of_bf16_divisor": (xf * (1.0 / bf.to(torch.bfloat16).float())).to(torch.bfloat16),
    }
    out[f"Q3_div_scalar_D{D}"] = {k: int((v != ref).sum()) for k, v in cand.items()}

# ------------------------------------------------------------------ Q4 mean
x = torch.rand(4, 7, 4096, device=dev).to(torch.bfloat16) * 3
ref = x.mean(1)
s = x.float()
acc = s[:, 0]
for i in range(1, 7):
    acc = acc + s[:, i]
out["Q4_mean7"] = {
    "seq_sum_mul_rcp": int(((acc * torch.tensor(1.0 / 7.0, device=dev, dtype=torch.float32)).to(torch.bfloat16) != ref).sum()),
    "seq_sum_true_div": int(((acc / 7.0).to(torch.bfloat16) != ref).sum()),
    "n": ref.numel(),
}
ref0 = ref.mean(0)
r = ref.float()
acc = r[0]
for i in range(1, 4):
    acc = acc + r[i]
out["Q4_mean4"] = {
    "seq_sum_mul_rcp": int(((acc * 0.25).to(torch.bfloat16) != ref0).sum()),
    "n": ref0.numel(),
}
# column sum over q (dim=1 of [H, Lq, Lk]) : order sensitivity
pm = torch.rand(4, 1024, 1024, device=dev).to(torch.bfloat16) * 0.01
ref = pm.sum(1)
seq = torch.zeros(4, 1024, device=dev)
for i in range(1024):
    seq = seq + pm[:, i].float()
dbl = pm.double().sum(1)
out["Q4_colsum"] = {
    "seq_order_mismatch": int((seq.to(torch.bfloat16) != ref).sum()),
    "double_mismatch": int((dbl.to(torch.bfloat16) != ref).sum()),
    "n": ref.numel(),
}

# --------------------------------------------------------------- Q5 softmax
for L in (1024, 4096):
    s = (torch.randn(64, L, device=dev) * 3).to(torch.bfloat16)
    ref = torch.softmax(s, dim=-1, dtype=torch.float32)
    sf = s.float()
    m = sf.max(-1, keepdim=True).values
    e = torch.exp(sf - m)
    l32 = e.sum(-1, keepdim=True)
    l64 = e.double().sum(-1, keepdim=True).float()
    p32 = e / l32
    p64 = e / l64
    out[f"Q5_softmax_L{L}"] = {
        "f32_exact_mismatch_sum32": int((p32 != ref).sum()),
        "f32_exact_mismatch_sum64": int((p64 != ref).sum()),
        "bf16_mismatch_sum32": int((p32.to(torch.bfloat16) != ref.to(torch.bfloat16)).sum()),
        "bf16_mismatch_sum64": int((p64.to(torch.bfloat16) != ref.to(torch.bfloat16)).sum()),
        "exp2_path_bf16_mismatch": int(((torch.exp2((sf - m) * 1.4426950408889634) / l64).to(torch.bfloat16) != ref.to(torch.bfloat16)).sum()),
        "n": ref.numel(),
    }

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_aten_cuda.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
