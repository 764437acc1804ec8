"""GPU probe #2: discriminate the fp32 accumulation orders that probe #1 left open.

  Q6  linalg_vector_norm on a contiguous bf16 row: vec4 or vec8 lane mapping?  (many rows)
  Q7  full cosine chain at scale (>= 100k outputs) for (norm vec, sum vec) candidates
  Q8  bf16 mean over 7 (dim=1 of [4,7,L]): sequential vs 4-accumulator interleave
  Q9  fp32 mean over N (sync mode, dis.mean(1)): vec4 lane mapping?
"""
import json
import os

import torch

dev = os.environ.get("PROBE_DEV", "cuda")
out = {}
torch.manual_seed(1)
BF = torch.bfloat16


def emulate_rowsum(xf, vec, square=False):
    R, C = xf.shape
    nv = C // vec
    v = xf.view(R, nv, vec)
    acc = torch.zeros(R, 32, vec, device=xf.device, dtype=torch.float32)
    k = 0
    while k * 32 < nv:
        chunk = v[:, k * 32:(k + 1) * 32]
        acc[:, :chunk.shape[1]] = acc[:, :chunk.shape[1]] + (chunk * chunk if square else chunk)
        k += 1
    lane = acc[:, :, 0].clone()
    for i in range(1, vec):
        lane = lane + acc[:, :, i]
    off = 16
    while off > 0:
        lane = lane + torch.cat([lane[:, off:], lane[:, 32 - off:]], dim=1)
        off >>= 1
    return lane[:, 0]


big = dev == "cuda"
# Q6
q6 = {}
for C, rows in ((256, 1 << (22 if big else 14)), (3584, 1 << (18 if big else 10)), (1152, 1 << (19 if big else 10))):
    mism = {4: 0, 8: 0, "seq": 0}
    done = 0
    step = min(rows, 1 << 16)
    while done < rows:
        x = (torch.randn(step, C, device=dev) * (1 + 3 * torch.rand(step, 1, device=dev))).to(BF)
        ref = torch.linalg.vector_norm(x, 2, dim=-1)
        xf = x.float()
        for vec in (4, 8):
            e = torch.sqrt(emulate_rowsum(xf, vec, square=True)).to(BF)
            mism[vec] += int((e != ref).sum())
        mism["seq"] += int((torch.sqrt((xf.double() ** 2).sum(-1)).float().to(BF) != ref).sum())
        done += step
    q6[f"C{C}"] = {"rows": rows, "vec4": mism[4], "vec8": mism[8], "double": mism["seq"]}
out["Q6_norm_vec"] = q6

# Q7
q7 = {}
shapes = ((129, 1024, 256), (65, 256, 3584), (33, 729, 1152)) if big else ((9, 64, 256),)
for (T, N, C) in shapes:
    ent = {"outputs": 0}
    for trial in range(4 if big else 1):
        scene = torch.randn(1, 1, N, C, device=dev)
        X = (scene * (trial % 2) + torch.randn(1, T, N, C, device=dev) * (0.3 if trial % 2 else 1.0)).to(BF)
        sim_ref = torch.nn.functional.cosine_similarity(X[:, :-1], X[:, 1:], dim=-1)[0]
        xf = X[0].float().reshape(T * N, C)
        for vec in (4, 8):
            n = torch.sqrt(emulate_rowsum(xf, vec, square=True)).to(BF).float()
            n = torch.clamp_min(n, torch.tensor(1e-8, device=dev).to(BF).float())
            u = (xf * (1.0 / n[:, None])).to(BF).float().view(T, N, C)     # rcp-mul form used by the kernel
            p = (u[:-1] * u[1:]).to(BF).float().reshape((T - 1) * N, C)
            for vec2 in (4, 8):
                sim = emulate_rowsum(p, vec2).to(BF).view(T - 1, N)
                key = f"norm_vec{vec}_sum_vec{vec2}"
                ent[key] = ent.get(key, 0) + int((sim != sim_ref).sum())
        ent["outputs"] += (T - 1) * N
    q7[f"T{T}_N{N}_C{C}"] = ent
out["Q7_cosine_chain_scale"] = q7

# Q8
L = 1 << (22 if big else 14)
x = (torch.rand(4, 7, L, device=dev) * 3).to(BF)
ref = x.mean(1)
s = x.float()
seq = s[:, 0]
for i in range(1, 7):
    seq = seq + s[:, i]
inter = (((s[:, 0] + s[:, 4]) + (s[:, 1] + s[:, 5])) + (s[:, 2] + s[:, 6])) + s[:, 3]
rc = torch.tensor(1.0, device=dev) / torch.tensor(7.0, device=dev)
out["Q8_mean7"] = {
    "sequential": int(((seq * rc).to(BF) != ref).sum()),
    "interleave4": int(((inter * rc).to(BF) != ref).sum()),
    "interleave4_truediv": int(((inter / torch.tensor(7.0, device=dev)).to(BF) != ref).sum()),
    "n": ref.numel(),
}
x4 = (torch.rand(4, L, device=dev) * 3).to(BF)
ref = x4.mean(0)
s = x4.float()
out["Q8_mean4"] = {
    "sequential": int(((((s[0] + s[1]) + s[2]) + s[3]) * 0.25).to(BF).ne(ref).sum()),
    "pairwise": int((((s[0] + s[1]) + (s[2] + s[3])) * 0.25).to(BF).ne(ref).sum()),
    "n": ref.numel(),
}

# Q9
q9 = {}
for N in (256, 64, 1024):
    d = torch.rand(1 << (16 if big else 10), N, device=dev)
    ref = d.mean(1)
    rc = torch.tensor(1.0, device=dev) / torch.tensor(float(N), device=dev)
    ent = {}
    for vec in (2, 4, 8):
        ent[f"vec{vec}_mulrcp"] = int(((emulate_rowsum(d, vec) * rc) != ref).sum())
    ent["torch_sum_mulrcp"] = int(((d.sum(1) * rc) != ref).sum())
    ent["n"] = ref.numel()
    q9[f"N{N}"] = ent
out["Q9_mean_f32_rows"] = q9

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_aten_cuda2.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
