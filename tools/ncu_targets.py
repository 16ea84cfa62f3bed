"""Kernels for `ncu --set full` captures (one launch of each after a warm-up), at the C2 benchmark shapes:

    ncu --set full --clock-control none --import-source on -k regex:'attn_|gemm_bf16|ce_rows' --launch-skip N -c M \
        -o gpurun_out/r02_kernels python tools/ncu_targets.py

Order of the profiled launches (after the warm-up pass): attention forward, attention backward (delta, pipelined kernel,
dq_finalize), GEMM forward c_fc, GEMM dgrad c_fc, the four weight gradients of a block in one launch, cross entropy."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dolomite_engine_b200 import kernels as k

hd = int(os.environ.get("NCU_HD", "80"))
S, B, nh = (4096, 2, 32) if hd == 80 else (8192, 1, 32)
T, H, F = S * B, nh * hd, 10240
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(T, 3 * H, device="cuda", generator=g).bfloat16()
dout = torch.randn(T, H, device="cuda", generator=g).bfloat16()
cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
x = (torch.randn(T, 2560, device="cuda", generator=g) * 0.1).bfloat16()
w_fc = (torch.randn(2 * F, 2560, device="cuda", generator=g) * 0.02).bfloat16()
d_fc = (torch.randn(T, 2 * F, device="cuda", generator=g) * 0.1).bfloat16()
probs = []
for M, N in ((2560, F), (2 * F, 2560), (2560, 2560), (7680, 2560)):
    probs.append(((torch.randn(T, M, device="cuda", generator=g) * 0.1).bfloat16(),
                  (torch.randn(T, N, device="cuda", generator=g) * 0.1).bfloat16(), torch.zeros(M, N, device="cuda"), 1.0, True))
V = 49152
logits = torch.randn(4096, V, device="cuda", generator=g).bfloat16()
labels = torch.randint(0, V, (4096,), device="cuda", generator=g)
scratch = k.cross_entropy_count(labels)
loss_tok = torch.empty(4096, device="cuda")


def once():
    out, lse = k.attn_varlen_fwd(qkv, cu, S, nh, 1, hd, hd**-0.5)
    k.attn_varlen_bwd(dout, qkv, out, lse, cu, S, nh, 1, hd, hd**-0.5)
    k.gemm(x, w_fc)
    k.gemm(d_fc, w_fc, b_mn=True)
    k.gemm_wgrad_multi(probs)
    k.cross_entropy_rows(logits, labels, loss_tok, scratch)


once()
torch.cuda.synchronize()
once()
torch.cuda.synchronize()
print("ncu targets done")
