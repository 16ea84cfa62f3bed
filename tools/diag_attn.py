"""diagnose one attention configuration: python tools/diag_attn.py {fwd|bwd} HD NG G LENS..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dolomite_engine_b200 import kernels as k

mode, hd, ng, g = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lens = [int(x) for x in sys.argv[5:]]
T = sum(lens)
torch.manual_seed(0)
qkv = torch.randn(T, ng * (g + 2) * hd, device="cuda").bfloat16()
cu = torch.tensor([0] + list(__import__("itertools").accumulate(lens)), dtype=torch.int32, device="cuda")
t0 = time.time()
out, lse = k.attn_varlen_fwd(qkv, cu, max(lens), ng, g, hd, hd ** -0.5)
torch.cuda.synchronize()
print(f"fwd done {time.time()-t0:.2f}s out absmax {out.float().abs().max().item():.3f} finite {torch.isfinite(out.float()).all().item()}", flush=True)
if mode == "bwd":
    dout = torch.randn(T, ng * g * hd, device="cuda").bfloat16()
    t0 = time.time()
    d = k.attn_varlen_bwd(dout, qkv, out, lse, cu, max(lens), ng, g, hd, hd ** -0.5)
    torch.cuda.synchronize()
    print(f"bwd done {time.time()-t0:.2f}s absmax {d.float().abs().max().item():.3f} finite {torch.isfinite(d.float()).all().item()}", flush=True)
