"""One forward + backward of the attention kernels at the Llama-3-8B micro-batch shape (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paddlenlp_b200 import ops  # noqa: E402

B, S, nh, kvh, d = 2, 4096, 32, 8, 128
ld = (nh + 2 * kvh) * d
qkv = torch.randn(B, S, ld, device="cuda").to(torch.bfloat16)
q = qkv[:, :, : nh * d].view(B, S, nh, d)
k = qkv[:, :, nh * d: (nh + kvh) * d].view(B, S, kvh, d)
v = qkv[:, :, (nh + kvh) * d:].view(B, S, kvh, d)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out, lse = ops.flash_attn_fwd(q, k, v)
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    dq = dqkv[:, :, : nh * d].view(B, S, nh, d)
    dk = dqkv[:, :, nh * d: (nh + kvh) * d].view(B, S, kvh, d)
    dv = dqkv[:, :, (nh + kvh) * d:].view(B, S, kvh, d)
    ops.flash_attn_bwd(q, k, v, out, dout, lse, dq, dk, dv)
torch.cuda.synchronize()
print("ok")
