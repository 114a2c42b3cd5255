"""One launch of the tcgen05 attention kernel per shape, for `ncu --set full -k regex:flash_attn_tc`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitron_b200 import ops

dev = torch.device("cuda:0")
ops.set_attention_impl(2)
for B, H, S, D, causal in ((16, 5, 2560, 64, False), (1, 32, 1728, 128, True)):
    q, k, v = (torch.randn((B, S, H, D), device=dev).to(torch.bfloat16) for _ in range(3))
    for _ in range(2):
        ops.attention(q, k, v, causal=causal)
    torch.cuda.synchronize()
print("watchdog", ops.attention_watchdog())
