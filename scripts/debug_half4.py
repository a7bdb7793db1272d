import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from acan_b200.modules import SADecoder
from acan_b200 import ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
hh, ww, b = 224, 304, 1
torch.manual_seed(0)
dec = SADecoder(height=hh, width=ww).cuda().eval().to(memory_format=torch.channels_last)
x = torch.relu(torch.randn(b, 2048, hh // 8, ww // 8, device="cuda"))
xh = x.half().contiguous(memory_format=torch.channels_last)
sa = dec.saconv
with torch.no_grad():
    f0 = sa.key_features(x).float(); v0 = sa.f_value(x).float()
    c0 = ops.attention_forward(f0, v0)["ctx"]
    with torch.autocast("cuda", dtype=torch.float16):
        o, _ = dec(xh)
    f1 = sa.key_features(x).float(); v1 = sa.f_value(x).float()
    print("feat same", float((f1 - f0).abs().max()), "value same", float((v1 - v0).abs().max()))
    c1 = ops.attention_forward(f1, v1)["ctx"]; print("ops fresh ws", float((c1 - c0).abs().max() / c0.abs().max()))
    c2 = ops.attention_forward(f1, v1, ws=sa._holder.ws)["ctx"]; print("ops holder ws", float((c2 - c0).abs().max() / c0.abs().max()))
    c3, _ = sa(x); print("saconv module", float((c3 - c0).abs().max() / c0.abs().max()))
    from acan_b200 import functions as fn
    class H: pass
    c4, _ = fn.Attention.apply(f1, v1, None, H()); print("Attention.apply fresh holder", float((c4 - c0).abs().max() / c0.abs().max()))
    c5, _ = fn.Attention.apply(f1, v1, None, sa._holder); print("Attention.apply module holder", float((c5 - c0).abs().max() / c0.abs().max()))
print("---- parameter integrity")
dec = SADecoder(height=hh, width=ww).cuda().eval().to(memory_format=torch.channels_last)
snap = {k: v.clone() for k, v in dec.state_dict().items()}
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    feat = dec.saconv.key_features(xh); value = dec.saconv.f_value(xh)
torch.cuda.synchronize()
print("after convs:", [k for k, v in dec.state_dict().items() if not torch.equal(v, snap[k])])
with torch.no_grad():
    out = ops.attention_forward_cl(feat, value)
torch.cuda.synchronize()
print("after attention_cl:", [k for k, v in dec.state_dict().items() if not torch.equal(v, snap[k])])
ic = dec.image_context
with torch.no_grad():
    ops.pool_branch_cl(xh, ic.linear1.weight, ic.linear1.bias, ic.linear2.weight, ic.linear2.bias, None)
torch.cuda.synchronize()
print("after pool_cl:", [k for k, v in dec.state_dict().items() if not torch.equal(v, snap[k])])
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    o, _ = dec(xh)
torch.cuda.synchronize()
print("after full half path:", [k for k, v in dec.state_dict().items() if not torch.equal(v, snap[k])])
print("x intact:", torch.equal(x, torch.relu(x)), float(x.sum()))
