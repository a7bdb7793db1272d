import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from acan_b200 import ops
torch.manual_seed(0)
B, h, w = 1, 28, 38
f = torch.relu(torch.randn(B, 512, h, w, device="cuda")); v = torch.relu(torch.randn(B, 2048, h, w, device="cuda"))
fh = f.half().contiguous(memory_format=torch.channels_last); vh = v.half().contiguous(memory_format=torch.channels_last)
ref = ops.attention_forward(f, v)["ctx"].clone()
def err(a): return float((a.float() - ref).abs().max() / ref.abs().max())
o1 = ops.attention_forward_cl(fh, vh); print("cl (own ws)", err(o1["ctx"]))
o2 = ops.attention_forward(f, v, ws=o1["ws"]); print("fp32 reusing cl ws", err(o2["ctx"]))
o3 = ops.attention_forward(f, v, ws=o1["ws"]); print("fp32 reusing again", err(o3["ctx"]))
o4 = ops.attention_forward(f, v); print("fp32 fresh ws", err(o4["ctx"]))
o5 = ops.attention_forward_cl(fh, vh, ws=o4["ws"]); print("cl reusing fp32 ws", err(o5["ctx"]))
o6 = ops.attention_forward(f, v, ws=o4["ws"]); print("fp32 reusing after cl", err(o6["ctx"]))
# channels-last fp32 inputs (strided) into the fp32 entry
fc = f.contiguous(memory_format=torch.channels_last); vc = v.contiguous(memory_format=torch.channels_last)
o7 = ops.attention_forward(fc, vc); print("fp32 entry with channels_last fp32 tensors", err(o7["ctx"]), o7["ctx"].stride())
