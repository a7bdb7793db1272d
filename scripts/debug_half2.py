import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from acan_b200.modules import SADecoder
g = np.load(os.path.join(ROOT, "tests/golden/sadecoder_1064.npz"))
hh, ww, b = (int(v) for v in g["hw"])
sd = synth.fill_state_dict({k: torch.empty(eval(s)) for k, s in zip(g["keys"], g["shapes"])})
def mk(cl):
    d = SADecoder(height=hh, width=ww); d.load_state_dict(sd); d = d.cuda().eval()
    return d.to(memory_format=torch.channels_last) if cl else d
x = synth.feature_map("sadecoder_1064.x", b, 2048, hh // 8, ww // 8).cuda()
xh = x.half().contiguous(memory_format=torch.channels_last)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    ref2, _ = mk(False)(x)
    dec = mk(True)
    with torch.autocast("cuda", dtype=torch.float16):
        o, _ = dec(xh)
    print("half vs ref2", synth.rel_err(o.float(), ref2))
    c, _ = dec.saconv(x); c2, _ = mk(False).saconv(x)
    print("saconv right after half path", synth.rel_err(c, c2))
    ra, _ = dec(x); print("1st fp32 call", synth.rel_err(ra, ref2))
    rb, _ = dec(x); print("2nd fp32 call", synth.rel_err(rb, ref2))
    dec = mk(True)
    with torch.autocast("cuda", dtype=torch.float16):
        o, _ = dec(xh)
    ra, _ = dec(x); print("fresh: 1st fp32 call directly after half", synth.rel_err(ra, ref2))
    dec = mk(True)
    with torch.autocast("cuda", dtype=torch.float16):
        o, _ = dec(xh)
        ra, _ = dec(x)
    print("fresh: fp32 input call INSIDE autocast", synth.rel_err(ra.float(), ref2), ra.dtype)
