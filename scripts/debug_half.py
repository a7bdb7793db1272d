import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from acan_b200.modules import SADecoder
from acan_b200 import ops
g = np.load(os.path.join(ROOT, "tests/golden/sadecoder_1064.npz"))
T = torch.from_numpy
hh, ww, b = (int(v) for v in g["hw"])
dec = SADecoder(height=hh, width=ww)
sd = synth.fill_state_dict({k: torch.empty(eval(s)) for k, s in zip(g["keys"], g["shapes"])})
dec.load_state_dict(sd)
dec = dec.cuda().eval().to(memory_format=torch.channels_last)
x = synth.feature_map("sadecoder_1064.x", b, 2048, hh // 8, ww // 8).cuda()
xh = x.half().contiguous(memory_format=torch.channels_last)
def chk(name, t, key):
    got = t.float().reshape(-1)[T(g[key + "_idx"]).cuda()]
    print(name, synth.rel_err(got, T(g[key + "_val"])))
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    feat = dec.saconv.key_features(xh); value = dec.saconv.f_value(xh)
    chk("feat", feat, "feat"); chk("value", value, "val")
    out = ops.attention_forward_cl(feat, value, want_sim_map=True)
    chk("ctx", out["ctx"].contiguous(), "ctx"); chk("sim", out["sim_map"], "sim")
    ic = dec.image_context
    _, _, gg = ops.pool_branch_cl(xh, ic.linear1.weight, ic.linear1.bias, ic.linear2.weight, ic.linear2.bias, None)
    print("g", synth.rel_err(gg, T(g["g"])))
    o, sim = dec(xh)
    chk("out", o.contiguous(), "out")
    print(o.shape, o.stride(), o.dtype)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    ref, _ = dec(x)
chk("ref(fp32 path, cl weights)", ref.contiguous(), "out")
print("out vs ref", synth.rel_err(o.float(), ref))
err = (o.float() - ref).abs()
idx = err.flatten().argmax().item()
print("worst at", idx, np.unravel_index(idx, tuple(ref.shape)), float(o.float().flatten()[idx]), float(ref.flatten()[idx]), "err count >1e-2*max", int((err > 1e-2 * ref.abs().max()).sum()))
dec2 = SADecoder(height=hh, width=ww); dec2.load_state_dict(sd); dec2 = dec2.cuda().eval()
with torch.no_grad():
    ref2, _ = dec2(x)
print("ref vs ref2(NCHW weights)", synth.rel_err(ref, ref2), " out vs ref2", synth.rel_err(o.float(), ref2))
with torch.no_grad():
    for name, fa, fb in (("f_key", dec.saconv.f_key, dec2.saconv.f_key), ("f_value", dec.saconv.f_value, dec2.saconv.f_value)):
        a, b2 = fa(x.clone()), fb(x.clone())
        print(name, "cl-weights vs nchw-weights", synth.rel_err(a, b2), a.stride(), b2.stride())
    a = dec.saconv.f_value.conv1(x); b2 = dec2.saconv.f_value.conv1(x); print("conv1 3x3", synth.rel_err(a, b2))
    xc = x.contiguous(memory_format=torch.channels_last)
    a = dec.saconv.f_value.conv1(xc); print("conv1 3x3 cl input", synth.rel_err(a, b2))
    print("weights equal", all(torch.equal(p, q) for p, q in zip(dec.state_dict().values(), dec2.state_dict().values())))
print("---- piecewise fp32 path, cl-weight module (after half path ran) vs fresh NCHW module")
with torch.no_grad():
    c1, s1 = dec.saconv(x); c2, s2 = dec2.saconv(x)
    print("context", synth.rel_err(c1, c2), c1.stride(), c2.stride())
    dec3 = SADecoder(height=hh, width=ww); dec3.load_state_dict(sd); dec3 = dec3.cuda().eval().to(memory_format=torch.channels_last)
    c3, _ = dec3.saconv(x)
    print("context fresh cl module", synth.rel_err(c3, c2))
    o3, _ = dec3(x); print("out fresh cl module", synth.rel_err(o3, ref2))
    from acan_b200 import functions as fn
    ic = dec.image_context
    g1 = fn.PoolBranch.apply(x.float(), ic.linear1.weight, ic.linear1.bias, ic.linear2.weight, ic.linear2.bias, None)
    ic2 = dec2.image_context
    g2 = fn.PoolBranch.apply(x.float(), ic2.linear1.weight, ic2.linear1.bias, ic2.linear2.weight, ic2.linear2.bias, None)
    print("g", synth.rel_err(g1, g2))
    w = dec.merge.conv1.weight.view(2048, -1); w2 = dec2.merge.conv1.weight.view(2048, -1)
    print("w view equal", torch.equal(w, w2), dec.merge.conv1.weight.stride(), dec2.merge.conv1.weight.stride())
    y1 = torch.nn.functional.conv2d(c2, dec._ctx_weight(w, 2048)); y2 = torch.nn.functional.conv2d(c2, dec2._ctx_weight(w2, 2048))
    print("merge conv", synth.rel_err(y1, y2))
