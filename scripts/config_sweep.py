#!/usr/bin/env python
"""One measured line per BASELINE.json config that is not bench.py's default (configs[1]) or its --workload train (configs[3]):
   configs[0]  CAM (SADecoder) forward, batch 1, 256x352 -- on the B200 (fp32 drop-in module, and the fp16 channels_last inference path)
               and the reference CPU path (oracle port) on the host cores beside it
   configs[2]  full ACAN forward, KITTI-shaped 160x512, batch 32, HARD ordinal inference (fused tail with bit-exact bin indices)
python scripts/config_sweep.py  -> JSON lines"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from acan_b200.modules import SADecoder
from acan_b200.inference import GraphedPredictor
from oracle import acan_oracle as O

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)

def ev(fn_, iters=20, warm=5):
    for _ in range(warm): fn_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn_()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3

# ---- configs[0]
torch.manual_seed(2019)
dec = SADecoder(height=256, width=352).to(dev).eval()
x = torch.relu(torch.randn(1, 2048, 32, 44, device=dev))
with torch.no_grad():
    t32 = ev(lambda: dec(x))
    dech = dec.half().to(memory_format=torch.channels_last)
    xh = x.half().contiguous(memory_format=torch.channels_last)
    t16 = ev(lambda: dech(xh))
sd = {k: v.detach().float().cpu() for k, v in dec.float().state_dict().items()}
xc = x.float().cpu()
threads = os.cpu_count() or 1
torch.set_num_threads(threads)
with torch.no_grad():
    O.sadecoder_forward(xc, sd)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); O.sadecoder_forward(xc, sd); ts.append(time.perf_counter() - t0)
print(json.dumps({"config": "configs[0]: CAM forward, batch 1, 256x352 (N=1408)", "gpu_fp32_module_ms": t32 * 1e3, "gpu_fp16_channels_last_ms": t16 * 1e3,
                  "cpu_reference_port_ms": min(ts) * 1e3, "cpu_cores": threads, "speedup_fp32_vs_cpu": min(ts) / t32}), flush=True)

# ---- configs[2]
bench.H, bench.W, bench.BATCH = 160, 512, 32
bench.DMIN, bench.DMAX = 1.98, 80.0
net = bench.build_net(dev)
net.inferenceType = "hard"
imgs = [torch.rand(32, 3, 160, 512, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(4)]
pred = GraphedPredictor(net, imgs[0])
i = [0]
def step():
    i[0] += 1
    return pred.predict(imgs[i[0] % 4])
t = ev(step)
host = [torch.rand(32, 3, 160, 512).contiguous(memory_format=torch.channels_last).pin_memory() for _ in range(4)]
for k in range(3): pred.submit(host[k % 4])
pred.flush(); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20): pred.submit(host[k % 4])
pred.flush(); torch.cuda.synchronize()
te = (time.perf_counter() - t0) / 20
print(json.dumps({"config": "configs[2]: full ACAN forward, KITTI 160x512, batch 32, hard ordinal inference (N=1280)", "ms_per_step": t * 1e3,
                  "images_per_s": 32 / t, "e2e_images_per_s": 32 / te}), flush=True)
