import torch
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
for cl_w in (True, False):
    conv = torch.nn.Conv2d(2048, 2048, 3, padding=1).cuda()
    if cl_w: conv = conv.to(memory_format=torch.channels_last)
    x = torch.relu(torch.randn(1, 2048, 28, 38, device="cuda"))
    xh = x.half().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        a = conv(x)
        with torch.autocast("cuda", dtype=torch.float16):
            h = conv(xh)
        b = conv(x)
        c = conv(x.contiguous(memory_format=torch.channels_last))
        ref = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double(), padding=1)
    e = lambda t: float((t.double() - ref).abs().max() / ref.abs().max())
    print(f"cl_weights={cl_w}: fp32 before {e(a):.2e}  fp16 {e(h):.2e}  fp32 after {e(b):.2e}  fp32 after (cl input) {e(c):.2e}")
