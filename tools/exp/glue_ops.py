"""Experiment: warm timings of the non-convolution glue ops of the DeepLab inference plan."""
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
B = 45
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
cl = torch.channels_last
x = torch.randn(B, 256, 32, 32, device="cuda").contiguous(memory_format=cl)
low = torch.randn(B, 48, 128, 128, device="cuda").contiguous(memory_format=cl)
stem = torch.randn(B, 64, 256, 256, device="cuda").contiguous(memory_format=cl)
dec = torch.randn(B, 9, 128, 128, device="cuda").contiguous(memory_format=cl)
print("upsample 32->128 nhwc 256ch : %.1f us" % t(lambda: F.interpolate(x, size=(128, 128), mode='bilinear', align_corners=True)))
up = F.interpolate(x, size=(128, 128), mode='bilinear', align_corners=True)
print("cat (256+48) nhwc           : %.1f us" % t(lambda: torch.cat((up, low), dim=1)))
print("maxpool 3x3 s2 nhwc 64ch    : %.1f us" % t(lambda: F.max_pool2d(stem, 3, stride=2, padding=1)))
print("dec.float().contiguous()    : %.1f us" % t(lambda: dec.float().contiguous()))
y = dec.float().contiguous()
print("final upsample 128->512 nchw: %.1f us" % t(lambda: F.interpolate(y, size=(512, 512), mode='bilinear', align_corners=True)))
print("ideal at 6.5 TB/s: up+cat %.0f us, maxpool %.0f us, final %.0f us" % ((896+142+47)/6.5, (755+189)/6.5, (425+26)/6.5))
