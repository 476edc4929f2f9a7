"""Experiment: warm timings of the odd-shaped convolutions of the inference plan (cuDNN, TF32, channels_last)."""
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
B = 45
cl = torch.channels_last
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
def conv(cin, cout, k, hw, stride=1, pad=0, relu=True):
    x = torch.randn(B, cin, hw, hw, device="cuda").contiguous(memory_format=cl)
    w = torch.randn(cout, cin, k, k, device="cuda").contiguous(memory_format=cl)
    b = torch.randn(cout, device="cuda")
    if relu:
        return t(lambda: torch.cudnn_convolution_relu(x, w, b, (stride, stride), (pad, pad), (1, 1), 1))
    return t(lambda: F.conv2d(x, w, b, stride, pad))
print("stem 7x7/2 3->64 @512      : %.0f us" % conv(3, 64, 7, 512, 2, 3))
print("dec_low 1x1 256->48 @128   : %.0f us" % conv(256, 48, 1, 128))
print("dec1 3x3 304->256 @128     : %.0f us" % conv(304, 256, 3, 128, 1, 1))
print("dec2 3x3 256->256 @128     : %.0f us" % conv(256, 256, 3, 128, 1, 1))
print("dec_out 1x1 256->9 @128    : %.0f us (no relu)" % conv(256, 9, 1, 128, relu=False))
print("layer1 1x1 64->256 @128    : %.0f us" % conv(64, 256, 1, 128))
print("layer1 3x3 64->64 @128     : %.0f us" % conv(64, 64, 3, 128, 1, 1))
print("s2d stem 4x4 12->64 @259   : %.0f us" % conv(12, 64, 4, 259))
print("s2d stem 4x4 16->64 @259   : %.0f us" % conv(16, 64, 4, 259))
print("stem 7x7/2 4->64 @512      : %.0f us" % conv(4, 64, 7, 512, 2, 3))
print("stem 7x7/2 8->64 @512      : %.0f us" % conv(8, 64, 7, 512, 2, 3))
print("dec1 3x3 320->256 @128     : %.0f us" % conv(320, 256, 3, 128, 1, 1))
print("dec1 3x3 384->256 @128     : %.0f us" % conv(384, 256, 3, 128, 1, 1))
