"""Experiment: ASPP tail, stock (up-sample pooled branch + concat of five + fused 1x1 conv) against the folded form."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pylc_b200.models.deeplab import DeepLab
from pylc_b200.models.fused import FusedDeepLab
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = DeepLab(9).cuda().eval().to(memory_format=torch.channels_last)
plan = FusedDeepLab(net, channels_last=True)
B = 45
brs = [torch.randn(B, 256, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last) for _ in range(4)]
pooled = torch.randn(B, 256, 1, 1, device="cuda")
def stock():
    p = F.interpolate(pooled, size=(32, 32), mode="bilinear", align_corners=True)
    return plan.aspp_out(torch.cat(brs + [p], dim=1))
def folded():
    return plan._aspp_out_folded(torch.cat(brs, dim=1), pooled)
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / n * 1e3
with torch.no_grad():
    a, b = stock(), folded()
    print("max diff %.3e of %.3e" % (float((a - b).abs().max()), float(a.abs().max())))
    print("stock  %.0f us" % t(stock)); print("folded %.0f us" % t(folded))
