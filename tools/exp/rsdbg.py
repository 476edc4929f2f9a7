import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pylc_oracle as orc
from pylc_b200 import ops, _lib
for (W, H, ch) in [(3000, 2000, 3), (6000, 4000, 1)]:
    w, h = orc.fit_dims(W, H, 512)
    raw, rp = ops.upload_image(orc.synth_image(2, W, H, ch))
    out = torch.empty((h, ops.pitch_for(w * ch)), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(5):
        flush.fill_(it)
        ops.fit_resize_area(raw, H, W, ch, rp, h, w, out=out)
    torch.cuda.synchronize()
    buf = np.zeros(8 * 2048, dtype=np.uint64)
    lib = _lib.load()
    lib.pylc_debug_read_rs.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    assert lib.pylc_debug_read_rs(buf.ctypes.data, buf.nbytes) == 0
    b = buf.reshape(2048, 8)
    n = int((b[:, 0] > 0).sum())
    b = b[:n].astype(np.int64)
    g = (b[:, :5] - b[:, 0].min()) / 1e3
    names = ["start", "tma issued", "setup done", "first row", "end"]
    print("%dx%d ch%d: %d CTAs" % (W, H, ch, n))
    for i, nm in enumerate(names):
        print("   %-12s min %6.2f  med %6.2f  max %6.2f us" % (nm, g[:, i].min(), np.median(g[:, i]), g[:, i].max()))
    print("   per-CTA lifetime med %.2f us; loop (first row -> end) med %.2f us" % (np.median(g[:, 4] - g[:, 0]), np.median(g[:, 4] - g[:, 3])))
