"""Experiment: TMA + f32x2 fit-resize (area_resize_x2_kernel) against the per-thread kernels and cv2."""
import os, sys, time
import numpy as np, torch, cv2
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pylc_oracle as orc
from pylc_b200 import ops
os.environ["PYLC_SYNC_CHECK"] = "1"
only = sys.argv[1:] or None
for (W, H, ch) in [(3000, 2000, 3), (2000, 1500, 1), (6000, 4000, 1), (1023, 700, 3), (640, 600, 1)]:
    if only and ("%dx%dx%d" % (W, H, ch)) not in only:
        continue
    w, h = orc.fit_dims(W, H, 512)
    rng = np.random.default_rng(W + ch)
    img = rng.integers(0, 256, size=(H, W) if ch == 1 else (H, W, 3), dtype=np.uint8)
    want = cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA)
    d_img, pitch = ops.upload_image(img)
    for no_tma in ("1", "0"):
        os.environ["PYLC_NO_TMA"] = no_tma
        try:
            out, po = ops.fit_resize_area(d_img, H, W, ch, pitch, h, w)
            torch.cuda.synchronize()
            got = out.cpu().numpy()[:, :w * ch].reshape(want.shape)
            ok = np.array_equal(got, want)
            nbad = int((got != want).sum())
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            os.environ["PYLC_SYNC_CHECK"] = "0"
            for _ in range(3):
                ops.fit_resize_area(d_img, H, W, ch, pitch, h, w, out=out)
            torch.cuda.synchronize()
            s.record()
            for _ in range(20):
                ops.fit_resize_area(d_img, H, W, ch, pitch, h, w, out=out)
            e.record(); e.synchronize()
            print("%dx%d ch%d -> %dx%d  NO_TMA=%s  exact=%s (bad %d)  %.1f us/launch (back to back)" % (W, H, ch, w, h, no_tma, ok, nbad, s.elapsed_time(e) / 20 * 1e3), flush=True)
        except Exception as ex:
            print("%dx%d ch%d NO_TMA=%s FAILED: %s" % (W, H, ch, no_tma, str(ex)[:200]), flush=True)
            sys.exit(1)
