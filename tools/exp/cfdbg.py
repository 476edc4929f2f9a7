import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pylc_oracle as orc
from pylc_b200 import ops, _lib
from pylc_b200.config import Parameters
pal = Parameters().palette_rgb; C = 9
for (Wb, Hb, wb, hb) in [(3000, 2000, 2560, 1536), (6000, 4000, 5632, 3584)]:
    dmb, dpb = ops.upload_image(orc.synth_mask(30, Wb, Hb, pal, skew=True))
    lab_b = torch.from_numpy(orc.synth_labels(31, wb, hb, C, skew=True, block=37)).cuda()
    maps_b = (torch.from_numpy(ops.nn_index_map(wb, Wb)).cuda(), torch.from_numpy(ops.nn_index_map(hb, Hb)).cuda())
    conf = torch.zeros((C, C), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(5):
        flush.fill_(it)
        ops.resample_encode_confusion(lab_b, Wb, Hb, gt_rgb=dmb, gt_pitch=dpb, palette=pal, n_inject=C, conf=conf, maps=maps_b)
    torch.cuda.synchronize()
    buf = np.zeros(8 * 512, dtype=np.uint64)
    lib = _lib.load()
    lib.pylc_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    assert lib.pylc_debug_read(buf.ctypes.data, buf.nbytes) == 0
    b = buf.reshape(512, 8)
    n = int((b[:, 0] > 0).sum())
    b = b[:n].astype(np.int64)
    t0 = b[:, 0].min()
    g = (b[:, :4] - t0) / 1e3
    c = (b[:, 4:] - b[:, 4:5]) / 1.9e3
    print(f"{Wb}x{Hb}: {n} CTAs; globaltimer us: start min/med/max {g[:,0].min():.2f}/{np.median(g[:,0]):.2f}/{g[:,0].max():.2f}; setup done {np.median(g[:,1]):.2f} (max {g[:,1].max():.2f}); first box med {np.median(g[:,2]):.2f} max {g[:,2].max():.2f}; loop end med {np.median(g[:,3]):.2f} max {g[:,3].max():.2f}")
    print(f"   per-CTA clock us (med): setup {np.median(c[:,1]):.2f}, first box {np.median(c[:,2]):.2f}, loop end {np.median(c[:,3]):.2f}")
