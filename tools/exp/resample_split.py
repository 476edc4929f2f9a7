"""Experiment: cost split of the resample+encode+confusion kernel (not part of the product)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pylc_b200 import ops
from pylc_b200.config import Parameters
import pylc_oracle as orc
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kbench import Timer
tm = Timer(20)
meta = Parameters(); pal, C = meta.palette_rgb, meta.n_classes
W, H, Wf, Hf = 6000, 4000, 5632, 3584
mask = orc.synth_mask(0, W, H, pal, skew=True)
d_mask, mp = ops.upload_image(mask)
maps = (torch.from_numpy(ops.nn_index_map(Wf, W)).cuda(), torch.from_numpy(ops.nn_index_map(Hf, H)).cuda())
conf = torch.zeros((C, C), dtype=torch.int64, device="cuda")
coherent = torch.from_numpy(orc.synth_labels(3, Wf, Hf, C, skew=True, block=37)).cuda()
palc, _ = ops._lib.palette_array(pal)
lib = ops._lib.load()
pred_full = torch.empty((H, W), dtype=torch.uint8, device="cuda")
def call(gt, cf, pf):
    ops._lib.check(lib.pylc_resample_encode_confusion(ops._p(coherent), Hf, Wf, ops._p(maps[0]), ops._p(maps[1]), H, W,
        ops._p(gt), mp if gt is not None else 0, palc if gt is not None else None, None, C, C if cf is not None else 0,
        ops._p(cf), ops._p(pf), None, None, ops._stream()), "resample")
for name, fn in [("full (encode+gather+count)", lambda: call(d_mask, conf, None)),
                 ("encode+gather, no count", lambda: call(d_mask, None, None)),
                 ("gather only + pred_full out", lambda: call(None, None, pred_full))]:
    med, best = tm.run(fn)
    print(name, "median %.1f us  min %.1f us" % (med * 1e3, best * 1e3), flush=True)
