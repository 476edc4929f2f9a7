"""
MultiLoss -- host-side mirror of PyLC's models/modules/loss.py (reference loss.py:23-215).

    L = ce_weight * CE + dice_weight * Dice + focal_weight * Focal        (loss.py:107-112)

The reference evaluates this with ~10 torch kernels forward (2 softmax, 2 one_hot to int64
[B,H,W,C], CE, pow/log/sum ...) plus autograd backward.  Here the forward is ONE streaming pass
over the logits (2C+3 partial sums) and the backward ONE more (closed-form gradient of all three
terms); Dice needs the batch-global sums before any gradient exists, so two passes
over the logits is the minimum.  On one GPU both passes run in a single cooperative launch
(pylc_multiloss_fwd_bwd: grid-wide barrier between them, gradient pass back to front so it starts
in L2) whenever the logits require a gradient.  Under data parallelism (`distributed=True`) the 2C+3
partials are all-reduced between the passes, which makes the loss and gradient those of the single large
batch: inside the same single launch, over peer-addressable memory (pylc_multiloss_fwd_bwd_dp, when the
process group has symmetric memory), else as two launches (pylc_multiloss_reduce / pylc_multiloss_grad)
with an NCCL all-reduce in between.

Interface kept: MultiLoss(loss_weights, schema); .forward(pred, target); .ce_loss / .dice_loss /
.focal_loss callable on their own (models/model.py:360-362); .ce / .dsc / .fl hold the last
component values; .print_settings().
"""
import numpy as np
import torch

from ... import dist as pdist
from ... import ops
from ...config import defaults


class _MultiLossFn(torch.autograd.Function):
    """pred [B,C,H,W] f32, target [B,H,W] i64/u8 -> out[4] = (loss, ce, dice, focal)."""

    @staticmethod
    def forward(ctx, pred, target, class_w, cfg, distributed, grad_mult):
        pred = pred.contiguous()
        target = target.contiguous()
        C = pred.shape[1]
        ctx.fused_grad = None
        ctx.grad_mult = float(grad_mult)
        ctx.class_w, ctx.cfg = class_w, cfg
        if ctx.needs_input_grad[0] and not (distributed and pdist.world_size() > 1):
            # single-GPU training step: forward and backward in ONE cooperative launch; the gradient is
            # stashed for the first backward(), which only rescales it if the upstream gradient is not 1.
            # A second backward over the same graph (retain_graph=True, checkpointing) recomputes the
            # gradient from the saved tensors with pylc_multiloss_grad.
            out, grad, partials = ops.multiloss_fwd_bwd(pred, target, cfg, class_w)
            ctx.fused_grad = grad
            ctx.n_px = target.numel()
            ctx.save_for_backward(pred, target, partials)
            ctx.mark_non_differentiable(target)
            return out
        if ctx.needs_input_grad[0] and distributed and pdist.world_size() > 1:
            ex = pdist.loss_exchange()
            if ex is not None:
                # data-parallel training step, still ONE cooperative launch: the 2C+3 partials are all-reduced inside
                # the kernel over peer-addressable memory (NVLink / NVSwitch), between the two passes
                out, grad, partials = ops.multiloss_fwd_bwd_dp(pred, target, cfg, ex["ptrs_dev"], ex["rank"], ex["world"],
                                                               ex["next_epoch"](), class_w)
                ctx.fused_grad = grad
                ctx.n_px = target.numel() * ex["world"]
                ctx.save_for_backward(pred, target, partials)
                ctx.mark_non_differentiable(target)
                return out
        # int64 targets (the reference dtype) are copied to one byte per pixel by the reduce pass; the
        # gradient pass, and autograd's saved tensors, keep the 8x smaller copy
        t8 = None
        if ctx.needs_input_grad[0] and target.dtype == torch.int64:
            t8 = torch.empty((target.numel(),), dtype=torch.uint8, device=target.device)
        partials = ops.multiloss_reduce(pred, target, cfg, class_w, target_u8_out=t8)
        n_px = target.numel()
        if distributed and pdist.world_size() > 1:
            pdist.all_reduce_(partials)
            n_px *= pdist.world_size()
        out = ops.multiloss_finalize(partials, C, n_px, cfg)
        ctx.save_for_backward(pred, target if t8 is None else t8.view(target.shape), partials)
        ctx.n_px = n_px
        ctx.mark_non_differentiable(target)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        # dL/dz from the kernel, scaled by the upstream gradient of out[0] (the weighted loss).
        # The component outputs out[1:4] are reporting values; gradients through them are dropped.
        g0 = grad_out[0].to(torch.float32).contiguous()
        if ctx.grad_mult != 1.0:
            g0 = g0 * ctx.grad_mult
        if ctx.fused_grad is not None:
            grad, ctx.fused_grad = ctx.fused_grad, None
            ops.scale_unless_one_(grad, g0)
            return grad, None, None, None, None, None
        pred, target, partials = ctx.saved_tensors
        grad = ops.multiloss_grad(pred, target, ctx.cfg, partials, ctx.n_px, ctx.class_w, grad_scale_dev=g0)
        return grad, None, None, None, None, None


class MultiLoss(torch.nn.Module):
    def __init__(self, loss_weights, schema, distributed=False, ddp_average=False):
        """distributed: all-reduce the 2C+3 partial sums between the two passes, so the loss is that of
        the single large batch and each rank's gradient is d(global loss)/d(local logits).
        ddp_average: the parameter gradients will be AVERAGED over ranks afterwards (stock
        DistributedDataParallel); the logit gradient is then multiplied by the world size so that the
        averaged parameter gradient equals the single-large-batch gradient (it matters: Model.train clips
        the gradient norm at 0.5 before the optimiser step, reference model.py:325)."""
        super(MultiLoss, self).__init__()
        self.ddp_average = ddp_average
        # True: every call reads the loss back (one 4-byte D2H) and raises on an out-of-range target like the
        # reference; False: no host sync -- the values are NaN instead (Model.log raises when it reads them)
        self.validate_targets = True
        self.n_classes = schema['n_classes']
        self.codes = schema['class_codes']
        self.categories = schema['class_labels']
        self.weighted = loss_weights['weighted']
        self.dsc_weight = loss_weights['dice']
        self.ce_weight = loss_weights['ce']
        self.fl_weight = loss_weights['focal']
        self.eps = 1e-8
        self.ce = 0.
        self.dsc = 0.
        self.fl = 0.
        self.distributed = distributed
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device(defaults.device)
        w = loss_weights.get('weights')
        if w is not None:
            self.weights = torch.tensor(np.array(w)).float().to(self.device)
        else:
            self.weights = torch.ones(self.n_classes).to(self.device)

    # -- kernel plumbing ---------------------------------------------------------------------
    def _cfg(self, ce, dice, focal):
        return ops.loss_cfg(ce=ce, dice=dice, focal=focal, smooth=defaults.dice_smooth, gamma=defaults.fl_gamma,
                            alpha=defaults.fl_alpha, eps=self.eps)

    def _class_w(self, pred):
        return self.weights.to(pred.device) if self.weighted else None

    def _check(self, pred, target):
        if not torch.is_tensor(pred):
            raise TypeError("Input type is not a torch.Tensor. Got {}".format(type(pred)))
        if not len(pred.shape) >= 2:
            raise ValueError("Invalid input shape, we expect BxCx*. Got: {}".format(pred.shape))
        assert pred.size(0) == target.size(0)
        assert pred.size(2) == target.size(1)
        assert pred.size(3) == target.size(2)
        if not pred.device == target.device:
            raise ValueError("input and target must be in the same device. Got: {} and {}".format(
                pred.device, target.device))
        if defaults.fl_reduction != 'mean':
            raise NotImplementedError("Invalid reduction mode: {}".format(defaults.fl_reduction))

    def _run(self, pred, target, ce, dice, focal):
        self._check(pred, target)
        mult = pdist.world_size() if (self.distributed and self.ddp_average) else 1
        out = _MultiLossFn.apply(pred.float(), target, self._class_w(pred), self._cfg(ce, dice, focal),
                                 self.distributed, mult)
        if self.validate_targets and bool(torch.isnan(out[0])):
            # the kernels poison the batch with NaN when a target lies outside [0, n_classes) -- where the
            # reference's CrossEntropyLoss / one_hot raise (loss.py:66-69,137)
            if int(target.min()) < 0 or int(target.max()) >= self.n_classes:
                raise IndexError("Target {} is out of bounds.".format(
                    int(target.max()) if int(target.max()) >= self.n_classes else int(target.min())))
            raise FloatingPointError("MultiLoss is NaN: non-finite logits")
        return out

    # -- reference API -----------------------------------------------------------------------
    def forward(self, pred, target):
        out = self._run(pred, target, self.ce_weight, self.dsc_weight, self.fl_weight)
        vals = out.detach()
        self.ce, self.dsc, self.fl = vals[1], vals[2], vals[3]
        self.last = vals            # (loss, ce, dice, focal): one D2H copy reads all four
        return out[0]

    def ce_loss(self, pred, target):
        return self._run(pred, target, 1.0, 0.0, 0.0)[0]

    def dice_loss(self, pred, target):
        return self._run(pred, target, 0.0, 1.0, 0.0)[0]

    def focal_loss(self, pred, target):
        return self._run(pred, target, 0.0, 0.0, 1.0)[0]

    def components(self, pred, target):
        """(ce, dice, focal) from ONE pass -- what Model.eval needs (models/model.py:360-362 makes
        three separate calls)."""
        with torch.no_grad():
            out = self._run(pred, target, self.ce_weight, self.dsc_weight, self.fl_weight)
        return out[1], out[2], out[3]

    def print_settings(self):
        hline = '_' * 40
        print('{:30s}{:<10s}'.format('Loss', 'Weight'))
        print(hline)
        print('{:30s}{:<10f}'.format('Cross-entropy', self.ce_weight))
        print('\tCE losses weighted by class.' if self.weighted else '\tCE losses not weighted by class.')
        print('{:30s}{:<10f}'.format('Dice Coefficient', self.dsc_weight))
        print('{:30s}{:<10f}'.format('Focal Loss', self.fl_weight))
        print()
        print('{:8s}{:22s}{:<10s}'.format('Class', 'Label', 'Weight'))
        print(hline)
        for i, w in enumerate(self.weights):
            print('{:8s}{:22s}{:<10f}'.format(self.codes[i], self.categories[i], w))
        print()


def __getattr__(name):
    """`RunningLoss` is importable from here as in the reference (models/modules/loss.py:218-327); it is defined next
    to Model (models/model.py), which imports this module, hence the lazy look-up."""
    if name == "RunningLoss":
        from ..model import RunningLoss
        return RunningLoss
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
