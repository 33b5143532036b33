"""Discriminative target model (drop-in for ``model/discriminator.py:11-227``).

``Discriminator`` keeps the reference's constructor, attributes (``project``, ``filter``, ``layer``, ``memory``,
``update_optimizer``, ``current_sample``, ``frame_num``) and methods (``init``, ``apply``, ``update``, ``forward``,
``compute_pixel_weights``).  All arithmetic runs in libfrtm_b200: the 1x1 projection is an NHWC implicit-GEMM conv,
the 3x3 filter a correlation kernel, the online learning the closed-form GN/CG of ``model/optimizer.py``.  The
``< 10 px`` gate of ``update`` (``:214``) is evaluated on the device, so tracking never synchronises with the host.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..lib.tensorlist import TensorList
from .memory import Memory
from .optimizer import MinimizationProblem, GaussNewtonCG


def conv(ic, oc, ksize, bias=True, dilation=1, stride=1):
    return nn.Conv2d(ic, oc, ksize, padding=ksize // 2, bias=bias, dilation=dilation, stride=stride)


class DiscriminatorLoss(MinimizationProblem):
    """Describes the weighted least-squares problem (``:11-64``); evaluated in closed form by GaussNewtonCG.

    ``x`` is the frame memory's sample buffer (filter-only problem) or the raw NHWC features of the augmented first
    frame (joint problem, ``x_nhwc``).  ``filter_regs`` / ``precond`` as in the reference."""

    def __init__(self, x, y, filter_regs, precond, sample_weights, net, pixel_weighting, compute_norm=False,
                 memory: Memory = None, x_nhwc=None, stencil=None, uty=None):
        super().__init__()
        self.training_samples = x
        self.training_labels = y
        self.y_size = y.shape[-2:]
        self.filter_regs = TensorList([float(v) for v in filter_regs])
        self.diag_M = TensorList([float(v) for v in precond])
        self.sample_weights = sample_weights
        self.pixel_weighting = pixel_weighting
        self.net = net
        self.memory = memory
        self.x_nhwc, self.stencil, self.uty = x_nhwc, stencil, uty

    def initialize(self):
        pass   # the kernels skip inactive samples (weight == 0) in place; nothing is gathered or copied (:38-43)


class Discriminator(nn.Module):

    def forward(self, x):
        """x (B,C,h,w) NCHW -> scores (B,1,h,w)."""
        cft = self._project_nchw(ops.nchw_to_nhwc(x))
        return ops.corr3x3(cft, self.filter.weight).unsqueeze(1)

    def __init__(self, in_channels=1024, c_channels=96, out_channels=1,
                 init_iters=(5, 10, 10, 10, 10), update_iters=(10,), update_filters=True,
                 filter_reg=(1e-4, 1e-2), precond=(1e-4, 1e-2), precond_lr=0.1, CG_forgetting_rate=75,
                 memory_size=80, train_skipping=8, learning_rate=0.1,
                 pixel_weighting=None, device=None, layer=None):
        super().__init__()
        if out_channels != 1:
            raise ValueError("out_channels must be 1")
        self.project = conv(in_channels, c_channels, 1, bias=False)
        self.filter = conv(c_channels, out_channels, 3, bias=False)
        for p in self.parameters():
            p.requires_grad_(False)
        self.layer = layer
        self.init_iters = init_iters
        self.update_iters = update_iters
        self.filter_reg = filter_reg
        self.precond = precond
        self.direction_forget_factor = (1 - precond_lr) ** CG_forgetting_rate
        self.train_skipping = train_skipping
        self.learning_rate = learning_rate
        self.memory_size = memory_size
        self.pw_params = pixel_weighting
        self.device = device
        self.update_filters = update_filters
        self.to(device)
        self.frame_num = 0
        self.update_optimizer = None
        self.current_sample = None
        self.memory = None
        self.min_px = 10

    # ---------------------------------------------------------------------------------------------------------
    def _project_nchw(self, x_nhwc):
        """(B,h,w,C) NHWC -> projected (B,c,h,w) NCHW (the layout memory samples are stored in)."""
        pc = ops.pack_conv_tc_1x1_device(self.project.weight)
        return ops.conv2d_tc(ops.split_f16(x_nhwc), pc, out_f32=False, nchw=True)["nchw"]

    def compute_pixel_weights(self, y, threshold=False):
        """Hinge weighting (``:107-152``); ``y`` (N,1,H,W)."""
        if self.pw_params is None or self.pw_params["method"] == "none":
            return torch.ones_like(y, dtype=torch.float32)
        assert self.pw_params["method"] == "hinge"
        return ops.pixel_weights(y.float().contiguous(), self.pw_params["tf"], threshold)

    def init(self, x, y, x_nhwc=None):
        """x (K,C,h,w) first-frame augmented features, y (K,1,H,W) masks (``:154-199``)."""
        self.init_joint(x, y, x_nhwc)
        self.init_memory()

    def init_joint(self, x, y, x_nhwc=None):
        """First half of ``init`` (``:154-175``): the joint Gauss-Newton fit of projection and filter on the raw features.
        It touches nothing but this object's own ``project`` / ``filter`` weights, so the tracker runs it while the previous
        sequence's last block is still executing; ``init_memory`` (pooled per-slot buffers) completes the initialisation."""
        if x_nhwc is None:
            x_nhwc = ops.nchw_to_nhwc(x)
        x_nhwc = x_nhwc.contiguous()
        K, h, w, C = x_nhwc.shape
        yf = y.float().contiguous()
        pw = self.compute_pixel_weights(yf)
        stencil, uty = ops.build_stencil(pw, yf, (h, w))
        sw_h = torch.full((K,), 1.0 / K)
        sw_h[0] = 2.0 / K
        sw_h = sw_h / sw_h.sum()
        sw = torch.empty(K, device=x_nhwc.device, dtype=torch.float32)
        ops.fill_small(fdst=sw, fvals=sw_h.tolist())                     # no synchronising H2D copy

        # joint optimisation of projection and filter on the raw features
        problem = DiscriminatorLoss(x=x_nhwc, y=yf, filter_regs=self.filter_reg, precond=self.precond, sample_weights=sw,
                                    net=nn.Sequential(self.project, self.filter), pixel_weighting=pw, x_nhwc=x_nhwc,
                                    stencil=stencil, uty=uty)
        optimizer = GaussNewtonCG(problem, TensorList([self.project.weight, self.filter.weight]), fletcher_reeves=False,
                                  standard_alpha=True, direction_forget_factor=self.direction_forget_factor)
        optimizer.run(self.init_iters)
        self._init_ctx = (x_nhwc, yf, pw, stencil, uty)

    def init_memory(self):
        """Second half of ``init`` (``:177-199``): re-project with the learned matrix, fill the memory, warm up the
        filter-only optimiser.  Uses the pooled per-slot buffers, so it runs once the slot's previous owner is done."""
        x_nhwc, yf, pw, stencil, uty = self._init_ctx
        self._init_ctx = None
        cx = self._project_nchw(x_nhwc)
        # ``buffer_pool`` (set by the tracker): the frame memory and the CG state of this object slot are reused across
        # sequences — same addresses, so a captured track-block graph stays valid — instead of being reallocated
        pool = getattr(self, "buffer_pool", None)
        memory = pool.get("memory") if pool is not None else None
        if memory is not None and memory.matches(self.memory_size, cx.shape[-3:], yf.shape[-3:], x_nhwc.device):
            memory.reset()
        else:
            memory = Memory(self.memory_size, cx.shape[-3:], yf.shape[-3:], x_nhwc.device, self.learning_rate)
            if pool is not None:
                pool["memory"] = memory
        memory.initialize(cx, yf, pw, stencil, uty)
        problem = DiscriminatorLoss(x=memory.samples, y=memory.labels, filter_regs=self.filter_reg[1:],
                                    precond=self.precond[1:], sample_weights=memory.weights, net=self.filter,
                                    pixel_weighting=memory.pixel_weights, memory=memory)
        optimizer = GaussNewtonCG(problem, TensorList([self.filter.weight]), fletcher_reeves=False, standard_alpha=True,
                                  direction_forget_factor=self.direction_forget_factor,
                                  cg_state=pool.get("cg_state") if pool is not None else None)
        if pool is not None:
            pool["cg_state"] = optimizer.cg_state
        optimizer.run(self.update_iters)
        self.memory = memory
        self.update_optimizer = optimizer

    def apply(self, ft):
        """ft (1,C,h,w) -> scores (1,1,h,w); remembers the projected sample for ``update`` (``:201-206``)."""
        self.frame_num += 1
        cft = self._project_nchw(ops.nchw_to_nhwc(ft))
        self.current_sample = cft
        return ops.corr3x3(cft, self.filter.weight).unsqueeze(1)

    def update(self, train_y, gate_count=None, pw=None, stencil=None, uty=None, run_optimizer=True):
        """train_y (1,1,H,W) merged soft mask (``:208-227``).  ``gate_count``: int32 device scalar with the number of
        pixels > 0.5 (computed here when not supplied by the tracker's merge kernel)."""
        if not self.update_filters or self.current_sample is None:
            return
        train_y = train_y.contiguous()
        if pw is None or gate_count is None:
            tf = self.pw_params["tf"] if self.pw_params else 0.0
            pw_h, cnt = ops.pixel_weights(train_y.float(), tf, True, return_count=True)
            if gate_count is None:
                gate_count = cnt.to(torch.int32)          # device scalar; no host sync
            if pw is None:
                pw = pw_h if (self.pw_params and self.pw_params["method"] == "hinge") else torch.ones_like(train_y)
        self.memory.update(self.current_sample, train_y, pw, stencil, uty, gate_count=gate_count, min_px=self.min_px)
        if self.frame_num % self.train_skipping != 0 or not run_optimizer:
            return
        self.update_optimizer.run(self.update_iters, gate_count=gate_count, min_px=self.min_px)
