"""Gauss-Newton + Polak-Ribiere conjugate gradient for the target model, in closed form on the device.

The reference (``model/optimizer.py:18-160``) linearises the residual with autograd and evaluates ``J^T J p`` by a
double backward pass — ~25 ATen launches per CG iteration, each a full pass over the (M,1,H,W) label-size tensors.
For FRTM's two problems the operator has an exact closed form, so here ``run`` enqueues the
library's streaming kernels instead (``frtm_gn_update`` for the filter-only problem, ``frtm_gn_init`` for the joint
project+filter problem).  The CG recurrences are the reference's: ``z = r/diag_M``, Polak-Ribiere ``beta`` clamped at
0, ``standard_alpha``, ``r`` not advanced on the last iteration, direction/``rho``/``r_prev`` persisting across
``run`` calls with ``rho /= direction_forget_factor`` on entry (``:98-153``).

Only ``DiscriminatorLoss`` problems are supported: a generic autograd ``MinimizationProblem`` would need a CPU /
autograd fallback, which this package does not ship by design.
"""
from __future__ import annotations

import collections
import ctypes

import torch

from ..lib.tensorlist import TensorList
from .._lib import lib, ptr, stream


# Staging buffers of the joint (project, filter) optimisation, one set per (shape, device, stream): the library replays the
# whole schedule as a CUDA graph when it sees the same buffers again.  A small LRU: datasets with mixed resolutions
# (YouTubeVOS) would otherwise pin one set (tens of MB) per resolution for the life of the process; an evicted set's graphs
# are released in the library before its memory is.  ``release_init_stages()`` (``Tracker.clear``) drops everything.
_INIT_STAGE = collections.OrderedDict()
_INIT_STAGE_CAP = 16          # = the library's graph cache; one set per object stream (10 objects in BASELINE config 5)


def _drop_stage(st):
    lib().gn_init_release(ptr(st["ws"]))


def release_init_stages():
    while _INIT_STAGE:
        _, st = _INIT_STAGE.popitem(last=False)
        _drop_stage(st)


def _init_stage(K, C, c, h, w, device):
    key = (K, C, c, h, w, str(device), torch.cuda.current_stream().cuda_stream)
    st = _INIT_STAGE.get(key)
    if st is None:
        while len(_INIT_STAGE) >= _INIT_STAGE_CAP:
            _, old = _INIT_STAGE.popitem(last=False)
            _drop_stage(old)
        nbytes = lib().gn_init_workspace(K, C, c, h, w)
        f = dict(device=device, dtype=torch.float32)
        st = dict(x=torch.empty((K, h, w, C), **f), stencil=torch.empty((K, 9, h, w), **f), uty=torch.empty((K, h, w), **f),
                  sw=torch.empty(K, **f), P=torch.empty((c, C), **f), F=torch.empty(c * 9, **f),
                  ws=torch.empty(nbytes // 4, **f), nbytes=nbytes)
        _INIT_STAGE[key] = st
    else:
        _INIT_STAGE.move_to_end(key)
    return st


class MinimizationProblem:
    """Interface kept for API compatibility (``model/optimizer.py:5-15``)."""

    def __call__(self, x: TensorList) -> TensorList:
        raise NotImplementedError

    def ip_input(self, a, b):
        return sum(a.view(-1) @ b.view(-1))

    def M1(self, x):
        return x

    def initialize(self):
        pass


class GaussNewtonCG:

    def __init__(self, problem, variable: TensorList, cg_eps=0.0, fletcher_reeves=True, standard_alpha=True,
                 direction_forget_factor=0, step_alpha=1.0, cg_state=None):
        from .discriminator import DiscriminatorLoss
        if not isinstance(problem, DiscriminatorLoss):
            raise NotImplementedError("GaussNewtonCG runs DiscriminatorLoss problems in closed form on the GPU; generic "
                                      "autograd problems are not supported (no CPU/autograd fallback)")
        if fletcher_reeves or not standard_alpha or step_alpha != 1.0 or cg_eps != 0.0:
            raise NotImplementedError("only the configuration FRTM uses is implemented: Polak-Ribiere, standard_alpha, "
                                      "step_alpha=1, cg_eps=0 (model/discriminator.py:172,192)")
        if direction_forget_factor < 0:
            raise ValueError("direction_forget_factor must be >= 0 (0 = reset the CG state at every run)")
        self.problem = problem
        self.x = variable
        self.direction_forget_factor = float(direction_forget_factor)
        self.joint = len(variable) == 2
        n = variable[-1].numel()
        dev = variable[-1].device
        # p | r_prev | rho | has_p | pad | pad   (filter-only problem; persists across run() calls)
        if cg_state is not None and cg_state.numel() == 2 * n + 4 and cg_state.device == dev:
            self.cg_state = cg_state.zero_()             # pooled buffer (stable address), fresh state
        else:
            self.cg_state = torch.zeros(2 * n + 4, device=dev, dtype=torch.float32)
        self._n = n
        self._ws = None
        self.operator_select = 0    # 0 = the library picks the operator kernel by shape (include/frtm_b200.h: frtm_gn_update)

    # -- persistent CG state, exposed like the reference's attributes ---------------------------------------------
    @property
    def p(self):
        return None if float(self.cg_state[2 * self._n + 1]) == 0.0 else TensorList([self.cg_state[:self._n].view_as(self.x[-1])])

    @property
    def r_prev(self):
        return None if float(self.cg_state[2 * self._n + 1]) == 0.0 else TensorList([self.cg_state[self._n:2 * self._n].view_as(self.x[-1])])

    @property
    def rho(self):
        return self.cg_state[2 * self._n]

    def set_state(self, p, r_prev, rho):
        """Inject (p, r_prev, rho) — used by the oracle-replay parity mode."""
        n = self._n
        self.cg_state[:n] = p.reshape(-1)
        self.cg_state[n:2 * n] = r_prev.reshape(-1)
        self.cg_state[2 * n] = float(rho)
        self.cg_state[2 * n + 1] = 1.0

    # -- run ----------------------------------------------------------------------------------------------------------
    def run(self, num_cg_iter, num_gn_iter=None, gate_count=None, min_px=10):
        if isinstance(num_cg_iter, int):
            if num_gn_iter is None:
                raise ValueError("Must specify number of GN iter if CG iter is constant")
            num_cg_iter = [num_cg_iter] * num_gn_iter
        num_cg_iter = [int(v) for v in num_cg_iter]
        if len(num_cg_iter) == 0:
            return
        iters = (ctypes.c_int * len(num_cg_iter))(*num_cg_iter)
        pr = self.problem
        L = lib()
        if self.joint:
            K, h, w, C = pr.x_nhwc.shape
            c = self.x[1].shape[1]
            # Stage through persistent buffers: the library recognises the repeated signature and replays the whole
            # ~650-launch optimisation as one CUDA graph from the second object on.
            st = _init_stage(K, C, c, h, w, pr.x_nhwc.device)
            st["x"].copy_(pr.x_nhwc); st["stencil"].copy_(pr.stencil); st["uty"].copy_(pr.uty)
            st["sw"].copy_(pr.sample_weights); st["P"].copy_(self.x[0].detach().reshape(c, C))
            st["F"].copy_(self.x[1].detach().reshape(-1))
            L.gn_init(ptr(st["x"]), ptr(st["stencil"]), ptr(st["uty"]), ptr(st["sw"]), K, C, c, h, w, ptr(st["P"]),
                      ptr(st["F"]), iters, len(num_cg_iter), pr.filter_regs[0], pr.filter_regs[1], pr.diag_M[0],
                      pr.diag_M[1], self.direction_forget_factor, ptr(st["ws"]), st["nbytes"], stream())
            self.x[0].data.copy_(st["P"].view_as(self.x[0]))
            self.x[1].data.copy_(st["F"].view_as(self.x[1]))
        else:
            mem = pr.memory
            cap, c, h, w = mem.samples.shape
            nbytes = L.gn_update_workspace(cap, c, h, w)
            if self._ws is None or self._ws.numel() * 4 < nbytes:
                self._ws = torch.empty(nbytes // 4, device=mem.samples.device, dtype=torch.float32)
            # The single-object entry point is the API-fidelity path (tests, smoke): callers may have written
            # ``memory.samples`` directly, so the split tile image is rebuilt here; the tracker's batched path relies on
            # the image maintained by ``Memory.update`` / ``Memory.initialize``.
            mem.refresh_split()
            L.gn_update(ptr(mem.samples), ptr(mem.split), ptr(mem.stencil), ptr(mem.uty), ptr(mem.weights), cap, c, h, w, ptr(self.x[0]),
                        ptr(self.cg_state), iters, len(num_cg_iter), pr.filter_regs[-1], pr.diag_M[-1],
                        self.direction_forget_factor, ptr(gate_count), int(min_px), int(self.operator_select), ptr(self._ws), nbytes,
                        stream())
        return [], [], None
