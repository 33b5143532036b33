"""Fixed-capacity frame memory of one target model (drop-in for ``model/memory.py:4-92``).

Same buffers and attributes as the reference (``samples (cap,c,h,w)``, ``labels``/``pixel_weights (cap,1,H,W)``,
``weights (cap)``), plus what the closed-form optimiser needs per sample: the 9-tap stencil of ``U^T pw^2 U`` and the
vector ``U^T pw^2 y`` on the feature grid, built once at insert time.  The replace-minimum-weight policy with its
sample-weight update (``:65-92``) runs in a one-thread kernel on the device, so an insert never synchronises with
the host (the reference pays a ``.item()`` per insert, ``:81``); ``current_size`` / ``previous_replace_ind`` are read
back lazily only when somebody asks for them.
"""
from __future__ import annotations

import torch

from .. import ops
from .._lib import lib, ptr, stream


class Memory:

    def __init__(self, capacity, feature_size, labels_size, device, learning_rates):
        self.samples = torch.zeros(capacity, *feature_size, device=device)
        self.weights = torch.zeros(capacity, device=device)
        self.labels = torch.zeros(capacity, *labels_size, device=device)
        self.pixel_weights = torch.zeros(capacity, *labels_size, device=device)
        h, w = feature_size[-2:]
        self.stencil = torch.zeros(capacity, 9, h, w, device=device)
        self.uty = torch.zeros(capacity, h, w, device=device)
        # operator images (split tile image of the sample + its stencil rows) for the tensor-core operator kernel (include/frtm_b200.h: frtm_split_samples)
        c = feature_size[0]
        self._split_ok = c % 8 == 0
        nb = lib().split_sample_bytes(c, h * w) if self._split_ok else 16
        self.split = torch.zeros(capacity, nb, dtype=torch.uint8, device=device)
        # {current_size, previous_replace_ind (-1 = None), slot of the last insert (-1 = skipped), inserts}
        self.state = torch.empty(4, dtype=torch.int32, device=device)
        ops.fill_small(idst=self.state, ivals=(0, -1, -1, 0))          # asynchronous: no pageable H2D copy
        self._capacity = capacity
        self.device = device
        self.learning_rates = learning_rates

    def reset(self):
        """Empty the memory in place (buffers and their addresses are kept): all sample weights zero, fresh policy state.
        Slots with weight 0 are skipped by every kernel, so their stale contents need not be cleared."""
        self.weights.zero_()
        ops.fill_small(idst=self.state, ivals=(0, -1, -1, 0))

    def matches(self, capacity, feature_size, labels_size, device):
        return (self._capacity == capacity and tuple(self.samples.shape[1:]) == tuple(feature_size)
                and tuple(self.labels.shape[1:]) == tuple(labels_size) and self.samples.device == torch.device(device))

    @property
    def capacity(self):
        return self._capacity

    @property
    def current_size(self):
        return int(self.state[0].item())

    @property
    def previous_replace_ind(self):
        v = int(self.state[1].item())
        return None if v < 0 else v

    def refresh_split(self, first=0, count=None):
        """Rebuild the operator images of slots [first, first+count) from ``samples`` / ``stencil`` / ``uty`` (after a
        direct write to any of them)."""
        if not self._split_ok:
            return
        count = self._capacity - first if count is None else count
        c, h, w = self.samples.shape[1:]
        lib().split_samples(ptr(self.samples[first:]), ptr(self.stencil[first:]), ptr(self.uty[first:]), int(count), c, h * w,
                            ptr(self.split[first:]), stream())

    def initialize(self, init_features, init_labels, pixel_weights, stencil=None, uty=None):
        """First K slots <- the augmented first-frame samples; weights 2/K, 1/K, ... normalised (``:33-46``)."""
        K = init_features.shape[0]
        assert init_labels.shape[0] == K
        labels = init_labels.float()
        if stencil is None:
            stencil, uty = ops.build_stencil(pixel_weights, labels, self.samples.shape[-2:])
        self.samples[:K] = init_features.detach()
        w = torch.full((K,), 1.0 / K)            # host arithmetic identical to the reference's, uploaded without a sync
        w[0] = 2.0 / K
        w = w / w.sum()
        assert K <= 16
        ops.fill_small(fdst=self.weights, fvals=w.tolist(), idst=self.state, ivals=(K, -1, -1, 0))
        self.labels[:K] = labels
        self.pixel_weights[:K] = pixel_weights
        self.stencil[:K] = stencil
        self.uty[:K] = uty
        self.refresh_split(0, K)

    def update(self, features, labels, pixel_weights, stencil=None, uty=None, gate_count=None, min_px=10):
        """Insert one sample (``:59-92``).  With ``gate_count`` (int32 device scalar) the insert — including the
        sample-weight update — is skipped on the device when ``gate_count < min_px`` (discriminator.py:214)."""
        if stencil is None:
            stencil, uty = ops.build_stencil(pixel_weights.reshape(1, 1, *pixel_weights.shape[-2:]),
                                             labels.reshape(1, 1, *labels.shape[-2:]), self.samples.shape[-2:])
        L = lib()
        L.memory_next_slot(ptr(self.weights), self._capacity, float(self.learning_rates), ptr(self.state), ptr(gate_count),
                           int(min_px), stream())
        hw = self.uty.shape[-1] * self.uty.shape[-2]
        HW = self.labels.shape[-1] * self.labels.shape[-2]
        L.memory_insert(ptr(features), features.numel(), ptr(labels), ptr(pixel_weights), HW, ptr(stencil), ptr(uty), hw,
                        ptr(self.samples), ptr(self.labels), ptr(self.pixel_weights), ptr(self.stencil), ptr(self.uty),
                        ptr(self.split) if self._split_ok else None, ptr(self.state), stream())
