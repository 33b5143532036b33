"""Multi-GPU plumbing: one process per GPU, sequences sharded across ranks, one gather of label maps at the end.

Objects of one sequence are coupled every frame by the mask merge (``model/tracker.py:214-221``), sequences are fully
independent, so the unit of sharding is the sequence.  There is no collective inside the frame loop;
the only communication is the end-of-batch ``all_gather`` of uint8 label maps (NCCL over NVLink on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def assign_sequences(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of sequence indices to ranks (cost ~ frames x objects).
    Deterministic: ties are broken by index, so every rank computes the same plan without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world_size
    plan: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += float(costs[i])
    for p in plan:
        p.sort()
    return plan


def gather_label_maps(local: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """(S,T,H,W) uint8 label maps of this rank's sequences -> (world,S,T,H,W) on every rank (equal S per rank)."""
    import torch.distributed as dist
    if world_size == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    out = torch.empty((world_size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)          # concatenated along dim 0 (accepted by gloo and NCCL)
    return out.view((world_size,) + tuple(local.shape))


_WAIT_FLAGS = {"auto": 0, "spin": 1, "yield": 2, "block": 4}      # CU_CTX_SCHED_*


def set_host_wait_policy(device_index: int, policy: str = "yield") -> int:
    """How this process's host threads wait for GPU ``device_index`` (stream / event synchronisation).

    The CUDA default spins.  With one process per GPU plus the augmentation workers, a host with fewer cores than
    (ranks x threads) then spends its cores on waiting threads instead of on the ranks that still have kernels to
    launch; ``"yield"`` gives the core away while waiting and costs nothing when cores are free.  Works before or after
    the device's primary context exists (driver API, no torch involvement).  Returns the flags now in force."""
    import ctypes
    cu = ctypes.CDLL("libcuda.so.1")
    dev, flags, active = ctypes.c_int(), ctypes.c_uint(), ctypes.c_int()

    def chk(rc, what):
        if rc != 0:
            raise RuntimeError("frtm_vos_b200.parallel.set_host_wait_policy: %s failed (CUresult %d)" % (what, rc))
    chk(cu.cuInit(0), "cuInit")
    chk(cu.cuDeviceGet(ctypes.byref(dev), int(device_index)), "cuDeviceGet")
    chk(cu.cuDevicePrimaryCtxGetState(dev, ctypes.byref(flags), ctypes.byref(active)), "cuDevicePrimaryCtxGetState")
    want = (flags.value & ~0x7) | _WAIT_FLAGS[policy]
    chk(cu.cuDevicePrimaryCtxSetFlags_v2(dev, ctypes.c_uint(want)), "cuDevicePrimaryCtxSetFlags")
    chk(cu.cuDevicePrimaryCtxGetState(dev, ctypes.byref(flags), ctypes.byref(active)), "cuDevicePrimaryCtxGetState")
    return int(flags.value)
