"""Host-side logic of the N>1 path on CPU: deterministic sequence sharding and the end-of-batch label gather
(gloo, world_size 2)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def test_assignment_is_balanced_and_deterministic():
    from frtm_vos_b200.parallel import assign_sequences
    costs = [33 * 4] * 64
    plan = assign_sequences(costs, 8)
    assert sorted(i for p in plan for i in p) == list(range(64))
    assert all(len(p) == 8 for p in plan)
    ragged = assign_sequences([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    loads = [sum([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1][i] for i in p) for p in ragged]
    assert loads == [10, 10] and ragged == assign_sequences([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert assign_sequences([], 4) == [[], [], [], []]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from frtm_vos_b200.parallel import assign_sequences, gather_label_maps
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = assign_sequences([5, 3, 3, 5], world)
    mine = plan[rank]
    # each "sequence" produces label maps that encode its global index
    local = torch.stack([torch.full((3, 4, 6), i, dtype=torch.uint8) for i in mine])
    allmaps = gather_label_maps(local, world)
    ok = allmaps.shape == (world, len(mine), 3, 4, 6)
    for r in range(world):
        for j, i in enumerate(plan[r]):
            ok = ok and bool((allmaps[r, j] == i).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_label_maps_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_host_wait_policy_rejects_unknown_and_fails_loudly_without_driver():
    """set_host_wait_policy: unknown policy names are a KeyError before anything is touched; without a CUDA driver
    (this container) it raises instead of silently doing nothing."""
    import ctypes.util
    from frtm_vos_b200.parallel import set_host_wait_policy, _WAIT_FLAGS
    assert _WAIT_FLAGS == {"auto": 0, "spin": 1, "yield": 2, "block": 4}      # CU_CTX_SCHED_* values
    import torch
    if torch.cuda.is_available():
        with pytest.raises(KeyError):
            set_host_wait_policy(0, "sleepy")
    else:
        with pytest.raises((OSError, RuntimeError)):
            set_host_wait_policy(0, "yield")
