"""The multi-GPU partition (independent frames / row bands, no collective on the data path) exercised with
two CPU processes over gloo: each rank derives its share from the C ABI's host-only plan functions, the shares are
gathered and must tile the job exactly; the timing reduction bench.py uses (MAX over ranks) is checked too."""
import os
import socket
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_frames: int, target_h: int):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "avisynth-jincresize_b200"))
    import torch
    import torch.distributed as dist

    from jinc_b200 import capi

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # frame-parallel: every frame has exactly one owner, shares differ by at most one frame
        mine = capi.frames_of(rank, world, n_frames)
        owned = torch.zeros(n_frames, dtype=torch.int32)
        owned[mine] = 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all()), owned
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
        assert max(int(c) for c in counts) - min(int(c) for c in counts) <= 1

        # row bands: the ranks' bands tile [0, target_h) exactly, in order, on 16-row boundaries
        bands = capi.row_bands(target_h, world)
        y0, y1 = bands[rank]
        covered = torch.zeros(target_h, dtype=torch.int32)
        covered[y0:y1] = 1
        dist.all_reduce(covered)
        assert bool((covered == 1).all())
        assert y0 % 16 == 0 and (y1 % 16 == 0 or y1 == target_h)
        edges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(edges, torch.tensor([y0, y1], dtype=torch.int64))
        for a, b in zip(edges[:-1], edges[1:]):
            assert int(a[1]) == int(b[0])

        # bench.py's reduction: the step time of the job is the slowest rank's
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,target_h", [(240, 2160), (7, 4320), (1, 90)])
def test_two_ranks_partition_frames_and_bands(native_built, n_frames, target_h):
    import torch.multiprocessing as mp

    lib = os.path.join(REPO, "avisynth-jincresize_b200", "libjinc_b200.so")
    if not os.path.exists(lib):
        pytest.skip("libjinc_b200.so not built (run __graft_entry__.build())")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, target_h), nprocs=2, join=True)


def test_plan_functions_single_process(native_built):
    sys.path.insert(0, os.path.join(REPO, "avisynth-jincresize_b200"))
    from jinc_b200 import capi

    if not os.path.exists(os.path.join(REPO, "avisynth-jincresize_b200", "libjinc_b200.so")):
        pytest.skip("libjinc_b200.so not built")
    assert capi.row_bands(2160, 8) == [(i * 272, min(2160, (i + 1) * 272)) for i in range(8)]
    assert capi.row_bands(90, 4) == [(0, 32), (32, 64), (64, 90), (90, 90)]
    assert [capi.frame_owner(n, 4) for n in range(6)] == [0, 1, 2, 3, 0, 1]
    with pytest.raises(capi.JincError):
        capi.row_bands(0, 2)
