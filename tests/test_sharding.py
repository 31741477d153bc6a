"""CPU tier: sample sharding across ranks, including a world_size-2 gloo run in which
each rank produces its share of a pass with the CPU oracle and the buffers are all-reduced."""
import ctypes as C
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_frames_for_rank_partitions_the_pass():
    from tryraytrace_b200.sharding import frames_for_rank
    for world in (1, 2, 3, 4, 8):
        for total in (0, 1, 5, 8, 64, 1023):
            seen = []
            for r in range(world):
                first, count, stride = frames_for_rank(7, total, r, world)
                seen += [first + k * stride for k in range(count)]
            assert sorted(seen) == list(range(7, 7 + total))
    with pytest.raises(ValueError):
        frames_for_rank(1, 4, 2, 2)


def _oracle_pass(first, count, stride, w=48, h=36):
    sys.path.insert(0, str(ROOT / "tests"))
    import reflib
    import tryraytrace_b200 as trt
    sc = trt.HostScene.from_config(1, ROOT / "assets")
    cam, _, _ = trt.config_camera(1, w, h)
    li = np.ascontiguousarray(sc.lights, dtype=np.int32)
    acc = np.zeros(w * h * 4, dtype=np.float32)
    for k in range(count):
        reflib.cpu().oracle_render(sc.objects.ctypes.data_as(C.c_void_p), sc.nodes.ctypes.data_as(C.c_void_p),
                                   li.ctypes.data_as(C.c_void_p), len(li), cam.ctypes.data_as(C.c_void_p), w, h, 1984,
                                   first + k * stride, 1, 30, 3, None, None, None, 0, 0, h,
                                   acc.ctypes.data_as(C.c_void_p), None, 2)
    return acc


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tryraytrace_b200.sharding import frames_for_rank
    first, count, stride = frames_for_rank(1, 4, rank, world)
    acc = torch.from_numpy(_oracle_pass(first, count, stride))
    dist.all_reduce(acc)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_pass_equals_single_rank(tmp_path):
    sys.path.insert(0, str(ROOT / "tests"))
    import reflib
    if not reflib.cpu_available() or not (ROOT / "assets" / "cube.obj").exists():
        pytest.skip("CPU oracle or assets not built")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    single = _oracle_pass(1, 4, 1)
    assert reduced.reshape(-1, 4)[:, :3].sum() > 0
    assert np.allclose(reduced, single, rtol=1e-6, atol=1e-6)
