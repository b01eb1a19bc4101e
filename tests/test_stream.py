"""Multi-GPU plumbing on CPU: the path shards by FRAME with one parameter broadcast and no pixel collective
(SURVEY.md 8e).  These tests run the N > 1 host logic with world_size 2 over gloo; the GPU variant of the
same property (frame k gives the same bytes whichever rank handles it) is in test_gpu_parity.py."""
import ctypes
import os
import socket

import numpy as np
import pytest

import transform360_b200 as t360
from transform360_b200.stream import StreamSpec, broadcast_parameters, frames_for_rank

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = spec = None
        if rank == 0:
            ctx = t360.make_context(interpolation_alg=t360.LANCZOS4, fixed_yaw=12.5, num_vertical_segments=15,
                                    num_horizontal_segments=32, output_layout=t360.LAYOUT_EAC_32)
            spec = StreamSpec(7680, 3840, 3840, 2560)
        ctx, spec = broadcast_parameters(ctx, spec, rank, world, device="cpu")
        # every rank plans locally from the broadcast bytes: the host plan must come out identical
        hp = t360.HostPlan(ctx, 640, 320, 240, 160)
        digest = int(np.frombuffer(hp.map.tobytes(), np.uint32).sum(dtype=np.uint64))
        q.put((rank, bytes(ctx), spec.to_ints(), digest, list(frames_for_rank(11, rank, world))))
    finally:
        dist.destroy_process_group()


def test_parameter_broadcast_and_frame_sharding_world2():
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, s0, d0, f0), (r1, c1, s1, d1, f1) = res
    assert c0 == c1 and len(c0) == 112, "ranks disagree on the 112-byte context"
    assert s0 == s1 == [7680, 3840, 3840, 2560, 1, 1, 3]
    assert d0 == d1, "ranks planned different maps from the same parameters"
    assert f0 == [0, 2, 4, 6, 8, 10] and f1 == [1, 3, 5, 7, 9]
    assert sorted(f0 + f1) == list(range(11)), "every frame is handled exactly once"


def test_stream_spec_plane_dims_follow_the_filter():
    s = StreamSpec(1921, 961, 769, 513)  # odd sizes: FF_CEIL_RSHIFT (vf_transform360.c:87-97)
    assert s.plane_dims(0) == (1921, 961, 769, 513, 0)
    assert s.plane_dims(1) == (961, 481, 385, 257, 1)
    assert s.plane_dims(2) == (961, 481, 385, 257, 1)
    assert s.input_pixels_per_frame() == 1921 * 961 + 2 * 961 * 481
    assert StreamSpec.from_ints(s.to_ints()) == s


def test_single_rank_broadcast_is_identity():
    ctx = t360.make_context()
    spec = StreamSpec(64, 32, 24, 16)
    c, s = broadcast_parameters(ctx, spec, 0, 1)
    assert c is ctx and s is spec
    assert ctypes.sizeof(c) == 112
