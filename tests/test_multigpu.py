"""Multi-process paths: tile-row partition + strip gather and the shared-host-frame delivery protocol
on CPU (gloo, world_size 2); the peer-store composite epilogue between two processes sharing the one
GPU of the test box; and, where the box has two GPUs, the real thing: NCCL, one GPU per rank, the
peer-assembled frame and the shared host frame against a frame rendered by one rank alone."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_frame(h, w):
    y, x = np.mgrid[0:h, 0:w]
    return np.stack([(x * 7 + y * 3) % 256, (x + y * 11) % 256, (x * y) % 256], -1).astype(np.uint8)


def _cpu_worker(rank, world, port, h, w, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from sage3d_official_b200.multigpu import gather_strips, pixel_rows, rows_of
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(_fake_frame(h, w))
    a, b = pixel_rows(rows_of(rank, h, world), h)
    frame = gather_strips(full[a:b].clone(), h, w, rank, world)
    depth = gather_strips(full[a:b, :, 0].float().clone(), h, w, rank, world)
    q.put((rank, bool(torch.equal(frame, full)), bool(torch.equal(depth, full[:, :, 0].float()))))
    dist.destroy_process_group()


@pytest.mark.parametrize("h,w", [(1080, 1920), (100, 64)])
def test_strip_gather_gloo_world2(h, w):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, 2, port, h, w, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]


def test_tile_row_cuts():
    from sage3d_official_b200.multigpu import tile_row_cuts, rows_of, pixel_rows
    assert tile_row_cuts(1080, 8) == [0, 9, 18, 27, 36, 44, 52, 60, 68]
    assert tile_row_cuts(1080, 1) == [0, 68]
    assert tile_row_cuts(480, 4) == [0, 8, 16, 23, 30]
    for world in (1, 2, 3, 8):
        rows = [rows_of(r, 1080, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 68
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    assert pixel_rows((60, 68), 1080) == (960, 1080)


def _gpu_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import sage3d_official_b200 as sage
    from sage3d_official_b200.multigpu import PeerFrames, rows_of
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # both ranks share cuda:0 here
    cloud, _ = sage.synthetic_room(0, 20_000)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 640, 480)
    scene = sage.GaussianScene(cloud)
    ref_rgb, ref_dep = (t.clone() for t in scene.render(cam))
    rgb = [torch.zeros_like(ref_rgb)]
    dep = [torch.zeros_like(ref_dep)]
    peers = PeerFrames(rgb, dep, rank, world)
    scene.render_peers(cam, rows_of(rank, cam.height, world), peers.rgb[0], peers.depth[0])
    peers.close_frame(0)            # flags in peer memory: returns once BOTH ranks' strips are in this buffer
    torch.cuda.synchronize()
    ok = bool(torch.equal(rgb[0], ref_rgb)) and bool(torch.equal(dep[0], ref_dep))
    # a second frame through the same slot: the counter moves on
    scene.render_peers(cam, rows_of(rank, cam.height, world), peers.rgb[0], peers.depth[0])
    peers.close_frame(0)
    torch.cuda.synchronize()
    ok = ok and int(peers._flags[0, :world].min().item()) == 2
    dist.barrier()
    peers.close()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_store_composite_two_processes_one_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_balanced_cuts_equalise_cost():
    from sage3d_official_b200.multigpu import balanced_cuts
    cost = np.array([10.0] * 20 + [100.0] * 30 + [40.0] * 18)
    for world in (1, 2, 4, 8):
        cuts = balanced_cuts(cost, world)
        assert cuts[0] == 0 and cuts[-1] == len(cost) and len(cuts) == world + 1
        assert all(b > a for a, b in zip(cuts[:-1], cuts[1:]))
        shares = [cost[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])]
        assert max(shares) <= 1.35 * cost.sum() / world
    assert balanced_cuts([1.0, 1.0, 1.0], 5)[-1] == 3


@pytest.mark.gpu
def test_row_instances_sum_to_frame_total():
    import sage3d_official_b200 as sage
    cloud, _ = sage.synthetic_room(0, 20_000)
    cam = sage.Camera.look_heading((1.5, 1.0, 1.2), np.radians(30.0), 640, 480)
    scene = sage.GaussianScene(cloud)
    scene.render(cam)
    rows = scene.row_instances(30)
    assert rows.sum() == scene.stats()["n_instances"] and (rows >= 0).all() and rows.max() > 0
    scene.render(cam, rows=(10, 20))
    part = scene.row_instances(10)
    assert part.sum() == scene.stats()["n_instances"]


class _FakeScene:
    """Stands in for GaussianScene in the CPU protocol test: 'renders' its tile rows of a known frame."""

    def __init__(self, h, w):
        self.full = _fake_frame(h, w)

    def render_host_async(self, cam, slot, rows=None, want_depth=True, out=None):
        a, b = rows[0] * 16, min(rows[1] * 16, self.full.shape[0])
        out[0][a:b] = (self.full[a:b].astype(np.int64) + cam) % 256          # cam = frame number here
        out[1][a:b] = self.full[a:b, :, 0].astype(np.float32) + cam

    def host_wait(self, slot):
        pass


def _shared_host_worker(rank, world, port, h, w, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from sage3d_official_b200.multigpu import SharedHostFrames, rows_of
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, rows, d, n = _FakeScene(h, w), rows_of(rank, h, world), 3, 11
    shf = SharedHostFrames(h, w, d, rank, world, timeout_s=60.0, pin=False)
    ok = True
    for i in range(n):                       # d frames in flight, as bench.py drives it
        if i >= d:
            shf.finish(scene, i - d, release=False)
            if rank == 0:                    # delivered: the WHOLE frame i-d is in the shared buffer right now
                r, dd = shf.frame(i - d)
                ok &= bool(np.array_equal(r, (scene.full.astype(np.int64) + (i - d)) % 256))
                ok &= bool(np.array_equal(dd, scene.full[:, :, 0].astype(np.float32) + (i - d)))
            shf.release(i - d)
        shf.submit(scene, i, rows, i)
    for i in range(n - d, n):
        shf.finish(scene, i)
    dist.barrier()
    shf.close()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shared_host_frames_protocol_gloo_world2():
    """Every rank lands its strip in one shared host frame; rank 0 sees a frame only when all strips are in,
    and no rank overwrites a slot before rank 0 has taken the frame it held."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_host_worker, args=(r, 2, port, 100, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _two_gpu_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import sage3d_official_b200 as sage
    from sage3d_official_b200.multigpu import PeerFrames, SharedHostFrames, rows_of
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cloud, layout = sage.synthetic_room(0, 30_000)
    cams = [sage.Camera.look_heading(p, h, 640, 480) for p, h in sage.camera_path(layout, 5, seed=2)]
    scene = sage.GaussianScene(cloud, device=rank)
    dev = torch.device("cuda", rank)
    rgb = [torch.zeros((480, 640, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    dep = [torch.zeros((480, 640), dtype=torch.float32, device=dev) for _ in range(2)]
    peers = PeerFrames(rgb, dep, rank, world)
    rows = rows_of(rank, 480, world)
    ok = True
    for i, cam in enumerate(cams):
        b = i % 2
        scene.render_peers(cam, rows, peers.rgb[b], peers.depth[b])
        peers.close_frame(b)
        torch.cuda.synchronize()
        dist.barrier()
        ref_rgb, ref_dep = scene.render(cam)
        torch.cuda.synchronize()
        ok &= bool(torch.equal(rgb[b], ref_rgb)) and bool(torch.equal(dep[b].view(torch.int32), ref_dep.view(torch.int32)))
        dist.barrier()
    peers.close()
    shf = SharedHostFrames(480, 640, 3, rank, world)
    for i, cam in enumerate(cams):
        shf.submit(scene, cam, rows, i)
        shf.finish(scene, i, release=False)
        if rank == 0:
            r, d_ = shf.frame(i)
            ref_rgb, ref_dep = scene.render_host(cam)
            ok &= bool(np.array_equal(r, ref_rgb)) and bool(np.array_equal(d_.view(np.uint32), ref_dep.view(np.uint32)))
        shf.release(i)
    dist.barrier()
    shf.close()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_peer_frame_and_shared_host_frame_equal_single_gpu_frame():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
