"""Test doubles that put the CPU oracle behind the façade's engine hooks
(``SimpleVLNEnv._scene_factory`` / ``_detector_factory``), so the host logic of the drop-ins — and the
reference's unmodified ``run_benchmark.py`` on top of them — can run where there is no GPU.
TEST INFRASTRUCTURE ONLY: the product never imports this module (or ``oracle``)."""
import numpy as np

import oracle


class _HostTensor:
    """Just enough of a pinned torch tensor for the façade's frame pool (``.numpy()``)."""

    def __init__(self, arr):
        self._a = arr

    def numpy(self):
        return self._a


class OracleScene:
    """GaussianScene stand-in: frames by ``oracle.render`` (the frozen spec)."""

    pipeline_depth = 3

    def __init__(self, cloud, device=0):
        self.cloud = cloud
        self.n = cloud.n
        self._clip = None
        self._stats = {"overflow": 0, "n_instances_wanted": 0}
        self.frames = 0

    def set_depth_clip(self, lo=1.0, hi=0.0):
        self._clip = (lo, hi) if lo <= hi else None

    def _render(self, cam):
        oc = oracle.make_camera(cam.R, cam.t, cam.campos, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.near, cam.far)
        rgb, dep, st = oracle.render(oc, *self.cloud.as_tuple())
        if self._clip is not None:
            dep = np.clip(dep, *self._clip)
        self.frames += 1
        return rgb, dep

    def alloc_host_frame(self, H, W):
        return (_HostTensor(np.zeros((H, W, 3), np.uint8)), _HostTensor(np.zeros((H, W), np.float32)))

    def render_host(self, cam, rows=None, want_depth=True, stream=None):
        return self._render(cam)

    def render_host_async(self, cam, slot, rows=None, want_depth=True, out=None):
        rgb, dep = self._render(cam)
        if out is not None:
            out[0].numpy()[...] = rgb
            out[1].numpy()[...] = dep
            return out[0].numpy(), out[1].numpy()
        return rgb, dep

    def host_wait(self, slot):
        pass

    def stats(self):
        return dict(self._stats)

    def reserve(self, n):
        pass

    def close(self):
        pass


def make_oracle_detector(np2=None):
    """Factory with the product detector's constructor signature; lookups by ``oracle.collide_grid``."""
    from sage3d_official_b200.collision_detector import SemanticMap2DCollisionDetector

    class OracleGrid:
        def __init__(self, det):
            self.d = det

        def check(self, pos, np2_promote=None, segments=None):
            d = self.d
            hit = oracle.collide_grid(d.obstacle_map, d.min_x, d.max_x, d.min_y, d.max_y, d.scale, pos, np2_promote=np2_promote)
            if segments is None:
                return hit
            firsts, s = [], 0
            for n in segments:
                idx = np.nonzero(hit[s:s + n])[0]
                firsts.append(int(idx[0]) if idx.size else -1)
                s += n
            return hit, np.asarray(firsts, np.int32)

    class OracleBackedDetector(SemanticMap2DCollisionDetector):
        def _upload(self):
            self._grid = OracleGrid(self)

    return OracleBackedDetector
