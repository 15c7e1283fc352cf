"""Map building and path planning on the GPU behind the reference's own names.

The reference's trajectory generator (Code/data_pipeline/trajectory_generation/
vln_trajectory_generator.py) builds the same inflated occupancy grid the collision detector uses
and then runs thousands of pure-Python A* searches per scene.  This module mirrors the functions
on that path — same names, arguments, return values and error behaviour —

    build_2d_map(sem_data, scale, robot_radius_m)      :487-559   (inflation on the GPU, K8)
    build_connectivity_map(grid_map, itemid2inst)      :639-689   (components on the GPU, K9)
    astar_pixel(grid, start, goal)                     :253-286   (K10)
    generate_trajectory_points(path, scale, ...)       :703-741   (host arithmetic)

and adds the batched form the GPU wants: :class:`GridPlanner` / :func:`astar_pixel_batch`.
Paths are the reference's paths point for point (tests/test_plan_gpu.py, goldens recorded from the
unmodified reference).  No CPU fallback: without the CUDA library these functions raise.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict, deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import SageError, check

SCALE_M_PER_PX = 0.05      # vln_trajectory_generator.py:65
ROBOT_RADIUS_M = 0.2       # vln_trajectory_generator.py:64
FIXED_Z = 0.5              # vln_trajectory_generator.py:66
SAMPLE_STEP = 1            # vln_trajectory_generator.py:67
OBSTACLE_LABELS = ("unable area", "wall")


def _u8(grid) -> np.ndarray:
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    if g.ndim != 2:
        raise ValueError("grid must be 2-D")
    return g


def inflate_grid(raw, scale: float, robot_radius_m: float, device: int = 0) -> np.ndarray:
    """``(distance_transform_edt(raw == 0, sampling=scale) <= robot_radius_m).astype(uint8)`` on the GPU."""
    _lib.require_gpu()
    raw = _u8(raw)
    out = np.empty_like(raw)
    check(_lib.lib().sage_grid_inflate(int(device), raw.ctypes.data, raw.shape[0], raw.shape[1], float(scale),
                                       float(robot_radius_m), out.ctypes.data, None))
    return out


def connectivity_labels(grid, device: int = 0) -> np.ndarray:
    """int32 [H,W]: smallest linear index of the cell's 4-connected free component, -1 for obstacles."""
    _lib.require_gpu()
    grid = _u8(grid)
    lab = np.empty(grid.shape, np.int32)
    check(_lib.lib().sage_grid_components(int(device), grid.ctypes.data, grid.shape[0], grid.shape[1],
                                          lab.ctypes.data, None))
    return lab


class GridPlanner:
    """Batched A* on one grid (handle of ``sage_planner``): ``slots`` searches run side by side."""

    #: device memory the search slots may take (each slot: g f64 + came_from i32 per cell, 16 B per heap entry)
    MEMORY_BUDGET = 16 << 30

    def __init__(self, grid, slots: int = 16384, max_path: int = 0, heap_capacity: int = 0, device: int = 0,
                 memory_budget: Optional[int] = None, expect_queries: int = 256):
        """``slots``: most searches run side by side (one per thread: throughput follows the searches in flight, run Z
        in profiles/r02_variants.md).  The device arrays are sized for ``expect_queries`` at first and grown when a
        larger batch arrives, so a single ``astar_pixel`` call does not allocate gigabytes."""
        _lib.require_gpu()
        if memory_budget is not None:           # bytes
            self.MEMORY_BUDGET = int(memory_budget)
        self.grid = _u8(grid)
        self.H, self.W = self.grid.shape
        self.max_slots, self.device = int(slots), int(device)
        self.slots = 0
        self._h = C.c_void_p()
        self._make(int(max_path), int(heap_capacity), int(expect_queries))

    def _fit_slots(self, heap_capacity: int, want: int) -> int:
        """As many side-by-side searches as asked for, as the memory budget holds (a 600 x 600 map at 4096 slots would
        be ~30 GB; after a capacity retry the heaps are 4x larger and the slot count shrinks accordingly)."""
        cells = self.H * self.W
        heap = heap_capacity if heap_capacity > 0 else cells // 2 + 1024          # sage_planner_create's default
        per_slot = 12 * cells + 16 * heap
        want = 1 << max(8, int(max(want, 1) - 1).bit_length())                     # >= 256, power of two
        return int(max(1, min(self.max_slots, want, self.MEMORY_BUDGET // max(per_slot, 1))))

    def _make(self, max_path: int, heap_capacity: int, want: int):
        self.close()
        self._asked = (max_path, heap_capacity)
        self.slots = self._fit_slots(heap_capacity, want)
        h = C.c_void_p()
        check(_lib.lib().sage_planner_create(C.byref(h), self.device, self.grid.ctypes.data, self.H, self.W,
                                             self.slots, max_path, heap_capacity))
        self._h = h
        mp, hc = C.c_int(0), C.c_int64(0)
        check(_lib.lib().sage_planner_limits(self._h, C.byref(mp), C.byref(hc)))
        self.max_path, self.heap_capacity = int(mp.value), int(hc.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().sage_planner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, starts, goals):
        Q = starts.shape[0]
        plen = np.zeros(Q, np.int32)
        pxy = np.zeros((Q, self.max_path, 2), np.int32)
        nexp = np.zeros(Q, np.int64)
        check(_lib.lib().sage_plan_astar(self._h, starts.ctypes.data, goals.ctypes.data, Q, plen.ctypes.data,
                                         pxy.ctypes.data, nexp.ctypes.data, None))
        return plen, pxy, nexp

    def plan(self, starts, goals, want_expansions: bool = False):
        """starts/goals: [Q,2] integer (x, y).  Returns a list of int32 arrays [n,2] (start -> goal)
        or None where the reference returns None.  Searches that outgrow the heap or path capacity
        are re-run with more room."""
        starts = np.ascontiguousarray(np.asarray(starts).reshape(-1, 2), dtype=np.int32)
        goals = np.ascontiguousarray(np.asarray(goals).reshape(-1, 2), dtype=np.int32)
        if starts.shape != goals.shape:
            raise ValueError("starts and goals must have the same length")
        Q = starts.shape[0]
        out: List[Optional[np.ndarray]] = [None] * Q
        nexp_all = np.zeros(Q, np.int64)
        todo = np.arange(Q)
        if Q > self.slots and self._fit_slots(self._asked[1], Q) > self.slots:     # a larger batch than any before: more slots
            self._make(self._asked[0], self._asked[1], Q)
        for _ in range(6):
            if todo.size == 0:
                break
            plen, pxy, nexp = self._run(starts[todo], goals[todo])
            redo = []
            for k, q in enumerate(todo):
                nexp_all[q] = nexp[k]
                if plen[k] > 0:
                    out[q] = pxy[k, :plen[k]].copy()
                elif plen[k] < 0:
                    redo.append(q)
            todo = np.asarray(redo, dtype=np.int64)
            if todo.size:
                self._make(min(self.max_path * 4, self.H * self.W), min(self.heap_capacity * 4, 8 * self.H * self.W + 16), todo.size)
        if todo.size:
            raise SageError("A* search exceeded the planner's capacity even at the theoretical maximum")
        return (out, nexp_all) if want_expansions else out


_planner_cache: dict = {}


def _planner_for(grid) -> GridPlanner:
    g = _u8(grid)
    key = (g.shape, hash(g.tobytes()))
    p = _planner_cache.get(key)
    if p is None:
        _planner_cache.clear()                  # one grid at a time, like the reference's per-scene loop
        p = _planner_cache[key] = GridPlanner(g)
    return p


def astar_pixel_batch(grid, starts, goals) -> List[Optional[List[Tuple[int, int]]]]:
    """Many ``astar_pixel`` queries on one grid in one launch."""
    res = _planner_for(grid).plan(starts, goals)
    return [None if p is None else [(int(x), int(y)) for x, y in p] for p in res]


def astar_pixel(grid, start: Tuple[int, int], goal: Tuple[int, int]) -> Optional[List[Tuple[int, int]]]:
    """Drop-in for vln_trajectory_generator.astar_pixel: list of (x, y) from start to goal, or None.

    One documented difference: a START outside the grid returns None here, while the reference - which keys its
    scores by coordinate tuples - would expand from it (its callers only pass cells of the map: instance centroids
    and free cells found on it, vln_trajectory_generator.py:309-344,487-559)."""
    return astar_pixel_batch(grid, [start], [goal])[0]


def build_2d_map(sem_data: List[Dict], scale: float = SCALE_M_PER_PX, robot_radius_m: float = ROBOT_RADIUS_M):
    """Drop-in for vln_trajectory_generator.build_2d_map: (grid_map, scale, min_x, min_y), or four
    Nones when the map has no usable coordinate or anything fails; fills inst['mask_coords']."""
    try:
        ys, xs = [], []
        for inst in sem_data:
            for y, x in inst.get("mask_coords_m", []):
                try:
                    yv, xv = float(y), float(x)
                except (ValueError, TypeError):
                    continue
                ys.append(yv)
                xs.append(xv)
        if not ys or not xs:
            return None, None, None, None
        min_y, max_y, min_x, max_x = min(ys), max(ys), min(xs), max(xs)
        h = int(np.ceil((max_y - min_y) / scale)) + 1
        w = int(np.ceil((max_x - min_x) / scale)) + 1
        for inst in sem_data:
            cells = []
            for y_m, x_m in inst.get("mask_coords_m", []):
                try:
                    xv, yv = float(x_m), float(y_m)
                except (ValueError, TypeError):
                    continue
                px, py = int(round((xv - min_x) / scale)), int(round((yv - min_y) / scale))
                if 0 <= py < h and 0 <= px < w:
                    cells.append((py, px))
            inst["mask_coords"] = cells
        grid = np.zeros((h, w), dtype=np.uint8)
        for inst in sem_data:
            if str(inst.get("category_label", "")).lower() in OBSTACLE_LABELS:
                for py, px in inst.get("mask_coords", []):
                    grid[py, px] = 1
        if robot_radius_m > 0:
            grid = inflate_grid(grid, scale, robot_radius_m)
        return grid, scale, min_x, min_y
    except Exception as e:
        print(f"[ERROR] Failed to build 2D map: {e}")
        return None, None, None, None


def boundary_pixels(mask_coords) -> List[Tuple[int, int]]:
    """Mask cells with a 4-neighbour outside the mask (iteration order = Python set order, as in
    vln_trajectory_generator.py:298-306)."""
    cells = set((int(y), int(x)) for (y, x) in mask_coords)
    return [(y, x) for (y, x) in cells
            if any(nb not in cells for nb in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)))]


def get_nearest_free_pixel_on_side(instance_mask, base_map: np.ndarray, towards_px=None, max_search_dist: int = 50):
    """Breadth-first search outwards from the instance boundary for the first free cell (optionally
    on the side facing ``towards_px``); (x, y) or None (vln_trajectory_generator.py:309-344)."""
    H, W = base_map.shape
    seeds = boundary_pixels(instance_mask)
    if not seeds:
        return None
    seen, queue = set(), deque()
    for by, bx in seeds:
        if 0 <= bx < W and 0 <= by < H:
            seen.add((bx, by))
            queue.append((bx, by, 0))
    while queue:
        x, y, d = queue.popleft()
        if d > max_search_dist:
            break
        if 0 <= x < W and 0 <= y < H and base_map[y, x] == 0:
            if towards_px is None:
                return (x, y)
            cx, cy = np.mean([(px, py) for (py, px) in instance_mask], axis=0)
            if np.dot(np.array([x - cx, y - cy]), np.array([towards_px[0] - cx, towards_px[1] - cy])) >= 0:
                return (x, y)
        for dx, dy in ((-1, 0), (1, 0), (0, -1), (0, 1)):
            nx, ny = x + dx, y + dy
            if 0 <= nx < W and 0 <= ny < H and (nx, ny) not in seen:
                seen.add((nx, ny))
                queue.append((nx, ny, d + 1))
    return None


def build_connectivity_map(grid_map: np.ndarray, itemid2inst: Dict[str, Dict]) -> Dict[Tuple[int, int], set]:
    """Drop-in for vln_trajectory_generator.build_connectivity_map: items grouped by the free-space
    component their nearest free cell lies in.  Keys are a representative (x, y) of the component
    (the reference's union-find root is an arbitrary member too; callers only use the groups)."""
    labels = connectivity_labels(grid_map)
    W = grid_map.shape[1]
    groups: Dict[Tuple[int, int], set] = defaultdict(set)
    for item_id, inst in itemid2inst.items():
        if "mask_coords" not in inst or not inst["mask_coords"]:
            continue
        free = get_nearest_free_pixel_on_side(inst["mask_coords"], grid_map)
        if free:
            root = int(labels[free[1], free[0]])
            groups[(root % W, root // W)].add(item_id)
    return dict(groups)


def generate_trajectory_points(path, scale: float, min_x: float, min_y: float, fixed_z: float = FIXED_Z,
                               sample_step: int = SAMPLE_STEP) -> List[Dict]:
    """Drop-in for vln_trajectory_generator.generate_trajectory_points (:703-741): path cells ->
    way-points in the trajectory JSON format (cell centres in metres, heading towards the next
    sampled point as a z-rotation quaternion [0, 0, sin(yaw/2), cos(yaw/2)]; the last point keeps
    yaw 0).  Pure host arithmetic in binary64, same expressions as the reference."""
    import math
    world = [(min_x + (x + 0.5) * scale, min_y + (y + 0.5) * scale) for x, y in path][::sample_step]
    points = []
    for j, (wx, wy) in enumerate(world):
        nx, ny = world[j + 1] if j < len(world) - 1 else world[j]
        yaw = math.atan2(ny - wy, nx - wx)
        points.append({
            "point": str(j),
            "position": [wx, wy, fixed_z],
            "rotation": [0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0)],
            "action": [],
            "camera_images": [],
            "focal_length": 7.0,
            "horizontal_aperture": 20.954999923706055,
            "vertical_aperture": 20.954999923706055,
            "focus_distance": 0.0,
            "clipping_range": [1.0, 1000000.0],
        })
    return points


def plan_trajectories(grid_map, starts, goals, scale: float, min_x: float, min_y: float) -> List[Optional[List[Dict]]]:
    """Many start/goal cell pairs -> way-point lists (None where no path exists): one batched A*
    launch, then the reference's path -> pose conversion."""
    return [None if p is None else generate_trajectory_points(p, scale, min_x, min_y)
            for p in astar_pixel_batch(grid_map, starts, goals)]


def are_items_connected(item1: str, item2: str, connectivity_map) -> bool:
    return any(item1 in items and item2 in items for items in connectivity_map.values())


__all__ = ["GridPlanner", "astar_pixel", "astar_pixel_batch", "build_2d_map", "build_connectivity_map",
           "generate_trajectory_points", "plan_trajectories",
           "are_items_connected", "get_nearest_free_pixel_on_side", "boundary_pixels", "inflate_grid",
           "connectivity_labels"]
