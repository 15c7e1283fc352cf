/*
 * plan_oracle.c — CPU ORACLE for the map/planning row.  TEST INFRASTRUCTURE
 * ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * What it restates (reference = /root/reference/Code):
 *  (1) so_grid_inflate — robot-radius inflation of the obstacle grid,
 *        dist_m = distance_transform_edt(grid == 0, sampling=scale); grid = dist_m <= r
 *      benchmark/environment_evaluation/collision_detector.py:119-127 and
 *      data_pipeline/trajectory_generation/vln_trajectory_generator.py:549-552.
 *      scipy evaluates the distance of a cell to its nearest obstacle cell as
 *      sqrt((dy*s)*(dy*s) + (dx*s)*(dx*s)) in binary64 (ndimage/_morphology.py: dt *= sampling;
 *      multiply(dt, dt, dt); add.reduce(axis 0); sqrt), so "dist <= r" is restated as "some obstacle
 *      cell has that expression <= r".  Degenerate case: a grid WITHOUT any obstacle cell — scipy's
 *      output is an artefact of its feature transform there; this restatement returns all-free.
 *  (2) so_grid_components — 4-connected components of the free cells, the partition
 *      build_connectivity_map's union-find produces (vln_trajectory_generator.py:639-666).  Label =
 *      smallest linear index (y*W + x) of the component, -1 for obstacle cells.
 *  (3) so_astar — astar_pixel (vln_trajectory_generator.py:253-286) statement for statement: 8
 *      neighbours in the order (-1,0),(1,0),(0,-1),(0,1),(-1,-1),(-1,1),(1,-1),(1,1) (dx,dy); no
 *      closed set; heap entries (f, (x, y)) popped in tuple order; g/f in binary64 with
 *      math.hypot(dx, dy) == sqrt((double)(dx*dx + dy*dy)) for integer arguments (checked
 *      exhaustively to 1100 in tests/test_plan_oracle.py); corner cutting allowed; start cell not
 *      tested for occupancy.  The heap here is a plain binary heap: the pop ORDER is fixed by the
 *      total order on (f, x, y) (duplicates are indistinguishable), not by the heap's layout.
 *
 * PINNED: tests/golden/plan_*.npz were recorded by importing the UNMODIFIED reference functions
 * (tests/golden/make_plan_golden.py) on seeded maps; tests/test_plan_oracle.py requires identical
 * grids, partitions and paths, and re-runs the live reference where /root/reference exists.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int so_grid_inflate(const uint8_t* raw, int H, int W, double s, double r, uint8_t* out) {
    if (H <= 0 || W <= 0) return 2;
    int any = 0;
    for (long i = 0; i < (long)H * W; ++i) any |= raw[i] != 0;
    if (!any || !(r > 0.0)) {
        for (long i = 0; i < (long)H * W; ++i) out[i] = raw[i] != 0;
        return 0;
    }
    int R = (int)floor(r / s) + 1;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int hit = 0;
            for (int dy = -R; dy <= R && !hit; ++dy) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -R; dx <= R; ++dx) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= W || !raw[(long)yy * W + xx]) continue;
                    double a = (double)dy * s, b = (double)dx * s;
                    double d = sqrt(a * a + b * b);
                    if (d <= r) {
                        hit = 1;
                        break;
                    }
                }
            }
            out[(long)y * W + x] = (uint8_t)hit;
        }
    return 0;
}

int so_grid_components(const uint8_t* grid, int H, int W, int32_t* labels) {
    long n = (long)H * W;
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    if (!stack) return 3;
    for (long i = 0; i < n; ++i) labels[i] = -1;
    for (long i = 0; i < n; ++i) {
        if (grid[i] != 0 || labels[i] != -1) continue;
        long top = 0;
        stack[top++] = (int32_t)i; /* i is the smallest index of its component (scan order) */
        labels[i] = (int32_t)i;
        while (top) {
            long c = stack[--top];
            int y = (int)(c / W), x = (int)(c % W);
            const int dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
            for (int k = 0; k < 4; ++k) {
                int nx = x + dx[k], ny = y + dy[k];
                if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
                long j = (long)ny * W + nx;
                if (grid[j] != 0 || labels[j] != -1) continue;
                labels[j] = (int32_t)i;
                stack[top++] = (int32_t)j;
            }
        }
    }
    free(stack);
    return 0;
}

typedef struct {
    double f;
    int32_t x, y;
} hent;

static int hless(const hent* a, const hent* b) {
    if (a->f < b->f) return 1;
    if (a->f > b->f) return 0;
    if (a->x != b->x) return a->x < b->x;
    return a->y < b->y;
}

/* Returns the path length (>= 1) with path_xy = (x0,y0,x1,y1,...) from start to goal, 0 if the
 * reference would return None, -3 if the path does not fit max_path points. */
int so_astar(const uint8_t* grid, int H, int W, int sx, int sy, int gx, int gy, int32_t* path_xy, int max_path,
             int64_t* expansions_out) {
    const long n = (long)H * W;
    double* g = (double*)malloc(sizeof(double) * (size_t)n);
    int32_t* came = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    long cap = 1024, hn = 0;
    hent* heap = (hent*)malloc(sizeof(hent) * (size_t)cap);
    for (long i = 0; i < n; ++i) {
        g[i] = INFINITY; /* "nb not in g_score" */
        came[i] = -1;
    }
    const int ddx[8] = {-1, 1, 0, 0, -1, -1, 1, 1}, ddy[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    int64_t expansions = 0;
    int result = 0;
    /* the reference indexes g_score by the tuple, so an out-of-grid start is legal there; the
     * callers never do that and this restatement requires the start inside the grid */
    if (sx < 0 || sx >= W || sy < 0 || sy >= H) goto done;
    g[(long)sy * W + sx] = 0.0;
    heap[hn++] = (hent){0.0, sx, sy};
    while (hn) {
        hent cur = heap[0];
        hent last = heap[--hn];
        long i = 0;
        for (;;) { /* sift down */
            long c = 2 * i + 1;
            if (c >= hn) break;
            if (c + 1 < hn && hless(&heap[c + 1], &heap[c])) ++c;
            if (!hless(&heap[c], &last)) break;
            heap[i] = heap[c];
            i = c;
        }
        if (hn) heap[i] = last;
        ++expansions;
        if (cur.x == gx && cur.y == gy) {
            long len = 1, c = (long)cur.y * W + cur.x;
            while (came[c] >= 0) {
                c = came[c];
                ++len;
            }
            if (len > max_path) {
                result = -3;
                goto done;
            }
            c = (long)cur.y * W + cur.x;
            for (long k = len - 1; k >= 0; --k) {
                path_xy[2 * k] = (int32_t)(c % W);
                path_xy[2 * k + 1] = (int32_t)(c / W);
                c = came[c];
            }
            result = (int)len;
            goto done;
        }
        const double gc = g[(long)cur.y * W + cur.x];
        for (int d = 0; d < 8; ++d) {
            int nx = cur.x + ddx[d], ny = cur.y + ddy[d];
            if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
            if (grid[(long)ny * W + nx] == 1) continue;
            double step = sqrt((double)(ddx[d] * ddx[d] + ddy[d] * ddy[d]));
            double tg = gc + step;
            long j = (long)ny * W + nx;
            if (tg < g[j]) { /* g[j] == inf <=> not in g_score */
                came[j] = (int32_t)((long)cur.y * W + cur.x);
                g[j] = tg;
                int hx = nx - gx, hy = ny - gy;
                double f = tg + sqrt((double)((long)hx * hx + (long)hy * hy));
                if (hn == cap) {
                    cap *= 2;
                    heap = (hent*)realloc(heap, sizeof(hent) * (size_t)cap);
                }
                long k = hn++;
                hent e = {f, nx, ny};
                while (k > 0) { /* sift up */
                    long p = (k - 1) / 2;
                    if (!hless(&e, &heap[p])) break;
                    heap[k] = heap[p];
                    k = p;
                }
                heap[k] = e;
            }
        }
    }
done:
    if (expansions_out) *expansions_out = expansions;
    free(g);
    free(came);
    free(heap);
    return result;
}
