"""The Morton-grid k-NN's exactness argument, modelled on the CPU (no CUDA involved).

fast_gicp_b200/csrc/vgicp_stage1.cu answers a query in three steps: (1) the 64 points around it in Morton order of the finest
grid cells give an upper bound B on the k-th distance (the kernel stops its bisection at any threshold that at least k window
points meet; the model uses the tightest one, the k-th smallest); (2) the finest level l (cell size s_l = s_0 2^l, not below
l_min) with 0.998 s_l >= B is selected: every point within B of the query lies in the 3x3x3 block of the query's cell there;
(3) the block's cells whose box is within B (+ 2e-3 s slack) are scanned and the k smallest (d2, index) keys kept.  Queries whose
bound exceeds the coarsest cells scan the whole cloud.  This file restates those rules with the kernel's float32 cell arithmetic and checks, on adversarial
clouds, that the rows equal the brute-force rows -- so a change of the rules (or of their rounding slack) that breaks exactness
is caught here, before a GPU is involved.  The CUDA kernels themselves are checked against the oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

f32 = np.float32


def d2_f32(q, pts):
    d = pts - q
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]  # (dx*dx + dy*dy) + dz*dz, float32, as knn_d2


def brute_rows(pts, k):
    out = np.empty((len(pts), k), dtype=np.int64)
    for i, q in enumerate(pts):
        d2 = d2_f32(q, pts)
        out[i] = np.lexsort((np.arange(len(pts)), d2))[:k]  # ascending (d2, index)
    return out


def spread3(v):
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    for sh, m in ((32, 0x1F00000000FFFF), (16, 0x1F0000FF0000FF), (8, 0x100F00F00F00F00F), (4, 0x10C30C30C30C30C3), (2, 0x1249249249249249)):
        x = (x | (x << np.uint64(sh))) & np.uint64(m)
    return x


class MortonGrid:
    def __init__(self, pts, k, levels=None, table_size=None):
        self.p = pts.astype(f32)
        self.k = k
        n = len(pts)
        L = 8
        m = 16384
        while m * 4 <= n and L < 12:
            L += 1
            m *= 4
        if n < 4096:
            L = 6
        self.L = levels or L
        self.mn = self.p.min(axis=0)
        e = max(f32((self.p.max(axis=0) - self.mn).max()), f32(1e-3))
        self.s0 = f32(np.ldexp(e, -(self.L + 1)))
        self.inv_s0 = f32(1.0) / self.s0
        self.cmax = (2 << self.L) - 1
        self.c0 = np.clip(np.floor((self.p - self.mn) * self.inv_s0).astype(np.int64), 0, self.cmax)
        code = (spread3(self.c0[:, 0]) << np.uint64(2)) | (spread3(self.c0[:, 1]) << np.uint64(1)) | spread3(self.c0[:, 2])
        self.order = np.argsort(code, kind="stable")  # LSD radix sort: stable
        self.code = code[self.order]
        self.sorted = self.p[self.order]
        self.pos_of = np.empty(n, dtype=np.int64)
        self.pos_of[self.order] = np.arange(n)
        # cells per level -> l_min (k_grid_levels)
        T = table_size or 1024
        while table_size is None and T < 4 * n:
            T *= 2
        cap, total, self.l_min = T // 2, 0, self.L - 1
        self.ranges = {}
        for l in range(self.L - 1, -1, -1):
            pre = self.code >> np.uint64(3 * l)
            starts = np.flatnonzero(np.r_[True, pre[1:] != pre[:-1]])
            if total + len(starts) > cap and l < self.L - 1:
                break
            total += len(starts)
            self.l_min = l
            ends = np.r_[starts[1:], n]
            self.ranges[l] = {int(pre[s]): (int(s), int(e_)) for s, e_ in zip(starts, ends)}

    def morton(self, c):
        a = np.array([c], dtype=np.int64)
        return int(((spread3(a[:, 0]) << np.uint64(2)) | (spread3(a[:, 1]) << np.uint64(1)) | spread3(a[:, 2]))[0])

    def query(self, i, stats=None):
        """Row of the point at sorted position i, by the kernel's rules; stats collects the number of candidates looked at."""
        k, n = self.k, len(self.p)
        q = self.sorted[i]
        w0 = max(min(i - 32, n - 64), 0)
        pos = np.arange(w0, min(w0 + 64, n))
        seen = len(pos)

        def best(pos):
            d2 = d2_f32(q, self.sorted[pos])
            idx = self.order[pos]
            o = np.lexsort((idx, d2))[:k]
            return pos[o], d2[o]

        pos, d2 = best(pos)
        B = f32(np.sqrt(d2[-1])) if len(pos) == k else f32(np.inf)
        T = d2[-1] if len(pos) == k else f32(np.inf)  # threshold on the squared distance: fixed while the block is scanned
        pos, d2 = pos[:0], d2[:0]  # the window only provides the bound; its points are found again in the block
        l, s = self.l_min, f32(np.ldexp(self.s0, self.l_min))
        while l < self.L and not (B <= f32(0.998) * s):
            l += 1
            s = f32(s * f32(2.0))
        if l >= self.L:  # whole cloud
            pos, d2 = best(np.arange(n))
            if stats is not None:
                stats.append((n, self.L))
            return self.order[pos]
        cq = self.c0[self.order[i]] >> l
        f = q - self.mn
        slack = f32(2e-3) * s
        for dx in (0, -1, 1):  # (the kernel walks the block nearest-first; any order gives the same rows)
            for dy in (0, -1, 1):
                for dz in (0, -1, 1):
                    c = cq + np.array([dx, dy, dz])
                    if (c < 0).any() or (c > (self.cmax >> l)).any():
                        continue
                    lo = c.astype(f32) * s
                    ex = np.maximum(np.maximum(lo - f, f - (lo + s)), f32(0.0)).astype(f32)
                    reach = B + slack
                    if f32((ex[0] * ex[0] + ex[1] * ex[1]) + ex[2] * ex[2]) > reach * reach:
                        continue
                    r = self.ranges[l].get(self.morton(c))
                    if r is None:
                        continue
                    cand = np.arange(r[0], r[1])
                    seen += len(cand)
                    cand = cand[d2_f32(q, self.sorted[cand]) <= T]
                    pos, d2 = best(np.concatenate([pos, cand]))
        if stats is not None:
            stats.append((seen, l))
        return self.order[pos]


def check(pts, k, sample=None, **kw):
    g = MortonGrid(pts, k, **kw)
    want = brute_rows(g.p, k)
    rng = np.random.default_rng(0)
    qs = range(len(pts)) if sample is None else rng.choice(len(pts), sample, replace=False)
    for i in qs:
        row = g.query(int(g.pos_of[i]))
        assert np.array_equal(row, want[i]), i


@pytest.mark.parametrize("shape", ["uniform", "line", "clusters", "outliers", "tiny", "identical", "surface"])
def test_model_rows_are_exact(shape):
    rng = np.random.default_rng(5)
    if shape == "uniform":
        pts = rng.uniform(-20, 20, size=(900, 3))
        pts[100:110] = pts[100]  # duplicates: ties broken by index
    elif shape == "line":
        pts = np.zeros((800, 3))
        pts[:, 0] = np.sort(rng.uniform(0, 500, 800))
    elif shape == "clusters":
        pts = np.concatenate([rng.normal(c, 0.01, size=(120, 3)) for c in rng.uniform(-300, 300, size=(6, 3))])
    elif shape == "outliers":
        pts = rng.normal(0, 1.0, size=(800, 3))
        pts[:5] = rng.uniform(2000, 5000, size=(5, 3))
    elif shape == "tiny":
        pts = rng.uniform(-1, 1, size=(23, 3))
    elif shape == "identical":
        pts = np.tile(np.array([[1.5, -2.0, 0.25]]), (300, 1))
    else:
        xy = rng.uniform(-30, 30, size=(1200, 2))
        pts = np.c_[xy, 0.05 * np.sin(xy[:, 0]) + rng.normal(0, 0.01, 1200)]
    check(pts.astype(f32), 20 if len(pts) > 23 else 7)


def test_model_on_the_fixture_with_a_small_table(pair02):
    """The reference's test cloud, full table and a table so small that the fine levels are dropped (l_min > 0)."""
    tgt, _ = pair02
    check(tgt, 20, sample=150)
    g = MortonGrid(tgt, 20, table_size=2048)
    assert g.l_min > 0
    check(tgt, 20, sample=60, table_size=2048)
