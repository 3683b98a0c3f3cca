"""The k-NN ladder's exactness argument, modelled on the CPU (no CUDA involved).

fast_gicp_b200/csrc/vgicp_stage1.cu answers every query from a ladder of uniform grids: a level certifies when the k-th distance of
its 3x3x3 block is <= 0.999 s; a block that holds k points but cannot certify them is completed by the 5x5x5 shell (cells whose box
is within the current k-th distance, valid while that distance is <= 2 s * 0.999); otherwise the walk moves to a coarser level and
keeps the k-th distance as a filter.  This file restates those rules with the kernel's float32 cell arithmetic and checks, on
adversarial clouds, that the rows equal the brute-force rows -- so a change of the rules (or of their rounding slack) that breaks
exactness is caught here, before a GPU is involved.  The CUDA kernels themselves are checked against the oracle in
tests/test_gpu_parity.py."""
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


class Ladder:
    def __init__(self, pts, k):
        self.p = pts.astype(f32)
        self.k = k
        n = len(pts)
        self.L = 6 if n < 4096 else 8
        self.mn = self.p.min(axis=0)
        e = max(f32((self.p.max(axis=0) - self.mn).max()), f32(1e-3))
        self.s_max = f32(e * f32(0.25))
        self.cells = []
        for l in range(self.L):
            s = self.size(l)
            inv = f32(1.0) / s
            c = np.floor((self.p - self.mn) * inv).astype(np.int64)
            d = {}
            for i, key in enumerate(map(tuple, c)):
                d.setdefault(key, []).append(i)
            self.cells.append((c, {key: np.asarray(v) for key, v in d.items()}))

    def size(self, l):
        return f32(np.ldexp(self.s_max, l - (self.L - 1)))

    def block(self, l, c, r):
        idx = [self.cells[l][1].get((c[0] + dx, c[1] + dy, c[2] + dz)) for dx in range(-r, r + 1) for dy in range(-r, r + 1) for dz in range(-r, r + 1)]
        idx = [v for v in idx if v is not None]
        return np.concatenate(idx) if idx else np.empty(0, dtype=np.int64)

    def topk(self, q, idx, bound=None):
        d2 = d2_f32(q, self.p[idx])
        order = np.lexsort((idx, d2))
        idx, d2 = idx[order], d2[order]
        if bound is not None:
            keep = (d2 < bound[0]) | ((d2 == bound[0]) & (idx <= bound[1]))
            idx, d2 = idx[keep], d2[keep]
        return idx[: self.k], d2[: self.k]

    def query(self, i, stats):
        q, k = self.p[i], self.k
        bound = None
        for l in range(self.L):
            s = self.size(l)
            if bound is not None:  # skip a level that cannot certify when the next one still would not be guaranteed to
                need = f32(np.sqrt(bound[0]))
                if l + 1 < self.L and s * f32(0.999) < need and self.size(l + 1) * f32(0.999) <= need:
                    continue
            c = self.cells[l][0][i]
            inner = self.block(l, c, 1)
            if len(inner) < k or (2 * len(inner) < 5 * k and l + 1 < self.L):
                continue
            idx, d2 = self.topk(q, inner, bound)
            r1 = s * f32(0.999)
            if len(idx) == k and d2[-1] <= r1 * r1:
                stats["certified"] += 1
                return idx
            if len(idx) == k:
                B = f32(np.sqrt(d2[-1]))
                if B <= f32(2.0) * s * f32(0.999):  # the ball of radius B lies inside the 5x5x5 block: complete from the shell
                    reach = B + f32(1e-3) * s
                    fq = q - self.mn
                    extra = []
                    for dx in range(-2, 3):
                        for dy in range(-2, 3):
                            for dz in range(-2, 3):
                                if max(abs(dx), abs(dy), abs(dz)) != 2:
                                    continue
                                cc = (c[0] + dx, c[1] + dy, c[2] + dz)
                                lo = np.asarray(cc, dtype=f32) * s
                                ex = np.maximum(np.maximum(lo - fq, fq - (lo + s)), f32(0.0))
                                if (ex[0] * ex[0] + ex[1] * ex[1] + ex[2] * ex[2]) <= reach * reach and cc in self.cells[l][1]:
                                    extra.append(self.cells[l][1][cc])
                    cand = np.concatenate([idx] + extra) if extra else idx
                    stats["extended"] += 1
                    return self.topk(q, cand)[0]
                bound = (d2[-1], idx[-1])
        stats["bruteforce"] += 1
        return self.topk(q, np.arange(len(self.p)))[0]


def clouds():
    rng = np.random.default_rng(5)
    surf = np.stack([rng.uniform(-30, 30, 1500), rng.uniform(-30, 30, 1500), rng.normal(0, 0.02, 1500)], axis=1)
    rings = np.concatenate([np.stack([r * np.cos(t), r * np.sin(t), np.full_like(t, -1.7)], axis=1) for r in (3.0, 4.5, 7.0, 12.0, 25.0) for t in [np.linspace(0, 2 * np.pi, 260)]])
    clusters = np.concatenate([c + rng.normal(0, 0.05, (120, 3)) for c in rng.uniform(-50, 50, (10, 3))])
    outliers = np.concatenate([rng.normal(0, 1.0, (900, 3)), rng.uniform(-400, 400, (25, 3))])
    grid = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), np.arange(8.0), indexing="ij"), axis=-1).reshape(-1, 3) * 0.5  # ties and points on cell boundaries
    return {"surface": surf, "rings": rings, "clusters": clusters, "outliers": outliers, "lattice": grid}


@pytest.mark.parametrize("name", ["surface", "rings", "clusters", "outliers", "lattice"])
@pytest.mark.parametrize("k", [5, 20])
def test_ladder_rules_give_the_exact_rows(name, k):
    pts = clouds()[name].astype(f32)
    lad = Ladder(pts, k)
    want = brute_rows(pts, k)
    stats = {"certified": 0, "extended": 0, "bruteforce": 0}
    step = max(1, len(pts) // 400)  # a few hundred queries per cloud keep the CPU suite quick
    for i in range(0, len(pts), step):
        assert np.array_equal(lad.query(i, stats), want[i]), (name, k, i, stats)
    assert stats["certified"] + stats["extended"] + stats["bruteforce"] == len(range(0, len(pts), step))
