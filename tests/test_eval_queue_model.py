"""Bookkeeping of the compacted evaluation kernel (fast_gicp_b200/csrc/vgicp_kernels.cuh: lin_accumulate_impl), modelled on the CPU:
lanes of a warp split the neighbour cells of 32/G points, every pass appends its hits to a 256-entry ring by ballot + prefix, full
batches of 32 are drained after each pass and the remainder after the last one.  The model checks that every (point, voxel) hit
is consumed exactly once, that the ring never holds more than its capacity, and that the lane mapping (contiguous z-columns for
DIRECT27, strided offsets otherwise) covers every (point, offset) pair exactly once -- the invariants a refactor must keep.
The arithmetic itself is checked on the GPU against the oracle (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

K_QUEUE = 256


def warp_pass_structure(mode, G, n_off):
    columns = mode == 27
    if columns:
        cells = 3
        n_pass = 9 // G
    else:
        per_lane = -(-n_off // G)
        cells = 4 if mode == 0 else min(per_lane, 4)
        n_pass = -(-n_off // (G * cells))
    return columns, cells, n_pass


def run_warp(hit, mode, G, n_points_total, base_task):
    """hit[point, offset] -> voxel id or -1.  Returns the list of consumed (point, voxel) pairs of one warp iteration."""
    n_off = hit.shape[1]
    columns, cells, n_pass = warp_pass_structure(mode, G, n_off)
    tpw = (32 // G) * G
    n_tasks = n_points_total * G
    ring = [None] * K_QUEUE
    head = tail = 0
    consumed, visited = [], []
    for p in range(n_pass):
        for j in range(cells):  # one ballot per cell slot
            entrants = []
            for lane in range(32):
                task = base_task + lane
                if not (lane < tpw and task < n_tasks):
                    continue
                point, sub = task // G, task % G
                o = sub * (27 // G) + p * 3 + j if columns else p * (G * cells) + sub + j * G
                if o >= n_off:
                    continue
                visited.append((point, o))
                if hit[point, o] >= 0:
                    entrants.append((lane // G, point, hit[point, o]))
            for e in entrants:  # ballot order = lane order
                ring[tail % K_QUEUE] = e
                tail += 1
            assert tail - head <= K_QUEUE
        while tail - head >= 32:
            consumed += [ring[(head + l) % K_QUEUE] for l in range(32)]
            head += 32
    consumed += [ring[(head + l) % K_QUEUE] for l in range(tail - head)]
    return consumed, visited


@pytest.mark.parametrize("mode,G,n_off", [(27, 1, 27), (27, 3, 27), (7, 1, 7), (7, 4, 7), (1, 1, 1), (0, 1, 33), (0, 8, 33), (0, 8, 123)])
@pytest.mark.parametrize("density", [0.0, 0.28, 1.0])
def test_every_hit_is_consumed_exactly_once(mode, G, n_off, density):
    rng = np.random.default_rng(n_off * 7 + G)
    n_points = 70  # not a multiple of the points per warp: the last warp is partial
    hit = np.where(rng.random((n_points, n_off)) < density, rng.integers(0, 1 << 20, (n_points, n_off)), -1)
    tpw = (32 // G) * G
    got, seen = [], []
    for base in range(0, n_points * G, tpw):
        c, v = run_warp(hit, mode, G, n_points, base)
        got += [(p, vox) for _, p, vox in c]
        seen += v
        for slot, p, _ in c:  # a queue entry names the staging slot of its point inside the warp
            assert slot == (p * G - base) // G
    want = [(p, hit[p, o]) for p in range(n_points) for o in range(n_off) if hit[p, o] >= 0]
    assert sorted(got) == sorted(want)
    assert sorted(seen) == [(p, o) for p in range(n_points) for o in range(n_off)]
