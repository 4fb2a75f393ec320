#!/usr/bin/env python
"""SIMT cost model of the closest-hit traversal loop of k_trace (extend items), run on the CPU.

Study tool (uses the oracle, so it is test infrastructure like tests/).  It takes the bounce-b closest-hit rays of one
pass from the oracle (orc_probe_path_rays), every ray's visit pattern in the reference's left-first order
(orc_closest_hits_pattern: box fail / box pass / triangle miss / triangle hit), groups the rays into warps of 32 in a
candidate slot order and counts what a lock-step warp issues for a candidate loop shape.

  python tools/simt_model.py orders    [scene] [h w]   ray orderings (pixel order, octant sorts, walk-length bound)
  python tools/simt_model.py schedules [scene] [h w]   loop shapes: stages per iteration ('B' box, 'T' triangle)

Loop shapes: 'X' = one visit per lane and iteration (the first k_trace); 'BT', 'BBT', ... = every iteration runs the
listed stages in order and a lane takes part in a stage if its next visit has that type (k_trace today: 'BBT');
'B*T' = while-while (box stage repeated until no lane of the warp is at an internal node, then one triangle stage).
A stage is charged once per warp and iteration if any lane takes part.  Block sizes (warp instructions) are the measured
ones of profiles/: box stage 45 (33 with octant nodes), triangle stage 40-80 depending on the early-outs, loop 4-8.

What it showed for CornellBox bounce-1 rays (diffuse, incoherent): sorting rays by direction octant gains < 10 %;
'X' -> 'BBT' halves the iterations per warp (38 -> 20); while-while is the worst shape (+30 %); lane refill from a
warp-private queue gains <= 10 % and a two-child ("pair") node layout nothing in issued instructions."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from lysref import oracle as orc  # noqa: E402

BOX, TRI, LOOP = 45, 60, 6


def bounce_rays(scene, h, w):
    d = np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', scene + '.npz'))
    st = orc.State.init(d['tris'], d['tri_mats'], d['mats'], h, w)
    rays, nr = st.probe_path_rays()
    rays, nr = rays.reshape(-1, 16, 6), nr.reshape(-1)
    print(f'{scene} {w}x{h}: closest-hit rays per path {nr.mean():.3f}')
    for b in (0, 1, 2):
        live = np.nonzero(nr > b)[0]                       # pixel order == the compacted slot order k_shade produces
        r = rays[live, b]
        pat, ln = st.closest_hits_pattern(r, 400)
        assert ln.max() <= 400
        yield b, r, pat, ln


def schedule_cost(pat, ln, order, stages, costs):
    """warp instructions per ray and loop iterations per warp of a loop that runs `stages` every iteration"""
    p = pat[order]
    l = ln[order].astype(np.int64)
    n = len(p)
    pad = (-n) % 32
    if pad:
        p = np.concatenate([p, np.full((pad, p.shape[1]), 255, np.uint8)])
        l = np.concatenate([l, np.zeros(pad, np.int64)])
    kind = np.where(p <= 1, 0, np.where(p <= 3, 1, 2)).astype(np.int8)     # 0 box, 1 triangle, 2 none
    kind = np.concatenate([kind, np.full((len(kind), 1), 2, np.int8)], axis=1)
    ptr = np.zeros(len(p), np.int64)
    rows = np.arange(len(p))
    total = 0
    iters = 0
    one_visit = stages == 'X'
    while True:
        alive = ptr < l
        aw = alive.reshape(-1, 32).any(axis=1)
        if not aw.any():
            break
        iters += aw.sum()
        total += costs['loop'] * aw.sum()
        if one_visit:
            k = kind[rows, ptr]
            for want, s in ((0, 'B'), (1, 'T')):
                total += costs[s] * (alive & (k == want)).reshape(-1, 32).any(axis=1).sum()
            ptr = ptr + alive
            continue
        seq = stages
        if stages == 'B*T':
            seq = 'B' * 64 + 'T'
        for s in seq:
            want = 0 if s == 'B' else 1
            take = alive & (kind[rows, ptr] == want)
            if not take.any():
                continue
            total += costs[s] * take.reshape(-1, 32).any(axis=1).sum()
            ptr = ptr + take
            alive = ptr < l
    return total / n, iters / (len(p) // 32)


def lane_stats(pat, order):
    p = pat[order]
    pad = (-len(p)) % 32
    if pad:
        p = np.concatenate([p, np.full((pad, p.shape[1]), 255, np.uint8)])
    p = p.reshape(-1, 32, p.shape[1])
    is_box, is_tri = (p <= 1), ((p == 2) | (p == 3))
    return is_box.sum() / max(1, is_box.any(axis=1).sum()), is_tri.sum() / max(1, is_tri.any(axis=1).sum())


def study_orders(scene, h, w):
    costs = {'B': BOX, 'T': TRI, 'loop': LOOP}
    for b, r, pat, ln in bounce_rays(scene, h, w):
        n = len(r)
        octant = ((r[:, 3] < 0).astype(np.int32) << 2) | ((r[:, 4] < 0).astype(np.int32) << 1) | (r[:, 5] < 0).astype(np.int32)
        dom = np.argmax(np.abs(r[:, 3:6]), axis=1)
        base = np.arange(n)
        orders = {'pixel order (today)': base}
        for blk in (256, 2048):
            orders[f'octant sort within blocks of {blk} live rays'] = np.argsort((base // blk) * 8 + octant, kind='stable')
        orders['global octant sort'] = np.argsort(octant, kind='stable')
        orders['global octant + dominant axis'] = np.argsort(octant * 3 + dom, kind='stable')
        orders['sorted by walk length (bound)'] = np.argsort(ln, kind='stable')
        print(f'bounce {b}: {n} rays, visits per ray mean {ln.mean():.1f} max {ln.max()}')
        for k, o in orders.items():
            c, it = schedule_cost(pat, ln, o, 'X', costs)
            lb, lt = lane_stats(pat, o)
            print(f'   {k:48s} {c:7.1f} warp-instr/ray  {it:5.1f} iters/warp  lanes per box stage {lb:4.1f}, per triangle stage {lt:4.1f}')


def study_schedules(scene, h, w):
    for b, r, pat, ln in bounce_rays(scene, h, w):
        base = np.arange(len(r))
        print(f'bounce {b}: {len(r)} rays, visits per ray mean {ln.mean():.1f}')
        for box, tri in ((45, 80), (45, 60), (33, 60)):
            costs = {'B': box, 'T': tri, 'loop': LOOP}
            for stages in ('X', 'BT', 'BBT', 'BBBT', 'BTT', 'BBTT', 'B*T'):
                c, it = schedule_cost(pat, ln, base, stages, costs)
                print(f'   box {box} tri {tri}  {stages:5s} {c:7.1f} warp-instr/ray  {it:5.1f} iters/warp')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'orders'
    scene = sys.argv[2] if len(sys.argv) > 2 else 'cornell'
    h, w = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (270, 480)
    (study_schedules if what == 'schedules' else study_orders)(scene, h, w)
