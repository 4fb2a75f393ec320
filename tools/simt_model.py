#!/usr/bin/env python
"""SIMT cost model of the closest-hit traversal loop of k_trace (extend items), run on the CPU.

Study tool (uses the oracle, so it is test infrastructure like tests/).  It takes the bounce-b closest-hit rays of one
pass from the oracle (orc_probe_path_rays), every ray's visit pattern in the reference's left-first order
(orc_closest_hits_pattern: box fail / box pass / triangle miss / triangle hit), groups the rays into warps of 32 in a
candidate slot order and counts what a lock-step warp issues for a candidate loop shape.

  python tools/simt_model.py orders    [scene] [h w]   ray orderings (pixel order, octant sorts, walk-length bound)
  python tools/simt_model.py schedules [scene] [h w]   loop shapes: stages per iteration ('B' box, 'T' triangle)
  scene: a bundled scene or `synthetic[:k]` (BASELINE config 5); LYS_MODEL_MAX_STEPS bounds the modelled walk length (400)

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


MAX_STEPS = 400


def load_scene(scene):
    """a bundled scene name, or `synthetic` / `synthetic:k` = BASELINE config 5 (every Cornell quad as a k x k grid, k = 151)"""
    name, _, k = scene.partition(':')
    d = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'scenes', ('cornell' if name == 'synthetic' else name) + '.npz')))
    if name == 'synthetic':
        sys.path.insert(0, ROOT)
        import importlib
        scenes = importlib.import_module('msc-futhark-ray-tracer_b200.scenes')
        d['tris'], d['tri_mats'] = scenes.synthetic_cornell(d['tris'], d['tri_mats'], int(k or 151))
    return d


def bounce_rays(scene, h, w):
    d = load_scene(scene)
    st = orc.State.init(d['tris'], d['tri_mats'], d['mats'], h, w)
    rays, nr = st.probe_path_rays()
    rays, nr = rays.reshape(-1, 16, 6), nr.reshape(-1)
    print(f'{scene} {w}x{h}: {len(d["tris"])} triangles, closest-hit rays per path {nr.mean():.3f}')
    for b in (0, 1, 2):
        live = np.nonzero(nr > b)[0]                       # pixel order == the compacted slot order k_shade produces
        r = rays[live, b]
        pat, ln = st.closest_hits_pattern(r, MAX_STEPS)
        if ln.max() > MAX_STEPS:
            print(f'   (bounce {b}: {int((ln > MAX_STEPS).sum())} walks longer than {MAX_STEPS} visits are truncated in the model)')
            ln = np.minimum(ln, MAX_STEPS)
        yield b, r, pat, ln


def morton_cells(p, bits):
    """Morton index of points on a 2^bits grid over their bounding box"""
    lo, hi = p.min(axis=0), p.max(axis=0)
    q = np.minimum(((p - lo) / np.maximum(hi - lo, 1e-20) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    code = np.zeros(len(p), np.int64)
    for bit in range(bits):
        for ax in range(3):
            code |= ((q[:, ax] >> bit) & 1) << (3 * bit + (2 - ax))
    return code


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


def refill_cost(pat, ln, order, slice_len, keep, costs, refill=30):
    """k_trace_refill (LYS_TRACE_MODE=1): a warp owns a contiguous slice of `slice_len` rays; every lane does ONE visit per
    step; when fewer than `keep` lanes are busy (or the warp is empty) the idle lanes pull the next rays of the slice
    (`refill` warp instructions per refill round).  Returns warp instructions per ray and steps per warp."""
    p = pat[order]
    l = ln[order].astype(np.int64)
    n = len(p)
    kind = np.where(p <= 1, 0, np.where(p <= 3, 1, 2)).astype(np.int8)
    n_warps = (n + slice_len - 1) // slice_len
    cursor = np.arange(n_warps, dtype=np.int64) * slice_len
    end = np.minimum(cursor + slice_len, n)
    ray = np.full((n_warps, 32), -1, np.int64)               # ray of each lane, -1 = idle
    ptr = np.zeros((n_warps, 32), np.int64)
    total = steps = 0
    while True:
        busy = ray >= 0
        nb = busy.sum(axis=1)
        want = ((nb < keep) | (nb == 0)) & (cursor < end)    # refill round
        if want.any():
            for wi in np.nonzero(want)[0]:
                idle = np.nonzero(ray[wi] < 0)[0]
                take = min(len(idle), int(end[wi] - cursor[wi]))
                ray[wi, idle[:take]] = cursor[wi] + np.arange(take)
                ptr[wi, idle[:take]] = 0
                cursor[wi] += take
            total += refill * int(want.sum())
            busy = ray >= 0
        # lanes whose ray has no visits at all finish at once
        if not busy.any():
            break
        rr = np.where(busy, ray, 0)
        k = np.where(busy, kind[rr, np.minimum(ptr, kind.shape[1] - 1)], 2)
        done0 = busy & (ptr >= l[rr])
        k = np.where(done0, 2, k)
        aw = busy.any(axis=1)
        steps += int(aw.sum())
        total += costs['loop'] * int(aw.sum()) + costs['B'] * int((k == 0).any(axis=1).sum()) + costs['T'] * int((k == 1).any(axis=1).sum())
        ptr = ptr + busy
        fin = busy & (ptr >= l[rr])
        ray = np.where(fin, -1, ray)
    return total / n, steps / n_warps


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
        for bits in (2, 4):                                 # spatial sorts: rays that start in the same cell and head into the same octant
            cell = morton_cells(r[:, :3].astype(np.float64), bits)
            orders[f'origin cell ({1 << bits}^3) then octant'] = np.argsort(cell * 8 + octant, kind='stable')
            orders[f'octant then origin cell ({1 << bits}^3)'] = np.argsort(octant * (1 << (3 * bits)) + cell, kind='stable')
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
            for slice_len in (128, 1024):
                for keep in (20, 28):
                    c, it = refill_cost(pat, ln, base, slice_len, keep, costs)
                    print(f'   box {box} tri {tri}  refill: slices of {slice_len}, refill below {keep} busy lanes {c:7.1f} warp-instr/ray  {it:7.1f} steps/warp')


def pair_cost(pat, ln, order, policy, costs, thr=8, nb=2):
    """Pair records (the round-2 layout that the stackless walk replaced; commit f572386): a failing box test costs no visit of its own, so a ray's stage sequence is its
    reference pattern without the 'box fail' events ('N' = enter a node: both children tested, 'T' = triangle test).
    policy 'static': every iteration runs nb N stages then one T stage (k_trace today);
    policy 'vote':   every iteration runs ONE stage chosen by a warp vote: T if at least `thr` lanes wait at a leaf or no lane
                     is at a node, else N (lanes at leaves wait).  Returns warp instructions per ray, iterations per warp,
                     mean active lanes per N stage and per T stage."""
    p = pat[order]
    n = len(p)
    pad = (-n) % 32
    if pad:
        p = np.concatenate([p, np.full((pad, p.shape[1]), 255, np.uint8)])
    # compress: drop box-fail events (code 0); 1 -> N, 2/3 -> T
    keep = (p == 1) | (p == 2) | (p == 3)
    L = keep.sum(axis=1)
    width = int(L.max()) + 1
    kind = np.full((len(p), width), 2, np.int8)
    idx = np.cumsum(keep, axis=1) - 1
    r, c = np.nonzero(keep)
    kind[r, idx[r, c]] = np.where(p[r, c] == 1, 0, 1)
    ptr = np.zeros(len(p), np.int64)
    rows = np.arange(len(p))
    total = iters = 0
    laneN = stageN = laneT = stageT = 0
    while True:
        k = kind[rows, ptr]
        isN, isT = (k == 0).reshape(-1, 32), (k == 1).reshape(-1, 32)
        alive = isN.any(axis=1) | isT.any(axis=1)
        if not alive.any():
            break
        iters += int(alive.sum())
        total += costs['loop'] * int(alive.sum())
        if policy == 'static':
            for _ in range(nb):
                k = kind[rows, ptr]
                take = (k == 0)
                tw = take.reshape(-1, 32)
                total += costs['N'] * int(tw.any(axis=1).sum()); laneN += int(take.sum()); stageN += int(tw.any(axis=1).sum())
                ptr = ptr + take
            k = kind[rows, ptr]
            take = (k == 1)
            tw = take.reshape(-1, 32)
            total += costs['T'] * int(tw.any(axis=1).sum()); laneT += int(take.sum()); stageT += int(tw.any(axis=1).sum())
            ptr = ptr + take
        else:
            nN, nT = isN.sum(axis=1), isT.sum(axis=1)
            runT = (nT >= thr) | (nN == 0)
            takeT = (isT & runT[:, None]).reshape(-1)
            takeN = (isN & ~runT[:, None]).reshape(-1)
            total += costs['T'] * int((runT & alive).sum()) + costs['N'] * int((~runT & alive).sum())
            laneT += int(takeT.sum()); stageT += int((runT & alive).sum()); laneN += int(takeN.sum()); stageN += int((~runT & alive).sum())
            ptr = ptr + takeT + takeN
    return total / n, iters / (len(p) // 32), laneN / max(stageN, 1), laneT / max(stageT, 1)


def pair_phased_cost(pat, ln, order, caps, costs, nb=2, requeue=120):
    """Pair records, static N..NT iterations, in PHASES: a warp stops after caps[k] iterations; rays that are not finished
    save their walk (stack, current node, tmax), are compacted, and the next phase resumes them in dense warps.
    `requeue` = warp instructions per warp of 32 resumed rays (save + restore).  Returns warp instructions per ray and the
    fraction of rays entering each phase."""
    p = pat[order]
    n = len(p)
    keep = (p == 1) | (p == 2) | (p == 3)
    L = keep.sum(axis=1)
    width = int(L.max()) + 1
    kind = np.full((n, width), 2, np.int8)
    idx = np.cumsum(keep, axis=1) - 1
    r, c = np.nonzero(keep)
    kind[r, idx[r, c]] = np.where(p[r, c] == 1, 0, 1)
    ptr = np.zeros(n, np.int64)
    live = np.arange(n)
    total = 0
    fracs = []
    for cap in list(caps) + [1 << 30]:
        if len(live) == 0:
            break
        fracs.append(len(live) / n)
        m = len(live)
        pad = (-m) % 32
        rows = np.concatenate([live, np.full(pad, -1, np.int64)])
        valid = rows >= 0
        rr = np.where(valid, rows, 0)
        pp = np.where(valid, ptr[rr], width - 1)
        if len(fracs) > 1:
            total += requeue * (len(rows) // 32)
        it = 0
        while it < cap:
            k = np.where(valid, kind[rr, pp], 2)
            alive = (k != 2).reshape(-1, 32).any(axis=1)
            if not alive.any():
                break
            total += costs['loop'] * int(alive.sum())
            for _ in range(nb):
                k = np.where(valid, kind[rr, pp], 2)
                take = k == 0
                total += costs['N'] * int(take.reshape(-1, 32).any(axis=1).sum())
                pp = pp + take
            k = np.where(valid, kind[rr, pp], 2)
            take = k == 1
            total += costs['T'] * int(take.reshape(-1, 32).any(axis=1).sum())
            pp = pp + take
            it += 1
        ptr[rows[valid]] = pp[valid]
        unfinished = valid & (np.where(valid, kind[rr, pp], 2) != 2)
        live = rows[unfinished]
    return total / n, fracs


def study_pair(scene, h, w):
    for b, r, pat, ln in bounce_rays(scene, h, w):
        base = np.arange(len(r))
        costs = {'N': 70, 'T': 75, 'loop': 8}
        print(f'bounce {b}: {len(r)} rays, reference visits per ray mean {ln.mean():.1f}')
        c, it = schedule_cost(pat, ln, base, 'BBT', {'B': 33, 'T': 75, 'loop': 6})
        print(f'   single-box records BBT            {c:7.1f} warp-instr/ray  {it:6.1f} iters/warp')
        for nb in (1, 2, 3):
            c, it, ln_, lt_ = pair_cost(pat, ln, base, 'static', costs, nb=nb)
            print(f'   pair static {"N" * nb}T               {c:7.1f} warp-instr/ray  {it:6.1f} iters/warp  lanes per N stage {ln_:4.1f}, per T stage {lt_:4.1f}')
        for caps in ((8,), (12,), (16,), (24,), (8, 16), (8, 24), (12, 24), (12, 32), (8, 16, 32), (8, 16, 32, 64)):
            c, fr = pair_phased_cost(pat, ln, base, caps, costs)
            print(f'   pair NNT in phases, caps {str(caps):18s} {c:7.1f} warp-instr/ray  rays entering the phases: ' + ' '.join(f'{x:.2f}' for x in fr))
        for thr in (2, 33):
            c, it, ln_, lt_ = pair_cost(pat, ln, base, 'vote', costs, thr=thr)
            print(f'   pair vote, T when >= {thr:2d} at leaves  {c:7.1f} warp-instr/ray  {it:6.1f} iters/warp  lanes per N stage {ln_:4.1f}, per T stage {lt_:4.1f}')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'orders'
    scene = sys.argv[2] if len(sys.argv) > 2 else 'cornell'
    h, w = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (270, 480)
    MAX_STEPS = int(os.environ.get('LYS_MODEL_MAX_STEPS', MAX_STEPS))
    {'schedules': study_schedules, 'pair': study_pair}.get(what, study_orders)(scene, h, w)
