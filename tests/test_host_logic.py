"""State machine of src/lib.fut as restated by the oracle: step / key / resize / render / sample_n_frames."""
import numpy as np
from conftest import bits_equal

K = dict(SPACE=0x20, K1=0x31, K2=0x32, a=0x61, d=0x64, i=0x69, k=0x6B, l=0x6C, m=0x6D, n=0x6E, o=0x6F, p=0x70,
         s=0x73, t=0x74, w=0x77, x=0x78, z=0x7A, RIGHT=0x4000004F, LEFT=0x40000050, DOWN=0x40000051, UP=0x40000052)


def test_init_state(orc, scenes):
    s = orc.State.init(*scenes['cornell'], 24, 32)
    sc = s.scalars()
    assert sc['n_frames'] == 0 and sc['mode'] == 0 and sc['subsampling'] == 1 and sc['render_mode'] == 0   # lib.fut:93-101
    assert sc['rng'] == 263559660
    assert s.dims() == (32, 24, 32, 24)
    assert not s.image().any()


def test_accumulation_discards_first_frame(orc, scenes):
    """step uses the OLD n_frames as weight (integrator.fut:184,191): with n_frames = 1 the weights are (0, 1)."""
    s = orc.State.init(*scenes['cornell'], 16, 16).key(K['m'])
    s1 = s.step()                      # n_frames 0 -> sample_frame
    s2 = s1.step()                     # accum with n = 1: image == second frame alone
    assert s1.scalars()['n_frames'] == 1 and s2.scalars()['n_frames'] == 2
    fresh = orc.State.init(*scenes['cornell'], 16, 16).step().step()           # mode off: plain second frame
    assert bits_equal(s2.image(), fresh.image())
    s3 = s2.step()
    f3 = orc.State.init(*scenes['cornell'], 16, 16).step().step().step().image()
    want = np.float32(0.5) * s2.image() + np.float32(0.5) * f3
    assert bits_equal(s3.image(), want.astype(np.float32))


def test_sample_n_frames_equals_steps(orc, scenes):
    s = orc.State.init(*scenes['mirrorbox'], 12, 20)
    img = s.sample_n_frames(4)
    st = s.key(K['m'])
    for _ in range(4):
        st = st.step()
    assert bits_equal(img, st.image())


def test_keys(orc, scenes):
    s = orc.State.init(*scenes['cornell'], 16, 16).key(K['m']).step().step()
    for key in ('w', 'a', 's', 'd', 'x', 'z', 'UP', 'DOWN', 'LEFT', 'RIGHT', 'K2', 'SPACE', 'n', 't'):
        assert s.key(K[key]).scalars()['n_frames'] == 0, key
    for key in ('m', 'i', 'k', 'o', 'l', 'p'):
        assert s.key(K[key]).scalars()['n_frames'] == 2, key
    assert s.key(K['w'], e=1).scalars()['n_frames'] == 2                       # key-up is ignored (lib.fut:121)
    assert s.key(K['K2']).scalars()['subsampling'] == 2 and s.key(K['K1']).scalars()['subsampling'] == 1
    assert s.key(K['K2']).key(K['K2']).key(K['K1']).scalars()['subsampling'] == 2
    c0 = s.scalars()['cam']
    cw = s.key(K['w']).scalars()['cam']
    assert np.allclose(cw[2:5], c0[2:5] + np.array([0, 0, -0.1], np.float32), atol=1e-7)   # forward is -z at yaw 0
    assert abs(s.key(K['UP']).scalars()['cam'][0] + 0.1) < 1e-7 and abs(s.key(K['RIGHT']).scalars()['cam'][1] - 0.1) < 1e-7
    assert abs(s.key(K['i']).scalars()['cam'][5] - 0.08) < 1e-7 and s.key(K['k']).scalars()['cam'][5] == 0.0
    t1 = s.key(K['t'])
    assert t1.scalars()['cam_conf_id'] == 1 and t1.scalars()['render_mode'] == 0
    t2 = t1.key(K['t'])
    assert t2.scalars()['cam_conf_id'] == 2 and t2.scalars()['render_mode'] == 1
    assert t2.key(K['t']).scalars()['cam_conf_id'] == 0
    sky = s.key(K['p']).scalars()['ambience']
    assert sky[1] > 0 and s.key(K['p']).key(K['p']).scalars()['ambience'][1] == 0
    assert s.key(K['SPACE']).scalars()['mode'] == 0 and s.key(K['n']).scalars()['mode'] == 0


def test_resize_and_subsampling(orc, scenes):
    s = orc.State.init(*scenes['cornell'], 16, 24).key(K['m'])
    r = s.resize(10, 14)
    assert r.dims() == (14, 10, 14, 10) and r.scalars()['mode'] == 0
    assert r.step().image().shape == (10, 14, 3)
    q = s.key(K['K2']).step()
    assert q.image().shape == (8, 12, 3) and q.render().shape == (16, 24)
    img, px = q.image(), q.render()
    c = np.clip(img, 0, 1) * np.float32(255)
    want = (255 << 24) | (c[..., 0].astype(np.uint32) << 16) | (c[..., 1].astype(np.uint32) << 8) | c[..., 2].astype(np.uint32)
    assert np.array_equal(px.view(np.uint32), np.repeat(np.repeat(want, 2, axis=0), 2, axis=1).astype(np.uint32))


def test_lidar_and_flash_modes_run(orc, scenes):
    t, tm, m = scenes['cornell']
    lid = orc.State.init(t, tm, m, 12, 16, cam_conf_id=2)
    st, pts = lid.sample_points_n(3)
    assert pts.shape == (12, 16, 4) and st.scalars()['rng'] != lid.scalars()['rng']
    valid = pts[..., 3] > 0
    assert valid.any() and np.all(pts[~valid][:, :3] == -1)
    img = lid.step().image()
    assert img.max() <= 1.0 and img.min() >= 0.0                          # hue colours
    fl = orc.State.init(t, tm, m, 12, 16, cam_conf_id=1).sample_n_frames(2)
    assert np.isfinite(fl).all() and fl.max() > 0


def test_path_len_knob(orc, scenes):
    t, tm, m = scenes['mirrorbox']
    s = orc.State.init(t, tm, m, 16, 16)
    full = s.probe_pass()['radiance']
    orc.set_path_len(5)
    try:
        short = s.probe_pass()['radiance']
    finally:
        orc.set_path_len(16)
    assert bits_equal(short[..., :4], full[..., :4]) and not short[..., 5:].any()


def test_roofline_traffic_matches_the_committed_ncu_list(tmp_path):
    """bench.py's roofline.traffic comes from profiles/r1_k_trace_dram.json; that file must be what tools/ncu_pass_summary.py
    derives from the committed ncu launch list (17 trace launches per pass, per-launch mean of DRAM read + write bytes)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md, js = tmp_path / 'x.md', tmp_path / 'x.json'
    subprocess.check_call([sys.executable, os.path.join(root, 'tools', 'ncu_pass_summary.py'),
                           os.path.join(root, 'profiles', 'r1_pass_launches_final.csv'), str(md), str(js)], stdout=subprocess.DEVNULL)
    new, old = json.load(open(js)), json.load(open(os.path.join(root, 'profiles', 'r1_k_trace_dram.json')))
    for seq in ('per_bounce_sequence', 'steady_state_sequence'):
        assert new[seq]['launches_per_pass'] == old[seq]['launches_per_pass']
        assert abs(new[seq]['dram_bytes_per_launch'] - old[seq]['dram_bytes_per_launch']) < 1.0
    assert old['per_bounce_sequence']['launches_per_pass'] == 17
    # algorithmic bytes per launch (bench.py) are an order of magnitude above the DRAM traffic: the BVH is cache resident
    assert old['per_bounce_sequence']['dram_bytes_per_launch'] < 0.2 * 211e6


def test_path_tile_is_a_bijection(tmp_path):
    """path_tile (csrc/lys_wavefront.h): path id -> (column, local row) with 8x4 pixel tiles per warp.  Compiled for the host
    (the header is plain C++ next to the emulator's cuda_runtime.h stand-in) and checked for every id of grids with ragged
    edges, incomplete last bands, single rows / columns: every pixel exactly once, and a full band's first ids are 8x4 tiles."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / 'tile.cpp'
    src.write_text(r"""
#include "lys_wavefront.h"
#include <cstdio>
#include <vector>
int main() {
    const int sizes[][2] = {{1, 1}, {7, 3}, {8, 4}, {9, 5}, {16, 8}, {17, 9}, {31, 2}, {33, 7}, {128, 96}, {5, 100}, {100, 5}, {1920, 1080}, {3, 4}, {8, 1}};
    for (auto &sz : sizes) {
        const int gw = sz[0], rows = sz[1];
        std::vector<int> seen((size_t)gw * rows, 0);
        for (int pid = 0; pid < gw * rows; pid++) {
            int col = -1, rl = -1;
            lys::path_tile(gw, rows, pid, col, rl);
            if (col < 0 || col >= gw || rl < 0 || rl >= rows) { printf("out of range %d %d pid %d -> %d %d\n", gw, rows, pid, col, rl); return 1; }
            if (seen[(size_t)rl * gw + col]++) { printf("twice %d %d pid %d\n", gw, rows, pid); return 1; }
            const int band = pid / (4 * gw), q = pid % (4 * gw);
            if (band * 4 + 4 <= rows && q < 32 * (gw / 8)) {       /* a full tile: 32 consecutive ids = 8 columns x 4 rows */
                if (col / 8 != q / 32 || rl / 4 != band) { printf("tile %d %d pid %d -> %d %d\n", gw, rows, pid, col, rl); return 1; }
            }
        }
    }
    printf("ok\n");
    return 0;
}
""")
    exe = tmp_path / 'tile'
    subprocess.check_call(['g++', '-std=c++17', '-O1', '-I', os.path.join(root, 'tests', 'simt_emu'), '-I', os.path.join(root, 'msc-futhark-ray-tracer_b200', 'csrc'),
                           '-DLYS_EMU_HOST_ONLY', str(src), '-o', str(exe)])
    assert subprocess.check_output([str(exe)], text=True).strip() == 'ok'
