"""The pin kit (oracle/pin/*.fut, tools/make_pin.py): the committed `futhark test` programs carry exactly what the current
oracle computes, and every third-party semantic sits behind a named switch of include/lys_pins.h that the kit's README maps to
a pin program.  (Nobody here can RUN them: no futhark compiler in the image; this test keeps them in sync with the oracle.)"""
import os
import re
import subprocess
import sys

from conftest import ROOT

PIN = os.path.join(ROOT, 'oracle', 'pin')


def test_pin_programs_are_in_sync_with_the_oracle(orc, tmp_path):
    before = {f: open(os.path.join(PIN, f)).read() for f in sorted(os.listdir(PIN))}
    assert {'pin_rand.fut', 'pin_stat.fut', 'pin_vec.fut', 'pin_argb.fut', 'pin_sort.fut', 'pin_bvh.fut', 'pin_shapes.fut', 'pin_render.fut', 'README.md'} <= set(before)
    subprocess.check_call([sys.executable, os.path.join(ROOT, 'tools', 'make_pin.py')], stdout=subprocess.DEVNULL)
    after = {f: open(os.path.join(PIN, f)).read() for f in sorted(os.listdir(PIN))}
    assert before == after, 'oracle/pin is stale: run python tools/make_pin.py and commit'


def test_every_pin_switch_is_named_in_the_kit():
    hdr = open(os.path.join(ROOT, 'include', 'lys_pins.h')).read()
    switches = set(re.findall(r'#define (LYS_PIN_[A-Z0-9_]+) ', hdr)) - {'LYS_PIN_SHR16'}
    assert len(switches) >= 6
    kit = ''.join(open(os.path.join(PIN, f)).read() for f in os.listdir(PIN))
    for s in switches:
        stem = s.replace('_MIN', '').replace('_MAX', '')
        assert stem in kit, s


def test_pin_blocks_are_well_formed():
    for f in os.listdir(PIN):
        if not f.endswith('.fut'):
            continue
        txt = open(os.path.join(PIN, f)).read()
        entries = set(re.findall(r'^entry (\w+)', txt, re.M))
        tested = re.findall(r'^-- entry: (\w+)\n-- input \{ .* \}\n-- output \{ .* \}$', txt, re.M)
        assert tested and set(tested) <= entries, f
        assert txt.count('-- ==') == len(tested), f


def test_hash_switch_changes_what_it_says(orc):
    """LYS_PIN_HASH_SHIFT_ARITHMETIC = 1: (x >> 16) sign-extends, so the hash differs from the unsigned reading as soon as an
    intermediate has its top bit set; index 0 and the first multiply of small indices are common to both readings."""
    L = orc.lib()

    def logical(x):
        x &= 0xFFFFFFFF
        x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
        x = (((x >> 16) ^ x) * 0x45d9f3b) & 0xFFFFFFFF
        return (x >> 16) ^ x
    assert L.orc_hash(0) == 0 == logical(0)
    differ = sum(L.orc_hash(i) != logical(i) for i in range(1, 2000))
    assert 400 < differ < 1999            # about three quarters of the streams depend on the reading: the pin is worth running
