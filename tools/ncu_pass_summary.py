#!/usr/bin/env python
"""Summarises an `ncu --csv` launch list of tools/prof_pass.py (metrics gpu__time_duration.sum, dram__bytes_*.sum,
smsp__inst_executed.sum, ...thread_inst_executed_per_inst_executed.ratio) into a per-launch table of the LAST sample pass
in the file and into profiles/<round>_k_trace_dram.json (what bench.py reports as roofline.traffic).

  python tools/ncu_pass_summary.py gpurun_out/final_pass_launches.csv profiles/r1_pass_launches_final.md profiles/r1_k_trace_dram.json"""
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    iK, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    d, names = {}, {}
    for r in rows[hi + 1:]:
        if len(r) <= iV:
            continue
        k = int(r[iID])
        d.setdefault(k, {})[r[iM]] = float(r[iV].replace(',', ''))
        names[k] = r[iK].split('(')[0].replace('lys::', '').replace('void ', '')
    return [(names[k], d[k]) for k in sorted(d)]


def table(last):
    lines = ['| # | kernel | us | DRAM MB (read + write) | warp instr (M) | threads / instr |', '|---|---|---|---|---|---|']
    tot_us = 0.0
    tr = []
    tail_us = 0.0
    winst = tinst = tr_winst = tr_tinst = 0.0
    for i, (n, m) in enumerate(last):
        us = m.get('gpu__time_duration.sum', 0.0) / 1000.0
        mb = (m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)) / 1e6
        tot_us += us
        wi = m.get('smsp__inst_executed.sum', 0.0)
        ti = wi * m.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0.0)
        winst += wi; tinst += ti
        if n.startswith('k_trace') or n.startswith('k_generate_trace'):      # the camera-ray launch carries trace(-1)
            tr.append((us, mb * 1e6))
            tr_winst += wi; tr_tinst += ti
        if n.startswith('k_tail'):
            tail_us += us
        lines.append(f"| {i} | {n} | {us:.1f} | {mb:.1f} | {m.get('smsp__inst_executed.sum', 0.0) / 1e6:.1f} | {m.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0.0):.1f} |")
    share = sum(u for u, _ in tr) / tot_us
    lines.append('')
    lines.append(f'{len(last)} launches, {tot_us:.0f} us of serialised kernel time; k_trace: {len(tr)} launches, {sum(u for u, _ in tr):.0f} us '
                 f'({share * 100:.0f} % of the pass), {sum(b for _, b in tr) / 1e6:.0f} MB of DRAM traffic = {sum(b for _, b in tr) / len(tr) / 1e6:.1f} MB per launch.')
    return lines, {'launches_per_pass': len(tr), 'dram_bytes_per_pass': sum(b for _, b in tr), 'dram_bytes_per_launch': sum(b for _, b in tr) / len(tr),
                   'k_trace_us': sum(u for u, _ in tr), 'k_tail_us': tail_us, 'traversal_us': sum(u for u, _ in tr) + tail_us, 'pass_us': tot_us, 'k_trace_share_of_pass_ncu': share,
                   # issue roofline inputs (what bench.py's roofline.issue is computed from): warp-level instructions of one pass and
                   # the instruction-weighted number of active threads per issued instruction
                   'warp_inst_per_pass': winst, 'threads_per_inst': tinst / winst if winst else None,
                   'k_trace_warp_inst_per_pass': tr_winst, 'k_trace_threads_per_inst': tr_tinst / tr_winst if tr_winst else None}


def main():
    src, md, js = sys.argv[1], sys.argv[2], sys.argv[3]
    L = load(src)
    gens = [i for i, (n, _) in enumerate(L) if n.startswith('k_generate')]
    first = L[gens[0]:gens[1]] if len(gens) > 1 else L[gens[0]:]     # no queue-length estimates yet: one launch per stage and bounce
    last = L[gens[-1]:]                                               # steady state: late grids sized, fused tail
    t1, j1 = table(first)
    t2, j2 = table(last)
    out = ['## first pass of a frame: one launch per stage and bounce (the sequence bench.py profiles per class)', ''] + t1 + \
          ['', '## steady-state pass: late grids sized from the previous pass, fused tail (k_tail)', ''] + t2
    open(md, 'w').write('\n'.join(out) + '\n')
    json.dump({'kernel': 'k_trace', 'per_bounce_sequence': j1, 'steady_state_sequence': j2,
               'source': 'ncu gpu__time_duration.sum, dram__bytes_read.sum + dram__bytes_write.sum over the launches of two 1080p CornellBox passes (%s)' % src},
              open(js, 'w'), indent=1)
    print(t1[-1]); print(t2[-1])


if __name__ == '__main__':
    main()
