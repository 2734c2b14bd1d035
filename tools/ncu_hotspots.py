#!/usr/bin/env python
"""Per-instruction stall samples of one kernel of an .ncu-rep (read here, no GPU needed):

    python tools/ncu_hotspots.py gpurun_out/r3c_prof_step2048.ncu-rep seg_bwd [min_pct]

Prints the kernel's stall mix, the instruction count by execution level and every SASS instruction holding at least
min_pct (default 0.3) per cent of the samples with its two top stall reasons.  This is what located the seg backward's
constant-bank reloads, the LBS backward's shuffle rounds and the seg forward's lane = part pruning loop in round 2."""
import collections
import csv
import io
import subprocess
import sys

STALLS = ['stall_barrier', 'stall_branch_resolving', 'stall_dispatch', 'stall_long_sb', 'stall_math', 'stall_mio',
          'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_short_sb', 'stall_wait', 'stall_lg']


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:1" % pat],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []
            blocks.append((r[1] if len(r) > 1 else '', cur))
            continue
        if cur is not None:
            cur.append(r)
    name, b = blocks[0]
    hdr, data = b[0], b[1:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot = sum(f(r, '# Samples') for r in data)
    ti = sum(f(r, 'Instructions Executed') for r in data)
    print(name[:150])
    print('samples %d, warp instructions %d, SASS instructions %d' % (tot, ti, len(data)))
    print('stall mix:', {s[6:]: round(sum(f(r, s) for r in data) / tot, 3) for s in STALLS})
    lv, lvs = collections.Counter(), collections.Counter()
    for r in data:
        e = int(f(r, 'Instructions Executed'))
        lv[e] += 1
        lvs[e] += f(r, '# Samples')
    print('by execution level (executions, SASS instructions, share of warp instructions, share of samples):')
    for e, c in sorted(lv.items(), key=lambda x: -x[0] * x[1])[:10]:
        print('  %9d x %4d  %5.1f%%  %5.1f%%' % (e, c, 100.0 * e * c / ti, 100.0 * lvs[e] / tot))
    a0 = int(data[0][ix['Address']], 16)
    for r in data:
        n = f(r, '# Samples')
        if 100.0 * n / tot < min_pct:
            continue
        top = sorted(((f(r, s), s[6:]) for s in STALLS), reverse=True)[:2]
        print('%6x %5.2f%% %9d  %-66s %s' % (int(r[ix['Address']], 16) - a0, 100.0 * n / tot, f(r, 'Instructions Executed'),
                                             r[ix['Source']][:66], ' '.join('%s=%d' % (s, v) for v, s in top if v > 0)))


if __name__ == "__main__":
    main()
