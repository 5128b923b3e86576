"""Shared-memory wavefronts of an `ncu --page source --csv --print-source sass` dump, summed per barrier-delimited
phase and listed per instruction for the worst offenders.  usage: ncu_smem.py src.csv [frames]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
seg = 0; segs = collections.OrderedDict(); worst = []
tot = 0
for r in rows[2:]:
    src = r[ix['Source']]
    wf = int(float(r[ix['L1 Wavefronts Shared']] or 0)); ideal = int(float(r[ix['L1 Wavefronts Shared Ideal']] or 0))
    d = segs.setdefault(seg, [0, 0, 0])
    d[0] += wf; d[1] += ideal; d[2] += int(r[ix['Instructions Executed']] or 0)
    tot += wf
    if wf - ideal > 0: worst.append((wf - ideal, wf, ideal, seg, src.strip()[:80]))
    if 'BAR.SYNC' in src or 'EXIT' in src: seg += 1
print('total shared wavefronts', tot, ('= %.1f per frame' % (tot / frames)) if frames else '')
for s, (wf, ideal, inst) in segs.items():
    if wf: print('segment %d: wavefronts %d (%.1f%%)%s ideal %d excess %d' % (s, wf, 100.0 * wf / tot, (' = %.1f/frame' % (wf / frames)) if frames else '', ideal, wf - ideal))
print('worst excess:')
for e in sorted(worst, reverse=True)[:14]: print('  excess %d wf %d ideal %d seg %d  %s' % e)
