import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
seg = 0; segs = collections.OrderedDict()
tot_inst = 0
for r in data:
    src = r[ix['Source']]
    inst = int(r[ix['Instructions Executed']] or 0)
    samp = int(r[ix['# Samples']] or 0)
    d = segs.setdefault(seg, dict(inst=0, samp=0, n=0, ops=collections.Counter(), stalls=collections.Counter()))
    d['inst'] += inst; d['samp'] += samp; d['n'] += 1
    op = src.split()[0] if not src.strip().startswith('@') else src.split()[1]
    d['ops'][op.split('.')[0]] += inst
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v: d['stalls'][c] += v
    tot_inst += inst
    if 'BAR.SYNC' in src or 'EXIT' in src:
        seg += 1
print('total warp-instr', tot_inst)
for s, d in segs.items():
    print('--- segment', s, 'sass lines', d['n'], 'warp-instr', d['inst'], '(%.1f%%)' % (100.0 * d['inst'] / tot_inst), 'samples', d['samp'])
    print('   ops:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(d['inst'], 1)) for k, v in d['ops'].most_common(10)))
    print('   stalls:', ', '.join('%s %d' % (k.replace('stall_', ''), v) for k, v in d['stalls'].most_common(6)))
