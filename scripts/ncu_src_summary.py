"""Summarise an `ncu --page source --csv --print-source sass` export: stall reasons, instruction mix, hottest SASS lines."""
import csv, sys, collections
def main(path, top=14):
    rows = list(csv.reader(open(path)))
    name = rows[0][1][:100]
    H = rows[1]; data = rows[2:]
    idx = {h: i for i, h in enumerate(H)}
    S = idx['Warp Stall Sampling (All Samples)']; X = idx['Instructions Executed']
    tot = sum(int(r[S]) for r in data); ex = sum(int(r[X]) for r in data)
    print('==', name); print('samples', tot, 'warp-instructions', ex)
    reasons = [h for h in H if h.startswith('stall_') and 'Not Issued' not in h]
    c = collections.Counter()
    for r in data:
        for h in reasons:
            c[h] += int(r[idx[h]] or 0)
    print('stalls:', ', '.join('%s %.0f%%' % (k[6:], 100.0 * v / max(1, sum(c.values()))) for k, v in c.most_common(8)))
    ops = collections.Counter()
    for r in data:
        t = r[1].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] += int(r[X])
    print('mix:', ', '.join('%s %.1f%%' % (k, 100.0 * v / ex) for k, v in ops.most_common(12)))
    for r in sorted(data, key=lambda r: -int(r[S]))[:top]:
        why = max(reasons, key=lambda h: int(r[idx[h]] or 0))
        print('  %5d (%4.1f%%) %-60s %s' % (int(r[S]), 100.0 * int(r[S]) / tot, r[1].strip()[:60], why[6:]))
for p in sys.argv[1:]:
    main(p)
