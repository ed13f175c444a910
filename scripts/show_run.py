"""developer: print the launch list around the rank kernel + the bench lines of a gpurun_out/<dir>"""
import csv, json, glob, sys
d = sys.argv[1]
for f in sorted(glob.glob(f'{d}/launches_*.csv')):
    rows = list(csv.reader(open(f)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID']
    if not hdr: continue
    H = rows[hdr[0]]; data = rows[hdr[0] + 2:]
    ki = H.index('Kernel Name'); vi = H.index('Metric Value')
    seq = [(r[ki][:70], float(r[vi].replace(',', ''))) for r in data if len(r) > vi]
    print(f)
    idx = [i for i, (n, v) in enumerate(seq) if 'rank_kernel' in n or 'zsl_tc' in n]
    if idx:
        i = idx[min(3, len(idx) - 1)]
        for n, v in seq[max(0, i - 7):i + 2]: print('   %-70s %9.1f us' % (n, v / 1000))
for f in sorted(glob.glob(f'{d}/*.json')):
    try:
        x = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms=%.4f' % x['ms_per_step'], 'kern=%.4f' % x['roofline']['kernel_ms'], 'frac=%.3f' % x['roofline']['frac'],
              'host_ms=%.4f' % (x.get('host_ms_per_step') or -1), 'e2e=%.4g' % x['e2e']['value'], 'launches', x['gpu_launches'])
    except Exception as e: print(f, e)
