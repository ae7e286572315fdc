"""Long-format ncu --metrics CSVs (one row per launch x metric) -> one table per file: kernel, launch id, metrics."""
import csv, collections, re, sys
for path in sys.argv[1:]:
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    if not lines or not lines[0].startswith('"ID"'):
        continue
    rows = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row['ID'], re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')[:60], row['Grid Size'], row['Block Size'])
        rows.setdefault(key, collections.OrderedDict())[row['Metric Name']] = (row['Metric Value'], row['Metric Unit'])
    print('## %s' % path.split('/')[-1])
    for (i, name, grid, block), m in rows.items():
        print('launch %s  %s  grid %s block %s' % (i, name, grid, block))
        for k, (v, u) in m.items():
            print('    %-72s %14s %s' % (k, v, u))
        try:
            t = float(m['gpu__time_duration.sum'][0].replace(',', '')) * (1e-9 if m['gpu__time_duration.sum'][1] in ('ns', 'nsecond') else 1e-6)
            rd = float(m['dram__bytes_read.sum'][0].replace(',', '')); wr = float(m['dram__bytes_write.sum'][0].replace(',', ''))
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            rd *= scale[m['dram__bytes_read.sum'][1]]; wr *= scale[m['dram__bytes_write.sum'][1]]
            print('    %-72s %14.1f GB/s (dram read+write / duration)' % ('-> dram throughput', (rd + wr) / t / 1e9))
        except Exception:
            pass
