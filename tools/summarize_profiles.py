"""Turn the ncu outputs under gpurun_out/ into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'          # name of the committed summary (round)
suffix = sys.argv[2] if len(sys.argv) > 2 else tag.replace('r0', 'r')   # suffix of the gpurun_out files
G, P = 'gpurun_out', 'profiles'
os.makedirs(P, exist_ok=True)
out = ['# ncu summaries, round %s' % tag, '',
       'All captured on a B200 through `gpurun` with `--clock-control none`; times under ncu are cold-cache and',
       'serialised (compare SHARES, not absolutes — bench.py measures the live numbers with CUDA events).', '']


def rows_of(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    return list(csv.DictReader(lines))


lst = os.path.join(G, 'launches_%s.csv' % suffix)
if os.path.exists(lst):
    rows = rows_of(lst)
    # one whole step = from one nchw_to_nhwc launch (first kernel of the encoder) to the next; prefer the last
    # complete one (the capture may be truncated by -c)
    starts = [i for i, r in enumerate(rows) if 'nchw_to_nhwc' in r['Kernel Name']]
    has_lbs = lambda seg: any('lbs_extra' in r['Kernel Name'] for r in seg)
    segs = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]
    step = [seg for seg in segs if has_lbs(seg)][-1]
    while 'lbs_extra' not in step[-1]['Kernel Name']:
        step = step[:-1]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in step:
        name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        v = float(r['Metric Value'].replace(',', ''))
        v = v / 1000 if r['Metric Unit'] == 'ns' else (v * 1000 if r['Metric Unit'] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out += ['## Launch list of one step (`ncu --metrics gpu__time_duration.sum`, command: `python tools/profile_step.py 2`)', '',
            'B=32 images, N=100 samples, ResNet-50: %d launches, %.1f us of kernel time under ncu.' % (len(step), tot), '',
            '| kernel | launches | us | share |', '|---|---:|---:|---:|']
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('| `%s` | %d | %.1f | %.1f%% |' % (k[:70], c, t, 100 * t / tot))
    out.append('')
    with open(os.path.join(P, '%s_launches.csv' % tag), 'w') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel', 'grid', 'block', 'duration', 'unit'])
        for r in step:
            w.writerow([r['ID'], r['Kernel Name'].split('(')[0], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit']])

lst2 = os.path.join(G, 'launches_e2e_%s.csv' % suffix)
if os.path.exists(lst2):
    rows = rows_of(lst2)
    starts = [i for i, r in enumerate(rows) if 'proxy_rep_kernel' in r['Kernel Name']]
    ends = [i for i, r in enumerate(rows) if 'samples_reduce_kernel' in r['Kernel Name']]
    if starts and ends and ends[-1] > starts[-1]:
        step = rows[starts[-1]:ends[-1] + 1]
        agg = collections.OrderedDict()
        tot = 0.0
        for r in step:
            name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
            v = float(r['Metric Value'].replace(',', ''))
            v = v / 1000 if r['Metric Unit'] == 'ns' else (v * 1000 if r['Metric Unit'] == 'ms' else v)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
            tot += v
        out += ['## Launch list of one END-TO-END step (`python tools/e2e_launches.py 2`: RGB + 2-D joints -> staged proxy representation -> the step -> per-image metric rows)', '',
                '%d launches, %.1f us of kernel time under ncu.' % (len(step), tot), '',
                '| kernel | launches | us | share |', '|---|---:|---:|---:|']
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append('| `%s` | %d | %.1f | %.1f%% |' % (k[:70], c, t, 100 * t / tot))
        out.append('')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
for name in ('lbs', 'flow', 'conv3x3'):      # conv3x3 = one representative launch: a layer3 3x3 convolution (256 -> 256 at 16x16)
    rep = os.path.join(G, 'prof_%s_%s.ncu-rep' % (name, suffix))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rr[0], rr[1], rr[2]
    out += ['## `%s` (`ncu --set full --import-source on`)' % vals[hdr.index('Kernel Name')].split('(')[0].replace('void <unnamed>::', ''), '',
            '| metric | value | unit |', '|---|---:|---|']
    for w_ in WANT:
        if w_ in hdr:
            i = hdr.index(w_)
            out.append('| %s | %s | %s |' % (w_, vals[i], units[i]))
    out.append('')
conv = os.path.join(G, 'prof_conv_%s.csv' % suffix)
if os.path.exists(conv):
    import json
    rows = rows_of(conv)
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(r['ID'], {'name': r['Kernel Name'].split('conv_tcgen05_kernel')[1].split('(')[0], 'grid': r['Grid Size']})
        d[r['Metric Name']] = (float(r['Metric Value'].replace(',', '')), r['Metric Unit'])

    def us(d):
        v, u = d['gpu__time_duration.sum']
        return v / 1000 if u == 'ns' else (v if u == 'us' else v * 1000)

    def mb(d, k):
        v, u = d[k]
        return {'byte': v / 1e6, 'Kbyte': v / 1e3, 'Mbyte': v, 'Gbyte': v * 1e3}[u]
    out += ['## `conv_tcgen05_kernel`, the 49 convolution launches (53 convolutions, the four downsample branches fused) of one ResNet-50 forward at B=32 (`ncu --metrics ...`)', '',
            '| # | template <BN,STAGES,SR> | CTAs | us | tensor pipe % | DRAM rd MB | DRAM wr MB | L2->SM MB | L2 hit % |',
            '|---:|---|---:|---:|---:|---:|---:|---:|---:|']
    tot = wsum = traffic = 0.0
    for n, d in enumerate(by.values()):
        t = us(d)
        tp = d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]
        tot += t
        wsum += t * tp
        traffic += mb(d, 'dram__bytes_read.sum') + mb(d, 'dram__bytes_write.sum')
        out.append('| %d | %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |' % (
            n, d['name'], d['grid'].strip('()').split(',')[0], t, tp, mb(d, 'dram__bytes_read.sum'), mb(d, 'dram__bytes_write.sum'),
            mb(d, 'l1tex__m_xbar2l1tex_read_bytes.sum'), d['lts__t_sector_hit_rate.pct'][0]))
    out += ['', 'Total %.1f us under ncu; time-weighted tensor-pipe activity %.1f%%; DRAM traffic of the 49 launches %.1f MB.'
            % (tot, wsum / tot, traffic), '']
    json.dump({'kernel': 'conv_tcgen05_kernel x49 (one ResNet-50 forward, B=32, 18ch, 256x256)', 'dram_bytes_per_step': traffic * 1e6,
               'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/%s_summary.md' % tag},
              open(os.path.join(P, '%s_traffic.json' % tag), 'w'))
sass = os.path.join(G, 'sass_mnemonics_%s.txt' % suffix)
if os.path.exists(sass):
    out += ['## Blackwell-native evidence: SASS mnemonics in libhumaniflow_b200.so (`cuobjdump -sass | grep`)', '', '```'] + \
           [l.rstrip() for l in open(sass)] + ['```', '']
open(os.path.join(P, '%s_summary.md' % tag), 'w').write('\n'.join(out))
print('\n'.join(out[:40]))
