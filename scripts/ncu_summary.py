#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the conv kernels (run HERE, no GPU needed):
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_conv.txt profiles/ncu_traffic.json "<note>"
Writes a markdown table (one row per captured launch) and the per-launch DRAM traffic bench.py reports as roofline.traffic."""
import csv, io, json, subprocess, sys

rep, out_txt, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time_us'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_active_%'),
        ('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_MB'), ('lts__t_sector_hit_rate.pct', 'l2_hit_%'),
        ('smsp__cycles_active.avg', 'smsp_cycles_active'), ('sm__cycles_elapsed.avg', 'sm_cycles_elapsed')]
lines = ['# ' + note, '# ncu --set full --clock-control none; per-launch values.  ncu flushes L2 before every replay pass, so DRAM reads include',
         '# each layer\'s input tensor, which is L2-resident in a real step.', '',
         '| ' + ' | '.join(n for _, n in want) + ' |', '|' + '---|' * len(want)]
tot_r = tot_w = tot_t = 0.0
for r in data:
    vals = []
    for h, n in want:
        v = r[hdr.index(h)] if h in hdr else ''
        if h == 'Kernel Name':
            v = v.split('(se3tn')[0].split('(const')[0].replace('(int)', '').replace('(bool)', '').replace('void ', '').replace('se3tn::', '').replace('<unnamed>::', '')
        else:
            try: v = '%.2f' % float(v)
            except ValueError: pass
        vals.append(v)
    tot_r += float(r[hdr.index('dram__bytes_read.sum')]); tot_w += float(r[hdr.index('dram__bytes_write.sum')]); tot_t += float(r[hdr.index('gpu__time_duration.sum')])
    lines.append('| ' + ' | '.join(vals) + ' |')
n = len(data)
lines += ['', '%d launches: time %.1f us, dram read %.1f MB + write %.1f MB -> %.1f MB per launch on average' % (n, tot_t, tot_r, tot_w, (tot_r + tot_w) / max(n, 1))]
open(out_txt, 'w').write('\n'.join(lines) + '\n')
json.dump({'kernel': 'conv_resident_kernel x8 + conv_trunk_kernel x1 (average over the %d conv launches of one step)' % n, 'launches': n,
           'dram_bytes_per_launch': (tot_r + tot_w) * 1e6 / max(n, 1), 'dram_read_MB_total': tot_r, 'dram_write_MB_total': tot_w,
           'source': out_txt, 'note': note}, open(out_json, 'w'), indent=1)
print('\n'.join(lines[-3:]))
