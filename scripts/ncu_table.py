#!/usr/bin/env python
"""One table row per captured launch of an `ncu --set full` report (run HERE, no GPU needed):
   python scripts/ncu_table.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_x.txt "<note>" """
import csv, io, subprocess, sys

rep, out_txt = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time_us'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__registers_per_thread', 'regs'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_active_%'),
        ('dram__bytes_read.sum', 'dram_read_MB'), ('dram__bytes_write.sum', 'dram_write_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_%'),
        ('smsp__inst_executed.sum', 'warp_insts'), ('sm__cycles_elapsed.avg', 'sm_cycles')]
lines = ['# ' + note, '| ' + ' | '.join(n for _, n in want) + ' |', '|' + '---|' * len(want)]
for r in data:
    vals = []
    for h, n in want:
        v = r[hdr.index(h)] if h in hdr else ''
        if h == 'Kernel Name':
            v = v.split('(')[0].replace('void ', '').replace('se3tn::', '').replace('<unnamed>::', '')
        else:
            try: v = '%.2f' % float(v)
            except ValueError: pass
        vals.append(v)
    lines.append('| ' + ' | '.join(vals) + ' |')
open(out_txt, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
