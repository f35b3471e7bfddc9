"""Key metrics of an `ncu --set full` report as a markdown table (reads `ncu -i REP --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM bytes'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
]
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
print('| metric | ' + ' | '.join(f'launch {i}' for i in range(len(data))) + ' |')
print('|---|' + '---|' * len(data))
name_i = hdr.index('Kernel Name')
print('| kernel | ' + ' | '.join(r[name_i].split('(')[0].replace('void ', '') for r in data) + ' |')
for k, label in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f'| {label} (`{k}`) | ' + ' | '.join(f'{r[i]} {units[i]}' for r in data) + ' |')
