"""One line per profiled launch from an `ncu --page raw --csv` export of an `ncu --set full` capture:
time, tensor-pipe utilisation, DRAM read / write, DRAM throughput %, L2 hit rate, SM clock.

    python tools/ncu_table.py gpurun_out/f_full_raw.csv > profiles/r02_ncu_full_step_kernels.txt"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, key, scale=1.0):
    try:
        return float(r[ix[key]].replace(',', '')) * scale
    except (KeyError, ValueError):
        return float('nan')


def mb(r, key):
    u = units[ix[key]]
    return val(r, key, {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0))


def us(r, key):
    u = units[ix[key]]
    return val(r, key, {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0))


print('# ncu --set full --clock-control none, one launch each from a steady-state ips() step of `python bench.py`')
print('# units: time us, dram read/write Mbyte; tensor = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')
print('%-48s %6s %9s %8s %9s %9s %6s %7s %5s' % ('kernel', 'grid', 'time', 'tensor%', 'dram rd', 'dram wr', 'dram%', 'L2hit%', 'GHz'))
for r in rows[2:]:
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')
    grid = re.sub(r'[^0-9,]', '', r[ix['Grid Size']]).split(',')[0]
    print('%-48s %6s %9.1f %8.1f %9.1f %9.1f %6.1f %7.1f %5.2f' % (
        name[:48], grid, us(r, 'gpu__time_duration.sum'), val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        mb(r, 'dram__bytes_read.sum'), mb(r, 'dram__bytes_write.sum'), val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        val(r, 'lts__t_sector_hit_rate.pct'), val(r, 'smsp__cycles_elapsed.avg.per_second')))
