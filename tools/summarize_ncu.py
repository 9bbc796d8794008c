"""ncu --set full report (.ncu-rep) -> the handful of raw metrics the roofline argument uses (text).
usage: python tools/summarize_ncu.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

KEYS = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_cycles_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit', 'smsp__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'smsp__warp_issue_stalled', 'derived__smsp__sass_thread_inst_executed_op_ffma')
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(path, 'no data'); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f'== {path}: {d.get("Kernel Name", "?")[:100]}  grid {d.get("Grid Size")} block {d.get("Block Size")}')
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k) for k in KEYS):
                print(f'   {h:80s} {v:>18s} {u}')
        try:
            t = float(d['gpu__time_duration.sum'].replace(',', ''))
            tu = units[hdr.index('gpu__time_duration.sum')]
            rd = float(d['dram__bytes_read.sum'].replace(',', '')); ru = units[hdr.index('dram__bytes_read.sum')]
            wr = float(d['dram__bytes_write.sum'].replace(',', '')); wu = units[hdr.index('dram__bytes_write.sum')]
            mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            tm = {'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'nsecond': 1e-9, 'ms': 1e-3, 'msecond': 1e-3}[tu]
            B = rd * mul[ru] + wr * mul[wu]
            print(f'   traffic (dram read+write) {B / 1e6:.2f} MB in {t * tm * 1e6:.1f} us -> {B / (t * tm) / 1e9:.0f} GB/s under ncu')
        except Exception as e:
            print('   (traffic summary failed:', e, ')')
