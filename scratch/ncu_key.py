"""key metrics of an ncu --set full capture: python scratch/ncu_key.py <rep> [more reps]"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit', 'launch__registers_per_thread', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'smem', 'shared']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print('==', rep, d.get('Kernel Name', '')[:70], d.get('Grid Size'), d.get('Block Size'))
        for k in h:
            if any(x.strip() in k for x in KEYS) and '.max' not in k.split('.')[-2:][0] and 'min' not in k and '.sum.pct' not in k:
                print(f"   {k:95s} {u[h.index(k)]:12s} {d[k]}")
