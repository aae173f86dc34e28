import csv, sys
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex.sum', 'lts__t_sectors.sum', 'sm__cycles_elapsed.avg.per_second', 'gpc__cycles_elapsed.avg.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print({k.split('.')[0][-28:]: (r[h.index(k)], rows[1][h.index(k)]) for k in keys if k in h})
