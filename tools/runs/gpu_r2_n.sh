#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/big_probe.py > gpurun_out/big_probe.json 2> gpurun_out/big_probe.err; cat gpurun_out/big_probe.json; tail -3 gpurun_out/big_probe.err
timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/bench_n.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1]); print('crowd: ms/step %.4f'%d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
timeout 200 python bench.py --no-e2e --no-cpu --no-others --no-sharded --instances 8 --steps 50 > gpurun_out/bench_n_8inst.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_n_8inst.json').read().strip().splitlines()[-1]); print('8 instances: ms/step %.4f'%d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
timeout 120 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum --clock-control none -k regex:"k_walk_uniform|k_frame|k_untile" -c 12 --csv --log-file gpurun_out/launches_n.csv python tools/big_probe.py > /dev/null 2>&1
awk -F'","' '{print $5, $13, $15}' gpurun_out/launches_n.csv | grep -E "k_walk|k_frame|k_untile" | sort | uniq -c | sort -k2 | head -60
