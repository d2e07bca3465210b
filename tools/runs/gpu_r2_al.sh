#!/bin/bash
# round 2, GPU call AL: one range per item (short CTAs: the copy-out of an instance follows its reds sooner) -- time and DRAM bytes
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product r1 r1_cop512 r1_cop128 r2; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_al_${ab}.json 2>gpurun_out/bench_al_${ab}.err
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/traffic_al_$ab.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > /dev/null 2>&1
  echo $ab $(grep k_frame gpurun_out/traffic_al_$ab.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' ')
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_al_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_al_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
