#!/bin/bash
# round 2, GPU call AE: ring sizes 32 / 48 / 64 MiB with the new copy schedule; role 0 by the first instead of the last CTAs; e2e with tangents
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product first first128; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  for ring in 32 48 64; do
    $B --ring-mib $ring > gpurun_out/bench_ae_${ab}_ring$ring.json 2>/dev/null
  done
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ae_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_ae_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
for ring in 48; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,smsp__inst_executed.sum --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/traffic_ae_ring$ring.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded --ring-mib $ring > /dev/null 2>&1
grep k_frame gpurun_out/traffic_ae_ring$ring.csv | awk -F'","' '{print $13, $15}'
done
timeout 300 python bench.py --no-cpu --no-others --no-sharded --steps 5 --warmup 3 > gpurun_out/bench_ae_e2e.json 2>gpurun_out/bench_ae_e2e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ae_e2e.json').read().strip().splitlines()[-1])
print(json.dumps(d['e2e'])[:900])
PY
