mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 2 -c 1 -o gpurun_out/prof_p8 -f python bench.py --steps 1 --warmup 3 --instances 8 --no-e2e --no-cpu > gpurun_out/ncu_p8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 2 -c 1 -o gpurun_out/prof_c32 -f python bench.py --steps 1 --warmup 3 --instances 8 --no-e2e --no-cpu --strategy count32 > gpurun_out/ncu_c32.log 2>&1
