# round-2 evidence batch: TMA A/B, bench both arms, ncu launch list + full capture of the bench kernel
set -x
tools/tma_ab --verify > gpurun_out/r2_tma_ab.csv 2> gpurun_out/r2_tma_ab.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_1080p.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r2_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bulk_planes -s 4 -c 3 -o gpurun_out/r2_ncu_i420_1080p -f python bench.py --steps 5 --warmup 3 --no-extras >> gpurun_out/r2_under_ncu.log 2>&1
ncu -i gpurun_out/r2_ncu_i420_1080p.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_i420_1080p.csv 2>/dev/null
