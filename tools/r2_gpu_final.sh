# final round-2 evidence on one B200: tests, sanitizer, bench (both arms), ncu traffic per workload, odd sizes, front-end fps
python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_gpu_api.py tests/test_gpu_nvdec_frontend.py -m gpu -q -x -k "display_delay or output_destinations or wait_event or batch_drain or format_change or host_side_pointer or nvenc_short" 2>&1 | tail -2
compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pack_batches or pad_zero or random_widths" 2>&1 | tail -2
grep -c "ERROR SUMMARY: 0 errors\|RACECHECK SUMMARY: 0 hazards" gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
python bench.py --impl reference > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python tools/odd_sizes.py > gpurun_out/r2_final_odd_sizes.txt 2>&1
JMC_PAD_ZERO=1 python tools/odd_sizes.py pack rgb2nv12 > gpurun_out/r2_final_odd_sizes_pad_zero.txt 2>&1
python tools/nvdec_frontend_fps.py > gpurun_out/r2_nvdec_frontend_fps.txt 2>&1
for wl in nv12_to_i420_1080p_x300_pitch2048:bulk_planes nv12_to_i420_4k_x64_pitch4096:bulk_planes i420_to_nv12_1080p_x300_pitch2048:bulk_planes nv12_to_rgb24_4k_x64_pitch4096:rgb nv12_to_i420_rgb24_4k_x64_pitch4096:rgb nv12_to_argb32_4k_x64_pitch4096:rgb; do
  w="${wl%%:*}"; k="${wl##*:}"
  ncu --set full --clock-control none -k regex:$k -s 4 -c 3 -f -o gpurun_out/r2_traffic_$w python bench.py --workload $w --steps 5 --warmup 3 --no-extras > /dev/null 2>&1
  ncu -i gpurun_out/r2_traffic_$w.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$w.csv 2>/dev/null
  rm -f gpurun_out/r2_traffic_$w.ncu-rep
done
ls gpurun_out/r2_ncu_full_*x*.csv
