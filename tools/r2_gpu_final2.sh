# after the rgb_bulk_pairs_kernel change: tests, smoke, racecheck of the new kernel, ncu traffic per bench workload
# (recorded before the bench so that its line carries roofline.traffic), bench, the odd sizes, the pairs sweep
python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -1 gpurun_out/r2_final_smoke.log
timeout 200 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck_pairs.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_row_pairs and (0-0 or 0-2 or 1-3)" 2>&1 | tail -2
grep -c "RACECHECK SUMMARY: 0 hazards" gpurun_out/r2_sanitizer_racecheck_pairs.log
for wl in nv12_to_i420_1080p_x300_pitch2048:bulk_planes nv12_to_i420_4k_x64_pitch4096:bulk_planes i420_to_nv12_1080p_x300_pitch2048:bulk_planes nv12_to_rgb24_4k_x64_pitch4096:rgb nv12_to_i420_rgb24_4k_x64_pitch4096:rgb nv12_to_argb32_4k_x64_pitch4096:rgb; do
  w="${wl%%:*}"; k="${wl##*:}"
  timeout 120 ncu --set full --clock-control none -k regex:$k -s 4 -c 3 -f -o gpurun_out/r2_traffic_$w python bench.py --workload $w --steps 5 --warmup 3 --no-extras > /dev/null 2>&1
  ncu -i gpurun_out/r2_traffic_$w.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$w.csv 2>/dev/null
  rm -f gpurun_out/r2_traffic_$w.ncu-rep
  python tools/record_traffic.py $w gpurun_out/r2_ncu_full_$w.csv > /dev/null 2>&1
done
ls gpurun_out/r2_ncu_full_*x*.csv | wc -l
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python tools/odd_sizes.py rgb fused argb > gpurun_out/r2_final_odd_sizes_rgb.txt 2>&1
python tools/rgb_bulk_sweep.py > gpurun_out/r2_rgb_pairs_final.txt 2>&1
tail -8 gpurun_out/r2_rgb_pairs_final.txt
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_final_bench.json') if l.startswith('{')][0])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline'].get('frac'), d['roofline'].get('traffic'), d['e2e']['value'])
PY
