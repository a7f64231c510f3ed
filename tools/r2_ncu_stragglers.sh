for t in "pack 854x480:pack854" "=rgb 1366x768 p1536:rgb1366" "=rgb2nv12 1366x768 p1536:rgb2nv12_1366" "=fused 1366x768 p1536:fused1366"; do
  name="${t%%:*}"; tag="${t##*:}"
  ncu --set full --clock-control none --import-source on -s 5 -c 1 -f -o gpurun_out/r2_ncu_$tag python tools/odd_sizes.py "$name" > gpurun_out/r2_ncu_$tag.log 2>&1
  ncu -i gpurun_out/r2_ncu_$tag.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_$tag.csv 2>/dev/null
  ncu -i gpurun_out/r2_ncu_$tag.ncu-rep --page source --csv > gpurun_out/r2_ncu_source_$tag.csv 2>/dev/null
done
ls -la gpurun_out/r2_ncu_*
