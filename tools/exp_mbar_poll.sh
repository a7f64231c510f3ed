echo "=== single poller + suspend hint (default build)"; python tools/odd_sizes.py > gpurun_out/r2_odd_sizes_poll1.txt 2>&1; cat gpurun_out/r2_odd_sizes_poll1.txt
echo "=== every thread polls (round-1 behaviour, -DJMC_MBAR_POLL=0)"; JMCODEC_B200_LIB=build_variants/libjmc_DJMC_MBAR_POLL_0.so python tools/odd_sizes.py > gpurun_out/r2_odd_sizes_poll0.txt 2>&1; cat gpurun_out/r2_odd_sizes_poll0.txt
