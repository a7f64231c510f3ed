# whole-box record with the fixed-window link probe (every GPU measured over the same 400 ms)
for n in 1 2 4 8; do tools/jm_link --gpus $n > gpurun_out/r2d_link_n$n.json; done
for m in e2e d2h; do for a in static dynamic; do tools/jm_streams --gpus 8 --streams 128 --frames 75 --batch 25 --mode $m --assign $a > gpurun_out/r2d_streams_${m}_n8_${a}.json; done; done
for n in 2 4; do for m in e2e d2h; do tools/jm_streams --gpus $n --streams 32 --frames 300 --batch 30 --mode $m > gpurun_out/r2d_streams_${m}_n$n.json; done; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras 2>/dev/null | grep '^{' > gpurun_out/r2d_bench_n8.json
for n in 1 2 4 8; do python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], {k:(v['box_h2d_gbs'],v['box_d2h_gbs']) for k,v in d.items() if isinstance(v,dict) and 'wc' not in k}); print('   per GPU bidir', d['bidirectional']['per_gpu_h2d_gbs'], d['bidirectional']['per_gpu_d2h_gbs'], 'd2h only', d['d2h_only']['per_gpu_d2h_gbs'])" gpurun_out/r2d_link_n$n.json; done
for f in gpurun_out/r2d_streams_*.json; do python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['frames_per_s'], [g['streams'] for g in d['per_gpu']], [round(g['frames_per_s']) for g in d['per_gpu']])" $f; done
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_n8.json')); print(d['value'], d['e2e'], d['e2e_device_resident_input'])"
