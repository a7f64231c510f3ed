for m in e2e d2h; do for a in static dynamic; do
  tools/jm_streams --gpus 8 --streams 32 --frames 300 --batch 30 --mode $m --assign $a > gpurun_out/r2_streams_${m}_n8_${a}.json
  tools/jm_streams --gpus 8 --streams 128 --frames 75 --batch 25 --mode $m --assign $a > gpurun_out/r2_streams_${m}_n8_${a}_128x75.json
done; done
tools/jm_link --gpus 8 > gpurun_out/r2_link_n8_c.json
for f in gpurun_out/r2_streams_*_n8_static*.json gpurun_out/r2_streams_*_n8_dynamic*.json; do python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['frames_per_s'], [g['streams'] for g in d['per_gpu']], [round(g['frames_per_s']) for g in d['per_gpu']])" $f; done
python -c "
import json,sys; d=json.load(open('gpurun_out/r2_link_n8_c.json')); print({k:(v['box_h2d_gbs'],v['box_d2h_gbs'],v['per_gpu_d2h_gbs']) for k,v in d.items() if isinstance(v,dict) and 'wc' not in k})"
