# whole-box follow-up: does the host footprint of the DMA buffers set the ceiling?
tools/jm_link --gpus 8 --mb 1024 --copies 2 > gpurun_out/r2_link_n8_1024mb.json
tools/jm_link --gpus 8 --mb 64 --copies 24 > gpurun_out/r2_link_n8_64mb.json
tools/jm_link --gpus 8 --mb 256 --copies 6 > gpurun_out/r2_link_n8_256mb_b.json
for m in e2e d2h; do
  tools/jm_streams --gpus 8 --streams 320 --frames 30 --batch 30 --mode $m > gpurun_out/r2_streams_${m}_n8_small_footprint.json
  tools/jm_streams --gpus 8 --streams 32 --frames 300 --batch 100 --mode $m > gpurun_out/r2_streams_${m}_n8_batch100.json
  tools/jm_streams --gpus 8 --streams 32 --frames 300 --batch 10 --mode $m > gpurun_out/r2_streams_${m}_n8_batch10.json
done
for f in gpurun_out/r2_link_n8_1024mb.json gpurun_out/r2_link_n8_64mb.json gpurun_out/r2_link_n8_256mb_b.json; do python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], {k:(v['box_h2d_gbs'],v['box_d2h_gbs']) for k,v in d.items() if isinstance(v,dict)})" $f; done
for f in gpurun_out/r2_streams_*_n8_*.json; do python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['frames_per_s'], [round(g['frames_per_s']) for g in d['per_gpu']])" $f; done
