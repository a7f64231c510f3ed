# whole-box evidence (8 GPUs): host link ceiling, config 5 (32 streams x 300 frames, stream % nGPU), bench at N=8
set -x
for n in 1 2 4 8; do tools/jm_link --gpus $n > gpurun_out/r2_link_n$n.json 2>> gpurun_out/r2_n8.err; done
for n in 1 2 4 8; do for m in e2e d2h device; do tools/jm_streams --gpus $n --streams 32 --frames 300 --batch 30 --mode $m > gpurun_out/r2_streams_${m}_n$n.json 2>> gpurun_out/r2_n8.err; done; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2>> gpurun_out/r2_n8.err
python bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8_reference_arm.json 2>> gpurun_out/r2_n8.err
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
