set -x
cd $GRAFT_REPO_ROOT
for w in replay routes15 crossroute map; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $w --steps 20 --warmup 3 > gpurun_out/r2_bench_${w}_8gpu.json 2> gpurun_out/r2_bench_${w}_8gpu.err
  echo "rc=$? $w"; tail -c 400 gpurun_out/r2_bench_${w}_8gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --impl reference --workload crossroute --steps 4 --warmup 1 > gpurun_out/r2_bench_crossroute_8gpu_ref.json 2>/dev/null
