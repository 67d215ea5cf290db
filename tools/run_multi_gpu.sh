# 8-GPU lines of the bench workloads (one box): bash tools/run_multi_gpu.sh [workloads...]
set -x
cd $GRAFT_REPO_ROOT
W=${@:-replay routes15 crossroute map}
for w in $W; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $w --steps 20 --warmup 3 > gpurun_out/r2_bench_${w}_8gpu.json 2> gpurun_out/r2_bench_${w}_8gpu.err
  echo "rc=$? $w"; tail -c 400 gpurun_out/r2_bench_${w}_8gpu.json
done
