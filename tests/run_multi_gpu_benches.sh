# usage: [WORKLOADS="..."] [EXTRA="--graphs"] bash tests/run_multi_gpu_benches.sh N   -- one bench line per workload on N GPUs of this box (torchrun, NCCL) under gpurun_out/
N=$1
WORKLOADS=${WORKLOADS:-Taobao-10 Taobao-10-batch Taobao-30 Taobao-20-star Amazon-13-mmoe-sharded Amazon-13-ple-sharded Amazon-13-sharded}
for w in $WORKLOADS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-micro --no-cpu $EXTRA > gpurun_out/r2_bench_${w}_n${N}.json 2> gpurun_out/r2_bench_${w}_n${N}.err
  python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']), d['ms_per_step'], d.get('phase_us_per_minibatch'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
" gpurun_out/r2_bench_${w}_n${N}.json
done
