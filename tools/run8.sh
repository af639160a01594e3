# 8-GPU pass on one box (gpurun --gpus 8): bench at 8 and 4 ranks, sharded bit-identity check
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3i_bench_n8.json 2> gpurun_out/r3i_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3i_bench_n4.json 2> gpurun_out/r3i_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3i_bench_n2.json 2> gpurun_out/r3i_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 tools/sharded_check.py > gpurun_out/r3i_sharded_check_n8.log 2>&1
tail -2 gpurun_out/r3i_bench_n8.err; tail -1 gpurun_out/r3i_sharded_check_n8.log
python -c "
import json
for n in (8,4,2):
    d=json.loads(open('gpurun_out/r3i_bench_n%d.json'%n).read().strip().splitlines()[-1]); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['full_rows']['value'], d['host_enqueue_ms_per_step'], d['cuda_mallocs_in_timed_region'], d['sustained']['value'], d.get('config4'))
"
