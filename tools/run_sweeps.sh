for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) tools/sweep.py > gpurun_out/r2l_sweep_n$n.jsonl 2> gpurun_out/r2l_sweep_n$n.err
tail -1 gpurun_out/r2l_sweep_n$n.err
done
python -c "
import json
for n in (8,4,2):
    for l in open('gpurun_out/r2l_sweep_n%d.jsonl'%n):
        d=json.loads(l); print(n, d['rods'], round(d['ms_per_step'],3), int(d['frames_per_s']), round(d['k_cost_frac'],3), round(d['k_cost_fp64_pipe_busy_est'],3))
"
