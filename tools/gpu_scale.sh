# sharded bench at N = 2, 4 and 8 (gpurun --gpus 8)
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value %.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], 'chains %.3e'%d['config']['independent_chains_events_per_s'], 'switches', d['config']['switches_per_step'])"
done
