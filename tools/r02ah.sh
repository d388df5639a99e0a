for c in 8 4; do
NCCL_MAX_CTAS=$c python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$c bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02ah_bench_n8_ctas$c.json 2> gpurun_out/r02ah_bench_n8_ctas$c.err
python -c "
import json
d=json.loads(open('gpurun_out/r02ah_bench_n8_ctas$c.json').read().strip().splitlines()[-1])
print('NCCL_MAX_CTAS=$c', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_last'])
"
done
