# usage: bash tools/final_scale.sh N   (inside gpurun --gpus N)
N=$1
for pol in auto regions; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --partition $pol --no-cpu-baseline > gpurun_out/f_n${N}_$pol.json 2> gpurun_out/f_n${N}_$pol.err
  tail -c 200 gpurun_out/f_n${N}_$pol.err
  python -c "
import json
d=json.loads(open('gpurun_out/f_n${N}_$pol.json').read().strip().splitlines()[-1])
print('$N $pol', d['ms_per_step'], d['e2e']['ms_per_step'], d['checks']['final_loss_sum'], d['checks'].get('merged_calls_on_rank0'))
"
done
