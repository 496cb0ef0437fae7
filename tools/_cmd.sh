for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/scale_err_$N.log | tail -1 > gpurun_out/scale_$N.json
python -c "
import sys,json; d=json.loads(open('gpurun_out/scale_$N.json').read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
