python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short 2>&1 | grep -v Warning | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --no-cpu-baseline --train-steps 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -3 gpurun_out/bench_n2.err
python -c "
import json; b=json.load(open('gpurun_out/bench_n2.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['e2e']['search_knn_value'], {k:round(v['ms_per_step'],2) for k,v in b['kernel_shares'].items()}); print(b['train_step']['ms_per_step'], b['train_step']['pairs_per_s'])"
