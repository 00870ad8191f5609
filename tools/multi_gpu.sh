#!/bin/bash
# N-GPU check: bench under torchrun + 1-vs-N image equality through render.py
N=${1:-2}
mkdir -p gpurun_out outputs
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 600 gpurun_out/bench_${N}gpu.err
timeout 120 python render.py --scene csphere --name balls-mono.xml --iter_num 15 --no_gui --save_hdr --img_name one --no_watermark > /dev/null 2>&1
timeout 200 $TR render.py --scene csphere --name balls-mono.xml --iter_num 15 --no_gui --save_hdr --img_name two --no_watermark > gpurun_out/render_${N}gpu.log 2>&1
python - <<PY
import numpy as np, json
a = np.load('outputs/one-balls-mono-pt.npy'); b = np.load('outputs/two-balls-mono-pt.npy')
print('1-vs-$N GPU image: rel L2', float(np.linalg.norm(a-b)/np.linalg.norm(a)), 'max abs', float(np.abs(a-b).max()), 'shape', a.shape)
for f in ('gpurun_out/bench_1gpu.json', 'gpurun_out/bench_${N}gpu.json'):
    try:
        j = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, 'value', round(j['value'],1), 'Mrays/s  spp/s', round(j['spp_per_s'],1), 'ms/step', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],1), j['config']['parallelism'])
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
cp outputs/one-balls-mono-pt.png gpurun_out/ 2>/dev/null
