#!/bin/bash
# N-GPU session: bash tools/multi_gpu.sh N
#   the multi-device handle (one process, tests/test_gpu_multi.py), bench.py under torchrun (weak and strong scaling, against N = 1 on the
#   same box), 1-vs-N image equality through render.py (--gpus N in one process, and under torchrun)
N=${1:-2}
mkdir -p gpurun_out outputs
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 300 2>&1 | tail -5 | tee gpurun_out/pytest_multi_${N}gpu.log
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu_weak.json 2> gpurun_out/bench_${N}gpu_weak.err
tail -c 400 gpurun_out/bench_${N}gpu_weak.err
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --spp-per-step 256 > gpurun_out/bench_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --spp-per-step 256 > gpurun_out/bench_1gpu_256spp.json 2> gpurun_out/bench_1gpu_256spp.err
timeout 200 python render.py --scene csphere --name balls-mono.xml --type pt --iter_num 15 --no_gui --save_hdr --img_name one --no_watermark > /dev/null 2>&1
timeout 300 python render.py --scene csphere --name balls-mono.xml --type pt --iter_num 15 --no_gui --save_hdr --img_name grp --no_watermark --gpus $N > gpurun_out/render_group_${N}gpu.log 2>&1
timeout 300 $TR render.py --scene csphere --name balls-mono.xml --type pt --iter_num 15 --no_gui --save_hdr --img_name two --no_watermark > gpurun_out/render_${N}gpu.log 2>&1
python - <<PY | tee gpurun_out/multi_${N}gpu_summary.txt
import numpy as np, json
a = np.load('outputs/one-balls-mono-pt.npy')
for tag, what in (('grp', 'one process, $N devices behind one handle'), ('two', 'torchrun, $N ranks + NCCL reduce')):
    try:
        b = np.load('outputs/%s-balls-mono-pt.npy' % tag)
        print('1-vs-$N GPU image (%s): rel L2' % what, float(np.linalg.norm(a-b)/np.linalg.norm(a)), 'max abs', float(np.abs(a-b).max()), 'shape', a.shape)
    except Exception as ex:
        print(tag, 'FAILED', ex)
base = {}
for f in ('gpurun_out/bench_1gpu.json', 'gpurun_out/bench_${N}gpu_weak.json', 'gpurun_out/bench_1gpu_256spp.json', 'gpurun_out/bench_${N}gpu_strong.json'):
    try:
        j = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, j['scaling'], 'n_gpus', j['n_gpus'], 'value', round(j['value'],1), 'Mrays/s  ms/step', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],1), 'e2e ms/step', round(j['e2e']['ms_per_step'],2), 'spp/step', j['run']['spp_per_step'])
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
