#!/bin/bash
# 8-GPU box, short scan with the final library: N = 1 and N = 8, weak and strong (256 spp per step)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scan_weak_1.json 2> gpurun_out/scan_weak_1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scan_weak_8.json 2> gpurun_out/scan_weak_8.err
timeout 400 $TR bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong --spp-per-step 256 > gpurun_out/scan_strong_8.json 2> gpurun_out/scan_strong_8.err
python - <<PY | tee gpurun_out/scan_summary_short.txt
import json
base = None
for kind, n in (("weak", 1), ("weak", 8), ("strong", 8)):
    try:
        j = json.loads(open(f"gpurun_out/scan_{kind}_{n}.json").read().strip().split("\n")[-1])
        if n == 1: base = (j["value"], j["e2e"]["value"])
        print(f"{kind:6s} N={n}: value {j['value']:9.1f} Mrays/s ({j['value'] / base[0] / n:5.3f} of ideal)  e2e {j['e2e']['value']:9.1f} ({j['e2e']['value'] / base[1] / n:5.3f})  "
              f"ms/step {j['ms_per_step']:7.2f}  e2e ms/step {j['e2e']['ms_per_step']:7.2f}  spp/step {j['run']['spp_per_step']} lanes {j['run']['lanes']}")
    except Exception as ex:
        print(kind, n, "FAILED", ex)
PY
