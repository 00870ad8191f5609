#!/bin/bash
# 8-GPU box: weak and strong scaling of bench.py at N = 1, 2, 4, 8 (same box, back to back), the multi-device handle on 8 devices
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/pytest_multi_8gpu.log
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scan_weak_1.json 2> gpurun_out/scan_weak_1.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --spp-per-step 256 > gpurun_out/scan_strong_1.json 2> gpurun_out/scan_strong_1.err
for N in 2 4 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scan_weak_$N.json 2> gpurun_out/scan_weak_$N.err
  timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --spp-per-step 256 > gpurun_out/scan_strong_$N.json 2> gpurun_out/scan_strong_$N.err
done
timeout 300 python render.py --scene cbox --name bunny90k.xml --type pt --iter_num 63 --no_gui --save_hdr --img_name g8 --no_watermark --gpus 8 > gpurun_out/render_group_8gpu.log 2>&1
cp outputs/g8-bunny90k-pt.metrics.json gpurun_out/ 2>/dev/null
python - <<PY | tee gpurun_out/scan_summary.txt
import json
base = {}
for kind in ("weak", "strong"):
    for n in (1, 2, 4, 8):
        try:
            j = json.loads(open(f"gpurun_out/scan_{kind}_{n}.json").read().strip().split("\n")[-1])
            if n == 1: base[kind] = (j["value"], j["e2e"]["value"])
            print(f"{kind:6s} N={n}: value {j['value']:9.1f} Mrays/s ({j['value'] / base[kind][0] / n:5.3f} of ideal)  e2e {j['e2e']['value']:9.1f} ({j['e2e']['value'] / base[kind][1] / n:5.3f})  "
                  f"ms/step {j['ms_per_step']:7.2f}  e2e ms/step {j['e2e']['ms_per_step']:7.2f}  spp/step {j['run']['spp_per_step']}")
        except Exception as ex:
            print(kind, n, "FAILED", ex)
PY
