#!/bin/bash
# The commands of the CURRENT gpurun call (rewritten per call; git history keeps the earlier ones).
# Call D (2 GPUs): sharded == unsharded over NCCL (resident and stage-3 modes), the attention kernels changed since call C
# (backward with TMEM loads pipelined one half step ahead; serial backward with TMA tile reduce-add for dQ), 2-GPU bench lines.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > gpurun_out/d_attn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/d_attn_tests.log
rm -f gpurun_out/d_probe.jsonl
for c in attn_bench_c2 attn_bench_hd128; do timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/d_probe.jsonl > /dev/null 2>&1; done
timeout 900 python -m pytest tests/test_nccl_parity.py -m gpu -q -x > gpurun_out/d_nccl_parity.log 2>&1
echo "rc=$?" >> gpurun_out/d_nccl_parity.log
for mode in reshard resident; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --fsdp-mode $mode > gpurun_out/d_bench_c2_2gpu_$mode.json 2> gpurun_out/d_bench_c2_2gpu_$mode.err
done
tail -c 700 gpurun_out/d_attn_tests.log
python - <<'PY'
import json
for l in open("gpurun_out/d_probe.jsonl"):
    d = json.loads(l); print({k: d[k] for k in ("case", "ok", "fwd_ms", "bwd_ms", "fwd_tflops_causal", "bwd_tflops_causal") if k in d}, d.get("vs_flash_dq"), d.get("error"))
PY
tail -c 1500 gpurun_out/d_nccl_parity.log
for mode in reshard resident; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/d_bench_c2_2gpu_$mode.json")); print("$mode", d["value"], d["ms_per_step"], d["peak_hbm_gb"], d["clocks"], d["loss"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/d_bench_c2_2gpu_$mode.err").read()[-1200:])
PY
done
