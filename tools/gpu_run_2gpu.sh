#!/bin/bash
# 2-GPU call: gemm_dynamic (cluster launch control) next to real NCCL collectives.
#   NCCL parity (sharded == unsharded) with gemm_dynamic = 1; C2 at 2 GPUs, stage 3: static grids with the 8-SM margin vs
#   dynamic grids (interleaved, two rounds), plus dynamic with 16 communication CTAs.
set -u
mkdir -p gpurun_out
run() {  # name, env assignments..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 2 --steps 6 --warmup 3 "$@" > gpurun_out/h_$name.json 2> gpurun_out/h_$name.err
  echo "$name rc=$?"
}
# one-GPU checks of what changed since call 85 (RMSNorm kernels, split-K gate gradient, get_option)
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_moe.py -m gpu -q -x -k "rmsnorm or split_k or moe or layernorm" > gpurun_out/h_kernel_tests.log 2>&1
echo "kernel tests rc=$?"; tail -3 gpurun_out/h_kernel_tests.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention or empty or schedules" > gpurun_out/h_attn_tests.log 2>&1
echo "attention tests (head-fastest CTA order) rc=$?"; tail -3 gpurun_out/h_attn_tests.log
rm -f gpurun_out/h_probe.jsonl
for c in elementwise_bench_c2 attn_order_c2 attn_order_c5 attn_order_c4 overlap_wgrad_elementwise; do
  timeout 300 python tools/gpu_probe.py --only $c --out gpurun_out/h_probe.jsonl > gpurun_out/h_probe.log 2>&1
done
echo "probe rc=$?"; cut -c1-1700 gpurun_out/h_probe.jsonl
DOLO_OPTIONS=gemm_dynamic=1 timeout 900 python -m pytest tests/test_nccl_parity.py -m gpu -q > gpurun_out/h_nccl_parity_dynamic.log 2>&1
echo "nccl parity (dynamic) rc=$?"; tail -3 gpurun_out/h_nccl_parity_dynamic.log
run c2_static_a DOLO_OPTIONS=gemm_dynamic=0 -- --fsdp-mode reshard
run c2_dynamic_a DOLO_OPTIONS=gemm_dynamic=1 -- --fsdp-mode reshard
run c2_static_b DOLO_OPTIONS=gemm_dynamic=0 -- --fsdp-mode reshard
run c2_dynamic_b DOLO_OPTIONS=gemm_dynamic=1 -- --fsdp-mode reshard
run c2_dynamic_ctas16 DOLO_OPTIONS=gemm_dynamic=1 DOLO_COMM_CTAS=16 -- --fsdp-mode reshard
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/h_c*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", "per-gpu", round(d["tokens_per_sec_per_gpu"]),
              "gemm", round(d["roofline"]["achieved"]), d["clocks"]["sm_mhz"], d["loss"])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
