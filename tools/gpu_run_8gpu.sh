#!/bin/bash
# 8-GPU call: NCCL parity incl. HSDP 2 x 2, then the three workloads at 8 GPUs:
#   C2 stage 3 (reshard, the default) with 8 and 4 communication CTAs, C2 resident mode, C4, C5 (checkpoint every 2).
set -u
mkdir -p gpurun_out
run() {  # name, env assignments..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus 8 --steps 4 --warmup 3 "$@" > gpurun_out/g_$name.json 2> gpurun_out/g_$name.err
  echo "$name rc=$?"
}
timeout 900 python -m pytest tests/test_nccl_parity.py -m gpu -q > gpurun_out/g_nccl_parity.log 2>&1
echo "nccl parity rc=$?"
run c2_reshard_ctas8 DOLO_COMM_CTAS=8 -- --fsdp-mode reshard
run c2_reshard_ctas4 DOLO_COMM_CTAS=4 -- --fsdp-mode reshard
run c2_resident_ctas8 DOLO_COMM_CTAS=8 -- --fsdp-mode resident
run c4_reshard DOLO_COMM_CTAS=8 -- --config c4
run c5_reshard DOLO_COMM_CTAS=8 -- --config c5
tail -4 gpurun_out/g_nccl_parity.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/g_c*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", "per-gpu", round(d["tokens_per_sec_per_gpu"]), "peak", round(d["peak_hbm_gb"], 1),
              "gemm", round(d["roofline"]["achieved"]), d["clocks"]["sm_mhz"], d["loss"])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
