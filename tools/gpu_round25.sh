#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== pretrain entry point with the Megatron data feed + checkpoint/resume ==="
python - <<'PY'
import os, yaml
d = yaml.safe_load(open("configs/c1_tiny.yml"))
d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron", class_args=dict(
    data_path=["tests/golden/data_feed/corpus_a"], split="100,0,0", sequence_length=128, seed=7))]
d["training_parameters"].update(num_training_steps=6, micro_batch_size=2, gradient_accumulation_steps=1)
d["model_args"]["pretrained_config"]["vocab_size"] = 5120
d["save_args"] = dict(save_path="/tmp/ckpt_run", save_interval=3)
d["logging_args"] = dict(log_interval=1)
yaml.safe_dump(d, open("/tmp/run_a.yml", "w"))
d["load_args"] = dict(load_path="/tmp/ckpt_run", iteration=3)
yaml.safe_dump(d, open("/tmp/run_b.yml", "w"))
PY
timeout 300 python -m dolomite_engine_b200.pretrain --config /tmp/run_a.yml 2>&1 | grep -E "^step|Error|error" | tail -8
echo "--- resumed from global_step3 ---"
timeout 300 python -m dolomite_engine_b200.pretrain --config /tmp/run_b.yml 2>&1 | grep -E "^step|Error|error" | tail -5
ls /tmp/ckpt_run /tmp/ckpt_run/global_step6
