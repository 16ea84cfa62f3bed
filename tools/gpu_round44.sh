#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu ==="
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== pretrain: Megatron feed + evaluation + checkpoint ==="
python - <<'PY'
import yaml
d = yaml.safe_load(open("configs/c1_tiny.yml"))
d["datasets"] = [dict(class_name="MegatronDataset", data_name="Megatron", class_args=dict(
    data_path=["tests/golden/data_feed/corpus_a"], split="80,20,0", sequence_length=64, seed=7, eval_steps=2))]
d["training_parameters"].update(num_training_steps=4, micro_batch_size=2, gradient_accumulation_steps=1, eval_interval=2,
                                eval_during_training=True)
d["model_args"]["pretrained_config"]["vocab_size"] = 5120
d["save_args"] = dict(save_path="/tmp/ckpt_eval", save_interval=100)
d["logging_args"] = dict(log_interval=1)
yaml.safe_dump(d, open("/tmp/run_eval.yml", "w"))
PY
timeout 300 python -m dolomite_engine_b200.pretrain --config /tmp/run_eval.yml 2>&1 | grep -E "^step|Error|error" | tail -10
echo "=== memcheck (kernel tests, small shapes) ==="
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --launch-timeout 0 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "not full_size" 2>&1 | tail -6 | tee gpurun_out/memcheck_kernels.txt
