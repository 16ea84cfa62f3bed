"""`--config <yaml>` argument tree of the reference's pretrain entry point (arguments.py:30-447, utils/pydantic.py:7-8,
utils/yaml.py:6-23) restricted to the knobs of the data-parallel training hot path.  Same key names, same
`extra="forbid"` strictness; keys that select out-of-scope subsystems are accepted only at their default values and
raise NotImplementedError otherwise, so an existing reference YAML either runs or fails loudly.

Two additions: `model_args.moe_implementation` and
`model_args.normalization_implementation` are accepted here because the reference's own
configs/testing/scattermoe.yml sets them although its ModelArgs rejects them."""

from __future__ import annotations

import re
from argparse import ArgumentParser
from typing import Any

import yaml
from pydantic import BaseModel, ConfigDict


class BaseArgs(BaseModel):
    model_config = ConfigDict(extra="forbid", protected_namespaces=())


def load_yaml(file_path: str) -> dict:
    """utils/yaml.py:6-23 -- YAML 1.1 loader fixed so that `1e-5` parses as a float"""
    loader = yaml.SafeLoader
    loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(
            """^(?:
    [-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
    |\\.[0-9_]+(?:[eE][-+][0-9]+)?
    |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\\.[0-9_]*
    |[-+]?\\.(?:inf|Inf|INF)
    |\\.(?:nan|NaN|NAN))$""",
            re.X,
        ),
        list("-+0123456789."),
    )
    with open(file_path) as f:
        return yaml.load(f, loader)


def _check_not_None(pairs) -> None:
    for obj, name in pairs:
        assert obj is not None, f"{name} cannot be None"


class RandomArgs(BaseArgs):
    seed: int = 42


class TokenizerArgs(BaseArgs):
    tokenizer_name: str | None = None
    additional_special_tokens: list[str] | None = None


class ModelArgs(BaseArgs):
    model_name: str | None = None
    pretrained_config: dict | None = None
    model_class: str = None
    trust_remote_code: bool = False
    attention_implementation: str | None = None
    use_padding_free_transformer: bool = False
    efficient_initialization: bool = False
    reset_attention_mask: bool = False
    reset_position_ids: bool = False
    moe_implementation: str | None = None
    normalization_implementation: str | None = None

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.model_class, "model_class")])
        if self.model_name is None:
            _check_not_None([(self.pretrained_config, "pretrained_config")])
        else:
            assert self.pretrained_config is None, "pretrained_config shouldn't be specified with model_name"
        assert self.model_class in ["AutoModelForCausalLM", "AutoModelForSeq2SeqLM"], f"unexpected model_class ({self.model_class})"
        if self.model_class != "AutoModelForCausalLM":
            raise NotImplementedError("only AutoModelForCausalLM is on the B200 hot path")


class TuningArgs(BaseArgs):
    tuning_method: str = None
    prompt_tuning_args: dict | None = None
    lora_args: dict | None = None

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.tuning_method, "tuning_method")])
        if self.tuning_method not in ("pretraining", "full_finetuning"):
            raise NotImplementedError(f"tuning_method={self.tuning_method}: PEFT is out of scope of the B200 hot path")


class TrainingParameters(BaseArgs):
    ignore_sampling_proportion_for_validation: bool = False
    num_training_steps: int | None = None
    gradient_accumulation_steps: int = 1
    eval_interval: int | None = None
    micro_batch_size: int = None
    eval_during_training: bool = True
    loss_mask: str = "output_only"
    gradient_clipping: float | None = 1

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.num_training_steps, "num_training_steps"), (self.micro_batch_size, "micro_batch_size")])
        if self.eval_during_training:
            _check_not_None([(self.eval_interval, "eval_interval")])


class SaveArgs(BaseArgs):
    save_path: str = None
    save_interval: int = None
    save_optimizer: bool = True

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.save_path, "save_path"), (self.save_interval, "save_interval")])


class LoadArgs(BaseArgs):
    load_path: str = None
    iteration: int | None = None
    load_optimizer: bool = True
    load_lr_scheduler: bool = True
    load_rng_state: bool = True
    load_dataloader_state: bool = True
    load_experiments_tracker_state: bool = True
    load_starting_iteration: bool = True
    resume_learning_rate: bool = True


class DatasetArgs(BaseArgs):
    class_name: str = None
    class_args: dict = {}
    data_name: str = None
    input_format: str = "__input__"
    output_format: str = "__output__"
    data_sampling_ratio: int | None = None
    max_input_tokens: int | None = None
    max_output_tokens: int | None = None

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.class_name, "dataset class_name"), (self.data_name, "data_name")])


class OptimizerArgs(BaseArgs):
    class_name: str = "TorchAdamW"
    params_group_method: str | None = None
    class_args: dict = {"lr": 1e-5, "weight_decay": 0.1, "betas": [0.9, 0.95], "eps": 1e-10}


class LRSchedulerArgs(BaseArgs):
    num_warmup_steps: int = 200
    num_constant_steps: int = 0
    num_decay_steps: int | None = None
    lr_decay_style: str = "cosine"
    lr_decay_factor: float = 0.1
    extra_lr_scheduler_args: dict = {}


class MixedPrecisionArgs(BaseArgs):
    dtype: str = "fp32"
    fp8_backend: str | None = None

    def model_post_init(self, __context: Any) -> None:
        self.dtype = {"bfloat16": "bf16", "float32": "fp32", "float16": "fp16"}.get(self.dtype, self.dtype)
        if self.fp8_backend is not None or self.dtype == "fp8":
            raise NotImplementedError("FP8 backends are out of scope of the B200 hot path (bf16 target)")


class ZeroTopologyArgs(BaseArgs):
    data_parallel_replication_world_size: int | None = None
    data_parallel_sharding_world_size: int | None = None


class DistributedArgs(BaseArgs):
    stage: int = 3
    distributed_backend: str = "torch"
    overlap_comm: bool = False
    contiguous_gradients: bool = False
    cpu_offload: bool = False
    gradient_checkpointing_method: str | None = None
    gradient_checkpointing_args: dict = {}
    zero_topology: ZeroTopologyArgs = ZeroTopologyArgs()
    zero_quantized_weights: bool = False
    zero_quantized_gradients: bool = False
    communication_dtype: str | None = None
    torch_compile: bool = False
    dispatching_dataloader: bool = False
    tensor_parallel_size: int = 1
    tensor_parallel_word_embeddings: bool = False
    sequence_parallel: bool = False
    data_parallel_size: int | None = None
    timeout_minutes: int | None = None
    fsdp_algorithm: int = 1
    # B200 extension: free gathered parameters after forward and re-gather in backward (FSDP stage-3 memory profile)
    reshard_after_forward: bool = False

    def model_post_init(self, __context: Any) -> None:
        if self.distributed_backend != "torch":
            raise NotImplementedError("no DeepSpeed / multi-backend dispatch on the B200 path (north_star)")
        for flag in ("cpu_offload", "zero_quantized_weights", "zero_quantized_gradients", "torch_compile",
                     "dispatching_dataloader", "tensor_parallel_word_embeddings", "sequence_parallel"):
            if getattr(self, flag):
                raise NotImplementedError(f"distributed_args.{flag} is out of scope of the data-parallel B200 path")
        if self.tensor_parallel_size != 1:
            raise NotImplementedError("tensor parallelism is out of scope of the data-parallel B200 path")
        if self.gradient_checkpointing_method is not None:
            # reference enum GradientCheckpointingMethod has the single member `block` (enums.py; gradient_checkpointing/__init__.py)
            if str(self.gradient_checkpointing_method).split(".")[-1] != "block":
                raise ValueError(f"unexpected gradient_checkpointing_method ({self.gradient_checkpointing_method})")
            extra = set(self.gradient_checkpointing_args) - {"checkpoint_every", "use_reentrant", "block_name"}
            if extra:
                raise ValueError(f"unexpected gradient_checkpointing_args {sorted(extra)}")
        zt = self.zero_topology
        if (zt.data_parallel_replication_world_size is None) != (zt.data_parallel_sharding_world_size is None):
            raise AssertionError("data_parallel_replication_world_size and data_parallel_sharding_world_size go together")
        if self.communication_dtype is not None:
            self.communication_dtype = {"bfloat16": "bf16", "float32": "fp32"}.get(self.communication_dtype, self.communication_dtype)
            assert self.communication_dtype in ("bf16", "fp32")


class LoggingArgs(BaseArgs):
    logging_level: str = "INFO"
    log_interval: int = 1
    aim_args: dict | None = None
    wandb_args: dict | None = None
    experiments_tracker_name: str | None = None
    use_colored_logs: bool = False
    torch_profiler_trace_path: str | None = None


class ResearchArgs(BaseArgs):
    neft_alpha: float | None = None


class TrainingArgs(BaseArgs):
    random_args: RandomArgs = RandomArgs()
    tokenizer_args: TokenizerArgs = TokenizerArgs()
    model_args: ModelArgs = None
    tuning_args: TuningArgs = None
    optimizer_args: OptimizerArgs = OptimizerArgs()
    lr_scheduler_args: LRSchedulerArgs = LRSchedulerArgs()
    datasets: list[DatasetArgs] = []
    save_args: SaveArgs = None
    load_args: LoadArgs | None = None
    training_parameters: TrainingParameters | None = None
    logging_args: LoggingArgs = LoggingArgs()
    mixed_precision_args: MixedPrecisionArgs = MixedPrecisionArgs()
    distributed_args: DistributedArgs = DistributedArgs()
    research_args: ResearchArgs = ResearchArgs()

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.model_args, "model_args"), (self.tuning_args, "tuning_args"),
                         (self.save_args, "save_args"), (self.datasets, "datasets")])
        if self.mixed_precision_args.dtype != "bf16":
            raise NotImplementedError("the B200 path trains in bf16 mixed precision (mixed_precision_args.dtype: bf16)")


class GenerationParameters(BaseArgs):
    """arguments.py:449-465"""

    batch_size: int = None
    do_sample: bool | None = None
    max_new_tokens: int = None
    temperature: float | None = None
    top_k: int | None = None
    top_p: float | None = None

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.batch_size, "batch_size"), (self.max_new_tokens, "max_new_tokens")])

    def to_dict(self) -> dict:
        return self.model_dump()


class InferenceArgs(BaseArgs):
    """arguments.py:468-503: either `model_args` (a fresh / pretrained model) or `load_args` (a training checkpoint)"""

    random_args: RandomArgs = RandomArgs()
    tokenizer_args: TokenizerArgs = TokenizerArgs()
    model_args: ModelArgs | None = None
    datasets: list[DatasetArgs] = []
    load_args: LoadArgs | None = None
    generation_parameters: GenerationParameters = None
    mixed_precision_args: MixedPrecisionArgs = MixedPrecisionArgs()
    logging_args: LoggingArgs = LoggingArgs()
    output_dir: str = None

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.datasets, "datasets"), (self.generation_parameters, "generation_parameters"),
                         (self.output_dir, "output_dir")])
        if self.load_args is None:
            assert self.model_args is not None, "model_args need to be specified if load_args are not specified"
        else:
            assert self.model_args is None, "model_args can't be specified with load_args"


class UnshardingArgs(BaseArgs):
    """arguments.py:506-517"""

    load_args: LoadArgs = None
    unsharded_path: str = None
    mixed_precision_args: MixedPrecisionArgs = MixedPrecisionArgs()
    logging_args: LoggingArgs = LoggingArgs()

    def model_post_init(self, __context: Any) -> None:
        _check_not_None([(self.load_args, "load_args"), (self.unsharded_path, "unsharded_path")])


_MODE_ARGS = {"training": TrainingArgs, "inference": InferenceArgs, "unsharding": UnshardingArgs}


def get_args_from_dict(config: dict, mode=None):
    return _MODE_ARGS[str(getattr(mode, "value", mode) or "training")](**config)


def get_args(mode=None):
    """arguments.py:527-547; `mode` is "training" (default), "inference" or "unsharding" (enums.Mode values)"""
    parser = ArgumentParser()
    parser.add_argument("--config", type=str, required=True, help="path for the config")
    a = parser.parse_args()
    return get_args_from_dict(load_yaml(a.config), mode)
