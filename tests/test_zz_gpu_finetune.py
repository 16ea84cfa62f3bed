"""GPU: supervised finetuning path end to end (data.finetuning.collate -> ModelWrapperForFinetuning -> train_step).  Kept in its
own, alphabetically last, file: it exercises the newest host code (padded batches, stream padding to 8 tokens)."""
import numpy as np
import pytest
import torch

import oracle.dolomite_oracle as O
from test_gpu_model import GPU_CONFIGS, gpu_config, oracle_params  # same directory (pytest puts tests/ on sys.path)

pytestmark = pytest.mark.gpu


def test_sft_feed_padded_and_padding_free_batches_give_the_same_loss_and_train():
    """finetune path: data.finetuning.collate -> ModelWrapperForFinetuning; the left-padded [B, S] batch and the padding-free
    list batch of the same examples must produce the same loss (same packed stream underneath), and steps reduce it"""
    from dolomite_engine_b200.data.finetuning import build_example, collate
    from dolomite_engine_b200.distributed import ShardedDataParallel
    from dolomite_engine_b200.model_wrapper import ModelWrapperForFinetuning
    from dolomite_engine_b200.optimization import get_optimizer
    from dolomite_engine_b200.train_utils import train_step

    kw = GPU_CONFIGS["hd80_bias"]
    rng = np.random.default_rng(9)
    tok = lambda text: [8 + (sum(map(ord, w)) % 1000) for w in text.split()]  # noqa: E731
    exs = [build_example(tok, 7, " ".join(f"w{rng.integers(0, 500)}" for _ in range(n_in)),
                         " ".join(f"r{rng.integers(0, 500)}" for _ in range(n_out)))
           for n_in, n_out in [(20, 5), (3, 30), (40, 1), (11, 11)]]
    losses = {}
    for pf in (True, False):
        w = ModelWrapperForFinetuning(pretrained_config=gpu_config(kw).to_dict(), use_padding_free_transformer=pf)
        w.model.load_state_dict(oracle_params(O.OracleConfig(**kw)))
        batch = collate(exs, 7, pf)
        losses[pf] = float(w(batch).item())
    assert abs(losses[True] - losses[False]) / losses[True] < 1e-6, losses
    sdp = ShardedDataParallel(w)
    opt = get_optimizer("DolomiteFusedAdamW", {"lr": 1e-3, "weight_decay": 0.0, "betas": [0.9, 0.95], "eps": 1e-10}, sdp)

    def loader():
        while True:
            yield collate(exs, 7, False)

    it = loader()
    hist = [train_step(sdp, opt, None, train_dataloader=it, gradient_accumulation_steps=1, gradient_clipping=1.0)[0] for _ in range(4)]
    assert hist[0] == pytest.approx(losses[False], rel=1e-5) and hist[-1] < hist[0]
