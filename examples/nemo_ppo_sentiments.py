"""PPO sentiments on the tensor/sequence/pipeline-parallel trainer (reference: examples/nemo_ppo_sentiments.py).

`NEMO_CONFIG=1.3B|2B|20B|65B` selects `configs/nemo_configs/megatron_<size>.yaml`.  Launch with one process per GPU, e.g.
`python -m torch.distributed.run --nproc-per-node 8 --standalone --local-addr 127.0.0.1 examples/nemo_ppo_sentiments.py`."""
import json
import os
import sys
from typing import List

import trlx_b200 as trlx
from examples._offline import load_imdb, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config


def main(hparams={}):
    default_config = TRLConfig.update(default_ppo_config().to_dict(), hparams)
    cfg_name = os.environ.get("NEMO_CONFIG", "1.3B")
    if cfg_name not in ("1.3B", "2B", "20B", "65B"):
        raise ValueError(f"Unknown NEMO_CONFIG: {cfg_name}")
    config = default_config.evolve(
        train=dict(total_steps=512, seq_length=2048, batch_size=32, epochs=100, eval_interval=64, trainer="NeMoPPOTrainer",
                   trainer_kwargs=dict(pretrained_model=f"/mnt/hdd/nemo-megatron-gpt-{cfg_name}/",
                                       megatron_cfg=f"megatron_{cfg_name.lower()}.yaml"),
                   checkpoint_interval=256, checkpoint_dir=f"nemo_{cfg_name}_ppo_sentiments", seed=2023, project_name="trlxnemo",
                   tags=["nemo", "ppo", "sentiments", cfg_name]),
        optimizer=dict(name="distributed_fused_adam", kwargs=dict(lr=6.001e-5, weight_decay=1e-06, eps=1.0e-8, betas=(0.9, 0.95))),
        scheduler=dict(name="CosineAnnealing"),
        model=dict(num_layers_unfrozen=2),
        method=dict(num_rollouts=128, init_kl_coef=0.05, scale_reward="ref", vf_coef=1,
                    gen_kwargs=dict(temperature=1.0, max_new_tokens=40), chunk_size=128, ppo_epochs=4),
    )
    config.scheduler.kwargs = dict(warmup_steps=0, constant_steps=1e12, min_lr=6.0e-5)
    config = TRLConfig.update(config, {k: v for k, v in hparams.items() if k.startswith(("train.", "model.", "method."))})
    local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("SLURM_LOCALID", 0)))
    sentiment_fn = sentiment_scorer(local_rank)

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
