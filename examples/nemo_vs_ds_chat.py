"""Throughput recipe matching the "NeMo vs DeepSpeed-Chat" benchmark (reference: examples/nemo_vs_ds_chat.py).

`NEMO_CONFIG=1.3B|6.7B|13B|20B|33B|66B` picks the model shape and tensor-parallel degree the reference uses (TP1 / TP1 / TP2 /
TP4 on one 8-GPU node, TP8 on four); PPO with seq 512, 256 new tokens (forced: min = max), one PPO epoch per rollout batch, and
a dummy reward model forward so the reward phase costs what a 350M scorer costs.  Launch with one process per GPU:
`python -m torch.distributed.run --nproc-per-node 8 --standalone --local-addr 127.0.0.1 examples/nemo_vs_ds_chat.py`."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import synthetic_dialogues
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config
from trlx_b200.models.modeling_base import build_base_model

SHAPES = {  # layers, hidden, ffn, heads, tp, batch, minibatch, chunk, unfrozen, nodes
    "1.3B": (24, 2048, 8192, 16, 1, 16, 16, 64, -1, 1),
    "6.7B": (32, 4096, 16384, 32, 1, 4, 4, 16, -1, 1),
    "13B": (40, 5120, 20480, 40, 2, 16, 4, 16, -1, 1),
    "20B": (44, 6144, 24576, 64, 4, 16, 2, 16, -1, 1),
    "33B": (48, 7168, 28672, 56, 8, 32, 4, 32, -1, 4),
    "66B": (64, 9216, 36864, 72, 8, 32, 2, 32, 32, 4),
}


def main(hparams={}):
    default_config = TRLConfig.update(default_ppo_config().to_dict(), hparams)
    cfg_name = os.environ.get("NEMO_CONFIG", "1.3B")
    if cfg_name not in SHAPES:
        raise ValueError(f"Unknown NEMO_CONFIG: {cfg_name}")
    L, H, F, heads, tp, batch_size, mini_batch_size, chunk_size, unfrozen, nodes = SHAPES[cfg_name]
    megatron_cfg = dict(name=f"megatron_gpt_{cfg_name.lower()}", trainer=dict(devices=8, num_nodes=nodes, precision="bf16"),
                        model=dict(num_layers=L, hidden_size=H, ffn_hidden_size=F, num_attention_heads=heads,
                                   tensor_model_parallel_size=tp, pipeline_model_parallel_size=1, sequence_parallel=tp > 1,
                                   max_position_embeddings=2048, encoder_seq_length=2048, vocab_size=50257))
    config = default_config.evolve(
        train=dict(total_steps=None, seq_length=512, batch_size=batch_size, minibatch_size=mini_batch_size, epochs=int(1e6),
                   eval_interval=int(1e6), trainer="NeMoPPOTrainer", trainer_kwargs=dict(pretrained_model=None, megatron_cfg=megatron_cfg),
                   checkpoint_interval=int(1e6), checkpoint_dir=f"nemo_{cfg_name}_ppo_ds_chat_benchmark", seed=2023,
                   project_name="trlxnemo", tags=["nemo", "ppo", "benchmark", cfg_name]),
        optimizer=dict(name="distributed_fused_adam", kwargs=dict(lr=6.001e-5, weight_decay=1e-06, eps=1.0e-8, betas=(0.9, 0.95))),
        scheduler=dict(name="CosineAnnealing"),
        model=dict(num_layers_unfrozen=unfrozen),
        method=dict(num_rollouts=chunk_size, init_kl_coef=0.05, scale_reward="ref", vf_coef=1,
                    gen_kwargs=dict(temperature=1.0, max_new_tokens=256, min_new_tokens=256), chunk_size=chunk_size, ppo_epochs=1),
    )
    config.scheduler.kwargs = dict(warmup_steps=0, constant_steps=1e12, min_lr=6.0e-5)
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    device = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    reward_model = build_base_model(dict(model_type="opt", vocab_size=50272, hidden_size=1024, num_hidden_layers=24,
                                         num_attention_heads=16, ffn_dim=4096, max_position_embeddings=2048,
                                         word_embed_proj_dim=512)).eval()  # OPT-350m-shaped scorer

    @torch.no_grad()
    def reward_fn(samples: List[str], tokenizer=None, **kwargs) -> List[float]:
        reward_model.to(device)
        mbs = max(1, config.method.chunk_size // 2)
        for i in range(0, len(samples) // mbs):
            ids = torch.randint(0, 50272, (mbs, 512), device=device)  # same token count the real scorer would see
            reward_model(input_ids=ids)
        reward_model.to("cpu")
        return [0.5 for _ in samples]

    prompts = [d["prompt"] for d in synthetic_dialogues(4096, seed=2023)]
    world = int(os.environ.get("WORLD_SIZE", 1))
    global_batch_size = config.train.batch_size * max(world // tp, 1)
    config.train.total_steps = max(len(prompts) // global_batch_size, 1)
    print(f"Total steps: {config.train.total_steps=} {len(prompts)=} {global_batch_size=}")
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
