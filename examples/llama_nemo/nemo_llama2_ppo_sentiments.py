"""PPO sentiments on a Llama-2-shaped model with the tensor/sequence-parallel trainer
(reference: examples/llama_nemo/nemo_llama2_ppo_sentiments.py + megatron_llama_cfg.yaml, 7B on 8 GPUs with TP).

The reference first converts HF Llama weights to a `.nemo` archive (`convert_llama_to_nemo.py`).  Here the trainer reads HF
checkpoints directly and shards them at load time (`apply_tensor_parallel`), so no conversion step exists; pass the checkpoint
directory as `model.model_path`.  `examples/llama_nemo/convert_llama.py` re-shards a TP checkpoint back into a single HF one."""
import json
import os
import sys
from typing import List

import trlx_b200 as trlx
from examples._offline import LLAMA_TINY, load_imdb, offline_model, sentiment_scorer
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config

LLAMA2_7B = dict(model_type="llama", vocab_size=32000, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=32, intermediate_size=11008, max_position_embeddings=4096, rms_norm_eps=1e-5)


def main(hparams={}):
    tp = int(os.environ.get("TENSOR_PARALLEL", 4))
    config = default_ppo_config().evolve(
        train=dict(total_steps=1600, seq_length=256, batch_size=16, epochs=100, eval_interval=100, trainer="NeMoPPOTrainer",
                   checkpoint_interval=400, checkpoint_dir="llama2_7b_ppo_sentiments", seed=2023, project_name="trlxnemo",
                   tags=["nemo", "ppo", "sentiments", "llama2-7b"],
                   parallel=dict(tensor_parallel=tp, sequence_parallel=tp > 1)),
        model=dict(model_path=offline_model("NousResearch/Llama-2-7b-hf", LLAMA2_7B if os.environ.get("FULL_SIZE") else LLAMA_TINY),
                   num_layers_unfrozen=2),
        tokenizer=dict(tokenizer_path="NousResearch/Llama-2-7b-hf"),
        optimizer=dict(name="distributed_fused_adam", kwargs=dict(lr=1.001e-5, weight_decay=1e-06, eps=1.0e-8, betas=(0.9, 0.95))),
        scheduler=dict(name="CosineAnnealing", kwargs=dict(warmup_steps=0, constant_steps=1e12, min_lr=1.0e-5)),
        method=dict(num_rollouts=128, init_kl_coef=0.05, vf_coef=1, scale_reward="ignored", gamma=1, lam=0.95,
                    gen_kwargs=dict(temperature=1.0, max_new_tokens=64), chunk_size=64, ppo_epochs=4),
    )
    config = TRLConfig.update(config, hparams)
    sentiment_fn = sentiment_scorer(int(os.environ.get("LOCAL_RANK", 0)))

    def reward_fn(samples: List[str], **kwargs) -> List[float]:
        return [s["POSITIVE"] for s in sentiment_fn(samples)]

    texts, _ = load_imdb()
    prompts = [" ".join(review.split()[:4]) for review in texts]
    return trlx.train(reward_fn=reward_fn, prompts=prompts, eval_prompts=["I don't know much about Hungarian underground"] * 256,
                      config=config)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
