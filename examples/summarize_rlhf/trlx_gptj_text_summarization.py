"""Stage 3: PPO against the learned reward model, reward = r(sample) − r(post + reference summary)
(reference: examples/summarize_rlhf/trlx_gptj_text_summarization.py)."""
import json
import os
import sys
from typing import List

import torch

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, GPTJ_TINY, offline_model, overlap_f1, synthetic_summaries
from examples.summarize_rlhf.reward_model.reward_model import GPTRewardModel
from trlx_b200.data.configs import ModelConfig, OptimizerConfig, SchedulerConfig, TokenizerConfig, TrainConfig, TRLConfig
from trlx_b200.models.modeling_ppo import PPOConfig
from trlx_b200.utils.tokenizer import load_tokenizer

REWARD_CHECKPOINT_PATH = os.environ.get("REWARD_CHECKPOINT", "rm_checkpoint/reward_model.pt")
SFT_MODEL_PATH = os.environ.get("SFT_MODEL", "CarperAI/openai_summarize_tldr_sft")

config = TRLConfig(
    train=TrainConfig(seq_length=550, epochs=50, total_steps=100000, batch_size=4, checkpoint_interval=10000, eval_interval=200,
                      pipeline="PromptPipeline", trainer="AcceleratePPOTrainer"),
    model=ModelConfig(model_path=offline_model(SFT_MODEL_PATH, GPT2_TINY), num_layers_unfrozen=8),
    tokenizer=TokenizerConfig(tokenizer_path="gpt2", truncation_side="right"),
    optimizer=OptimizerConfig(name="adamw", kwargs={"lr": 5.0e-6, "betas": [0.9, 0.999], "eps": 1.0e-8, "weight_decay": 0.01}),
    scheduler=SchedulerConfig(name="cosine_annealing", kwargs={"T_max": 100000, "eta_min": 5.0e-6}),
    method=PPOConfig(name="PPOConfig", num_rollouts=128, chunk_size=16, ppo_epochs=4, init_kl_coef=0.1, target=6, horizon=10000,
                     gamma=1, lam=0.95, cliprange=0.2, cliprange_value=0.2, vf_coef=0.2, scale_reward=None, ref_mean=None,
                     ref_std=None, cliprange_reward=10, gen_kwargs={"max_new_tokens": 50}),
)


def main(hparams={}):
    cfg = TRLConfig.update(config, hparams)
    rw_tok = load_tokenizer("EleutherAI/gpt-j-6B")
    rw_tok.pad_token = rw_tok.eos_token
    rw_model = GPTRewardModel(offline_model(SFT_MODEL_PATH, GPTJ_TINY), rw_tok.pad_token_id)
    if os.path.exists(REWARD_CHECKPOINT_PATH):
        rw_model.load_state_dict(torch.load(REWARD_CHECKPOINT_PATH, map_location="cpu"), strict=False)
    rw_device = torch.device("cuda", torch.cuda.device_count() - 1) if torch.cuda.is_available() else torch.device("cpu")
    rw_model = (rw_model.half() if rw_device.type == "cuda" else rw_model).to(rw_device).eval()

    @torch.no_grad()
    def get_scores(samples: List[str]) -> torch.Tensor:
        out = []
        for i in range(0, len(samples), 2):
            batch = ["<|startoftext|>" + s + "<|endoftext|>" for s in samples[i:i + 2]]
            enc = rw_tok(batch, truncation=True, max_length=cfg.train.seq_length, padding="max_length", return_tensors="pt")
            ids = enc.input_ids.repeat(2, 1).to(rw_device)  # identical halves ⇒ the model's inference branch
            out.append(rw_model(ids, enc.attention_mask.repeat(2, 1).to(rw_device))["chosen_end_scores"].float().cpu())
        return torch.cat(out)

    data = synthetic_summaries(2048)
    post_summary = {d["prompt"].strip(): d["label"] for d in data}

    def reward_fn(samples: List[str], prompts: List[str], outputs: List[str], **kwargs):
        originals = [p + post_summary.get(p.strip(), "") for p in prompts]
        return get_scores(samples) - get_scores(originals)  # normalised by the reference summary's score

    def metric_fn(samples, prompts, outputs, **kw):
        return {"overlap_f1": [overlap_f1(o, post_summary.get(p.strip(), "")) for p, o in zip(prompts, outputs)]}

    return trlx.train(reward_fn=reward_fn, metric_fn=metric_fn, prompts=[d["prompt"] for d in data[:-128]],
                      eval_prompts=[d["prompt"] for d in data[-128:]][:1000], config=cfg)


if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
