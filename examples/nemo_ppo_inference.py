"""Load a model-parallel PPO checkpoint (`mp_rank_XX[_YYY]/model_weights.ckpt`) and sample from it
(reference: examples/nemo_ppo_inference.py).  Run with as many ranks as `TP x PP` of the recipe:

    python -m torch.distributed.run --nproc-per-node 4 --standalone --local-addr 127.0.0.1 \\
        examples/nemo_ppo_inference.py configs/nemo_configs/megatron_20b.yaml <checkpoint dir>
"""
import sys

from trlx_b200.data.default_configs import default_ppo_config
from trlx_b200.utils.loading import get_trainer


def build_trainer(megatron_cfg_path: str, trainer_name: str = "NeMoPPOTrainer", base=None):
    base = base or default_ppo_config()
    config = base.evolve(train=dict(trainer=trainer_name, tracker=None,
                                    trainer_kwargs=dict(pretrained_model=None, megatron_cfg=megatron_cfg_path)))
    trainer = get_trainer(trainer_name)(config=config, reward_fn=None, metric_fn=None, stop_sequences=[])
    if trainer.runtime.dp_size != 1:
        raise ValueError("Inference only supports data parallel world size of 1")
    return trainer


def main(megatron_cfg_path: str, checkpoint_path: str, prompts=("I don't know much about Hungarian underground",)):
    trainer = build_trainer(megatron_cfg_path)
    trainer.load_from_pretrained(checkpoint_path)
    enc = trainer.tokenizer(list(prompts), return_tensors="pt", padding=True)
    out = trainer.generate_eval(enc.input_ids, enc.attention_mask, max_new_tokens=40, min_new_tokens=0)
    if trainer.runtime.is_main_process:
        print(trainer.tokenizer.batch_decode(out, skip_special_tokens=True))
    return out


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
