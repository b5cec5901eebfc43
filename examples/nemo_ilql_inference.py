"""Sample from a model-parallel ILQL checkpoint with the advantage-shifted policy (reference: examples/nemo_ilql_inference.py)."""
import sys

from examples.nemo_ppo_inference import build_trainer
from trlx_b200.data.default_configs import default_ilql_config


def main(megatron_cfg_path: str, checkpoint_path: str, prompts=("I don't know much about Hungarian underground",), beta: float = 2.0):
    trainer = build_trainer(megatron_cfg_path, "NeMoILQLTrainer", default_ilql_config())
    trainer.load_from_pretrained(checkpoint_path)
    enc = trainer.tokenizer(list(prompts), return_tensors="pt", padding=True)
    out = trainer.generate_eval(enc.input_ids, enc.attention_mask, max_new_tokens=40, beta=beta, temperature=0.9)
    if trainer.runtime.is_main_process:
        print(trainer.tokenizer.batch_decode(out, skip_special_tokens=True))
    return out


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
