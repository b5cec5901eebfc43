"""PPO for program synthesis grounded in an interpreter: +1 when the generated program reproduces the requested output,
−0.5 when it runs but is wrong, −1 when it does not parse (reference: examples/experiments/grounded_program_synthesis/train_trlx.py)."""
import ast
import json
import sys

import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model
from examples.experiments.grounded_program_synthesis.lang import Interpreter, create_synthetic_dataset
from trlx_b200.data.default_configs import TRLConfig, default_ppo_config

interpreter = Interpreter()


def reward_fn(samples, **kwargs):
    rewards = []
    for sample in samples:
        try:
            code = sample.split("Function:")[1].strip()
            wanted = ast.literal_eval(sample.split("Output:")[1].split("Function:")[0].strip())
        except (IndexError, ValueError, SyntaxError):
            rewards.append(-1.0)
            continue
        got = interpreter(code)
        rewards.append(-1.0 if got == "ERROR" else (1.0 if got == wanted else -0.5))
    return rewards


def main(hparams={}):
    config = default_ppo_config().evolve(
        train=dict(seq_length=256, batch_size=32, total_steps=6000, eval_interval=200, checkpoint_dir="ckpts/program_synthesis"),
        model=dict(model_path=offline_model("reshinthadith/codegen_350M_list_manip_5_len", GPT2_TINY), num_layers_unfrozen=2),
        tokenizer=dict(tokenizer_path="reshinthadith/codegen_350M_list_manip_5_len"),
        method=dict(num_rollouts=128, chunk_size=16, init_kl_coef=0.2, gen_kwargs=dict(max_new_tokens=36, top_k=20, top_p=1.0)))
    config = TRLConfig.update(config, hparams)
    data = create_synthetic_dataset(1100)
    trainer = trlx.train(reward_fn=reward_fn, prompts=[d["input"] for d in data[:1000]], eval_prompts=[d["input"] for d in data[1000:]],
                         config=config)
    trainer.save_pretrained("dataset/trained_model")
    return trainer


if __name__ == "__main__":
    assert reward_fn(["Input: 1 Output: [-4,-5,-2] Function: div_n(reverse([-2, -5, -4]),1)"]) == [1.0]
    assert reward_fn(["Input: 1 Output: [-4,-5,-2] Function: div_n(reverse([-2, -5, -a]),1)"]) == [-1.0]
    assert reward_fn(["Input: 1 Output: [-4,-5,-2] Function: div_n(reverse([-2, -5, -3]),1)"]) == [-0.5]
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
