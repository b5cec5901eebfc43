"""Toy reward: fewer rooms in a generated floor-plan description (reference: examples/architext.py)."""
import trlx_b200 as trlx
from examples._offline import GPT2_TINY, offline_model
from trlx_b200.data.default_configs import default_ppo_config


def reward_fn(samples, **kwargs):
    "Gives a negative count of rooms for each sample"
    return [-sample.count(":") for sample in samples]


prompts = [f"[prompt] {a} {room} is {neg}adjacent to the {other} [layout]"
           for neg in ("", "not ") for a, room, other in (("the", "bedroom", "living room"), ("a", "bedroom", "living room"),
                                                           ("the", "bedroom", "kitchen"), ("a", "bedroom", "kitchen"),
                                                           ("the", "kitchen", "bathroom"), ("a", "bathroom", "living room"),
                                                           ("the", "bathroom", "living room"))]


def main(hparams={}):
    config = default_ppo_config()
    if hparams:
        from trlx_b200.data.configs import TRLConfig

        config = TRLConfig.update(config, hparams)
    return trlx.train(model_path=offline_model("architext/gptj-162M", GPT2_TINY), reward_fn=reward_fn, prompts=prompts, config=config)


if __name__ == "__main__":
    import json
    import sys

    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
