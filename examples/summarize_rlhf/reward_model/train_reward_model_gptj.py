"""Name of this stage in the reference (examples/summarize_rlhf/reward_model/train_reward_model_gptj.py); the implementation
lives in ``train_reward_model.py``."""
import json
import sys

from examples.summarize_rlhf.reward_model.train_reward_model import main  # noqa: F401

if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
