"""Name of this stage in the reference (examples/summarize_rlhf/sft/train_gptj_summarize.py); the implementation lives in
``train_sft_summarize.py`` (the framework's own SFT trainer instead of the HF ``Trainer`` + DeepSpeed)."""
import json
import sys

from examples.summarize_rlhf.sft.train_sft_summarize import main  # noqa: F401

if __name__ == "__main__":
    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
