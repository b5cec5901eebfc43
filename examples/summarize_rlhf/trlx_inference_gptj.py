"""Name of the evaluation script in the reference (examples/summarize_rlhf/trlx_inference_gptj.py); see ``trlx_inference.py``."""
import runpy

if __name__ == "__main__":
    runpy.run_module("examples.summarize_rlhf.trlx_inference", run_name="__main__")
