"""Name of the conversion script in the reference (examples/llama_nemo/convert_llama_to_nemo.py: HF LLaMA → NeMo tensor-parallel
checkpoint).  The implementation is ``convert_llama.py`` (``shard`` / ``unshard`` between an HF directory and
``mp_rank_XX/model_weights.ckpt``):

    python examples/llama_nemo/convert_llama_to_nemo.py shard <hf_dir> <out_dir> --tp 4
"""
from examples.llama_nemo.convert_llama import main

if __name__ == "__main__":
    main()
