"""ncu target: a few decode steps of the GPT-2 124M rollout engine with the megakernel, eager launches (no CUDA graph)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200.engine.rollout import RolloutEngine  # noqa: E402
from trlx_b200.models.modeling_ppo import AutoModelForCausalLMWithHydraValueHead  # noqa: E402
from trlx_b200.utils.modeling import freeze_bottom_causal_layers  # noqa: E402

torch.manual_seed(0)
m = AutoModelForCausalLMWithHydraValueHead.from_pretrained("gpt2", num_layers_unfrozen=2)
freeze_bottom_causal_layers(m.base_model, 2)
m = m.cuda().to(torch.bfloat16).eval()
V = m.base_model.config.vocab_size
new = int(os.environ.get("NCU_NEW", 24))
gen = dict(max_new_tokens=new, do_sample=True, eos_token_id=V - 1, pad_token_id=V - 1, top_k=0, top_p=1.0, min_new_tokens=new)
eng = RolloutEngine(m, V - 1, V - 1, gen, seed=0, use_cuda_graph=False)
ids = torch.randint(0, V - 2, (128, 8), device="cuda")
eng.rollout(ids, torch.ones_like(ids))
torch.cuda.synchronize()
