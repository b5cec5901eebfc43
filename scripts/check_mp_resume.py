"""Save / resume round trip of a model-parallel trainer (run under torch.distributed.run; PP=<stages>, CKPT_DIR=<dir>): one
optimizer step, ``save`` + ``save_pretrained``, a fresh trainer ``load``s and must reproduce the logits exactly."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trlx_b200.data.default_configs import default_sft_config
from trlx_b200.utils import set_seed
from trlx_b200.utils.loading import get_trainer
import torch.distributed as dist
arch = dict(model_type="gpt_neox", vocab_size=512, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128, max_position_embeddings=128, rotary_pct=0.25)
pp = int(os.environ.get("PP", "1")); world = int(os.environ["WORLD_SIZE"]); tp = int(os.environ.get("TP", world // pp))
d = os.environ["CKPT_DIR"]
cfg = default_sft_config().evolve(train=dict(seq_length=32, batch_size=4, trainer="NeMoSFTTrainer", tracker=None, checkpoint_dir=d, checkpoint_interval=10**9, eval_interval=10**9, total_steps=10**9, parallel=dict(tensor_parallel=tp, pipeline_parallel=pp, sequence_parallel=pp == 1)), model=dict(model_path=arch), tokenizer=dict(tokenizer_path="toy://bpe?vocab=512"))
def make():
    set_seed(cfg.train.seed, cfg.train.parallel)
    t = get_trainer(cfg.train.trainer)(config=cfg, **cfg.train.trainer_kwargs)
    return t
t = make()
texts = ["the movie was really quite good and long"] * 16
t.make_experience(texts, cfg.train.seq_length)
from trlx_b200.pipeline import MiniBatchIterator
mb = next(iter(MiniBatchIterator(t.create_train_dataloader(), t.mb_size, t.num_mb)))
t.train_step(mb)
ids = torch.arange(24).view(2, 12) % 500
with torch.no_grad():
    a = t.model(ids, attention_mask=torch.ones_like(ids)).logits
t.save(os.path.join(d, "ck"))
t.save_pretrained(os.path.join(d, "hf"))
dist.barrier()
t2 = make()
t2.load(os.path.join(d, "ck"))
with torch.no_grad():
    b = t2.model(ids, attention_mask=torch.ones_like(ids)).logits
err = (a - b).abs().max().item()
# the optimizer state came back too: one more identical step on both trainers must land on the same weights
t.train_step(mb)
t2.train_step(mb)
with torch.no_grad():
    a2 = t.model(ids, attention_mask=torch.ones_like(ids)).logits
    b2 = t2.model(ids, attention_mask=torch.ones_like(ids)).logits
step_err = (a2 - b2).abs().max().item()
assert step_err < 1e-5 and (a2 - a).abs().max().item() > 0, (step_err, "weights did not move" if (a2 - a).abs().max().item() == 0 else "")
print(f"rank {dist.get_rank()} resume_err {err:.3e} iter {t2.iter_count} files {sorted(os.listdir(os.path.join(d,'ck')))[:6]} hf {sorted(os.listdir(os.path.join(d,'hf')))[:6]}", flush=True)
assert err < 1e-5
# weights-only export / import in the per-rank layout (``mp_rank_XX[_YYY]/model_weights.ckpt``) of the model-parallel trainers
if hasattr(t2, "load_from_pretrained") and (tp > 1 or pp > 1):
    t3 = make()
    t3.load_from_pretrained(os.path.join(d, "hf"))
    with torch.no_grad():
        c = t3.model(ids, attention_mask=torch.ones_like(ids)).logits
    assert (c - a).abs().max().item() < 1e-5, (c - a).abs().max().item()
    print(f"rank {dist.get_rank()} pretrained_roundtrip ok", flush=True)
