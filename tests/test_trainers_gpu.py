"""Trainers end to end on the GPU kernels: whole-step CUDA graph vs eager step, ILQL (gathered-form loss through the fused
LM-head kernel), SFT, RFT and the encoder-decoder PPO path."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GPT2 = dict(model_type="gpt2", vocab_size=512, n_embd=128, n_layer=4, n_head=2, n_positions=64, eos_token_id=256, bos_token_id=256)
T5 = dict(model_type="t5", vocab_size=512, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_heads=4, eos_token_id=256, pad_token_id=0,
          decoder_start_token_id=0)
PROMPTS = ["hello world", "the quick", "a", "brown fox jumps"] * 4


def _ppo_cfg(tmp, **train):
    from trlx_b200.data.default_configs import default_ppo_config

    return default_ppo_config().evolve(
        train=dict(dict(seq_length=32, batch_size=8, total_steps=4, epochs=2, checkpoint_interval=100, eval_interval=100, tracker=None,
                        checkpoint_dir=str(tmp), seed=3), **train),
        model=dict(model_path=GPT2, num_layers_unfrozen=2), tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(num_rollouts=16, chunk_size=8, ppo_epochs=2, gen_kwargs=dict(max_new_tokens=8, top_k=0, top_p=1.0, do_sample=True)))


def test_graphed_train_step_matches_eager_step(tmp_path, monkeypatch):
    """Same weights, same batch: one optimizer step replayed from the captured CUDA graph == the eager step."""
    from trlx_b200.pipeline import MiniBatchIterator
    from trlx_b200.utils.loading import get_pipeline, get_trainer

    def build():
        torch.manual_seed(0)
        cfg = _ppo_cfg(tmp_path)
        tr = get_trainer(cfg.train.trainer)(config=cfg, reward_fn=lambda samples, **kw: [float(len(s)) for s in samples],
                                            metric_fn=None, stop_sequences=[])
        tr.add_prompt_pipeline(get_pipeline("PromptPipeline")(PROMPTS, 16, tr.tokenizer))
        return tr

    a = build()
    a.make_experience(8)
    assert a._graphs_enabled()
    loader = a.store.create_loader(8, shuffle=False, static_shapes=True)
    batch = next(iter(MiniBatchIterator(loader, a.mb_size, a.num_mb)))
    assert hasattr(batch[0], "width"), "the micro-batch splitter must keep the batch annotations the graph path keys on"
    before = [p.detach().float().clone() for p in a.model.parameters() if p.requires_grad]
    stats_g = a.train_step(batch)
    assert len(a._graphed_steps) == 1, "the whole-step CUDA graph was not engaged"
    after_g = [p.detach().float().clone() for p in a.model.parameters() if p.requires_grad]

    monkeypatch.setenv("TRLX_B200_TRAIN_GRAPH", "0")
    b = build()
    b.model.load_state_dict(a.model.state_dict(), strict=False)
    with torch.no_grad():
        for p, w in zip([p for p in b.model.parameters() if p.requires_grad], before):
            p.copy_(w.to(p.dtype))
    b.opt = b.setup_optimizer()
    b.scheduler = b.setup_scheduler()
    stats_e = b.train_step(batch)
    assert not b._graphed_steps
    after_e = [p.detach().float() for p in b.model.parameters() if p.requires_grad]
    moved = 0.0
    for g, e, w in zip(after_g, after_e, before):
        moved = max(moved, (e - w).abs().max().item())
        assert (g - e).abs().max().item() <= 2e-2 * max((e - w).abs().max().item(), 1e-6) + 1e-3
    assert moved > 0
    assert abs(float(stats_g["losses/total_loss"]) - float(stats_e["losses/total_loss"])) < 5e-2


def test_ilql_trainer_on_gpu(tmp_path):
    import trlx_b200 as trlx
    from trlx_b200.data.default_configs import default_ilql_config

    cfg = default_ilql_config().evolve(
        train=dict(seq_length=32, batch_size=8, total_steps=4, epochs=4, checkpoint_interval=100, eval_interval=2, tracker=None,
                   checkpoint_dir=str(tmp_path), seed=3),
        model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes"),
        method=dict(gen_kwargs=dict(max_new_tokens=6, top_k=4, beta=1, temperature=1.0)))
    samples = [[p, " yes it is"] for p in PROMPTS]
    trainer = trlx.train(samples=samples, rewards=[float(i % 3) for i in range(len(samples))], eval_prompts=["hello", "the"], config=cfg)
    assert trainer.iter_count == 4
    for p in trainer.model.parameters():
        assert torch.isfinite(p).all()


def test_sft_and_rft_trainers_on_gpu(tmp_path):
    import trlx_b200 as trlx
    from trlx_b200.data.configs import TRLConfig
    from trlx_b200.data.default_configs import default_sft_config
    from trlx_b200.trainer.accelerate_rft_trainer import RFTConfig

    common = dict(seq_length=32, batch_size=8, total_steps=3, epochs=3, checkpoint_interval=100, eval_interval=100, tracker=None,
                  checkpoint_dir=str(tmp_path), seed=3)
    sft = default_sft_config().evolve(train=common, model=dict(model_path=GPT2), tokenizer=dict(tokenizer_path="toy://bytes"),
                                      method=dict(gen_kwargs=dict(max_new_tokens=4)))
    t1 = trlx.train(samples=[p + " indeed" for p in PROMPTS], eval_prompts=["hello"], config=sft)
    assert t1.iter_count == 3
    rft = TRLConfig.from_dict(dict(sft.to_dict(), method=RFTConfig(name="RFTConfig", n_generations_per_prompt=2, start_percentile=0.5,
                                                                   end_percentile=0.9, n_improve_steps=1,
                                                                   gen_kwargs=dict(max_new_tokens=4, do_sample=True, top_k=0, top_p=1.0)).to_dict()))
    rft = rft.evolve(train=dict(trainer="AccelerateRFTTrainer"))
    t2 = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples], prompts=PROMPTS, eval_prompts=["hello"], config=rft)
    assert t2.iter_count >= 1


def test_seq2seq_ppo_on_gpu(tmp_path):
    import trlx_b200 as trlx

    cfg = _ppo_cfg(tmp_path).evolve(model=dict(model_path=T5, model_arch_type="seq2seq", num_layers_unfrozen=-1),
                                    tokenizer=dict(padding_side="right"))
    trainer = trlx.train(reward_fn=lambda samples, **kw: [float(len(s)) for s in samples], prompts=PROMPTS, eval_prompts=["hi"] * 2,
                         config=cfg)
    assert trainer.iter_count == 4


def test_bf16_gradient_accumulation_error_is_bounded():
    """`.grad` tensors are views of the optimizer's flat bf16 buffer, so micro-batch accumulation rounds to bf16 after every
    backward.  Over 8 micro-batches the accumulated gradient must stay within bf16 rounding noise of an fp32 accumulation of
    the same per-micro-batch gradients (relative L2 error and cosine per tensor)."""
    from trlx_b200.parallel.optim import FusedAdamW

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.GELU(), torch.nn.Linear(512, 256), torch.nn.GELU(),
                              torch.nn.Linear(256, 16)).cuda().to(torch.bfloat16)
    opt = FusedAdamW(net.parameters(), lr=1e-3).prepare()
    assert opt._flat is not None and all(p.grad.dtype == torch.bfloat16 for p in net.parameters())
    fp32_acc = [torch.zeros_like(p, dtype=torch.float32) for p in net.parameters()]
    n_mb = 8
    for i in range(n_mb):
        x = torch.randn(64, 256, device="cuda").to(torch.bfloat16)
        y = torch.randn(64, 16, device="cuda")
        loss = (net(x).float() - y).pow(2).mean() / n_mb
        # per-micro-batch gradient in fp32 (autograd.grad does not touch .grad), accumulated in fp32 ...
        gs = torch.autograd.grad(loss, list(net.parameters()), retain_graph=True)
        for acc, g in zip(fp32_acc, gs):
            acc += g.float()
        loss.backward()  # ... while .grad accumulates in the flat bf16 buffer
    for p, ref in zip(net.parameters(), fp32_acc):
        got = p.grad.float()
        rel = (got - ref).norm() / ref.norm().clamp_min(1e-12)
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
        assert rel < 1.5e-2 and cos > 0.9999, (tuple(p.shape), rel.item(), cos.item())
