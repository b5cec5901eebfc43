from .randomwalks import generate_random_walks  # noqa: F401


def online_task(seed: int):
    """``dict(reward_fn=…, metric_fn=…, prompts=…, eval_prompts=…)`` for the online trainers (PPO, RFT): the reward is the
    optimality of a walk (1 = shortest path from its start node), the prompts are the start nodes."""
    metric, prompts, *_ = generate_random_walks(seed=seed)
    return dict(reward_fn=lambda samples, **_: metric(samples)["optimality"], metric_fn=lambda samples, **_: metric(samples),
                prompts=prompts, eval_prompts=prompts)
