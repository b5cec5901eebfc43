"""Toy offline/online RL task: find short paths to node `a` in a random directed graph (walks are strings of letters).

Task definition follows the reference's ``examples/randomwalks/randomwalks.py`` (21 nodes, edge probability 0.1, goal
node 0 absorbing, walks of at most 10 nodes, optimality = (max_len − len) / (max_len − shortest_len)); this
implementation uses a plain BFS for shortest paths (no networkx) and vectorised validity checks.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Dict, List, Tuple

import numpy as np
import torch


def _shortest_to_goal(adj: np.ndarray, goal: int, cap: int) -> List[int]:
    """Number of nodes on the shortest path from every node to ``goal`` (BFS on reversed edges), capped at ``cap``."""
    n = adj.shape[0]
    dist = [None] * n
    dist[goal] = 1
    queue = deque([goal])
    while queue:
        v = queue.popleft()
        for u in np.nonzero(adj[:, v])[0]:
            if dist[u] is None:
                dist[u] = dist[v] + 1
                queue.append(int(u))
    return [min(d, cap) if d is not None else cap for d in dist]


def generate_random_walks(n_nodes: int = 21, max_length: int = 10, n_walks: int = 1000, p_edge: float = 0.1, seed: int = 1002,
                          gpt2_tokenizer: bool = False
                          ) -> Tuple[Callable[[List[str]], Dict[str, List[float]]], List[str], List[str], torch.Tensor]:
    """Returns ``(metric_fn, eval_prompts, sample_walks, logit_mask)``."""
    assert n_nodes <= 26
    rng = np.random.RandomState(seed)
    while True:
        adj = rng.rand(n_nodes, n_nodes) > (1 - p_edge)
        np.fill_diagonal(adj, 0)
        if np.all(adj.sum(1)):  # every node has an outgoing edge
            break
    goal = 0
    adj[goal, :] = 0
    adj[goal, goal] = 1
    letters = [chr(ord("a") + i) for i in range(n_nodes)]
    index = {c: i for i, c in enumerate(letters)}
    delimiter = "|" if gpt2_tokenizer else ""

    walks: List[str] = []
    for _ in range(n_walks):
        node = goal
        while node == goal:
            node = rng.randint(n_nodes)
        path = [node]
        for _step in range(max_length - 1):
            node = int(rng.choice(np.nonzero(adj[node])[0]))
            path.append(node)
            if node == goal:
                break
        walks.append(delimiter.join(letters[i] for i in path))

    shortest = _shortest_to_goal(adj, goal, max_length)
    invalid = 100.0

    def metric_fn(samples: List[str], **_unused) -> Dict[str, List[float]]:
        lengths, optimal = [], []
        for s in samples:
            if gpt2_tokenizer:
                s = s.replace("|", "")
            nodes = [index.get(c, 1000) for c in s]
            length = invalid
            for i, v in enumerate(nodes):
                if v >= n_nodes or (i > 0 and not adj[nodes[i - 1], v]):
                    break
                if v == goal:
                    length = float(i + 1)
                    break
            lengths.append(length)
            first = nodes[0] if nodes and nodes[0] < n_nodes else 1
            optimal.append(shortest[first])
        lt = torch.tensor(lengths)
        bound = torch.where(lt.eq(invalid), torch.tensor(float(max_length)), lt)
        opt = torch.tensor(optimal, dtype=torch.float)
        optimality = (max_length - bound) / (max_length - opt).clamp_min(1e-6)
        return {"lengths": lengths, "optimality": optimality.tolist()}

    eval_prompts = [p + delimiter for p in sorted({w[0] for w in walks})]
    return metric_fn, eval_prompts, walks, torch.tensor(adj)


ALPHABET = "abcdefghijklmnopqrstu"
TOKENIZER = f"toy://chars?alphabet={ALPHABET}"
# tiny GPT-2 of the size the reference's ILQL twin uses (examples/randomwalks/ilql_randomwalks.py:25,46)
MODEL = dict(model_type="gpt2", n_layer=6, n_embd=144, n_head=12, vocab_size=23, n_positions=16, bos_token_id=21, eos_token_id=21)
