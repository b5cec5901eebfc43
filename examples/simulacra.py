"""ILQL on (caption, aesthetic rating) pairs (reference: examples/simulacra.py, which reads the Simulacra Aesthetic Captions
sqlite dump).  Point `SAC_SQLITE` at a local copy of `sac_public_2022_06_29.sqlite`; otherwise synthetic rated captions are used."""
import os
import random
import sqlite3

import trlx_b200 as trlx
from trlx_b200.data.default_configs import default_ilql_config


def load_ratings():
    dbpath = os.environ.get("SAC_SQLITE", "sac_public_2022_06_29.sqlite")
    if os.path.exists(dbpath):  # pragma: no cover - needs the dataset
        c = sqlite3.connect(dbpath).cursor()
        c.execute("SELECT prompt, rating FROM ratings JOIN images ON images.id=ratings.iid "
                  "JOIN generations ON images.gid=generations.id WHERE rating IS NOT NULL;")
        return tuple(map(list, zip(*c.fetchall())))
    rng = random.Random(0)
    subjects = ["an astronaut", "a lighthouse", "a fox", "a city street", "a mountain lake"]
    styles = [("oil painting, highly detailed", 8), ("blurry phone photo", 2), ("studio lighting, 4k", 7), ("crayon scribble", 3)]
    prompts, ratings = [], []
    for _ in range(2048):
        s, (st, r) = rng.choice(subjects), rng.choice(styles)
        prompts.append(f"{s}, {st}")
        ratings.append(max(1, min(10, r + rng.randint(-1, 1))))
    return prompts, ratings


def main(hparams={}):
    prompts, ratings = load_ratings()
    config = default_ilql_config()
    if hparams:
        from trlx_b200.data.configs import TRLConfig

        config = TRLConfig.update(config, hparams)
    return trlx.train(config=config, samples=prompts, rewards=ratings, eval_prompts=["An astronaut riding a horse"] * 64)


if __name__ == "__main__":
    import json
    import sys

    main({} if len(sys.argv) == 1 else json.loads(sys.argv[1]))
