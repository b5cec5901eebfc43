"""Package the HH reward model for serving (reference: examples/hh/to_triton.py, which traces the GPT-J reward model and
writes a Triton Inference Server model repository with a ``config.pbtxt``).

This image has no Triton server; the serving stack here is the FastAPI / uvicorn app in ``examples/hh/reward.py``.  This
script produces the equivalent *model repository* — ``<out>/<name>/1/`` with the reward-model weights, the architecture
config and a ``serving.json`` describing inputs/outputs, batch limit and port — and can start the server on it:

    python -m examples.hh.to_triton --checkpoint rm_checkpoint --out model_store --name gptj-rm-static [--serve --port 8000]

Training processes then score through ``REWARD_HOST=host:port`` (the role of ``TRITON_HOST`` in the reference)."""
import argparse
import json
import os
import shutil


def build_repository(checkpoint: str, out: str, name: str, max_batch_size: int = 25, port: int = 8000) -> str:
    version_dir = os.path.join(out, name, "1")
    os.makedirs(version_dir, exist_ok=True)
    copied = []
    if checkpoint and os.path.isdir(checkpoint):
        for fn in os.listdir(checkpoint):
            if fn.endswith((".bin", ".safetensors", ".json", ".ckpt", ".model", ".txt")):
                shutil.copy2(os.path.join(checkpoint, fn), os.path.join(version_dir, fn))
                copied.append(fn)
    elif checkpoint and os.path.isfile(checkpoint):
        shutil.copy2(checkpoint, os.path.join(version_dir, os.path.basename(checkpoint)))
        copied.append(os.path.basename(checkpoint))
    spec = dict(name=name, platform="trlx_b200_fastapi", max_batch_size=max_batch_size, port=port, files=sorted(copied),
                input=[dict(name="samples", data_type="TYPE_STRING", dims=[-1])],
                output=[dict(name="rewards", data_type="TYPE_FP32", dims=[-1])],
                endpoint="POST /score  {\"samples\": [...]} -> {\"rewards\": [...]}")
    with open(os.path.join(out, name, "serving.json"), "w") as fh:
        json.dump(spec, fh, indent=2)
    # for deployments that do have a Triton server: the matching ``config.pbtxt`` (a traced ``reward-model.pt`` goes into ``1/``)
    template = os.path.join(os.path.dirname(os.path.abspath(__file__)), "triton_config.pbtxt")
    if os.path.exists(template):
        from string import Template

        with open(template) as fh, open(os.path.join(out, name, "config.pbtxt"), "w") as dst:
            dst.write(Template(fh.read()).substitute(model_name=name, max_batch_size=max_batch_size))
    return version_dir


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--checkpoint", default=os.environ.get("REWARD_CHECKPOINT"), help="reward-model checkpoint (dir or file)")
    ap.add_argument("--out", default="model_store")
    ap.add_argument("--name", default="gptj-rm-static")
    ap.add_argument("--max_batch_size", type=int, default=25)
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--serve", action="store_true", help="start the reward server on the packaged model")
    a = ap.parse_args(argv)
    version_dir = build_repository(a.checkpoint, a.out, a.name, a.max_batch_size, a.port)
    print(f"model repository entry written to {os.path.join(a.out, a.name)}")
    if a.serve:  # pragma: no cover - network service
        from examples.hh.reward import serve

        serve(a.port, version_dir if os.listdir(version_dir) else None)


if __name__ == "__main__":
    main()
