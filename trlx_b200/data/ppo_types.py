from trlx_b200.data.types import PPORLBatch, PPORLElement  # noqa: F401
