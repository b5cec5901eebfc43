from .randomwalks import generate_random_walks  # noqa: F401
