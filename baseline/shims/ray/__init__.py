def is_initialized():
    return False
