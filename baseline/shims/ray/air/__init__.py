class _Session:
    def get_checkpoint(self):
        return None

    def report(self, *a, **k):
        pass


session = _Session()


class Checkpoint:
    pass
