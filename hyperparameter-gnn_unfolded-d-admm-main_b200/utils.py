"""Import shim: ``unfolded_train_new.py`` does ``import utils`` (reference unfolded_train_new.py:3) without
using it.  The reference's ``utils.py`` holds the legacy MNIST / tensorboardX training helpers, which
are out of scope of the B200 hot path."""


class IOStream:
    """Minimal log-file helper with the reference's interface (utils.py IOStream)."""

    def __init__(self, path):
        self.f = open(path, "a")

    def cprint(self, text):
        print(text)
        self.f.write(text + "\n")
        self.f.flush()

    def close(self):
        self.f.close()
