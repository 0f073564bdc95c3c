"""Stand-in for the `lru-dict` package (test infrastructure for the reference install)."""
from collections import OrderedDict


class LRU(OrderedDict):
    def __init__(self, size):
        super().__init__()
        self._size = int(size)

    def __getitem__(self, key):
        val = super().__getitem__(key)
        self.move_to_end(key)
        return val

    def __setitem__(self, key, val):
        super().__setitem__(key, val)
        self.move_to_end(key)
        while len(self) > self._size:
            self.popitem(last=False)
