"""Stand-in for astropy.table.Table: construction from columns, sort, column access (see ../README.md)."""
import numpy as np


class Table(object):
    def __init__(self, cols, names):
        self.colnames = list(names)
        self._c = {n: np.array(c, dtype=np.float64) for n, c in zip(names, cols)}

    def sort(self, key):
        order = np.argsort(self._c[key], kind='stable')
        for n in self.colnames:
            self._c[n] = self._c[n][order]

    def __getitem__(self, name):
        return self._c[name]

    def __len__(self):
        return len(self._c[self.colnames[0]])
