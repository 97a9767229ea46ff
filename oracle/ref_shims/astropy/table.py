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


def _read_ascii(path):
    """Table.read(path, format='ascii') for whitespace-separated numeric tables with a header line."""
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.strip() and not l.startswith('#')]
    names = lines[0].split()
    data = np.array([[float(x) for x in l.split()] for l in lines[1:]], dtype=np.float64)
    return Table([data[:, i] for i in range(len(names))], names)


Table.read = staticmethod(lambda path, format='ascii': _read_ascii(path))
