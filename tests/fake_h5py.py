"""
A recording stand-in for h5py (absent from this image) -- TEST INFRASTRUCTURE.

It keeps "files" in memory as nested dicts and LOGS every call that shapes the file (create_group, create_dataset with its
keyword arguments, scalar assignment, attribute assignment with the attribute's Python type).  The log of the
reference's own `mca_out_ng.dump` (er3t/rtm/mca/mca_out.py:209-233), recorded by tests/golden/make_h5_golden.py in the
build container, is committed as tests/golden/h5_dump_calls.json; tests/test_host_api.py replays this repo's `dump`
against the same fake and compares the logs: same groups, same dataset options, same attributes, same types.
"""

import numpy as np

FILES = {}
LOG = []


def _describe(v):
    if isinstance(v, np.ndarray):
        return ['ndarray', str(v.dtype), list(v.shape)]
    if isinstance(v, (bytes, np.bytes_)):
        return ['bytes', bytes(v).decode()]
    if isinstance(v, (str, np.str_)):
        return ['str', str(v)]
    if isinstance(v, (bool, np.bool_)):
        return ['bool', bool(v)]
    if isinstance(v, (int, np.integer)):
        return ['int', int(v)]
    if isinstance(v, (float, np.floating)):
        return ['float', float(v)]
    return [type(v).__name__, str(v)]


class _Attrs(dict):
    def __init__(self, path):
        super().__init__()
        self._path = path

    def __setitem__(self, k, v):
        LOG.append(['attr', self._path, k, _describe(v)])
        super().__setitem__(k, v)


class Dataset:
    def __init__(self, path, data):
        self._data = np.asarray(data)
        self.attrs = _Attrs(path)

    def __getitem__(self, key):
        return self._data[key]

    @property
    def shape(self):
        return self._data.shape


class Group:
    def __init__(self, path):
        self._path = path
        self._items = {}
        self.attrs = _Attrs(path)

    def create_group(self, name):
        LOG.append(['create_group', self._path, name])
        g = Group(self._path + '/' + name)
        self._items[name] = g
        return g

    def create_dataset(self, name, data=None, **kw):
        LOG.append(['create_dataset', self._path, name, _describe(np.asarray(data)), {k: (v if not isinstance(v, np.generic) else v.item()) for k, v in sorted(kw.items())}])
        d = Dataset(self._path + '/' + name, data)
        self._items[name] = d
        return d

    def __setitem__(self, name, value):
        LOG.append(['setitem', self._path, name, _describe(value)])
        self._items[name] = Dataset(self._path + '/' + name, value)

    def __getitem__(self, name):
        node = self
        for part in name.strip('/').split('/'):
            node = node._items[part]
        return node

    def keys(self):
        return self._items.keys()


class File(Group):
    def __init__(self, fname, mode='r'):
        if mode == 'w':
            super().__init__('')
            FILES[str(fname)] = self
        else:
            src = FILES[str(fname)]
            self.__dict__ = src.__dict__

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def reset():
    FILES.clear()
    del LOG[:]
