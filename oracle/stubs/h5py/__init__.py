"""Test-infrastructure stub for h5py (absent from this image).

Only used by oracle/ref_harness.py so that the unmodified reference package
under /root/reference imports and runs.  An in-memory store keyed by file name:
enough for the estimator writer (`fh5[dset] = data`) and nothing else.
This is NOT part of the product path.
"""
import numpy

_STORE = {}


class _Group(dict):
    def create_dataset(self, name, shape=None, dtype=None, data=None):
        if data is None:
            data = numpy.zeros(shape, dtype=dtype)
        self[name] = numpy.array(data)
        return self[name]


class File(object):
    def __init__(self, name, mode='r', **kwargs):
        self.name = name
        if mode == 'w' or name not in _STORE:
            if mode == 'r':
                raise OSError("stub h5py: no such file %s" % name)
            _STORE[name] = _Group()
        self._g = _STORE[name]

    def __enter__(self):
        return self

    def __exit__(self, *args):
        return False

    def __setitem__(self, key, value):
        self._g[key] = value

    def __getitem__(self, key):
        return self._g[key]

    def __contains__(self, key):
        return key in self._g

    def keys(self):
        return self._g.keys()

    def create_dataset(self, *a, **k):
        return self._g.create_dataset(*a, **k)

    def close(self):
        pass
