"""A minimal in-memory stand-in for the part of the h5py API an MDTraj-style writer uses (File as a context
manager, create_group, create_dataset with data / dtype / chunks / compression, attrs, item access by
path).  h5py is not installed in this image (the reference guards it the same way, PyCD/core.py:28-32), so
tests/test_host.py::test_hdf5_layout_through_the_h5py_api runs pycd_b200.hdf5_io against this recorder and
checks the layout of PyCD/hdf5_io.py:59-101 / docs/hdf5_format.md:31-49: names, shapes, dtypes, units and
the unit conversions.  It does NOT prove byte-level HDF5 compatibility; the real-h5py test
(test_hdf5_writer_layout) does that wherever h5py exists."""
import numpy as np

FILES = {}


class Dataset:
    def __init__(self, data, dtype=None, chunks=None, compression=None, maxshape=None):
        self._a = np.array(data, dtype=dtype) if dtype is not None else np.array(data)
        self.chunks, self.compression, self.maxshape = chunks, compression, maxshape
        self.attrs = {}

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)

    def __getitem__(self, key):
        return self._a[key]

    def __setitem__(self, key, value):
        self._a[key] = value

    def resize(self, size, axis=None):
        shape = list(self._a.shape)
        if axis is None:
            shape = list(size)
        else:
            shape[axis] = size
        new = np.zeros(shape, dtype=self._a.dtype)
        sl = tuple(slice(0, min(a, b)) for a, b in zip(self._a.shape, shape))
        new[sl] = self._a[sl]
        self._a = new


class Group:
    def __init__(self):
        self._items = {}
        self.attrs = {}

    def create_group(self, name):
        g = Group()
        self._items[name] = g
        return g

    def create_dataset(self, name, shape=None, dtype=None, data=None, chunks=None, compression=None,
                       maxshape=None, **_):
        if data is None:
            data = np.zeros(shape, dtype=dtype)
        d = Dataset(data, dtype, chunks, compression, maxshape)
        self._items[name] = d
        return d

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in path.strip('/').split('/'):
            node = node._items[part]
        return node

    def __delitem__(self, name):
        del self._items[name]

    def keys(self):
        return self._items.keys()


class File(Group):
    def __init__(self, path, mode='r'):
        super().__init__()
        self.path, self.mode = str(path), mode
        if mode in ('r', 'r+', 'a') and self.path in FILES:
            self._items = FILES[self.path]._items
        FILES[self.path] = self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass
