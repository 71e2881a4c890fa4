"""In-memory stand-in for the h5py package (absent from this image; there is no network to install it).

Used ONLY by tests/golden/make_golden_io.py so that the reference's dataset.py / util.py /
est_lands_csv.py import and run UNMODIFIED in the authoring container: a "file" is a pickled dict of
numpy arrays / scalars / strings keyed by the HDF5 path.  Never imported by tests, bench or the package."""
import pickle

import numpy as np


class _Dataset:
    def __init__(self, store, key):
        self._s, self._k = store, key

    def __getitem__(self, idx):
        v = self._s[self._k]
        if isinstance(idx, tuple) and len(idx) == 0:
            return v
        return np.asarray(v)[idx]

    def __setitem__(self, idx, val):
        self._s[self._k][idx] = val

    @property
    def shape(self):
        return np.asarray(self._s[self._k]).shape


class File:
    def __init__(self, path, mode="r"):
        self.path, self.mode = path, mode
        if mode == "r":
            with open(path, "rb") as f:
                self.store = pickle.load(f)
        else:
            self.store = {}

    def __getitem__(self, key):
        return _Dataset(self.store, key)

    def __contains__(self, key):
        return key in self.store

    def create_dataset(self, name, shape, dtype="f4", **_unused):
        self.store[name] = np.zeros(shape, dtype=dtype)
        return _Dataset(self.store, name)

    def close(self):
        if self.mode != "r":
            with open(self.path, "wb") as f:
                pickle.dump(self.store, f)
