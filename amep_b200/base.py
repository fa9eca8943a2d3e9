"""``BaseEvaluation`` look-alike (amep/base.py:1808-1960): dict-style access to
the result properties, a ``name`` and ``save``.  The reference's HDF5 database
layer (amep/base.py:2048-2184) is storage and out of scope; ``save`` writes the
same layout (one group per evaluation, one dataset per property) when h5py is
installed and says so when it is not."""
from __future__ import annotations

import inspect
import os

import numpy as np


class BaseEvaluation:
    def __init__(self) -> None:
        self.__name = ""

    def __getitem__(self, key: str):
        return getattr(self, key)

    def keys(self) -> list[str]:
        return [name for (name, _) in inspect.getmembers(type(self), lambda x: isinstance(x, property))]

    def values(self):
        return [self.__getitem__(key) for key in self.keys()]

    def items(self):
        return [i for i in zip(self.keys(), self.values())]

    @property
    def name(self) -> str:
        return self.__name

    @name.setter
    def name(self, x: str) -> None:
        if type(x) == str:
            self.__name = x

    def save(self, path: str, backup: bool = True, database: bool = False, name: str | None = None) -> None:
        """Store the evaluation in an HDF5 file (needs h5py, like the reference)."""
        try:
            import h5py
        except ImportError as exc:
            raise ImportError("BaseEvaluation.save needs h5py, which is not installed") from exc
        directory, filename = os.path.split(path)
        if filename == "":
            filename = self.name + ".h5"
        if not filename.endswith(".h5"):
            raise ValueError(f"Wrong file extension {os.path.splitext(filename)[1]}. Use '.h5'.")
        if directory and not os.path.isdir(directory):
            raise ValueError(f"The directory {directory} does not exist.")
        current = os.path.join(directory, filename)
        if name is None:
            name = self.name
        mode = "a" if (database and os.path.exists(current)) else "w"
        if mode == "w" and backup and os.path.exists(current):
            from datetime import datetime
            import shutil
            stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
            shutil.copy(current, os.path.join(directory, "#" + filename[:-3] + "_" + stamp + ".h5"))
        with h5py.File(current, mode) as root:
            if name in root:
                del root[name]
            grp = root.create_group(name)
            grp.attrs["type"] = type(self).__name__
            for key, val in self.items():
                if key == "name":
                    continue
                arr = np.asarray(val)
                if arr.dtype.kind == "f":
                    arr = arr.astype(np.float32)  # amep/base.py:2145-2157 stores float32
                grp.create_dataset(key, data=arr)
