"""An in-memory object with the part of the h5py File / Group / Dataset API that the reference's `Module.save / load`,
`Container.save / load` and `Optimizer.save / load` call (Modules/Module.py:179-283, Containers/Container.py:138-203,
Optimizers/Optimizer.py:202-246).  h5py is not in this image; the reference accepts an already-open file object (`ensureHdf` returns
anything that is not a path / bytes as it is), so the checkpoint path above the seam can be exercised end to end -- what reaches the
backend is `GPUArray.get()` and `GPUArray.set()` only.  The on-disk HDF5 encoding itself is h5py's business, not the backend's."""
import numpy as np


class Dataset:
	def __init__(self, value):
		self.value = np.array(value)

	def __getitem__(self, key):
		return self.value[key]

	def __array__(self, dtype=None, copy=None):
		return self.value if dtype is None else self.value.astype(dtype)

	@property
	def shape(self):
		return self.value.shape


class Group:
	def __init__(self):
		self.members = {}

	def require_group(self, name):
		return self.members.setdefault(name, Group())

	def create_group(self, name):
		if name in self.members:
			raise ValueError("group %s exists" % name)
		return self.require_group(name)

	def create_dataset(self, name, shape=None, dtype=None, data=None, compression=None):
		if name in self.members:
			raise ValueError("dataset %s exists" % name)
		self.members[name] = Dataset(data)
		return self.members[name]

	def __setitem__(self, name, value):
		self.members[name] = Dataset(value)

	def __getitem__(self, name):
		return self.members[name]

	def __contains__(self, name):
		return name in self.members

	def items(self):
		return self.members.items()

	def keys(self):
		return self.members.keys()


class File(Group):
	def __init__(self):
		super().__init__()
		self.closed = 0

	def flush(self):
		pass

	def close(self):
		self.closed += 1
