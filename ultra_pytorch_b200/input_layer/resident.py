"""Device-resident data sets (batch assembly without moving feature rows).

The reference copies the feature rows of every sampled list into a fresh `letor_features` array per batch
(click_simulation_feed.py:66-99), the algorithm gathers and converts them again and ships them to the device
(base_algorithm.py:148-152, DNN.py:72-75): 11 MB of float64 per 256-query batch at config 2.  A B200 has 180 GB of HBM:
the WHOLE feature matrix of MSLR-WEB30K (3.7 M documents x 136 features = 2 GB in fp32) fits hundreds of times.

`ResidentFeatures` is an ndarray VIEW (no copy) of the data set's complete feature matrix.  A feed that emits it as
`input_feed["letor_features"]` uses GLOBAL row ids in `docid_input{l}` (PAD id = number of rows), which is still a valid
reference-format feed - the reference's own algorithms run on it unchanged, they just `np.take` from a bigger matrix.
The B200 engine recognises the type, uploads the matrix ONCE (converted to fp32, plus the zero PAD row) and afterwards a
training step moves only the ids and labels (8 L B bytes: 80 KB at config 2); the kernels gather rows by id as before.
"""
import numpy as np


class ResidentFeatures(np.ndarray):
    def __new__(cls, array):
        a = np.asarray(array)
        if a.ndim != 2 or a.dtype != np.float64 or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a, dtype=np.float64)
        return a.view(cls)

    def resident_key(self):
        """Identity of the underlying matrix (address + shape): the engine re-uploads when it changes."""
        return (self.ctypes.data, self.shape)


class DeviceFeed(dict):
    """An `input_feed` whose doc ids and labels were produced ON THE DEVICE (csrc/sampling.cu: click_batch_kernel).

    A B200 learning algorithm consumes `.docid` (i32 [L, B]) / `.labels` (f32 [B, L]) in place - no host work, no
    H2D copy.  For everybody else it still behaves like the reference's dict: the first access to a
    `docid_input{l}` / `label{l}` key copies the batch to the host once and materialises the numpy arrays
    (`letor_features` is the ResidentFeatures view, available without any copy)."""

    def __init__(self, model, features, docid, labels, query_idx, n_rows):
        dict.__init__(self)
        self.model = model
        self.docid, self.labels, self.query_idx = docid, labels, query_idx     # device tensors
        self.L, self.B = docid.shape
        self.n_rows = n_rows
        dict.__setitem__(self, model.letor_features_name, features)
        self._names = set(model.docid_inputs_name[:self.L]) | set(model.labels_name[:self.L])
        self._materialised = False

    def materialise(self):
        if not self._materialised:
            d = self.docid.cpu().numpy().astype(np.float32)           # [L, B]
            y = np.ascontiguousarray(self.labels.cpu().numpy().T)     # [L, B]
            for l in range(self.L):
                dict.__setitem__(self, self.model.docid_inputs_name[l], d[l])
                dict.__setitem__(self, self.model.labels_name[l], y[l])
            self._materialised = True
        return self

    def __getitem__(self, key):
        if not self._materialised and key in self._names:
            self.materialise()
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return key in self._names or dict.__contains__(self, key)

    def keys(self):
        return self.materialise() and dict.keys(self)

    def items(self):
        return self.materialise() and dict.items(self)

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return 1 + len(self._names)
