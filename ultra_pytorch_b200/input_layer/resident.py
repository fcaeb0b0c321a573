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
