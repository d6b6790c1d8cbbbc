"""Host-side frame-index stream with the semantics of the reference's tf.data pipeline
(DINCAE/__init__.py:399-401): ``from_generator -> repeat() -> shuffle(buffer) -> batch(B)``.

Only integers travel: the frame index, the index of the frame whose missing mask is laid
over it as added clouds (drawn with Python's ``random.randrange`` in generator order, exactly
like :214), and the sequence number of the generated frame (key of the device-side jitter
noise).  The pixel data never leaves the GPU.  TensorFlow's own shuffle RNG cannot be
reproduced; the *algorithm* is (fill a ``buffer``-slot reservoir from the sequential
stream, emit a uniformly chosen slot, refill it), driven by a numpy RandomState.
"""
import random as _pyrandom

import numpy as np


class TrainSampler:
    def __init__(self, ntime, batch_size, shuffle_buffer_size=45, seed=None, py_random=None,
                 train=True):
        self.ntime = int(ntime)
        self.batch = int(batch_size)
        self.bufsize = max(1, int(shuffle_buffer_size))
        self.rs = np.random.RandomState(seed)
        self.py_random = py_random if py_random is not None else _pyrandom
        self.train = train
        self._next_frame = 0          # position in the repeated sequential stream
        self._buffer = []

    def _generate(self):
        """Next element of the (infinitely repeated) generator: (frame, cloud frame, seq)."""
        seq = self._next_frame
        i = seq % self.ntime
        self._next_frame += 1
        cloud = self.py_random.randrange(0, self.ntime) if self.train else 0
        return (i, cloud, seq)

    def next_batch(self, n=None):
        """-> int32 arrays frame_idx[n], cloud_idx[n] and int64 seq[n] (n defaults to B)."""
        n = self.batch if n is None else n
        out = np.empty((n, 3), dtype=np.int64)
        for k in range(n):
            while len(self._buffer) < self.bufsize:
                self._buffer.append(self._generate())
            j = int(self.rs.randint(0, len(self._buffer)))
            out[k] = self._buffer[j]
            self._buffer[j] = self._generate()
        return out[:, 0].astype(np.int32), out[:, 1].astype(np.int32), out[:, 2].copy()


def sequential_batches(ntime, batch_size):
    """Test iterator: frames in order, last batch partial (:408-410, :680)."""
    for start in range(0, ntime, batch_size):
        yield np.arange(start, min(ntime, start + batch_size), dtype=np.int32)
