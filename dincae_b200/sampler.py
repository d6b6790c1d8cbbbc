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

    def _generate(self, n):
        """Next ``n`` elements of the (infinitely repeated) generator, each packed into one
        integer ``seq * ntime + cloud`` (frame = seq % ntime); the added-cloud frames are drawn
        one by one in generator order (:214)."""
        seq0 = self._next_frame
        self._next_frame += n
        T = self.ntime
        if self.train:
            # random.randrange(0, T) spelled out (CPython's _randbelow_with_getrandbits: draw
            # bit_length(T) bits, reject values >= T) -- the same stream at a quarter of the call
            # cost; tests/test_host_logic.py checks it against randrange itself
            gb, kb = self.py_random.getrandbits, T.bit_length()
            out = []
            for k in range(n):
                r = gb(kb)
                while r >= T:
                    r = gb(kb)
                out.append((seq0 + k) * T + r)
            return out
        return [(seq0 + k) * T for k in range(n)]

    def next_batch(self, n=None):
        """-> int32 arrays frame_idx[n], cloud_idx[n] and int64 seq[n] (n defaults to B).

        Plain-Python inner loop on a list of packed integers (no numpy scalar indexing, no
        tuples): ~50 us for a batch of 50, so that one host thread can produce eight GPUs'
        worth of indices per step."""
        n = self.batch if n is None else n
        buf = self._buffer
        if len(buf) < self.bufsize:
            buf.extend(self._generate(self.bufsize - len(buf)))
        # the slot picks do not depend on the buffer's content: one vectorised draw gives the
        # same stream as n scalar randint calls
        picks = self.rs.randint(0, self.bufsize, size=n).tolist()
        fresh = self._generate(n)
        out = [0] * n
        for k in range(n):
            j = picks[k]
            out[k] = buf[j]
            buf[j] = fresh[k]
        out = np.array(out, dtype=np.int64)
        seq = out // self.ntime
        return ((seq % self.ntime).astype(np.int32), (out - seq * self.ntime).astype(np.int32), seq)


def sequential_batches(ntime, batch_size):
    """Test iterator: frames in order, last batch partial (:408-410, :680)."""
    for start in range(0, ntime, batch_size):
        yield np.arange(start, min(ntime, start + batch_size), dtype=np.int32)
