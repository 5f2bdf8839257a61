"""Input pipeline for the reference's dataset files (/root/reference/input/flow_input.py:14-60).

The reference reads `../data/*.tfrecords` through a TF filename queue, parses each tf.train.Example
({'boundary': uint8[H*W] bytes, 'sflow': float32[H*W*2] bytes}, utils/createTFRecords.py:45-56) and batches through
`tf.train.shuffle_batch(capacity = min_queue_examples + 3*batch, min_after_dequeue = min_queue_examples)` with one
preprocessing thread.  Here the framing / proto parsing is native host code in libflownet.so (csrc/tfrecord.cpp, no
TensorFlow), the shuffle buffer has the same capacity rule, one background thread fills pinned host batches, and the
boundary stays uint8 all the way to the GPU (the first conv casts it; 4x fewer bytes than the reference's float32).
"""
from __future__ import annotations

import ctypes as C
import glob as _glob
import mmap
import os
import queue
import threading

import numpy as np
import torch

import flownet_native as native
from model import flags

FLAGS = flags.FLAGS
flags.DEFINE_integer('min_queue_examples', 1000, """ min examples to queue up""")


class TFRecordFile:
    """One memory-mapped .tfrecords file with its record index."""

    def __init__(self, path, check_crc=True):
        self.path = path
        self._f = open(path, 'rb')
        size = os.fstat(self._f.fileno()).st_size
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size else None
        self._buf = np.frombuffer(self._mm, dtype=np.uint8) if size else np.zeros(0, np.uint8)
        n = native.lib.fn_tfrecord_index(self._buf.ctypes.data, self._buf.size, None, None, 0, int(check_crc))
        if n < 0:
            raise ValueError("%s: truncated or corrupt TFRecord file (code %d)" % (path, n))
        self.offsets = np.zeros(n, np.int64)
        self.lengths = np.zeros(n, np.int64)
        if n:
            native.lib.fn_tfrecord_index(self._buf.ctypes.data, self._buf.size, self.offsets.ctypes.data, self.lengths.ctypes.data, n, 0)

    def __len__(self):
        return len(self.offsets)

    def read_batch_into(self, indices, boundary, sflow, nthreads=4):
        """Parse records `indices` into boundary[k], sflow[k] (contiguous numpy batches) with native host threads."""
        idx = np.asarray(indices, dtype=np.int64)
        offs = np.ascontiguousarray(self.offsets[idx])
        lens = np.ascontiguousarray(self.lengths[idx])
        rc = native.lib.fn_tfrecord_parse_flow_batch(self._buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(idx),
                                                      boundary.ctypes.data, boundary[0].nbytes, sflow.ctypes.data, sflow[0].nbytes,
                                                      int(nthreads))
        if rc != 0:
            raise ValueError("%s: a record is not a {'boundary','sflow'} example of %d / %d bytes" %
                             (self.path, boundary[0].nbytes, sflow[0].nbytes))

    def read_into(self, i, boundary, sflow):
        """Parse record i into boundary (uint8 [H,W,1]) and sflow (float32 [H,W,2]) -- numpy views, contiguous."""
        rc = native.lib.fn_tfrecord_parse_flow_example(self._buf.ctypes.data + int(self.offsets[i]), int(self.lengths[i]),
                                                        boundary.ctypes.data, boundary.nbytes, sflow.ctypes.data, sflow.nbytes)
        if rc != 0:
            raise ValueError("%s record %d: not a {'boundary','sflow'} example of %d / %d bytes" %
                             (self.path, i, boundary.nbytes, sflow.nbytes))


def read_data(path, shape=(128, 256)):
    """reference :14-30: yields (boundary uint8 [H,W,1], sflow float32 [H,W,2]) for every record of one file."""
    f = TFRecordFile(path)
    for i in range(len(f)):
        b = np.empty((shape[0], shape[1], 1), np.uint8)
        s = np.empty((shape[0], shape[1], 2), np.float32)
        f.read_into(i, b, s)
        yield b, s


def write_tfrecords(path, boundaries, sflows):
    """The writer side (utils/createTFRecords.py:45-56): boundaries uint8 [N,H,W(,1)], sflows float32 [N,H,W,2]."""
    boundaries = np.ascontiguousarray(boundaries, dtype=np.uint8)
    sflows = np.ascontiguousarray(sflows, dtype=np.float32)
    with open(path, 'wb') as f:
        for b, s in zip(boundaries, sflows):
            need = native.lib.fn_tfrecord_write_flow_example(None, 0, b.ctypes.data, b.nbytes, s.ctypes.data, s.nbytes)
            out = np.empty(need, np.uint8)
            got = native.lib.fn_tfrecord_write_flow_example(out.ctypes.data, need, b.ctypes.data, b.nbytes, s.ctypes.data, s.nbytes)
            if got != need:
                raise RuntimeError("fn_tfrecord_write_flow_example failed (%d)" % got)
            f.write(out.tobytes())


class FlowInputQueue:
    """`tf.train.shuffle_batch` over the records of some .tfrecords files, endlessly (the filename queue of the
    reference cycles forever): a buffer of parsed examples is kept between min_after_dequeue and capacity, every
    batch draws uniformly from it.  next() returns CUDA tensors (boundary uint8 [B,H,W,1], sflow float32 [B,H,W,2])
    whose host->device copies were issued from pinned memory on a side stream."""

    def __init__(self, files, batch_size, shape=(128, 256), min_after_dequeue=None, seed=0, device="cuda", shuffle=True,
                 prefetch=2, parse_threads=8, shard=(0, 1)):
        if not files:
            raise FileNotFoundError("no .tfrecords files")
        self.files = [TFRecordFile(p) for p in files]
        self.total = sum(len(f) for f in self.files)
        if self.total == 0:
            raise ValueError("the .tfrecords files hold no records")
        self.B, self.shape = batch_size, shape
        self.min_after = FLAGS.min_queue_examples if min_after_dequeue is None else min_after_dequeue
        self.capacity = self.min_after + 3 * batch_size                         # reference :38
        self.rng = np.random.default_rng(seed)
        self.shuffle = shuffle
        self.device = device
        self.parse_threads = parse_threads
        # the buffer holds record REFERENCES (file, index) -- the same sliding shuffle window over the record stream as
        # buffering parsed examples, but each example is parsed exactly once, straight into its pinned batch slot
        self._pool = []
        self._cursor = (0, 0)                                                   # (file, record)
        # data-parallel training: rank r of R takes every R-th record of the stream, so the ranks' shuffle windows are
        # disjoint (with one shared stream and different seeds they would draw from nearly the same window)
        self._shard = (int(shard[0]), max(1, int(shard[1])))
        if not 0 <= self._shard[0] < self._shard[1]:
            raise ValueError("shard = (rank, world) with 0 <= rank < world")
        if self.total < self._shard[1]:
            raise ValueError("fewer records than ranks")
        self._n = 0
        self._q = queue.Queue(maxsize=prefetch)
        self._stop = False
        self._copy_done = {}
        self._cuda = torch.cuda.is_available() and str(device).startswith("cuda")
        self._stream = torch.cuda.Stream() if self._cuda else None
        self._thread = threading.Thread(target=self._producer, daemon=True)     # num_preprocess_threads = 1 (reference :34)
        self._thread.start()

    def _next_ref(self):
        rank, world = self._shard
        while True:
            fi, ri = self._cursor
            while ri >= len(self.files[fi]):
                fi, ri = (fi + 1) % len(self.files), 0
            self._cursor = (fi, ri + 1)
            mine = self._n % world == rank
            self._n += 1
            if mine:
                return fi, ri

    def _producer(self):
        H, W = self.shape
        ring, turn = [], 0
        try:
            while not self._stop:
                want = self.capacity if self.shuffle else self.B
                while len(self._pool) < want:
                    self._pool.append(self._next_ref())
                # pinned staging buffers are recycled: queue depth + the batch being copied + the one being filled
                if len(ring) < self._q.maxsize + 3:
                    ring.append((torch.empty((self.B, H, W, 1), dtype=torch.uint8, pin_memory=self._cuda),
                                 torch.empty((self.B, H, W, 2), dtype=torch.float32, pin_memory=self._cuda)))
                hb, hs = ring[turn % len(ring)] if len(ring) == self._q.maxsize + 3 else ring[-1]
                turn += 1
                done = self._copy_done.pop(id(hb), None)
                if done is not None:
                    done.synchronize()                      # the upload that last read this buffer has finished
                nb, ns = hb.numpy(), hs.numpy()
                refs = []
                for i in range(self.B):
                    if self.shuffle:                        # swap-remove a uniformly drawn buffered example
                        j = int(self.rng.integers(0, len(self._pool)))
                        self._pool[j], self._pool[-1] = self._pool[-1], self._pool[j]
                        refs.append(self._pool.pop())
                    else:                                   # FIFO
                        refs.append(self._pool.pop(0))
                # parse runs of consecutive slots that come from the same file in one multi-threaded native call
                i = 0
                while i < self.B:
                    j = i
                    while j < self.B and refs[j][0] == refs[i][0]:
                        j += 1
                    self.files[refs[i][0]].read_batch_into([r for _, r in refs[i:j]], nb[i:j], ns[i:j], self.parse_threads)
                    i = j
                self._q.put((hb, hs))
        except BaseException as e:                          # surface parser errors in the consumer
            self._q.put(e)

    def next(self):
        item = self._q.get()
        if isinstance(item, BaseException):
            raise item
        hb, hs = item
        if not self._cuda:
            return hb, hs
        with torch.cuda.stream(self._stream):
            db = hb.to(self.device, non_blocking=True)
            ds = hs.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
            self._copy_done[id(hb)] = ev
        torch.cuda.current_stream().wait_stream(self._stream)
        db.record_stream(torch.cuda.current_stream())
        ds.record_stream(torch.cuda.current_stream())
        self._keep = (hb, hs)                               # pinned sources stay alive until the next batch
        return db, ds

    __next__ = next

    def __iter__(self):
        return self

    def close(self):
        self._stop = True
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass


def flow_inputs(batch_size, data_glob='../data/*.tfrecords', shape=(128, 256), seed=0, device="cuda", shard=(0, 1)):
    """reference :42-60.  Returns a FlowInputQueue; `boundarys, sflows = q.next()` per training step.
    shard = (rank, world): this process's share of the record stream in data-parallel training."""
    return FlowInputQueue(sorted(_glob.glob(data_glob)), batch_size, shape=shape, seed=seed, device=device, shard=shard)
