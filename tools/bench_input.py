"""Measures the dataset path: native TFRecord/Example parse rate (single thread, CRC on and off),
and images/s delivered by the shuffle queue into GPU memory."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "steady-state-flow-with-neural-nets_b200"))
sys.path.insert(0, ROOT)
import input.flow_input as fi  # noqa: E402
from utils.synthetic import synth_boundary, synth_sflow  # noqa: E402

N = int(os.environ.get("N", "512"))
d = tempfile.mkdtemp(dir=os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None)
path = os.path.join(d, "train.tfrecords")
b = synth_boundary(N, 128, 256, seed=0)
s = synth_sflow(b)
t0 = time.time(); fi.write_tfrecords(path, b, s); t_write = time.time() - t0
size = os.path.getsize(path)
out = {"records": N, "file_MB": size / 1e6, "write_MBps": size / 1e6 / t_write}
for crc in (1, 0):
    t0 = time.time(); f = fi.TFRecordFile(path, check_crc=bool(crc)); t_idx = time.time() - t0
    bb = np.empty((128, 256, 1), np.uint8); ss = np.empty((128, 256, 2), np.float32)
    t0 = time.time()
    for i in range(len(f)):
        f.read_into(i, bb, ss)
    t_parse = time.time() - t0
    out["index_crc%d_MBps" % crc] = size / 1e6 / t_idx
    out["parse_records_per_s"] = N / t_parse
    out["parse_MBps"] = size / 1e6 / t_parse
f = fi.TFRecordFile(path, check_crc=False)
nb = np.empty((min(N, 64), 128, 256, 1), np.uint8); ns = np.empty((min(N, 64), 128, 256, 2), np.float32)
for th in (1, 4, 8, 16):
    t0 = time.time(); reps = 0
    while time.time() - t0 < 0.5:
        f.read_batch_into(np.arange(len(nb)) % len(f), nb, ns, th); reps += 1
    out["batch_parse_records_per_s_%dthreads" % th] = reps * len(nb) / (time.time() - t0)
if torch.cuda.is_available():
    q = fi.FlowInputQueue([path], 64, min_after_dequeue=256, device="cuda")
    for _ in range(3):
        q.next()
    torch.cuda.synchronize(); t0 = time.time()
    steps = 20
    for _ in range(steps):
        db, ds = q.next()
    torch.cuda.synchronize()
    out["queue_images_per_s_to_gpu"] = steps * 64 / (time.time() - t0)
    q.close()
os.remove(path); os.rmdir(d)
print(json.dumps(out))
