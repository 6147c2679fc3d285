#!/usr/bin/env python
"""Developer tool: where the time of the streaming loop (model.forward_stream, pdsc_forward_host_submit / _wait) goes at the bench
size.  Prints per-step host durations of submit / wait through the raw C ABI, and the loop's throughput with the results kept or
dropped, on the default (legacy) stream and on a side stream."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointdsc_b200 import PointDSC, _capi

B, N, K = 256, 1000, 10
m = PointDSC(num_layers=12, precision="fp16x3")
m.load_state_dict(bench.load_snapshot("3dmatch"), strict=False)
m = m.cuda().eval()
host = bench.make_inputs(N, B, "3dmatch", 0)
pin = {k: host[k].pin_memory() for k in ("corr_pos", "src_keypts", "tgt_keypts")}
hd = dict(pin, testing=True)
dev = [host[k].cuda() for k in ("corr_pos", "src_keypts", "tgt_keypts")]
for _ in range(3):
    m.run(*dev)
torch.cuda.synchronize()


def loop(keep, steps=K):
    for _ in m.forward_stream(hd for _ in range(3)):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kept = []
    for o in m.forward_stream(hd for _ in range(steps)):
        if keep:
            kept.append(o)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return B * steps / dt, (kept[0]["final_labels"].is_pinned() if kept else None)


print("forward_stream, results dropped :", loop(False))
print("forward_stream, results kept    :", loop(True))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    print("forward_stream on a side stream, dropped / kept:", loop(False), loop(True))

# raw C ABI with preallocated pinned result buffers: host time of every submit and wait
lib = m._ensure_engine()
outs = [(torch.empty(B, 4, 4).pin_memory(), torch.empty(B, N).pin_memory()) for _ in range(2)]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def submit(i):
    slot = C.c_int32(-1)
    _capi.check(lib.pdsc_forward_host_submit(m._engine, B, N, C.c_void_p(pin["corr_pos"].data_ptr()), C.c_void_p(pin["src_keypts"].data_ptr()),
                                             C.c_void_p(pin["tgt_keypts"].data_ptr()), C.c_void_p(outs[i & 1][0].data_ptr()),
                                             C.c_void_p(outs[i & 1][1].data_ptr()), st, C.byref(slot)))
    return slot.value


torch.cuda.synchronize()
rows = []
t_all = time.perf_counter()
pend = None
for i in range(K):
    t0 = time.perf_counter()
    s = submit(i)
    t1 = time.perf_counter()
    if pend is not None:
        _capi.check(lib.pdsc_forward_host_wait(m._engine, pend))
    t2 = time.perf_counter()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    pend = s
_capi.check(lib.pdsc_forward_host_wait(m._engine, pend))
torch.cuda.synchronize()
dt = time.perf_counter() - t_all
print("raw submit/wait: sets/s", B * K / dt, " per step (submit ms, wait ms):", [(round(a, 2), round(b, 2)) for a, b in rows])
