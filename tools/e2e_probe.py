import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import bench, argparse
from pointdsc_b200 import PointDSC
a = argparse.Namespace(n=1000, batch=256, dataset="3dmatch", precision="fp16x3")
m = PointDSC(num_layers=12, precision="fp16x3"); m.load_state_dict(bench.load_snapshot("3dmatch"), strict=False); m = m.cuda().eval()
host = bench.make_inputs(a, 0, 1)
pin = {k: host[k].pin_memory() for k in ("corr_pos", "src_keypts", "tgt_keypts")}
dev = {k: host[k].cuda() for k in pin}
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("device forward      ms", t(lambda: m.run(dev["corr_pos"], dev["src_keypts"], dev["tgt_keypts"])))
print("host   forward      ms", t(lambda: m.run(pin["corr_pos"], pin["src_keypts"], pin["tgt_keypts"])))
print("H2D of the 3 inputs ms", t(lambda: [pin[k].cuda(non_blocking=True) for k in pin]))
x = torch.empty(256, 1000, dtype=torch.float32, device="cuda"); hp = torch.empty(256, 1000).pin_memory()
print("D2H labels          ms", t(lambda: hp.copy_(x, non_blocking=True)))
# finer: device path with a synchronize per call, and GPU-side time of the host call
print("device fwd + sync   ms", t(lambda: (m.run(dev["corr_pos"], dev["src_keypts"], dev["tgt_keypts"]), torch.cuda.synchronize())))
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
gpu, wall = [], []
for _ in range(5):
    t0 = time.perf_counter(); e0.record(); m.run(pin["corr_pos"], pin["src_keypts"], pin["tgt_keypts"]); e1.record(); torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
print("host call: wall ms", np.round(wall, 3), "gpu-side ms", np.round(gpu, 3))
import ctypes as C
from pointdsc_b200 import _capi
lib = _capi.load()
B, N = 256, 1000
tr = torch.empty(B, 4, 4).pin_memory(); lb = torch.empty(B, N).pin_memory()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def raw():
    _capi.check(lib.pdsc_forward_host(m._engine, B, N, C.c_void_p(pin["corr_pos"].data_ptr()), C.c_void_p(pin["src_keypts"].data_ptr()),
                                      C.c_void_p(pin["tgt_keypts"].data_ptr()), C.c_void_p(tr.data_ptr()), C.c_void_p(lb.data_ptr()), st))
print("raw pdsc_forward_host ms", t(raw))
