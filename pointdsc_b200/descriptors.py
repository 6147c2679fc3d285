"""Device-side descriptor front end: PLY reader, voxel down-sampling, normals, FPFH.

Mirrors the open3d 0.9 calls the reference makes before matching (misc/cal_fpfh.py:21-26, demo_registration.py:37-44):

    pcd = orig_pcd.voxel_down_sample(voxel_size)                                   -> voxel_down_sample(points, voxel_size)
    pcd.estimate_normals(KDTreeSearchParamHybrid(radius=2 * voxel, max_nn=30))     -> estimate_normals(keypts, 2 * voxel, 30)
    compute_fpfh_feature(pcd, KDTreeSearchParamHybrid(radius=5 * voxel, max_nn=100)) -> compute_fpfh(keypts, normals, 5 * voxel, 100)

open3d itself is not part of the reference tree or of this image: the kernels follow its published algorithms and are checked
against the CPU restatement under oracle/ (parity unpinned, see its header).  Everything runs on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _capi

_STATUS = {1: "more than 2^21 voxels along an axis, or a non-finite coordinate",
           2: "a neighbourhood holds more than 4096 points inside the search radius (the search is sized for down-sampled clouds)"}


def _ctx(points: torch.Tensor):
    if points.device.type != "cuda":
        raise _capi.PdscError("pointdsc_b200.descriptors runs on a B200 only: pass CUDA tensors (there is no CPU fallback)")
    if points.dim() != 2 or points.shape[1] != 3 or points.shape[0] < 1:
        raise ValueError(f"expected points [n,3] with n >= 1, got {tuple(points.shape)}")
    dev = points.device
    lib = _capi.load()
    engine = _capi.utility_engine(dev.index if dev.index is not None else torch.cuda.current_device())
    return dev, lib, engine, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _raise_status(status: int) -> None:
    if status:
        raise _capi.PdscError("; ".join(msg for bit, msg in _STATUS.items() if status & bit))


def read_ply(path: str) -> np.ndarray:
    """Vertex positions [n,3] float32 of a PLY file (host memory) — `np.asarray(o3d.io.read_point_cloud(path).points)`."""
    lib = _capi.load()
    n = C.c_int64(0)
    _capi.check(lib.pdsc_read_ply(path.encode(), None, 0, C.byref(n)))
    out = np.empty((n.value, 3), np.float32)
    _capi.check(lib.pdsc_read_ply(path.encode(), out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return out


@torch.no_grad()
def voxel_down_sample(points: torch.Tensor, voxel_size: float) -> torch.Tensor:
    """[n,3] -> [m,3] float32: the mean of the points of every occupied voxel, rows in ascending (ix, iy, iz) order."""
    dev, lib, engine, stream = _ctx(points)
    pts = points.to(torch.float32).contiguous()
    n = int(pts.shape[0])
    out = torch.empty(n, 3, dtype=torch.float32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)          # [count, status]
    scratch = torch.empty(int(lib.pdsc_voxel_down_sample_scratch_bytes(n)) + 8, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _capi.check(lib.pdsc_voxel_down_sample(engine, n, C.c_void_p(pts.data_ptr()), float(voxel_size), C.c_void_p(out.data_ptr()),
                                               C.c_void_p(meta.data_ptr()), C.c_void_p(meta.data_ptr() + 4),
                                               C.c_void_p((scratch.data_ptr() + 7) // 8 * 8), scratch.numel() - 8, stream))
    m, status = (int(v) for v in meta.tolist())                   # the one host read: it fixes the output shape
    _raise_status(status)
    return out[:m].clone()


def _search_buffers(lib, dev, m: int, max_nn: int):
    scratch = torch.empty(int(lib.pdsc_fpfh_scratch_bytes(m, max_nn)) + 8, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    return scratch, status


@torch.no_grad()
def estimate_normals(points: torch.Tensor, radius: float, max_nn: int = 30) -> torch.Tensor:
    """[m,3] -> unit normals [m,3] float64 (largest-magnitude component positive; (0,0,1) below three neighbours)."""
    dev, lib, engine, stream = _ctx(points)
    pts = points.to(torch.float32).contiguous()
    m = int(pts.shape[0])
    normals = torch.empty(m, 3, dtype=torch.float64, device=dev)
    scratch, status = _search_buffers(lib, dev, m, max_nn)
    with torch.cuda.device(dev):
        _capi.check(lib.pdsc_estimate_normals(engine, m, C.c_void_p(pts.data_ptr()), float(radius), int(max_nn),
                                              C.c_void_p(normals.data_ptr()), C.c_void_p(status.data_ptr()),
                                              C.c_void_p((scratch.data_ptr() + 7) // 8 * 8), scratch.numel() - 8, stream))
    _raise_status(int(status.item()))
    return normals


@torch.no_grad()
def compute_fpfh(points: torch.Tensor, normals: torch.Tensor, radius: float, max_nn: int = 100, normalise: bool = False) -> torch.Tensor:
    """[m,3], [m,3] -> FPFH [m,33] float64 (`np.array(fpfh.data).T`); normalise=True applies x / (||x|| + 1e-6) per row."""
    dev, lib, engine, stream = _ctx(points)
    pts = points.to(torch.float32).contiguous()
    m = int(pts.shape[0])
    if tuple(normals.shape) != (m, 3):
        raise ValueError(f"normals must be [{m},3], got {tuple(normals.shape)}")
    nrm = normals.to(device=dev, dtype=torch.float64).contiguous()
    out = torch.empty(m, 33, dtype=torch.float64, device=dev)
    scratch, status = _search_buffers(lib, dev, m, max_nn)
    with torch.cuda.device(dev):
        _capi.check(lib.pdsc_compute_fpfh(engine, m, C.c_void_p(pts.data_ptr()), C.c_void_p(nrm.data_ptr()), float(radius), int(max_nn),
                                          1 if normalise else 0, C.c_void_p(out.data_ptr()), C.c_void_p(status.data_ptr()),
                                          C.c_void_p((scratch.data_ptr() + 7) // 8 * 8), scratch.numel() - 8, stream))
    _raise_status(int(status.item()))
    return out


@torch.no_grad()
def fpfh_descriptors(points: torch.Tensor, voxel_size: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """misc/cal_fpfh.py:21-26 + the row normalisation of demo_registration.py:43: (key points [m,3] float32, FPFH [m,33] float64),
    ready for `pointdsc_b200.frontend.match`."""
    keypts = voxel_down_sample(points, voxel_size)
    normals = estimate_normals(keypts, 2.0 * voxel_size, 30)
    feat = compute_fpfh(keypts, normals, 5.0 * voxel_size, 100, normalise=True)
    return keypts, feat
