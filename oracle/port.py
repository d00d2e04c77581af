"""ORACLE / CPU BASELINE (test infrastructure only).  ctypes face of oracle/sdtree_port.c: the
reference's per-vertex SD-tree operations in C + OpenMP on the reference's own SoA layout (the
23 npz arrays).  Used by bench.py's cpu_baseline / --impl reference legs and checked against the
numpy oracle in tests/test_oracle_port.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsdtree_port.so")


class _Tree(C.Structure):
    _fields_ = [("n_kd", C.c_uint32), ("kd_bmin", C.c_void_p), ("kd_bmax", C.c_void_p), ("kd_leaf", C.c_void_p),
                ("kd_root", C.c_void_p), ("kd_left", C.c_void_p), ("kd_right", C.c_void_p), ("kd_count", C.c_void_p),
                ("n_q", C.c_uint32), ("q_rootnode", C.c_void_p), ("q_bmin", C.c_void_p), ("q_bmax", C.c_void_p),
                ("q_leaf", C.c_void_p), ("q_child", C.c_void_p * 4), ("q_energy", C.c_void_p)]


def build():
    src = os.path.join(HERE, "sdtree_port.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


class PortTree:
    """d: mapping with the 23 npz keys (e.g. oracle KDTree.to_arrays() or SDTree.download())"""

    def __init__(self, d):
        self.lib = C.CDLL(build())
        self.lib.port_max_threads.restype = C.c_int
        f32 = lambda k, sh=None: np.ascontiguousarray(np.asarray(d[k], np.float32).reshape(sh) if sh else np.asarray(d[k], np.float32))
        u32 = lambda k: np.ascontiguousarray(np.asarray(d[k]).astype(np.uint32))
        u8 = lambda k: np.ascontiguousarray(np.asarray(d[k]).astype(np.uint8))
        self.a = dict(kd_bmin=f32('kdtree_bbox_min', (-1, 3)), kd_bmax=f32('kdtree_bbox_max', (-1, 3)), kd_leaf=u8('kdtree_isLeaf'),
                      kd_root=u32('kdtree_quadTreeRootIndex'), kd_left=u32('kdtree_child_left_index'),
                      kd_right=u32('kdtree_child_right_index'), kd_count=f32('kdtree_vertCount').copy(),
                      q_rootnode=u32('quadtree_rootNodeIndex'), q_bmin=f32('quadtree_bbox_min', (-1, 2)),
                      q_bmax=f32('quadtree_bbox_max', (-1, 2)), q_leaf=u8('quadtree_isLeaf'),
                      q_energy=f32('quadtree_irradiance').copy())
        self.qc = [u32(f'quadtree_child_{k}_index') for k in (1, 2, 3, 4)]
        t = _Tree()
        t.n_kd = self.a['kd_leaf'].shape[0]
        t.n_q = self.a['q_leaf'].shape[0]
        for k, v in self.a.items():
            setattr(t, k, v.ctypes.data)
        for k in range(4):
            t.q_child[k] = self.qc[k].ctypes.data
        self.t = t

    def threads(self):
        return int(self.lib.port_max_threads())

    def set_threads(self, n):
        """OpenMP team size for the following calls (torchrun sets OMP_NUM_THREADS=1 in every worker)"""
        self.lib.port_set_threads(C.c_int(int(n)))
        return self.threads()

    def sample(self, pos, seed, lane_offset=0, active=None, debug=False):
        pos = np.ascontiguousarray(pos, np.float32)
        n = pos.shape[0]
        d = np.empty((n, 3), np.float32)
        p = np.empty(n, np.float32)
        dbg = np.empty((n, 4), np.uint32) if debug else None
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        self.lib.port_sample(C.byref(self.t), C.c_uint32(n), C.c_void_p(pos.ctypes.data), C.c_void_p(None if act is None else act.ctypes.data),
                             C.c_uint32(seed), C.c_uint32(lane_offset), C.c_void_p(d.ctypes.data), C.c_void_p(p.ctypes.data),
                             C.c_void_p(None if dbg is None else dbg.ctypes.data))
        return (d, p, dbg) if debug else (d, p)

    def pdf(self, pos, dirs, active=None, debug=False):
        pos = np.ascontiguousarray(pos, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = pos.shape[0]
        p = np.empty(n, np.float32)
        dbg = np.empty((n, 3), np.uint32) if debug else None
        act = None if active is None else np.ascontiguousarray(active, np.uint8)
        self.lib.port_pdf(C.byref(self.t), C.c_uint32(n), C.c_void_p(pos.ctypes.data), C.c_void_p(dirs.ctypes.data),
                          C.c_void_p(None if act is None else act.ctypes.data), C.c_void_p(p.ctypes.data),
                          C.c_void_p(None if dbg is None else dbg.ctypes.data))
        return (p, dbg) if debug else p

    def splat(self, pos, dir2, radiance, wo_pdf):
        """accumulates into self.a['kd_count'] / self.a['q_energy'] (every visited node, like the reference)"""
        pos = np.ascontiguousarray(pos, np.float32)
        dir2 = np.ascontiguousarray(dir2, np.float32)
        radiance = np.ascontiguousarray(radiance, np.float32)
        wo_pdf = np.ascontiguousarray(wo_pdf, np.float32)
        self.lib.port_splat(C.byref(self.t), C.c_uint32(pos.shape[0]), C.c_void_p(pos.ctypes.data), C.c_void_p(dir2.ctypes.data),
                            C.c_void_p(radiance.ctypes.data), C.c_void_p(wo_pdf.ctypes.data))
