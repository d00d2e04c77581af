// libsdtree.so -- the SD-tree hot path of the Mitsuba 3 "Practical Path Guiding" lab as
// hand-written sm_100a CUDA kernels behind the C ABI of include/sdtree.h.
//   sdt_query.inl   locate / sample / pdf / guided bounce / MIS   (wavefront kernels)
//   sdt_splat.inl   radiance-record splat + bottom-up sweeps
//   sdt_refine.inl  device-side per-iteration refine (scan-built tree rebuild)
//   sdt_io.inl      lifetime, npz-schema upload/download, thresholds, L2 probe
//   sdt_nccl.inl    one all-reduce per training iteration
// Build: practical_path_guiding_lab_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a
// -fmad=false; IEEE div/sqrt -- no -use_fast_math: integer outputs must be bit-exact).
#include <math.h>
#include <stdio.h>

#include "sdt_impl.h"

#include "sdt_query.inl"
#include "sdt_splat.inl"
#include "sdt_refine.inl"
#include "sdt_nccl.inl"
#include "sdt_io.inl"
