"""rive-runtime_b200: a B200-native (sm_100a CUDA) backend for Rive's GPU vector
renderer, behind the reference's RenderContextImpl boundary.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/rivecuda.h) -> librivecuda.so
             and the ABI call recorder -> librivecuda_trace.so
  host/      C++ RenderContextCUDAImpl (the reference-facing plugin) + scene player
  abi.py     ctypes binding of the C ABI (device memory stays behind the ABI)
  trace.py   reader for recorded flush traces (the reference front end's output)
  replay.py  drives a trace through the ABI: the public Python entry point
  sharding.py  multi-GPU partitioning (frames per GPU; screen bands + gather)

The product path has NO CPU fallback: importing `abi` raises if librivecuda.so
is missing, and rivecuda_create() fails without a CUDA device.
"""
__version__ = "0.1.0"
