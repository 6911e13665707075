"""TEST INFRASTRUCTURE ONLY.

Everything under oracle/ exists to check the CUDA path, never to produce a result a user sees:

  refcpu/, refcpu.py            CPU restatement of the reference's shaders and fixed-function raster rules
  front_end_host/, front_end_host.py   host build of the GPU front end's per-contour core (csrc/front_end_core.h)
  front_end_ref.py              numpy restatement of the front end's per-curve arithmetic
  ref/                          recipe that compiles the reference's own front end and core runtime in place
                                (outputs in oracle/_ref/, git-ignored)

Only tests/ (including the developer scripts in tests/tools/), __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py import, link or execute it. The product --
rive-runtime_b200/ and include/ -- never does: it fails loudly when librivecuda.so or a CUDA
device is missing, and has no CPU fallback.
"""
