"""Reference-side binding: run the UNMODIFIED reference host code (synchrad/calc.py) on libsynchrad_b200.so.

The reference's device boundary is PyOpenCL: `cl.Program(ctx, src).build()` and the per-particle kernel calls of
`_process_track` (calc.py:324-353).  This directory holds import-compatible stand-ins for `pyopencl` (and for
`mako.template`, which the reference only uses to substitute two tokens into its kernel sources) whose kernels
marshal the reference's positional argument list into the C ABI (`srb_grid`, `srb_tracks`, include/synchrad_b200.h)
and call `srb_integrate_host` -- one particle per call, exactly the reference's launch granularity:

    PYTHONPATH=<repo>/synchrad_b200/compat:<reference checkout>:<repo>  python tests/test_undulator_analytic.py

`pyopencl.array` arrays live in host memory; the library uploads, integrates on the B200 and adds into them.
This is the COMPATIBILITY path (a launch, a cudaMalloc and two copies per particle, like the reference's own
loop); the fast path is `synchrad_b200.calc.SynchRad`, which packs all tracks and launches once.
dtype='float' maps to the all-fp32 reproduction of the reference kernels (SRB_DTYPE_F32_LITERAL).
"""
