"""CPU oracle for the DeFlow hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU
baseline.  ``deflow_b200`` never imports this package.

Parity status
-------------
* ``mmcv_ext_oracle`` restates the three ``mmcv._ext`` functions on the path
  (reference: OpenSceneFlow/assets/cuda/mmcv/*.cu, *.cuh).  The reference ships
  **no tests and no golden vectors** for them, and the
  extension registers CUDA kernels only, so it cannot execute in the (GPU-less)
  build container.  It is **pinned** against outputs of the reference's own CUDA
  extension, compiled from ``/root/reference`` into ``oracle/_ref`` by
  ``oracle/build_ref.py`` and run on a B200: ``tests/golden/make_ext_golden_gpu.py``
  -> ``tests/golden/ext_gpu_ref.npz`` (checked on CPU by
  ``tests/test_oracle_ext_golden.py``); the CUDA path is additionally compared
  with those reference kernels live on the GPU (``tests/test_gpu_vs_reference_ext.py``).
* ``deflow_oracle`` restates the pure-PyTorch part of the path as plain fp32
  functional torch code.  It is pinned against the reference's own Python
  modules imported from ``/root/reference`` in the build container
  (``tests/golden/make_golden.py`` -> ``tests/golden/deflow_*.npz``).
"""
