"""adapt_b200 -- B200-native wavefront path tracer behind AdaPT's `pt` renderer interface.

Host side (this package): AdaPT-compatible XML/OBJ scene ingestion (``parsers``), emitter / BxDF
descriptors (``emitters``, ``bxdf``), camera maths (``la``) and the ``Renderer`` class
(``renderer.vanilla_renderer``) that the reference's ``render.py`` drives.  Device side:
``csrc/`` -- hand-written sm_100a CUDA kernels behind the C ABI of ``include/adapt_b200.h``,
loaded with ctypes (``_lib``).  There is no CPU fallback in this package.
"""
__version__ = "0.1.0"
