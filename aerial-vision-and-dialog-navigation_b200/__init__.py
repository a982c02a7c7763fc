"""B200-native hot path of AVDN's HAA-Transformer (import as ``avdn_b200``).

Host side: Python + PyTorch for device memory, streams, autograd plumbing and
``torch.distributed``.  Device side: hand-written sm_100a CUDA kernels in
``csrc/`` behind the C ABI declared in ``include/avdn.h`` (``csrc/libavdn.so``,
loaded with ctypes by ``_lib``).  There is no CPU fallback: every op raises if
the library or an sm_100 device is missing.
"""
__version__ = "0.1.0"
