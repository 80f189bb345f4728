"""uforecon_b200 - B200-native per-ray rendering hot path of UFORecon (drop-in for code1/model.py's infer).

Layout: ``csrc/`` holds the CUDA kernels and the C ABI (``include/uforecon_b200.h``); ``renderer`` and
``costvolume`` mirror the reference's Python interface on top of it; ``checkpoint`` is the state-dict
layout contract; ``synthetic`` generates DTU-shaped inputs; ``dist`` shards rays/images over GPUs.
"""
__version__ = "0.1.0"
