"""facenet_b200 -- B200-native embedding-evaluation hot path of sMedX/FaceNet.

``facenet_b200.statistics`` mirrors ``facenet/statistics.py`` of the reference;
``facenet_b200.mining`` provides the triplet-mining entry point; ``facenet_b200._capi`` is the
ctypes binding of ``include/facenet_b200.h``.  Build the CUDA library with
``python -m facenet_b200.build`` (or ``__graft_entry__.build()``).
"""
__version__ = '0.1.0'
