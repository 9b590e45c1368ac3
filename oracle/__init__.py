"""CPU oracle for the embedding-evaluation hot path of sMedX/FaceNet.

TEST INFRASTRUCTURE ONLY.  Nothing in ``facenet_b200/`` (the product) may
import this package.  The only callers allowed are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker or as the timed CPU
baseline -- never as a fallback for the CUDA path.

Contents
--------
``statistics_oracle``   NumPy restatement of ``facenet/statistics.py`` (reference
                        lines cited per function): the literal per-class-pair
                        loop and a vectorised form of the same arithmetic.
``mining_oracle``       NumPy statement of the repo-defined triplet-mining
                        semantics (the reference fork has no mining code, see
                        SURVEY.md section 0 R1) -- PARITY UNPINNED by reference.
``reference_loader``    Loads ``/root/reference/facenet/statistics.py`` UNMODIFIED
                        under a stub ``facenet`` package (build container only;
                        the path does not exist on the GPU box).
``gen_golden``          Script that ran the literal reference here and wrote
                        ``tests/golden/*.npz``.

Pinning status: the reference ships no tests and no golden vectors
(SURVEY.md section 0 R8).  The restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF executed in the build container (``gen_golden.py`` ->
``tests/golden``) with one shim: ``scipy.interpolate.interp1d(kind='slinear')``
raises on duplicate abscissae under scipy 1.18 (pinned scipy was 1.4.1), so
that single call is restated as ordinary piecewise-linear interpolation
(SURVEY.md section 8c).  Mining: parity unpinned.
"""
