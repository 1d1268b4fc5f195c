"""ORACLE — test infrastructure only.

CPU restatement of the reference's algorithm for the face path (pyannote/video/face/face.py,
tracking.py, face/clustering.py and the dlib / scipy / pyannote.algorithms internals they call).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package — as the checker or the timed CPU baseline, never as the product path.
PARITY UNPINNED for the dlib internals (dlib 19.12, its weights and its tests are absent here);
pinned where the reference's own Python (tracking.py) or scipy (pdist, linkage) can be executed.
"""
