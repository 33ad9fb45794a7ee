"""
oracle/ -- CPU restatement of the reference's NMF fitting hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and there only as the checker / the thing timed as "the reference's CPU path".
Nothing under ``salamander_b200/`` imports this package; the product path fails
loudly when the CUDA library is missing.

Everything here is plain numpy float64, written from the mathematics of the
reference (parklab/Salamander v0.4.2), each function citing the reference
``file:line`` it restates.  Parity is PINNED: ``tests/test_oracle_golden.py``
checks every function against the reference's own golden fixtures
(``tests/golden/`` = the reference's ``tests/test_data``), and
``tests/golden/trajectories/`` holds multi-iteration outputs of the live
reference generated in the build container by ``oracle/make_golden.py``.
"""

import numpy as np

#: clip constant used everywhere on the path: float32 machine epsilon
#: (reference: models/_utils_klnmf.py:7, utils.py:13, initialization/initialize.py:30)
EPSILON = float(np.finfo(np.float32).eps)
