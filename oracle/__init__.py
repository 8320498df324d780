"""CPU oracle for the qmps classical hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline.  The product (``qmps_b200``) never
imports this package and fails loudly when its CUDA library is missing.

The oracle is a numpy/scipy restatement (complex128) of the reference's
algorithm for the path SURVEY.md section 8 names, each function citing the
reference ``file:line`` it follows.

Parity pinning status (see DESIGN.md section "Oracle"):

* PINNED against the reference's own code run in the build container
  (``oracle/make_golden.py`` imports ``/root/reference/qmps/tools.py`` and
  ``qmps/loschmidts/exact_loschmidt.py`` under stub modules and records their
  outputs in ``tests/golden/``): ``unitary_to_tensor``, ``tensor_to_unitary``,
  ``unitary_extension``, ``environment_to_unitary``,
  ``environment_from_unitary``, ``from_real_vector``/``to_real_vector``,
  ``direct_sum``, ``cT``, ``double_rotosolve``, ``get_env_exact`` (with the
  un-vendored ``TransferMatrix`` supplied by this oracle), the exact TFIM
  Loschmidt rate function.
* PINNED against literals the reference's tests hold: the TFIM matrix of
  ``tests/test_ground_state.py:29-38``, the exact ``E0(g)`` integral
  ``tests/test_ground_state.py:101-102``, ``D2_gse`` of
  ``scripts/noisy_optimization.py:93``, the known-answer environment of
  ``new_tdvp/testTDVPStripped.py:156-170``.
* PINNED, energy route: ``energy_of_unitary`` / ``energy_transfer`` against the reference's own cirq-free
  script ``scripts/ground_state_finding.py:83-128`` run under stubs (``oracle/make_golden_gs.py`` ->
  ``tests/golden/ref_ground_state_script.npz``; Pauli / CNOT matrices and ``TransferMatrix`` supplied).
* PINNED, definitions cut out of reference modules with ``ast`` and executed unmodified
  (``oracle/make_golden_misc.py`` -> ``tests/golden/ref_misc.npz``): ``merge``, ``put_env_on_*_site``,
  ``get_env_off_*``, ``Hamiltonian.to_matrix``, ``rotosolve`` / ``double_rotosolve`` of qmps/rotosolve.py.
* PINNED, ansatz gate lists: the reference's own ``_decompose_`` methods (qmps/represent.py:268-442) run
  against a recording cirq stand-in (``oracle/make_golden_gates.py`` -> ``tests/golden/ref_gate_lists.json``);
  the gate matrices themselves (cirq conventions) remain stated in ``oracle/gates.py``.
* PINNED, Loschmidt / TDVP-step cost: the reference's own ``obj(p, A, WW)`` (qmps/loschmidts/time_evo.py:75-116)
  executed unmodified on a minimal cirq stand-in with the oracle's xmps parts (``oracle/make_golden_obj.py`` ->
  ``tests/golden/ref_loschmidt_obj.npz``).
* PINNED, brick-wall family (``oracle/brickwall.py``): every function against outputs of the
  reference's unmodified ``new_tdvp/ClassicalTDVPStripped.py`` run under stub modules
  (``oracle/make_golden_bw.py`` -> ``tests/golden/ref_brickwall.npz``).
* PARITY UNPINNED: everything whose arithmetic lives in the un-vendored,
  un-pinned third-party packages ``xmps`` (``TransferMatrix.eigs``,
  ``Map.right_fixed_point``/``left_fixed_point``, ``iMPS.left_canonicalise``,
  ``iMPS.overlap``, ``iMPS.mixed``, ``iMPS.Es`` -- ``oracle/canonical.py``) and ``cirq`` (gate matrices, circuit simulation).  Those
  are restated from their published definitions and anchored on the
  reference's call sites (eigen-equation, Hermitian PD ``r``, unit-Frobenius
  fixed points, big-endian qubit order) -- see SURVEY.md A.2 / A.5.
"""

from .tensors import *      # noqa: F401,F403
from .gates import *        # noqa: F401,F403
from .costs import *        # noqa: F401,F403
from .canonical import *    # noqa: F401,F403
from .brickwall import *    # noqa: F401,F403
from .stacked import *      # noqa: F401,F403
from .tdvp import *         # noqa: F401,F403
from .scars import *        # noqa: F401,F403
