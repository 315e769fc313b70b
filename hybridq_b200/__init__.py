"""hybridq_b200 -- B200-native (sm_100a) state-vector evolution core for HybridQ.

Importing this package loads ``hybridq_b200/lib/libhybridq_b200.so`` and raises if it is
missing: there is no CPU fallback.  Public surface:

* :func:`simulate`            mirror of ``hybridq.circuit.simulation.simulate(optimize='evolution')``
* :func:`expectation_value`   mirror of ``hybridq.circuit.simulation.expectation_value`` (device reduction)
* :func:`dot`, :func:`transpose`  mirrors of ``hybridq.utils.dot`` / ``hybridq.utils.transpose``
* :class:`DeviceState`, :class:`Plan`, :class:`PlanOptions`  device-resident state and circuit plans
* :mod:`hybridq_b200.circuits`  seeded synthetic circuits and the ``GateApply`` gate stand-in
* :data:`DROPIN_DIR`          directory to put on ``LD_LIBRARY_PATH`` so that the unmodified
                              reference loads this library as ``hybridq.so`` / ``hybridq_swap.so``
"""
from ._lib import lib, PlanOptions, HybridQB200Error, DROPIN_DIR, LIBPATH  # noqa: F401
from .state import DeviceState, Plan, BitPermPlan  # noqa: F401
from .simulate import simulate, expectation_value, clear_caches, sharded_runner  # noqa: F401
from .dot import dot, transpose, to_complex  # noqa: F401
from . import circuits  # noqa: F401

__version__ = "0.1.0"
