"""amep_b200 -- B200-native RDF and static structure factor for AMEP trajectories.

Drop-in for one hot path of AMEP 1.3.0: ``amep.spatialcor.rdf / sfiso / sf2d``
and ``amep.evaluate.RDF / SFiso / SF2d``, plus the first function of the next
row of pair observables, ``spatialcor.pcf2d`` / ``evaluate.PCF2d``.  The work runs in hand-written
sm_100a CUDA kernels behind a C ABI (include/amepb.h, amep_b200/lib/libamepb.so);
there is no CPU fallback.
"""
from . import dist, evaluate, spatialcor, utils  # noqa: F401
from ._lib import AmepbError, LIB_PATH  # noqa: F401
from .trajectory import TensorFrame, TensorTrajectory  # noqa: F401

__version__ = "0.1.0"
