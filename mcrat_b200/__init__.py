"""mcrat_b200 -- B200-native (sm_100a, FP64, hand-written CUDA) implementation of the
photon-propagation / scattering hot path of MCRaT behind the reference's own C function
surface.  The product is ``csrc/libmcrat_b200.so`` and its C ABI (``include/mcrat_b200.h``);
this package holds the Python harness around it: a ctypes binding (:mod:`mcrat_b200.lib`),
synthetic BASELINE workloads (:mod:`mcrat_b200.synth`) and shard helpers
(:mod:`mcrat_b200.shard`)."""
from . import synth  # noqa: F401
from .lib import Comm, HotPath, McratB200Error, build, comm_unique_id, load  # noqa: F401
