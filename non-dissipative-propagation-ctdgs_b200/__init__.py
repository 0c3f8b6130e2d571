"""ctan_b200 -- B200-native (sm_100a) implementation of CTAN's per-event-batch hot path.

Import as ``ctan_b200`` (see ``/ctan_b200.py``).  Sub-modules are imported lazily so that
host-only helpers (``synth``) work without a GPU; anything that computes needs the C-ABI
library ``csrc/libctan_sm100.so`` and raises loudly when it is missing."""
__version__ = "0.1.0"
