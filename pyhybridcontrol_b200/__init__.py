"""pyhybridcontrol_b200 -- B200-native (sm_100a) per-step hybrid-MPC solve behind the pyhybridcontrol API.

Only the hot path named in BASELINE.json:north_star lives here: batched MLD condensing, the batched
mixed-integer solve over the binaries, the batched MLD simulation step and the aggregate-power exchange.
The CUDA kernels are reached through the C-ABI library ``csrc/libhmpc.so`` (see include/hmpc.h); there is
no CPU fallback -- importing ``pyhybridcontrol_b200.cabi`` without the built library raises.
"""
__version__ = "0.1.0"
