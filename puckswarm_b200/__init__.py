"""puckswarm_b200 -- B200-native batched engine for PuckSwarm's per-tick hot path (RunOffline.step).

csrc/    CUDA kernels (sm_100a) + the C-ABI of include/puckswarm_b200.h (libpuckswarm_b200.so)
abi.py   ctypes mirror of the header; engine.py the 1:1 binding; host.py / host/ the mirrors of the reference's
         Experiment / Arena / RunOffline / Suite / Controller interfaces; dist.py arena sharding + statistics reduce
There is no CPU path: without the built library and a B200 every entry point fails loudly.
"""
