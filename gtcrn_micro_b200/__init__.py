"""gtcrn_micro_b200 -- B200-native (sm_100a) inference path for GTCRN-Micro.

Drop-in for the reference's `GTCRNMicro` / `StreamGTCRNMicro` / `convert_to_stream` (same constructor,
forward contract and state_dict keys) in front of hand-written CUDA kernels reached through the C ABI of
`libgtcrn_b200.so` (include/gtcrn_b200.h).  Build the library with `python -m gtcrn_micro_b200.build`.
"""
from .lib import GtcrnB200Error  # noqa: F401
from .modules import (GTCRNMicro, StreamGTCRNMicro, StreamPool, StreamState, convert_to_stream,  # noqa: F401
                      load_checkpoint, load_state_dict_npz, param_spec, pinned_empty)
from .full import GTCRN, GTCRNFull, StreamGTCRN, StreamGTCRNFull  # noqa: F401
from .pack import blob_size, pack_state_dict  # noqa: F401

__all__ = ["GTCRN", "GTCRNFull", "StreamGTCRN", "StreamGTCRNFull", "GTCRNMicro", "StreamGTCRNMicro", "StreamPool", "StreamState", "convert_to_stream", "load_checkpoint",
           "load_state_dict_npz", "pack_state_dict", "blob_size", "param_spec", "pinned_empty", "GtcrnB200Error"]
