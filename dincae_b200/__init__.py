"""dincae_b200 -- B200-native (sm_100a) implementation of DINCAE's training and
reconstruction hot path behind the reference's own Python API.

    import dincae_b200 as DINCAE        # or: import DINCAE (shim package in this repo)
    DINCAE.reconstruct_gridded_nc(filename, varname, outdir)

Importing the package does not need a GPU; calling any compute entry point does, and it
fails loudly if the CUDA extension (libdincae_b200.so) is missing -- there is no CPU path.
"""
from .api import (FrameSource, NEAREST_NEIGHBOR, data_generator, data_generator_list, identity,
                  load_checkpoint, load_gridded_nc, reconstruct, reconstruct_gridded_files,
                  reconstruct_gridded_nc, save_checkpoint, savesample, set_precision, get_precision, set_dropout_active, sinv)
from .ensemble import EnsembleAccumulator, average_reconstructions

__all__ = ["reconstruct", "load_gridded_nc", "data_generator", "reconstruct_gridded_nc"]
__version__ = "0.1.0"
